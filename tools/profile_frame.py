"""Minimal driver for ncu: build a bench workload, render it `--frames` times with the frame left on the device.
    python tools/profile_frame.py --workload c3 --frames 3 [--fma] [--detailed]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import WORKLOADS, build_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--fma", action="store_true")
ap.add_argument("--detailed", action="store_true")
a = ap.parse_args()
api = rt.new_session()
api.set_render_options(device_ids=[0], fma=a.fma)
cam, world, depth, desc = build_scene(api, a.workload)
p = cam.prepare(world)
for i in range(a.frames):
    p.render(depth, want_rgb=False, want_u8=False, fma=a.fma, detailed=a.detailed)
    st = p.last_stats
    print(f"frame {i}: {st.kernel_ms:.3f} ms kernel, {st.rays} rays, {st.rays / st.kernel_ms / 1e3:.1f} Mrays/s", flush=True)
if a.detailed:
    print(st.as_dict())
p.release()
