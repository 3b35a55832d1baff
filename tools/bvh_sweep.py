"""Sweep the BVH leaf size on a bench workload (device-resident frames)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c5"
lib = rt.device_library()
api = rt.new_session()
cam, world, depth, desc = build_scene(api, workload)
for leaf in (1, 2, 4, 8):
    os.environ["RTC_BVH_LEAF"] = str(leaf)
    p = cam.prepare(world)
    for i in range(3):
        p.render(depth, want_rgb=False, want_u8=False)
    ms = p.last_stats.kernel_ms
    p.render(depth, want_rgb=False, want_u8=False, detailed=True)
    st = p.last_stats
    print(f"{workload} leaf={leaf}: {ms:.3f} ms  nodes/ray {st.node_visits / st.rays:.1f}  prim tests/ray {sum(st.prim_tests) / st.rays:.2f}",
          flush=True)
    p.release()
