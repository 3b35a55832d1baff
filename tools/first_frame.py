"""First-frame kernel time of every shard of an N-way split, on ONE GPU (each shard = what one GPU of N would render):
a fresh commit per measurement, L2 flushed, the render's own CUDA-event time.  With RTC_PROXY_ORDER=0 / 1 and
RTC_ADAPTIVE_ORDER as set by the caller.   python tools/first_frame.py --workload c3 --shards 1 2 4 8"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import WORKLOADS, build_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
ap.add_argument("--shards", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
api = rt.new_session()
api.set_render_options(device_ids=[0])
cam, world, depth, desc = build_scene(api, a.workload)
warm = cam.prepare(world)
for _ in range(3):
    warm.render(depth, want_rgb=False, want_u8=False)
warm.release()
for n in a.shards:
    worst = []
    for rep in range(a.reps):
        times = []
        for k in range(n):
            p = cam.prepare(world)
            p.flush_l2()
            p.render(depth, want_rgb=False, want_u8=False, shard=k if n > 1 else 0, n_shards=n if n > 1 else 0)
            times.append(p.last_stats.kernel_ms)
            p.release()
        worst.append(max(times))
    print(f"{a.workload} proxy={os.environ.get('RTC_PROXY_ORDER', '1')} shards={n}: first-frame kernel ms, max over shards: "
          + " ".join(f"{t:.4f}" for t in worst), flush=True)
