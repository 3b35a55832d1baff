"""Kernel time of one shard of an N-way split with and without the L2 flush bench.py does between iterations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
cam, world, depth, desc = build_scene(api, sys.argv[1] if len(sys.argv) > 1 else "c3")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
p = cam.prepare(world)
for flush in (False, True, False, True):
    ts = []
    for i in range(12):
        if flush:
            p.flush_l2()
        p.render(depth, want_rgb=False, want_u8=False, shard=0, n_shards=n)
        ts.append(p.last_stats.kernel_ms)
    print("flush" if flush else "no flush", " ".join(f"{t:.4f}" for t in ts), flush=True)
