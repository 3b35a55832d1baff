"""Kernel time of every shard of an N-way split, one after another on one GPU: how even is the split?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
w = sys.argv[1] if len(sys.argv) > 1 else "c3"
cam, world, depth, desc = build_scene(api, w)
p = cam.prepare(world)
for n in [int(a) for a in sys.argv[2:]] or [2, 4, 8]:
    best = []
    for shard in range(n):
        ts = []
        for i in range(10):
            p.flush_l2()
            p.render(depth, want_rgb=False, want_u8=False, shard=shard, n_shards=n)
            ts.append(p.last_stats.kernel_ms)
        best.append(sorted(ts)[len(ts) // 2])
    print(w, n, "max %.4f mean %.4f" % (max(best), sum(best) / n), " ".join(f"{t:.4f}" for t in best), flush=True)
