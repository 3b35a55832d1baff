import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt
from bench import build_scene
api = rt.new_session()
cam, world, depth, _ = build_scene(api, "c3")
p = cam.prepare(world)
def t(shard, n):
    ts = []
    for i in range(8):
        p.render(depth, want_rgb=False, want_u8=False, shard=shard, n_shards=n)
        ts.append(p.last_stats.kernel_ms)
    return sorted(ts[2:])[len(ts[2:]) // 2], p.last_stats.rays
for n in (8, 16, 64):
    res = [t(k, n) for k in range(n)]
    print(f"n_shards={n}: per-shard ms min {min(r[0] for r in res):.4f} max {max(r[0] for r in res):.4f} sum {sum(r[0] for r in res):.3f}; rays min {min(r[1] for r in res)} max {max(r[1] for r in res)}")
# contiguous slabs: 270 bands -> 34 bands each, shard arithmetic with n_shards=1 cannot do that; use many shards to see per-band cost profile
n = 270
res = [t(k, n) for k in range(0, n, 10)]
print("per-band ms (every 10th band):", " ".join(f"{r[0]*1000:.0f}" for r in res))
p.release()
