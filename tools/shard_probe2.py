"""Kernel time of shard 0 of n for n = 1, 2, 4, 8 with the adaptive band order on / off (strong-scaling estimate on one GPU).
    python tools/shard_probe2.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt
from bench import build_scene
api = rt.new_session()
cam, world, depth, _ = build_scene(api, sys.argv[1] if len(sys.argv) > 1 else "c3")
p = cam.prepare(world)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for n in (1, 2, 4, 8):
    for adaptive in (0, 1):
        p.set_option(5, adaptive)
        ts = []
        for i in range(reps):
            p.render(depth, want_rgb=False, want_u8=False, shard=0, n_shards=n)
            ts.append(p.last_stats.kernel_ms)
        ts = sorted(ts[2:])
        print(f"shards={n} adaptive={adaptive}: median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f}  rays {p.last_stats.rays} (x{n} = {ts[len(ts)//2]*n:.3f})", flush=True)
p.release()
