"""How much of a 1/8-frame shard's kernel time is cold-start (L2 flushed: code, constants and the scene refetched
from DRAM) — the term that limits strong scaling at 8 GPUs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt
from bench import build_scene
api = rt.new_session()
cam, world, depth, _ = build_scene(api, sys.argv[1] if len(sys.argv) > 1 else "c3")
p = cam.prepare(world)
for n in (1, 2, 4, 8):
    for flush in (False, True):
        ts = []
        for i in range(12):
            if flush:
                p.flush_l2()
            p.render(depth, want_rgb=False, want_u8=False, shard=0, n_shards=n)
            ts.append(p.last_stats.kernel_ms)
        ts = sorted(ts[2:])
        print(f"shards={n} flush={flush}: median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f}  (x{n} = {ts[len(ts)//2]*n:.3f})", flush=True)
p.release()
