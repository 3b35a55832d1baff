"""Does an idle gap before a render slow its kernel down?  Prepared scene, `gap` ms of host sleep before every render."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
w = sys.argv[1] if len(sys.argv) > 1 else "c4"
cam, world, depth, desc = build_scene(api, w)
p = cam.prepare(world)
for i in range(5):
    p.render(depth, want_rgb=False, want_u8=False)
for gap in (0, 1, 5, 20, 40, 100):
    for copies in (False, True):
        ts = []
        for i in range(6):
            time.sleep(gap * 1e-3)
            p.render(depth, want_rgb=False, want_u8=copies)
            ts.append(p.last_stats.kernel_ms)
        print(f"{w} gap {gap:3d} ms u8-copy {int(copies)}: kernel", " ".join(f"{t:.3f}" for t in ts), flush=True)
