import sys, time
sys.path.insert(0, ".")
import ray_tracer_challenge_b200 as rt
from bench import build_scene
api = rt.new_session(); cam, world, depth, _ = build_scene(api, "c3"); p = cam.prepare(world)
for i in range(3): p.render(depth, want_rgb=False, want_u8=False)
for label, fl in (("noflush", False), ("flush", True)):
    t0 = time.perf_counter(); k = 0
    for i in range(20):
        if fl: p.flush_l2()
        p.render(depth, want_rgb=False, want_u8=False); k += p.last_stats.kernel_ms
    dt = (time.perf_counter() - t0) * 1e3 / 20
    print(label, "wall/step %.3f ms, kernel %.3f ms, total_ms(last) %.3f" % (dt, k / 20, p.last_stats.total_ms))
