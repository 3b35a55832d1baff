"""One-shot Camera::render_b200 of the big BASELINE scenes with either tree builder: wall time, kernel time, and (with
RTC_TIMING=1) the phases of the host half of the commit.   python tools/one_shot_times.py [c4 c5]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
for w in (sys.argv[1:] or ["c4", "c5"]):
    cam, world, depth, _ = build_scene(api, w)
    for mode in (0, 1):
        api.set_bvh_builder(mode)
        for i in range(3):
            if i == 2:
                print(f"---- {w} builder {mode}", file=sys.stderr, flush=True)
            t0 = time.perf_counter()
            cam.render_b200(world, depth, want_u8=True)
            dt = time.perf_counter() - t0
        st = cam.last_rtc_stats
        print(f"{w} builder {mode}: one-shot {dt * 1e3:.1f} ms (kernel {st.kernel_ms:.2f} ms, rtc_render total {st.total_ms:.1f} ms)", flush=True)
