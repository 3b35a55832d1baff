"""Where a one-shot Camera::render_b200_u8 (flatten + commit + render + copy into a pinned 8-bit canvas) spends its wall
time:  RTC_TIMING=1 python tools/one_shot_phases.py [c4 c5]   (phases on stderr, totals on stdout)"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
lib = rt.load_device_lib() if hasattr(rt, "load_device_lib") else C.CDLL(rt.LIB_DEVICE)
lib.rtc_host_alloc.restype = C.c_void_p
lib.rtc_host_alloc.argtypes = [C.c_size_t]
for w in (sys.argv[1:] or ["c4", "c5"]):
    cam, world, depth, _ = build_scene(api, w)
    W, H = cam.width_pixels, cam.height_pixels
    p = lib.rtc_host_alloc(W * H * 3)
    u8 = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(H, W, 3))
    for i in range(4):
        if i == 3:
            print(f"---- {w}", file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        cam.render_b200_u8(world, depth, out_u8=u8)
        dt = time.perf_counter() - t0
    st = cam.last_rtc_stats
    print(f"{w}: one-shot {dt * 1e3:.1f} ms (kernel {st.kernel_ms:.2f} ms, inside the library {st.total_ms:.1f} ms)", flush=True)
