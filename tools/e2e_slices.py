"""One-shot render_b200 (8-bit canvas into pinned memory) of a workload for several RTC_RENDER_SLICES values.
    python tools/e2e_slices.py c3 3 6 10 16"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 2:
    for n in sys.argv[2:]:
        subprocess.run([sys.executable, __file__, sys.argv[1]], env=dict(os.environ, RTC_RENDER_SLICES=n), check=True)
    sys.exit(0)
import numpy as np  # noqa: E402

import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402
from ray_tracer_challenge_b200.api import U8P  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
lib = rt.device_library()
lib.rtc_host_alloc.restype = C.c_void_p
lib.rtc_host_alloc.argtypes = [C.c_size_t]
cam, world, depth, _ = build_scene(api, sys.argv[1])
w, h = cam.width_pixels, cam.height_pixels
if os.environ.get("PAGEABLE"):  # what the reference-shaped API hands over: a plain heap array (Canvas is a Vec)
    u8 = np.zeros((h, w, 3), np.uint8)
else:
    u8 = np.ctypeslib.as_array(C.cast(lib.rtc_host_alloc(w * h * 3), C.POINTER(C.c_uint8)), shape=(h, w, 3))
stats = rt.SgStats()
for _ in range(5):
    api.check(api.lib.sg_camera_render_shard(api.ctx, cam.handle, world.handle, depth, 0, 1, None, u8.ctypes.data_as(U8P), C.byref(stats)))
t0 = time.perf_counter()
n = 20
for _ in range(n):
    api.check(api.lib.sg_camera_render_shard(api.ctx, cam.handle, world.handle, depth, 0, 1, None, u8.ctypes.data_as(U8P), C.byref(stats)))
print(f"{sys.argv[1]} slices={os.environ.get('RTC_RENDER_SLICES', 'default')}: one-shot e2e {(time.perf_counter() - t0) / n * 1e3:.3f} ms "
      f"(kernel span {api.last_rtc_stats().kernel_ms:.3f} ms, {api.last_rtc_stats().launches} launches)", flush=True)
