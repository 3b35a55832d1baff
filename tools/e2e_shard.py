"""One-shot render_b200 of ONE shard of an N-way split (what every rank of a torchrun job does), 8-bit plane into pinned
memory:   python tools/e2e_shard.py c3 8      (RTC_TIMING=1 for the phases)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402
from ray_tracer_challenge_b200.api import U8P  # noqa: E402

api = rt.new_session()
api.set_render_options(device_ids=[0])
lib = rt.device_library()
lib.rtc_host_alloc.restype = C.c_void_p
lib.rtc_host_alloc.argtypes = [C.c_size_t]
w_name, n_sh = sys.argv[1], int(sys.argv[2])
cam, world, depth, _ = build_scene(api, w_name)
w, h = cam.width_pixels, cam.height_pixels
u8 = np.ctypeslib.as_array(C.cast(lib.rtc_host_alloc(w * h * 3), C.POINTER(C.c_uint8)), shape=(h, w, 3))
stats = rt.SgStats()


def one():
    api.check(api.lib.sg_camera_render_shard(api.ctx, cam.handle, world.handle, depth, 0, n_sh, None, u8.ctypes.data_as(U8P), C.byref(stats)))


for _ in range(5):
    one()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    one()
dt = (time.perf_counter() - t0) / n * 1e3
st = api.last_rtc_stats()
print(f"{w_name} shard 0 of {n_sh}: one-shot e2e {dt:.3f} ms (kernel span {st.kernel_ms:.3f} ms, {st.launches} launches)", flush=True)
