"""The drop-in boundary: every entry point declared in include/*.h is exported by the library that implements it
(loaded with ctypes — no compute calls, no GPU needed), and the product fails loudly without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", text)))


def test_device_library_exports_the_whole_c_abi():
    import ray_tracer_challenge_b200 as rt

    names = declared("rtc_b200.h", "rtc_")
    assert len(names) >= 20 and "rtc_render" in names and "rtc_scene_commit" in names
    lib = C.CDLL(rt.LIB_DEVICE)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"librtc_b200.so lacks {missing}"


def test_host_library_exports_the_scene_api():
    import ray_tracer_challenge_b200 as rt

    names = declared("rtc_scene.h", "sg_")
    assert "sg_camera_render" in names and "sg_group_add_child" in names
    lib = C.CDLL(rt.LIB_HOST)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"librtc_host.so lacks {missing}"


def test_oracle_exports_the_same_scene_api(oracle):
    names = declared("rtc_scene.h", "sg_")
    missing = [n for n in names if not hasattr(oracle.lib, n)]
    assert not missing, f"the oracle lacks {missing}"


def test_libraries_do_not_link_the_oracle():
    """The product must never route through the CPU oracle: neither library may depend on it."""
    import subprocess

    import ray_tracer_challenge_b200 as rt

    for lib in (rt.LIB_DEVICE, rt.LIB_HOST):
        out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in out, out
        syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
        assert "orc_" not in syms


def test_shard_band_rule_matches_python_mirror():
    import ray_tracer_challenge_b200 as rt
    from ray_tracer_challenge_b200.sharding import BAND_ROWS, bands_of, rows_of

    lib = C.CDLL(rt.LIB_DEVICE)
    lib.rtc_shard_bands.argtypes = [C.c_uint32, C.c_int32, C.c_int32, C.POINTER(C.c_uint32)]
    for height in (1, 7, 8, 9, 77, 400, 2160):
        for n in (1, 2, 3, 8):
            seen = []
            for shard in range(n):
                buf = (C.c_uint32 * 512)()
                cnt = lib.rtc_shard_bands(height, shard, n, buf)
                assert cnt == len(bands_of(height, shard, n))
                assert [buf[i] for i in range(cnt)] == [b * BAND_ROWS for b in bands_of(height, shard, n)]
                for r in rows_of(height, shard, n):
                    seen.extend(r)
            assert sorted(seen) == list(range(height)), (height, n)  # every row exactly once
    assert lib.rtc_shard_bands(100, 2, 2, None) < 0  # bad shard index is an error, not a silent no-op


def test_no_device_is_a_loud_error():
    """Without a GPU the product refuses to render: there is no CPU fallback."""
    import ray_tracer_challenge_b200 as rt

    lib = C.CDLL(rt.LIB_DEVICE)
    if lib.rtc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    session = rt.new_session()
    cam = session.Camera(8, 8, 1.0, session.identity_4x4())
    world = session.World.default()
    with pytest.raises(rt.RtcError):
        cam.render_b200(world, 1)
    with pytest.raises(rt.RtcError):
        cam.prepare(world)
