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


def test_raw_c_abi_validates_scenes_without_a_device():
    """The C ABI driven the way a foreign host (the Rust -sys crate) drives it: plain structs in, error codes out.
    rtc_scene_inspect runs the host half of the commit, so bad references are reported without a GPU."""
    import ray_tracer_challenge_b200 as rt

    lib = C.CDLL(rt.LIB_DEVICE)
    lib.rtc_last_error.restype = C.c_char_p
    scene = C.c_void_p()
    assert lib.rtc_scene_create(C.byref(scene)) == 0
    ident = (C.c_float * 16)(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)
    lib.rtc_set_camera.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]
    assert lib.rtc_set_camera(scene, 64, 32, 1.0, 0.5, 2.0 / 64, ident) == 0
    info = rt.RtcCommitInfo()
    # no light yet: the reference's own failure (world.rs:66), as an error code
    assert lib.rtc_scene_inspect(scene, C.byref(info)) < 0 and b"World light should be set" in lib.rtc_last_error()
    pos, rgb = (C.c_float * 3)(-10, 10, -10), (C.c_float * 3)(1, 1, 1)
    assert lib.rtc_set_point_light(scene, pos, rgb) == 0

    class RtcMaterial(C.Structure):
        _fields_ = [("color", C.c_float * 3), ("v", C.c_float * 7), ("pattern", C.c_int32)]

    class RtcPattern(C.Structure):
        _fields_ = [("kind", C.c_int32), ("mapping", C.c_int32), ("uv", C.c_int32 * 6), ("inv", C.c_float * 16),
                    ("a", C.c_float * 3), ("b", C.c_float * 3)]

    class RtcUvPattern(C.Structure):
        _fields_ = [("kind", C.c_int32), ("params", C.c_float * 15)]

    class RtcTexture(C.Structure):
        _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rgb", C.POINTER(C.c_float))]

    prim = rt.RtcPrim()
    prim.type, prim.material, prim.casts_shadow, prim.parent = 0, 0, 1, -1
    prim.inv[:] = list(ident)
    prim.bbox_min[:], prim.bbox_max[:] = [-1, -1, -1], [1, 1, 1]
    assert lib.rtc_set_primitives(scene, 1, C.byref(prim)) == 0
    mat = RtcMaterial()
    mat.color[:], mat.v[:], mat.pattern = [1, 1, 1], [0.1, 0.9, 0.9, 200.0, 0.0, 0.0, 1.0], 0
    assert lib.rtc_set_materials(scene, 1, C.byref(mat)) == 0
    pat = RtcPattern()
    pat.kind, pat.mapping = 6, 0  # RTC_PAT_TEXTURE_MAP, spherical
    pat.uv[:] = [0, -1, -1, -1, -1, -1]
    pat.inv[:] = list(ident)
    uv = RtcUvPattern()
    uv.kind = 2  # RTC_UV_IMAGE
    uv.params[0] = 3.0  # texture 3 of none
    assert lib.rtc_set_patterns(scene, 1, C.byref(pat), 1, C.byref(uv)) == 0
    assert lib.rtc_scene_inspect(scene, C.byref(info)) < 0 and b"texture" in lib.rtc_last_error()
    pixels = (C.c_float * 12)(*([0.5] * 12))
    tex = RtcTexture(2, 2, pixels)
    uv.params[0] = 0.0
    assert lib.rtc_set_patterns(scene, 1, C.byref(pat), 1, C.byref(uv)) == 0
    assert lib.rtc_set_textures(scene, 1, C.byref(tex)) == 0
    assert lib.rtc_scene_inspect(scene, C.byref(info)) == 0
    assert (info.small_n, info.n_positions, info.filter_ok, info.n_bvh_nodes) == (1, 1, 1, 0)
    mat.pattern = 5  # a pattern that does not exist
    assert lib.rtc_set_materials(scene, 1, C.byref(mat)) == 0
    assert lib.rtc_scene_inspect(scene, C.byref(info)) < 0
    if lib.rtc_device_count() == 0:
        assert lib.rtc_scene_commit(scene, 1, None) < 0 and b"no CPU fallback" in lib.rtc_last_error()
    lib.rtc_scene_destroy(scene)


def test_ctypes_mirrors_have_the_layout_of_the_header(tmp_path):
    """Every ctypes structure that crosses the C ABI has the size and the field offsets gcc gives the header's struct
    (the Rust `#[repr(C)]` mirrors in rust/rtc-b200-sys follow the same field lists)."""
    import shutil
    import subprocess

    import ray_tracer_challenge_b200 as rt

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    mirrors = [getattr(rt, n) for n in ("RtcStats", "RtcCommitInfo", "RtcPrim", "RtcNode", "RtcMaterial", "RtcPattern",
                                        "RtcUvPattern", "RtcTexture") if hasattr(rt, n)]
    assert len(mirrors) >= 4
    from ray_tracer_challenge_b200.api import SgStats

    c_names = {m: m.__name__ for m in mirrors}
    mirrors.append(SgStats)  # include/rtc_scene.h
    c_names[SgStats] = "sg_stats"
    lines = []
    for m in mirrors:
        lines.append(f'printf("{m.__name__} %zu", sizeof({c_names[m]}));')
        for field, _ in m._fields_:
            c_field = "type" if field == "type_" else field
            lines.append(f'printf(" %zu", offsetof({c_names[m]}, {c_field}));')
        lines.append('printf("\\n");')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rtc_b200.h"\n#include "rtc_scene.h"\nint main(void) {\n'
                   + "\n".join(lines) + "\nreturn 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for m, line in zip(mirrors, out):
        name, size, *offsets = line.split()
        assert name == m.__name__
        assert int(size) == C.sizeof(m), (name, size, C.sizeof(m))
        assert [int(o) for o in offsets] == [getattr(m, f).offset for f, _ in m._fields_], name
