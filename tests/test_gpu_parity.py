"""GPU parity tests proper: the CUDA path (through librtc_host.so -> the C ABI of include/rtc_b200.h) against
the CPU oracle on the same scenes.  Run on the B200 box: `pytest tests -m gpu`."""
import json
import os

import numpy as np
import pytest

from ray_tracer_challenge_b200 import scenes
from tests.parity import assert_parity, assert_unrendered_border, compare_frames

pytestmark = pytest.mark.gpu

SCENES = {
    "default_world": (scenes.default_world, dict(width=64, height=48)),
    "soft_shadows_table": (scenes.soft_shadows, dict(width=250, height=100, u_steps=4, v_steps=4)),
    "soft_shadows_constant": (scenes.soft_shadows, dict(width=200, height=80, u_steps=3, v_steps=2, jitter="constant")),
    "soft_shadows_counter_rng": (scenes.soft_shadows, dict(width=200, height=80, u_steps=3, v_steps=3, jitter=None, seed=7)),
    # as shipped (soft_shadows.rs:72-82): 10 x 10 cells, jitter None -> four 32-cell chunks of drawn samples
    "soft_shadows_as_shipped": (scenes.soft_shadows, dict(width=150, height=60, u_steps=10, v_steps=10, jitter=None, seed=3)),
    "reflect_refract": (scenes.reflect_refract, dict(width=320, height=160)),
    "reflect_refract_csg": (scenes.reflect_refract, dict(width=240, height=120, with_csg=True)),
    "hexagons": (scenes.hexagons, dict(width=240, height=120)),
    "shapes_zoo": (scenes.shapes_zoo, dict(width=320, height=200)),
    "shapes_zoo_area": (scenes.shapes_zoo, dict(width=200, height=120, area_light=True)),
    "csg_gallery": (scenes.csg_gallery, dict(width=320, height=200)),
    "dragon_element": (scenes.dragon_element, dict(width=240, height=135, n_u=32, n_v=16)),
    "dragon_smooth_nodivide": (scenes.dragon_element, dict(width=160, height=90, n_u=16, n_v=8, divide=0, smooth=True)),
    "textured": (scenes.textured, dict(width=320, height=200)),
    # the remaining camera demos of demos/src/bin/ at reduced size (full-size c1-c4 are in test_gpu_full_size.py)
    "first_scene": (scenes.first_scene, dict(width=300, height=150)),
    "first_plane": (scenes.first_plane, dict(width=100, height=50)),
    "first_patterns": (scenes.first_patterns, dict(width=200, height=100)),
    "first_textures": (scenes.first_textures, dict(width=300, height=150, u_steps=4, v_steps=4)),
    "skybox": (scenes.skybox, dict(width=320, height=160)),
    "here_be_dragons": (scenes.here_be_dragons, dict(width=300, height=120)),
    "filter_zoo_area": (scenes.filter_zoo, dict(width=320, height=200)),
    "filter_zoo_table": (scenes.filter_zoo, dict(width=200, height=120, jitter="table")),
    "filter_zoo_point": (scenes.filter_zoo, dict(width=320, height=200, area_light=False)),
    "stress_small": (scenes.stress, dict(width=240, height=135, n_spheres=3000, n_each=8, n_csg=6)),
}


# Scenes whose reference arithmetic is itself noise-dominated somewhere (the f32 discriminant of a far, tiny
# sphere; the CSG normal bug of SURVEY Q7 that puts over_point inside the surface): only a bit-exact evaluation
# reproduces them, so the FMA-contracted build is reported but not held to the 99.9 % bar there.
ILL_CONDITIONED = {"stress_small", "csg_gallery", "reflect_refract_csg"}


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


@pytest.fixture(scope="module")
def report():
    rows = {}
    yield rows
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as fh:
        json.dump(rows, fh, indent=1, sort_keys=True)


@pytest.mark.parametrize("name", sorted(SCENES))
def test_frame_parity(name, gpu, oracle, report):
    build, kw = SCENES[name]
    oracle.probe.set_threads(oracle.probe.max_threads())
    ocam, oworld = build(oracle, **kw)
    want = ocam.render(oworld, 5)
    gcam, gworld = build(gpu, **kw)
    prepared = gcam.prepare(gworld)
    try:
        # ---- the product path (IEEE build, the default): bit-for-bit the oracle apart from libm
        got = prepared.render(5)
        rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
        st = prepared.last_stats
        rep["rays_gpu"], rep["rays_oracle"] = int(st.rays), int(ocam.last_stats.rays)
        report[f"{name}:ieee"] = rep
        assert_unrendered_border(got.to_u8(), got.data)
        assert_parity(rep, min_within=0.9999, max_gross=0.0001, label=f"{name} (IEEE build)")
        assert rep["rays_gpu"] == rep["rays_oracle"], rep  # identical ray trees
        assert rep["bit_exact_f32"] >= 0.90, rep           # the rest differs only through powf / cosf / atan2f
        # ---- the optional FMA-contracted build: the documented tolerance, on well-conditioned scenes only
        got = prepared.render(5, fma=True)
        rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
        rep["rays_gpu"], rep["rays_oracle"] = int(prepared.last_stats.rays), int(ocam.last_stats.rays)
        report[f"{name}:fma"] = rep
        assert_unrendered_border(got.to_u8(), got.data)
        if name not in ILL_CONDITIONED:
            assert_parity(rep, min_within=0.998, label=f"{name} (FMA build)")  # optional build: contraction patterns vary with inlining
    finally:
        prepared.release()


def test_one_shot_render_b200_matches_prepared(gpu, oracle):
    """Camera::render_b200 (flatten + commit + render + release in one call) gives the same canvas."""
    cam, world = scenes.reflect_refract(gpu, width=160, height=80)
    a = cam.render_b200(world, 5)
    p = cam.prepare(world)
    b = p.render(5)
    p.release()
    assert np.array_equal(a.data.view(np.uint32), b.data.view(np.uint32))
    assert np.array_equal(a.to_u8(), b.to_u8())
    st = cam.last_rtc_stats
    assert st.primary_rays == 159 * 79 and st.kernel_ms > 0


def test_band_sharding_is_bit_identical(gpu):
    """N interleaved-band shards reassemble to exactly the single-launch frame."""
    cam, world = scenes.soft_shadows(gpu, width=203, height=77, u_steps=3, v_steps=3)
    p = cam.prepare(world)
    try:
        full = p.render(5)
        for n_shards in (2, 3, 8):
            rgb = np.full((77, 203, 3), np.nan, np.float32)
            u8 = np.full((77, 203, 3), 7, np.uint8)
            rays = 0
            for shard in range(n_shards):
                p.render(5, out_rgb=rgb, out_u8=u8, shard=shard, n_shards=n_shards)
                rays += p.last_stats.rays
            assert np.array_equal(rgb.view(np.uint32), full.data.view(np.uint32)), n_shards
            assert np.array_equal(u8, full.to_u8()), n_shards
        p.render(5)
        assert rays == p.last_stats.rays
    finally:
        p.release()


def test_longest_first_band_order_changes_no_pixel(gpu):
    """RTC_OPT_ADAPTIVE_ORDER: after the render that records the tile costs, a shard launches its tiles most-expensive-first."""
    cam, world = scenes.soft_shadows(gpu, width=203, height=177, u_steps=2, v_steps=2)
    p = cam.prepare(world)
    try:
        p.set_option(4, 1)  # RTC_OPT_RENDER_SLICES = 1: one launch, so the learnt order is used with a host canvas too
        for shard, n_shards in ((0, 0), (1, 3)):
            p.set_option(5, 0)
            ref = p.render(5, shard=shard, n_shards=n_shards)
            p.set_option(5, 1)
            # natural order recording the costs, natural again, two trial renders in the learnt order, the decided one
            frames = [p.render(5, shard=shard, n_shards=n_shards) for _ in range(5)]
            for got in frames:
                assert np.array_equal(got.data.view(np.uint32), ref.data.view(np.uint32))
                assert np.array_equal(got.to_u8(), ref.to_u8())
    finally:
        p.release()


FILTERED = ["default_world", "soft_shadows_table", "soft_shadows_constant", "soft_shadows_counter_rng", "soft_shadows_as_shipped",
            "filter_zoo_area",
            "filter_zoo_table", "filter_zoo_point"]


@pytest.mark.parametrize("name", FILTERED)
def test_shadow_filter_changes_no_pixel(name, gpu):
    """RTC_OPT_SHADOW_FILTER: the filter decides a shadow ray only when every comparison clears its error bound, so a
    frame is bit-identical with the filter on (default) and off (every shadow ray through the reference arithmetic)."""
    build, kw = SCENES[name]
    kw = dict(kw, width=kw["width"] * 2, height=kw["height"] * 2)
    cam, world = build(gpu, **kw)
    p = cam.prepare(world)
    try:
        for fma in (False, True):
            p.set_option(6, 0)
            off = p.render(5, fma=fma)
            rays_off = p.last_stats.rays
            p.set_option(6, 1)
            on = p.render(5, fma=fma, detailed=True)
            st = p.last_stats
            assert st.rays == rays_off
            assert np.array_equal(on.data.view(np.uint32), off.data.view(np.uint32)), (name, fma)
            assert np.array_equal(on.to_u8(), off.to_u8())
            # the filter is in use and hands only a small fraction of the shadow rays to the exact test
            fallbacks = st.prim_tests[7]
            assert fallbacks <= 0.05 * st.shadow_rays, (name, fallbacks, st.shadow_rays)
            p.set_option(6, 0)
            p.render(5, fma=fma, detailed=True)
            assert p.last_stats.prim_tests[7] == 0
            p.set_option(6, 1)
    finally:
        p.release()


def test_shadow_filter_fuzz(gpu, oracle):
    """Random filter-eligible scenes: the frame is bit-identical with the filter on and off, and (every third scene)
    identical to the oracle's in all 8-bit channels."""
    checked = 0
    for seed in range(24):
        cam, world = scenes.random_filter_scene(gpu, seed)
        p = cam.prepare(world)
        try:
            p.set_option(6, 0)
            off = p.render(5)
            rays_off = p.last_stats.rays
            p.set_option(6, 1)
            on = p.render(5, detailed=True)
            st = p.last_stats
            assert st.rays == rays_off, seed
            assert np.array_equal(on.data.view(np.uint32), off.data.view(np.uint32)), seed
            checked += st.prim_tests[7] < st.shadow_rays  # the filter decided something
            if seed % 3 == 0:
                ocam, oworld = scenes.random_filter_scene(oracle, seed)
                want = ocam.render(oworld, 5)
                rep = compare_frames(on.to_u8(), want.to_u8(), on.data, want.data)
                assert rep["exact_u8"] == 1.0 and st.rays == ocam.last_stats.rays, (seed, rep)
        finally:
            p.release()
    assert checked >= 20


@pytest.mark.parametrize("extreme", ["millimetre", "kilometre", "far light", "tiny spheres", "stretched", "grazing"])
def test_shadow_filter_at_scale_extremes(gpu, oracle, extreme):
    """The fuzz above in the regimes the filter's relative bounds must survive (scenes.random_filter_scene: whole scene
    x 1e-3 / x 1e3, light 1e4 away, radius-1e-3 spheres, condition ~60 ellipsoids, a light grazing the floor): on sm_100a
    the frame is bit-identical with the filter on and off AND equal to the oracle's in every 8-bit channel."""
    decided = 0
    for seed in range(300, 310):
        cam, world = scenes.random_filter_scene(gpu, seed, extreme=extreme)
        p = cam.prepare(world)
        try:
            p.set_option(6, 0)
            off = p.render(5)
            rays_off = p.last_stats.rays
            p.set_option(6, 1)
            on = p.render(5, detailed=True)
            st = p.last_stats
            assert st.rays == rays_off, (extreme, seed)
            assert np.array_equal(on.data.view(np.uint32), off.data.view(np.uint32)), (extreme, seed)
            decided += st.prim_tests[7] < st.shadow_rays
            ocam, oworld = scenes.random_filter_scene(oracle, seed, extreme=extreme)
            want = ocam.render(oworld, 5)
            rep = compare_frames(on.to_u8(), want.to_u8(), on.data, want.data)
            assert rep["exact_u8"] == 1.0 and st.rays == ocam.last_stats.rays, (extreme, seed, rep)
        finally:
            p.release()
    assert decided >= 5, (extreme, decided)


def test_depth_semantics(gpu, oracle):
    """remaining-depth guards (world.rs:126,140): depth 0 and 1 frames match the oracle."""
    for depth in (0, 1, 2):
        ocam, ow = scenes.reflect_refract(oracle, width=120, height=60)
        want = ocam.render(ow, depth)
        gcam, gw = scenes.reflect_refract(gpu, width=120, height=60)
        got = gcam.render_b200(gw, depth)
        assert_parity(compare_frames(got.to_u8(), want.to_u8()), label=f"depth {depth}")


def test_errors_are_loud(gpu):
    import ray_tracer_challenge_b200 as rt

    cam, world = scenes.default_world(gpu)
    with pytest.raises(rt.RtcError):
        cam.render_b200(world, 24)  # deeper than the device's bounce stack
    with pytest.raises(rt.RtcError):
        cam.render_b200(gpu.World([gpu.Sphere()]), 5)  # "World light should be set" (world.rs:66)


def test_render_b200_u8_is_the_8_bit_plane_of_render_b200(gpu, oracle):
    """The demos' render -> to_ppm flow without the f32 plane leaving the device: same bytes, same PPM."""
    cam, world = scenes.soft_shadows(gpu, width=160, height=64, u_steps=4, v_steps=4)
    full = cam.render_b200(world, 5)
    only_u8 = cam.render_b200_u8(world, 5)
    assert np.array_equal(only_u8.to_u8(), full.to_u8())
    assert only_u8.to_ppm() == full.to_ppm()
    ocam, oworld = scenes.soft_shadows(oracle, width=160, height=64, u_steps=4, v_steps=4)
    assert only_u8.to_ppm() == ocam.render(oworld, 5).to_ppm()


def test_in_process_multi_device_render_is_bit_identical(gpu):
    """rtc_scene_commit(scene, N, ids) + one rtc_render: N replicas in ONE process, bands interleaved over the devices,
    every device's copies queued after all devices have their kernels.  Needs >= 2 GPUs (skipped on a 1-GPU box; the
    one-process-per-GPU path is what bench.py and test_band_sharding_is_bit_identical exercise)."""
    import ctypes as C

    import ray_tracer_challenge_b200 as rt

    n = rt.device_library().rtc_device_count()
    if n < 2:
        pytest.skip("one GPU visible")
    one = rt.new_session()
    one.set_render_options(device_ids=[0])
    many = rt.new_session()
    many.set_render_options(device_ids=list(range(min(n, 4))))
    for name, kw in (("soft_shadows", dict(width=333, height=123, u_steps=4, v_steps=4)),
                     ("stress", dict(width=200, height=120, n_spheres=2000, n_each=4, n_csg=2))):
        cam1, w1 = getattr(scenes, name)(one, **kw)
        camn, wn = getattr(scenes, name)(many, **kw)
        a = cam1.render_b200(w1, 5)
        b = camn.render_b200(wn, 5)
        assert camn.last_rtc_stats.n_devices == min(n, 4)
        assert np.array_equal(a.data.view(np.uint32), b.data.view(np.uint32)), name
        assert np.array_equal(a.to_u8(), b.to_u8()), name
        assert cam1.last_rtc_stats.rays == camn.last_rtc_stats.rays
