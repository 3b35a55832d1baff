"""Edge cases of the render path on sm_100a, each against the oracle: degenerate canvases (the reference never renders
the last row and column, camera.rs:80-81 — a 1x1 frame renders nothing), ragged sizes (tiles cut by the right / bottom
edge, widths that are not a multiple of 8: the per-pixel store path), an empty world, a world of non-casters only, the
boundary between the small-scene table (16 items) and the tree (17), and the deepest recursion the device accepts."""
import math

import numpy as np
import pytest

from ray_tracer_challenge_b200 import scenes
from tests.parity import compare_frames

pytestmark = pytest.mark.gpu
PI = float(np.float32(math.pi))


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


def both(gpu, oracle, build, depth=5):
    cam, world = build(gpu)
    ocam, oworld = build(oracle)
    got = cam.render_b200(world, depth)
    want = ocam.render(oworld, depth)
    return got, want, cam.last_rtc_stats, ocam.last_stats


@pytest.mark.parametrize("w,h", [(1, 1), (2, 1), (1, 2), (2, 2), (3, 5), (7, 9), (8, 4), (9, 5), (17, 9), (33, 17), (100, 3)])
def test_degenerate_and_ragged_canvases(gpu, oracle, w, h):
    def build(rt):
        return scenes.soft_shadows(rt, width=w, height=h, u_steps=2, v_steps=2)
    got, want, st, ost = both(gpu, oracle, build)
    assert got.data.shape == (h, w, 3)
    assert not got.data[h - 1].any() and not got.data[:, w - 1].any()  # camera.rs:80-81
    assert st.primary_rays == (w - 1) * (h - 1) and st.rays == ost.rays
    rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
    assert rep["exact_u8"] == 1.0, rep


def test_empty_world_and_world_of_non_casters(gpu, oracle):
    def empty(rt):
        cam = rt.Camera(40, 24, PI / 3, rt.view_transform((0, 1, -5), (0, 1, 0), (0, 1, 0)))
        return cam, rt.World([], rt.PointLight((-10, 10, -10), (1, 1, 1)))
    got, want, st, ost = both(gpu, oracle, empty)
    assert not got.data.any() and st.shades == 0 and st.rays == ost.rays == 39 * 23

    def ghosts(rt):
        cam = rt.Camera(40, 24, PI / 3, rt.view_transform((0, 1, -5), (0, 1, 0), (0, 1, 0)))
        floor = rt.Plane.build(rt.identity_4x4(), rt.Material())
        ball = rt.Sphere.build(rt.translation(0, 1, 0), rt.Material(color=(1, 0.2, 0.2)))
        for s in (floor, ball):
            s.set_casts_shadow(False)
        return cam, rt.World([floor, ball], rt.PointLight((-10, 10, -10), (1, 1, 1)))
    got, want, st, ost = both(gpu, oracle, ghosts)
    rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
    assert rep["exact_u8"] == 1.0 and st.rays == ost.rays, rep


@pytest.mark.parametrize("n_spheres", [15, 16, 17, 40])
def test_small_table_to_tree_boundary(gpu, oracle, n_spheres):
    """16 top-level items still ride in the shared-memory table, 17 go to the tree: the frame must not notice."""
    def build(rt):
        cam = rt.Camera(96, 48, PI / 3, rt.view_transform((0, 3, -9), (0, 1, 0), (0, 1, 0)))
        objs = [rt.Plane.build(rt.identity_4x4(), rt.Material(reflective=0.2))]
        for i in range(n_spheres - 1):
            a = 2.0 * PI * i / max(n_spheres - 1, 1)
            objs.append(rt.Sphere.build(rt.translation(3.0 * math.cos(a), 0.6, 3.0 * math.sin(a)) * rt.scaling(0.6, 0.6, 0.6),
                                        rt.Material(color=(0.3 + 0.04 * (i % 10), 0.5, 0.9 - 0.05 * (i % 10)), reflective=0.3 if i % 3 == 0 else 0.0)))
        return cam, rt.World(objs, rt.PointLight((-10, 10, -10), (1, 1, 1)))
    cam, world = build(gpu)
    plan = gpu.inspect(cam, world)
    assert (plan["small_n"] > 0) == (n_spheres <= 16), plan
    got, want, st, ost = both(gpu, oracle, build)
    rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
    assert rep["exact_u8"] == 1.0 and st.rays == ost.rays, rep


def test_deepest_recursion(gpu, oracle):
    """reflection_recursion_depth = 23, the deepest bounce stack the device holds (24 is an error): two facing mirrors."""
    def build(rt):
        cam = rt.Camera(48, 32, PI / 3, rt.view_transform((0, 0.5, -2.5), (0, 0.5, 0), (0, 1, 0)))
        mirror = rt.Material(reflective=0.9, diffuse=0.1, ambient=0.05)
        left = rt.Plane.build(rt.translation(-3, 0, 0) * rt.rotation_z(-PI / 2), mirror)
        right = rt.Plane.build(rt.translation(3, 0, 0) * rt.rotation_z(PI / 2), mirror)
        floor = rt.Plane.build(rt.identity_4x4(), rt.Material(pattern=rt.Checkers((1, 1, 1), (0.2, 0.2, 0.2))))
        ball = rt.Sphere.build(rt.translation(0, 0.5, 0) * rt.scaling(0.5, 0.5, 0.5), rt.Material(color=(0.9, 0.3, 0.1)))
        return cam, rt.World([left, right, floor, ball], rt.PointLight((0, 4, -3), (1, 1, 1)))
    got, want, st, ost = both(gpu, oracle, build, depth=23)
    rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
    assert rep["exact_u8"] >= 0.999 and st.rays == ost.rays and st.secondary_rays > 5 * st.primary_rays, rep


def test_pageable_and_pinned_destinations_receive_the_same_frame(gpu):
    """A caller's heap array (the reference's Canvas is a Vec) is filled through a pinned staging frame by the host's
    threads, slice by slice; pinned memory is written by the copy engine directly.  Same bytes either way, for the whole
    frame and for the interleaved bands of a 3-way shard split."""
    import ctypes as C

    import ray_tracer_challenge_b200 as rt

    cam, world = scenes.soft_shadows(gpu, width=2048, height=1025, u_steps=2, v_steps=2)  # 6.3 MB / 25 MB planes, ragged last band
    w, h = cam.width_pixels, cam.height_pixels
    p = cam.prepare(world)
    lib = rt.device_library()
    lib.rtc_host_alloc.restype = C.c_void_p
    lib.rtc_host_alloc.argtypes = [C.c_size_t]
    lib.rtc_host_free.argtypes = [C.c_void_p]
    a_rgb, a_u8 = lib.rtc_host_alloc(w * h * 12), lib.rtc_host_alloc(w * h * 3)
    pin_rgb = np.ctypeslib.as_array(C.cast(a_rgb, C.POINTER(C.c_float)), shape=(h, w, 3))
    pin_u8 = np.ctypeslib.as_array(C.cast(a_u8, C.POINTER(C.c_uint8)), shape=(h, w, 3))
    pin_rgb[:], pin_u8[:] = -1.0, 7
    p.render(3, out_rgb=pin_rgb, out_u8=pin_u8)
    heap_rgb, heap_u8 = np.full((h, w, 3), -1.0, np.float32), np.full((h, w, 3), 7, np.uint8)
    p.render(3, out_rgb=heap_rgb, out_u8=heap_u8)
    assert np.array_equal(heap_u8, pin_u8) and np.array_equal(heap_rgb.view(np.uint32), pin_rgb.view(np.uint32))
    assert heap_u8.max() > 7  # something was rendered
    # three shards into one heap canvas: every band arrives exactly once
    shard_rgb, shard_u8 = np.full((h, w, 3), -1.0, np.float32), np.full((h, w, 3), 7, np.uint8)
    for k in range(3):
        p.render(3, out_rgb=shard_rgb, out_u8=shard_u8, shard=k, n_shards=3)
    assert np.array_equal(shard_u8, pin_u8) and np.array_equal(shard_rgb.view(np.uint32), pin_rgb.view(np.uint32))
    # the 8-bit plane alone (render_b200_u8's path)
    only_u8 = np.full((h, w, 3), 7, np.uint8)
    p.render(3, out_u8=only_u8, want_rgb=False)
    assert np.array_equal(only_u8, pin_u8)
    p.release()
    del pin_rgb, pin_u8
    lib.rtc_host_free(a_rgb), lib.rtc_host_free(a_u8)
