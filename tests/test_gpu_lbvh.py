"""K3 lbvh_build (csrc/rtc_lbvh.cu): the tree built on the device renders the pixels the host-built tree renders — the tree
only selects which primitives get the reference's exact test — for meshes (4-triangle leaves), sphere fields with CSG
and unbounded shapes around them, and degenerate inputs; and the commit says which builder made the tree."""
import ctypes as C

import numpy as np
import pytest

from ray_tracer_challenge_b200 import scenes
from tests.parity import compare_frames

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


def render_with(gpu, cam, world, builder, depth=5):
    gpu.set_bvh_builder(builder)
    try:
        p = cam.prepare(world)
        try:
            img = p.render(depth, detailed=True)
            return img, p.last_stats
        finally:
            p.release()
    finally:
        gpu.set_bvh_builder(-1)


@pytest.mark.parametrize("name,kw", [
    ("dragon_element", dict(width=240, height=135, n_u=96, n_v=48)),               # 9 k triangles, 4 per leaf
    ("dragon_element", dict(width=160, height=90, n_u=64, n_v=32, smooth=True)),
    ("stress", dict(width=192, height=108, n_spheres=6000, n_each=8, n_csg=4)),    # spheres + CSG + cylinders / cones / cubes
    ("here_be_dragons", dict(width=200, height=80, n_u=24, n_v=12)),
])
def test_device_tree_renders_the_same_frame(gpu, oracle, name, kw):
    cam, world = getattr(scenes, name)(gpu, **kw)
    host_img, host_st = render_with(gpu, cam, world, 0)
    dev_img, dev_st = render_with(gpu, cam, world, 1)
    # the same rays, the same pixels, bit for bit; only the tree walk differs
    assert dev_st.rays == host_st.rays
    assert np.array_equal(dev_img.data.view(np.uint32), host_img.data.view(np.uint32))
    assert dev_st.node_visits != host_st.node_visits, "the two builders should not produce the same tree"
    assert dev_st.node_visits < 4 * host_st.node_visits, "an LBVH is worse than SAH, but not this much"
    ocam, oworld = getattr(scenes, name)(oracle, **kw)
    want = ocam.render(oworld, 5)
    rep = compare_frames(dev_img.to_u8(), want.to_u8(), dev_img.data, want.data)
    assert rep["exact_u8"] >= 0.9995, rep


def test_one_shot_render_of_a_big_scene_uses_the_device_builder(gpu, oracle):
    """Camera::render_b200 (one shot) switches to the device builder from 10 000 primitives; the frame is the prepared
    (host-tree) frame."""
    kw = dict(width=160, height=90, n_u=160, n_v=80)  # 25 k triangles
    cam, world = scenes.dragon_element(gpu, **kw)
    one_shot = cam.render_b200(world, 5)
    prepared, _ = render_with(gpu, cam, world, 0)
    assert np.array_equal(one_shot.data.view(np.uint32), prepared.data.view(np.uint32))


def test_degenerate_input_falls_back_to_the_host_builder(gpu):
    """4 096 spheres with ONE centre: equal Morton codes everywhere — the position-based tie-break still yields a tree
    (depth 12), and it renders what the host tree renders.  Whatever the device builder returns is either within the
    traversal stack's depth or refused in favour of the balancing host builder: never silently truncated."""
    spheres = [gpu.Sphere.build(gpu.translation(0.0, 0.0, 0.0) * gpu.scaling(1.0 + 1e-4 * (i % 7), 1.0, 1.0), gpu.Material())
               for i in range(4096)]
    world = gpu.World(spheres, gpu.PointLight((-10, 10, -10), (1, 1, 1)))
    cam = gpu.Camera(48, 32, 1.0, gpu.view_transform((0, 0, -4), (0, 0, 0), (0, 1, 0)))
    host_img, _ = render_with(gpu, cam, world, 0, depth=1)
    dev_img, _ = render_with(gpu, cam, world, 1, depth=1)
    assert np.array_equal(dev_img.data.view(np.uint32), host_img.data.view(np.uint32))
    assert dev_img.data.max() > 0.1
