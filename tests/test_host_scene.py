"""Host-side logic of the product (librtc_host.so: scene model, cofactor inverse, group baking, divide, OBJ
loader, flattener) checked against the CPU oracle — no GPU needed.  The two libraries share no code, so equality
here is an independent check of both restatements of the reference's scene construction."""
import math

import numpy as np
import pytest

from ray_tracer_challenge_b200 import scenes


@pytest.fixture(scope="module")
def host():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def test_matrix_routines_are_bit_identical(host, oracle):
    rng = np.random.default_rng(7)
    for _ in range(200):
        a = rng.normal(size=(4, 4)).astype(np.float32)
        b = rng.normal(size=(4, 4)).astype(np.float32)
        a[3] = b[3] = (0, 0, 0, 1)
        ha, oa = host.Matrix(a), oracle.Matrix(a)
        hb, ob = host.Matrix(b), oracle.Matrix(b)
        assert np.array_equal(bits(ha.inverse().m), bits(oa.inverse().m))
        assert np.array_equal(bits((ha * hb).m), bits((oa * ob).m))
        assert bits(ha.determinant()) == bits(oa.determinant())
    for args in ((0.3,), (-2.1,), (math.pi / 5,)):
        for fn in ("rotation_x", "rotation_y", "rotation_z"):
            assert np.array_equal(bits(getattr(host, fn)(*args).m), bits(getattr(oracle, fn)(*args).m))
    h = host.view_transform((1, 3, 2), (4, -2, 8), (1, 1, 0))
    o = oracle.view_transform((1, 3, 2), (4, -2, 8), (1, 1, 0))
    assert np.array_equal(bits(h.m), bits(o.m))


def walk(shape):
    """Depth-first (kind, transform bits, bbox bits) of a shape tree."""
    out = [(shape.kind(), bits(shape.transformation().m).tobytes(), bits(np.concatenate(shape.bounding_box())).tobytes())]
    if shape.kind() in (7, 8):
        for c in shape.get_children():
            out += walk(c)
    return out


@pytest.mark.parametrize("name,kw", [
    ("hexagons", dict(width=32, height=16)),
    ("shapes_zoo", dict(width=32, height=16)),
    ("csg_gallery", dict(width=32, height=16)),
    ("dragon_element", dict(width=32, height=16, n_u=12, n_v=6)),
    ("dragon_element", dict(width=32, height=16, n_u=12, n_v=6, smooth=True, divide=2)),
    ("stress", dict(width=32, height=16, n_spheres=300, n_each=3, n_csg=3, divide=4)),
])
def test_scene_construction_matches_oracle(host, oracle, name, kw):
    """Group baking, cached boxes, divide() structure and OBJ loading agree object by object."""
    _, hw = getattr(scenes, name)(host, **kw)
    _, ow = getattr(scenes, name)(oracle, **kw)
    assert len(hw.objects) == len(ow.objects)
    for a, b in zip(hw.objects, ow.objects):
        assert walk(a) == walk(b)


def test_flattener_emits_depth_first_leaves(host):
    cam, world = scenes.csg_gallery(host, width=16, height=8)
    prims, nodes, refs, shapes, counts = host.flatten(world)

    def leaves(shape):
        if shape.kind() in (7, 8):
            return [h for c in shape.get_children() for h in leaves(c)]
        return [shape.handle]

    expected = [h for o in world.objects for h in leaves(o)]
    assert shapes == expected  # RtcPrim order == the reference's emission (tie-break) order
    # every CSG node has exactly two children and a valid operator; parents form a forest
    for n in nodes:
        assert n.parent < len(nodes)
        if n.kind == 1:
            assert n.child_count == 2 and 0 <= n.op <= 2
    # child references resolve to primitives or nodes that name this node as their parent
    for idx, n in enumerate(nodes):
        for r in refs[n.child_begin:n.child_begin + n.child_count]:
            if r >= 0:
                assert prims[r].parent == idx
            else:
                assert nodes[~r].parent == idx


def test_flattener_deduplicates_materials_and_lowers_smooth_triangles(host):
    cam, world = scenes.dragon_element(host, width=16, height=8, n_u=8, n_v=4, smooth=True)
    prims, nodes, refs, shapes, counts = host.flatten(world)
    n_tri = sum(1 for p in prims if p.type == 5)
    assert n_tri == 2 * 8 * 4  # fan triangulation of the quads; SmoothTriangle lowered to the flat triangle (Q5)
    assert counts[3] <= 5  # mesh, case, pedestal, floor materials (+ default)
    tri = next(p for p in prims if p.type == 5)
    p1, e1, e2, nrm = np.asarray(tri.params[:]).reshape(4, 3)
    n = np.cross(e2, e1)
    assert np.allclose(n / np.linalg.norm(n), nrm, atol=1e-6)  # triangle.rs:23


def test_flattened_boxes_and_inverses_match_oracle(host, oracle):
    _, hw = scenes.shapes_zoo(host, width=16, height=8)
    _, ow = scenes.shapes_zoo(oracle, width=16, height=8)
    prims, *_ = host.flatten(hw)

    def leaves(shape):
        if shape.kind() in (7, 8):
            return [x for c in shape.get_children() for x in leaves(c)]
        return [shape]

    oleaves = [x for o in ow.objects for x in leaves(o)]
    assert len(prims) == len(oleaves)
    for p, o in zip(prims, oleaves):
        assert np.array_equal(bits(p.inv[:]), bits(o.transformation_inverse().m).reshape(-1))
        mn, mx = o.parent_space_bounding_box()
        with np.errstate(invalid="ignore"):
            assert np.array_equal(bits(p.bbox_min[:]), bits(mn)) and np.array_equal(bits(p.bbox_max[:]), bits(mx))


def test_host_rejects_the_oracle_only_test_double(host):
    import ray_tracer_challenge_b200 as rt

    with pytest.raises(rt.RtcError):
        host.TestShape()
    with pytest.raises(rt.RtcError):
        g = host.GroupShape()
        s = host.Sphere()
        g.add_child(s)
        g.add_child(s)  # already owned (Rust: moved)


def test_commit_plan_without_a_device(host):
    """rtc_scene_inspect: the host half of the commit (tree, small-scene table, shadow-filter eligibility, kernel build
    choice) for the BASELINE workloads and a few demo scenes — no GPU needed."""
    from bench import build_scene

    plans = {}
    for workload in ("c1", "c2", "c3", "c4"):
        cam, world, _, _ = build_scene(host, workload)
        plans[workload] = host.inspect(cam, world)
    # c1 / c3: four objects (cube, plane, two spheres), table-mode area light: every fast path applies
    for w in ("c1", "c3"):
        p = plans[w]
        assert (p["small_n"], p["n_bvh_nodes"], p["filter_ok"], p["cell_masks"], p["plane_cells"], p["converge"]) == (4, 0, 1, 1, 1, 0), p
        assert 0 < p["tol_sphere"] < 1e-4 and p["light_ball"][3] > 1.0
    # c2: cylinder and cone -> exact shadow test; the glass sphere is reflective and transparent -> converging build
    p = plans["c2"]
    assert (p["small_n"], p["filter_ok"], p["cell_masks"], p["converge"]) == (6, 0, 0, 1), p
    # c4: 102 k triangles in one tree with 4-triangle leaves, the plane in the linear list, 4 distinct transforms
    p = plans["c4"]
    assert p["small_n"] == 0 and p["n_bvh_nodes"] > 20000 and p["bvh_leaf_size"] == 4 and p["n_linear"] == 1 and p["n_xforms"] == 4, p
    assert p["n_positions"] == 102403
    # soft_shadows as shipped (jitter None): filter + drawn-sample cell masks, no staged plane constants
    cam, world = scenes.soft_shadows(host, width=100, height=40, jitter=None)
    p = host.inspect(cam, world)
    assert (p["filter_ok"], p["cell_masks"], p["plane_cells"]) == (1, 1, 0), p
    # first_scene: the walls are spheres squashed 1000:1 -> condition number far above the filter's limit
    cam, world = scenes.first_scene(host, width=100, height=50)
    assert host.inspect(cam, world)["filter_ok"] == 0
    # a sphere field goes to the tree with one sphere per leaf
    cam, world = scenes.stress(host, width=64, height=36, n_spheres=500, n_each=2, n_csg=1)
    p = host.inspect(cam, world)
    assert p["small_n"] == 0 and p["bvh_leaf_size"] == 1 and p["converge"] == 1, p


def test_parallel_tree_build_is_deterministic(host):
    """The BVH's top levels are built by several threads; the tree (node count, and so the render) does not depend on
    their timing."""
    cam, world = scenes.stress(host, width=64, height=36, n_spheres=20000, n_each=4, n_csg=2)
    seen = {(p["n_bvh_nodes"], p["n_positions"], p["n_xforms"], p["digest"]) for p in (host.inspect(cam, world) for _ in range(4))}
    assert len(seen) == 1, seen  # digest: every byte of the heads, records, transforms and tree nodes
    # the digest does see geometry: the same field with one sphere fewer hashes differently
    cam2, world2 = scenes.stress(host, width=64, height=36, n_spheres=19999, n_each=4, n_csg=2)
    assert host.inspect(cam2, world2)["digest"] not in {s[3] for s in seen}


def test_empty_worlds_commit_and_render_black(host, oracle):
    """Edge cases of World::objects: no objects at all, and only an empty group (group.rs:199-206 — an empty group has
    no intersections).  The host half of the commit accepts both (nothing to trace: no tree, no table), and the oracle's
    frame is the reference's black canvas."""
    from ray_tracer_challenge_b200.scenes import PI

    for make in (lambda api: [], lambda api: [api.GroupShape()]):
        cam = host.Camera(16, 8, PI / 2.0, host.view_transform((0, 0, -5), (0, 0, 0), (0, 1, 0)))
        world = host.World(make(host), host.PointLight((-10, 10, -10), (1, 1, 1)))
        p = host.inspect(cam, world)
        assert (p["n_positions"], p["n_bvh_nodes"], p["n_linear"], p["small_n"]) == (0, 0, 0, 0), p
        ocam = oracle.Camera(16, 8, PI / 2.0, oracle.view_transform((0, 0, -5), (0, 0, 0), (0, 1, 0)))
        oworld = oracle.World(make(oracle), oracle.PointLight((-10, 10, -10), (1, 1, 1)))
        frame = ocam.render(oworld, 5)
        assert not frame.data.any()
