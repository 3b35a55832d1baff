"""The reference's per-shape hit-distance tables replayed ON sm_100a through rtc_trace_rays (nearest t >= 0 of
World::intersect + Intersection::hit), in both kernel builds and through both device paths: the small-scene table
(one shape in the world) and the BVH (the same shape among enough far-away filler spheres to force a tree).

    cube.rs:136-230 · cylinder.rs:160-363 · cone.rs:189-298 · triangle.rs:100-176 · csg.rs:271-362 ·
    bounding_box.rs:135-287 (the slab test, through a CSG's own box cull and Cube::local_intersect)

A table row gives the intersections the reference's `local_intersect` reports; the device returns what
`Intersection::hit` picks from them: the smallest non-negative one (or a miss).  Rows whose distances the reference
only counts are pinned against the oracle's intersection list for the same ray (bit-equal in the IEEE build)."""
import math

import numpy as np
import pytest

from tests.helpers import F32_EPSILON

PI = float(np.float32(math.pi))
UNION, INTERSECTION, DIFFERENCE = 0, 1, 2


def norm3(v):
    v = np.asarray(v, np.float32)
    m = np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
    return tuple(float(c) for c in (v / m).astype(np.float32))


def cube(rt):
    return rt.Cube()


def cylinder(lo=None, hi=None, closed=False):
    def build(rt):
        c = rt.Cylinder()
        if lo is not None:
            c.minimum_y, c.maximum_y = lo, hi
        c.closed = closed
        return c
    return build


def cone(lo=None, hi=None, closed=False):
    def build(rt):
        c = rt.Cone()
        if lo is not None:
            c.minimum_y, c.maximum_y = lo, hi
        c.closed = closed
        return c
    return build


def triangle(rt):
    return rt.Triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0))


def smooth_triangle(rt):
    return rt.SmoothTriangle((0, 1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0))


def csg_sphere_cube(rt):
    return rt.CSG(UNION, rt.Sphere(), rt.Cube())


def csg_two_spheres(rt):
    return rt.CSG(UNION, rt.Sphere(), rt.Sphere.build(rt.translation(0, 0, 0.5), rt.Material()))


def csg_box(lo, hi):
    """A CSG whose own bounding box (csg.rs:90-93 culls with it) is [lo, hi]: the union of a cube with itself."""
    def build(rt):
        c = [(a + b) / 2 for a, b in zip(lo, hi)]
        h = [(b - a) / 2 for a, b in zip(lo, hi)]
        m = rt.translation(*c) * rt.scaling(*h)
        return rt.CSG(UNION, rt.Cube.build(m, rt.Material()), rt.Cube.build(m, rt.Material()))
    return build


# (label, shape builder, [(origin, direction, expected)], normalise the direction like the reference's test does)
# expected: a list of the reference's asserted distances, an int = the asserted COUNT (distances from the oracle), or
# a bool = hit / miss only (bounding-box tables)
TABLES = [
    ("cube.rs:136-176 hits", cube, [
        ((5, 0.5, 0), (-1, 0, 0), [4.0, 6.0]), ((-5, 0.5, 0), (1, 0, 0), [4.0, 6.0]), ((0.5, 5, 0), (0, -1, 0), [4.0, 6.0]),
        ((0.5, -5, 0), (0, 1, 0), [4.0, 6.0]), ((0.5, 0, 5), (0, 0, -1), [4.0, 6.0]), ((0.5, 0.5, -5), (0, 0, 1), [4.0, 6.0]),
        ((0, 0.5, 0), (0, 0, 1), [-1.0, 1.0])], False),
    ("cube.rs:178-204 misses", cube, [
        ((-2, 0, 0), (0.2673, 0.5345, 0.8018), []), ((0, -2, 0), (0.8018, 0.2673, 0.5345), []),
        ((0, 0, -2), (0.5345, 0.8018, 0.2673), []), ((0, 0, 2), (0, 0, 1), []), ((2, 0, 2), (0, 0, -1), []),
        ((0, 2, 2), (0, -1, 0), []), ((2, 2, 0), (-1, 0, 0), [])], False),
    ("cylinder.rs:160-181 misses", cylinder(), [
        ((1, 0, 0), (0, 1, 0), []), ((0, 0, 0), (0, 1, 0), []), ((0, 0, -5), (1, 1, 1), [])], True),
    ("cylinder.rs:183-211 sides", cylinder(), [
        ((1, 0, -5), (0, 0, 1), [5.0, 5.0]), ((0, 0, -5), (0, 0, 1), [4.0, 6.0]),
        ((0.5, 0, -5), (0.1, 1, 1), [6.808006, 7.0886984])], True),
    ("cylinder.rs:240-273 truncated", cylinder(1.0, 2.0), [
        ((0, 1.5, 0), (0.1, 1, 0), 0), ((0, 3, -5), (0, 0, 1), 0), ((0, 0, -5), (0, 0, 1), 0), ((0, 2, -5), (0, 0, 1), 0),
        ((0, 1, -5), (0, 0, 1), 0), ((0, 1.5, -2), (0, 0, 1), 2)], True),
    ("cylinder.rs:283-313 caps", cylinder(1.0, 2.0, True), [
        ((0, 3, 0), (0, -1, 0), 2), ((0, 3, -2), (0, -1, 2), 2), ((0, 4, -2), (0, -1, 1), 2), ((0, 0, -2), (0, 1, 2), 2),
        ((0, -1, -2), (0, 1, 1), 2)], True),
    ("cone.rs:189-232 sides", cone(), [
        ((0, 0, -5), (0, 0, 1), [5.0, 5.0]), ((0, 0, -4.999999), (1, 1, 1), [8.660253, 8.660253]),
        ((1, 1, -5), (-0.5, -1, 1), [4.5500546, 49.449955])], True),
    ("cone.rs:234-246 parallel to one half", cone(), [((0, 0, -1), (0, 1, 1), [0.35355338])], True),
    ("cone.rs:248-275 caps", cone(-0.5, 0.5, True), [
        ((0, 0, -5), (0, 1, 0), 0), ((0, 0, -0.25), (0, 1, 1), 2), ((0, 0, -0.25), (0, 1, 0), 4)], True),
    ("triangle.rs:128-176", triangle, [
        ((0, -1, -2), (0, 1, 0), []), ((1, 1, -2), (0, 0, 1), []), ((-1, 1, -2), (0, 0, 1), []), ((0, -1, -2), (0, 0, 1), []),
        ((0, 0.5, -2), (0, 0, 1), [2.0])], False),
    ("smooth_triangle.rs:69-108 (flat hit, Q5)", smooth_triangle, [((-0.2, 0.3, -2), (0, 0, 1), [2.0])], False),
    ("csg.rs:271-283 miss", csg_sphere_cube, [((0, 2, -5), (0, 0, 1), [])], False),
    ("csg.rs:285-303 hit", csg_two_spheres, [((0, 0, -5), (0, 0, 1), [4.0, 6.5])], False),
    ("bounding_box.rs:135-189 unit box", csg_box((-1, -1, -1), (1, 1, 1)), [
        ((5, 0.5, 0), (-1, 0, 0), True), ((-5, 0.5, 0), (1, 0, 0), True), ((0.5, 5, 0), (0, -1, 0), True),
        ((0.5, -5, 0), (0, 1, 0), True), ((0.5, 0, 5), (0, 0, -1), True), ((0.5, 0, -5), (0, 0, 1), True),
        ((0, 0.5, 0), (0, 0, 1), True), ((-2, 0, 0), (2, 4, 6), False), ((0, -2, 0), (6, 2, 4), False),
        ((0, 0, -2), (4, 6, 2), False), ((2, 0, 2), (0, 0, -1), False), ((0, 2, 2), (0, -1, 0), False),
        ((2, 2, 0), (-1, 0, 0), False)], True),
    ("bounding_box.rs:191-245 non-cubic box", csg_box((5, -2, 0), (11, 4, 7)), [
        ((15, 1, 2), (-1, 0, 0), True), ((-5, -1, 4), (1, 0, 0), True), ((7, 6, 5), (0, -1, 0), True),
        ((9, -5, 6), (0, 1, 0), True), ((8, 2, 12), (0, 0, -1), True), ((6, 0, -5), (0, 0, 1), True),
        ((8, 1, 3.5), (0, 0, 1), True), ((9, -1, -8), (2, 4, 6), False), ((8, 3, -4), (6, 2, 4), False),
        ((9, -1, -2), (4, 6, 2), False), ((4, 0, 9), (0, 0, -1), False), ((8, 6, -1), (0, -1, 0), False),
        ((12, 5, 4), (-1, 0, 0), False)], True),
]


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


def hit_of(ts):
    """Intersection::hit (intersection.rs:30-35): the smallest non-negative distance, or None."""
    ts = [float(t) for t in ts if t >= 0.0]
    return min(ts) if ts else None


def world_of(rt, build, with_tree):
    shapes = [build(rt)]
    if with_tree:  # > kSmallCap bounded items: the commit builds a BVH and the rays go through dev_bvh.cuh
        for i in range(24):
            shapes.append(rt.Sphere.build(rt.translation(1000.0 + 3.0 * i, 1000.0, 1000.0), rt.Material()))
    return rt.World(shapes, rt.PointLight((-10, 10, -10), (1, 1, 1)))


@pytest.mark.parametrize("label,build,rows,normalise", TABLES, ids=[t[0] for t in TABLES])
def test_tables_hold_on_the_oracle(oracle, label, build, rows, normalise):
    """CPU leg: the transcribed rows are what the oracle's World::intersect reports (so the GPU leg checks the device
    against rows that are known to be the reference's)."""
    ow = world_of(oracle, build, False)
    for i, (o, d, expected) in enumerate(rows):
        d = norm3(d) if normalise else d
        ots, _ = oracle.probe.world_intersect(ow, o, d)
        if isinstance(expected, bool):
            assert (hit_of(ots) is not None) == expected, (label, i)
        elif isinstance(expected, int):
            assert len(ots) == expected, (label, i)
        else:
            assert len(ots) == len(expected), (label, i)
            for got, want in zip(ots, expected):
                assert abs(float(got) - want) <= F32_EPSILON * max(1.0, abs(want)), (label, i, float(got), want)


@pytest.mark.gpu
@pytest.mark.parametrize("with_tree", [False, True], ids=["small-scene path", "bvh path"])
@pytest.mark.parametrize("label,build,rows,normalise", TABLES, ids=[t[0] for t in TABLES])
def test_shape_table_on_gpu(gpu, oracle, label, build, rows, normalise, with_tree):
    gw, ow = world_of(gpu, build, with_tree), world_of(oracle, build, with_tree)
    cam = gpu.Camera(4, 4, PI / 2, gpu.identity_4x4())
    assert (gpu.inspect(cam, gw)["small_n"] == 0) == with_tree, "the world did not take the intended device path"
    info = cam.prepare(gw)
    try:
        origins = [r[0] for r in rows]
        dirs = [norm3(r[1]) if normalise else r[1] for r in rows]
        for strict in (True, False):
            _, t, _ = info.trace_rays(origins, dirs, 0, fma=not strict)
            for i, (o, d, expected) in enumerate(zip(origins, dirs, [r[2] for r in rows])):
                ots, _ = oracle.probe.world_intersect(ow, o, d)
                want = hit_of(ots)
                got = float(t[i])
                where = f"{label} row {i} origin {o} strict={strict}"
                if not strict and isinstance(expected, list) and len(expected) == 2 and expected[0] == expected[1]:
                    continue  # a tangent ray (discriminant exactly 0 in the reference's arithmetic): any contraction of
                    # b*b - 4ac may flip its sign, which is why the FMA build is opt-in and not the parity build
                if isinstance(expected, bool):
                    assert (got >= 0.0) == expected, where
                elif isinstance(expected, int):
                    assert len(ots) == expected, where + " (oracle count)"
                else:
                    pinned = hit_of(expected)
                    if pinned is None:
                        assert got == -1.0, where
                    else:  # the reference's own tolerance for these rows: f32 epsilon (assert_abs_diff_eq! / debug_assert!)
                        assert abs(got - pinned) <= (F32_EPSILON if strict else 2e-5) * max(1.0, abs(pinned)), where
                # and the device agrees with the oracle's list for the same ray: bit for bit in the IEEE build
                if want is None:
                    assert got == -1.0, where
                elif strict:
                    assert np.float32(got) == np.float32(want), where + f": {got!r} vs oracle {want!r}"
                else:
                    assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), where
    finally:
        info.release()


@pytest.mark.gpu
def test_random_rays_against_every_shape_kind(gpu, oracle):
    """2 000 random rays per kind through the IEEE build: hit distance and hit object identical to the oracle's."""
    rng = np.random.default_rng(20)
    builders = [cube, cylinder(-1.5, 0.75, True), cylinder(), cone(-1.0, 0.5, True), cone(-2.0, 2.0, False), triangle,
                csg_two_spheres, lambda rt: rt.Sphere.build(rt.scaling(1.0, 0.4, 2.0) * rt.rotation_z(0.7), rt.Material()),
                lambda rt: rt.Plane.build(rt.rotation_x(0.3), rt.Material())]
    for bi, build in enumerate(builders):
        for with_tree in (False, True):
            gw, ow = world_of(gpu, build, with_tree), world_of(oracle, build, with_tree)
            info = gpu.Camera(4, 4, PI / 2, gpu.identity_4x4()).prepare(gw)
            n = 2000 if not with_tree else 500
            origins = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
            targets = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32)
            dirs = np.asarray([norm3(t - o) for o, t in zip(origins, targets)], np.float32)
            _, t, _ = info.trace_rays(origins, dirs, 0)
            info.release()
            hits = 0
            for i in range(n):
                ots, _ = oracle.probe.world_intersect(ow, origins[i], dirs[i])
                want = hit_of(ots)
                if want is None:
                    assert t[i] == -1.0, (bi, with_tree, i)
                else:
                    hits += 1
                    assert np.float32(t[i]) == np.float32(want), (bi, with_tree, i, float(t[i]), want)
            assert hits > n // 20, "the random rays should hit the shape often enough to mean something"
