"""Golden vectors of camera.rs, ray.rs, shape/{shape,sphere,plane,cube,cylinder,cone,triangle,smooth_triangle}.rs
replayed against the CPU oracle."""
import math

import numpy as np

from tests.helpers import F32_EPSILON, assert_abs_diff_eq, assert_eq

PI = float(np.float32(math.pi))
FRAC_1_SQRT_2 = float(np.float32(0.70710678118654752440))
SQRT_2 = float(np.float32(1.41421356237309504880))
INF = float("inf")


def norm3(v):
    """Tuple::norm on a vector (tuple.rs:34-43) in f32, left to right."""
    v = np.asarray(v, np.float32)
    m = np.sqrt(np.float32(np.float32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]))
    return (v / m).astype(np.float32)


# ------------------------------------------------------------------ camera.rs:108-167
def test_camera_pixel_size(rt):
    for w, h in ((200, 125), (125, 200)):
        c = rt.Camera(w, h, PI / 2.0, rt.identity_4x4())
        assert_eq(rt.probe.camera_info(c)["pixel_size"], 0.01)


def test_ray_through_canvas_center(rt):
    c = rt.Camera(201, 101, PI / 2.0, rt.identity_4x4())
    o, d = rt.probe.camera_ray(c, 100, 50)
    assert_eq(o, (0, 0, 0))
    assert_abs_diff_eq(d, (0, 0, -1))


def test_ray_through_canvas_corner(rt):
    c = rt.Camera(201, 101, PI / 2.0, rt.identity_4x4())
    o, d = rt.probe.camera_ray(c, 0, 0)
    assert_eq(o, (0, 0, 0))
    assert_abs_diff_eq(d, (0.6651864, 0.33259323, -0.66851234))


def test_ray_with_transformed_camera(rt):
    c = rt.Camera(201, 101, PI / 2.0, rt.rotation_y(PI / 4.0) * rt.translation(0.0, -2.0, 5.0))
    o, d = rt.probe.camera_ray(c, 100, 50)
    assert_abs_diff_eq(o, (0, 2, -5), epsilon=10.0 * F32_EPSILON)
    assert_abs_diff_eq(d, (FRAC_1_SQRT_2, 0, -FRAC_1_SQRT_2))


def test_render_world(rt):
    c = rt.Camera(11, 11, PI / 2.0, rt.view_transform((0, 0, -5), (0, 0, 0), (0, 1, 0)))
    image = c.render(rt.World.default(), 5)
    assert_abs_diff_eq(image.pixel_at(5, 5), (0.38063288, 0.47579104, 0.28547466))
    # camera.rs:80-81 — the last row and the last column are never rendered
    assert not image.data[10, :, :].any() and not image.data[:, 10, :].any()


# ------------------------------------------------------------------ ray.rs:74-154
def test_sphere_intersections(rt):
    s = rt.Sphere()
    cases = [((0, 0, -5), [4.0, 6.0]), ((0, 1, -5), [5.0, 5.0]), ((0, 2, -5), []), ((0, 0, 0), [-1.0, 1.0]),
             ((0, 0, 5), [-6.0, -4.0])]
    for origin, expected in cases:
        ts, _, _ = rt.probe.intersect(s, origin, (0, 0, 1))
        assert_eq(ts, expected, msg=str(origin))


def test_reflect(rt):
    assert_eq(rt.probe.reflect((1, -1, 0), (0, 1, 0)), (1, 1, 0))
    assert_abs_diff_eq(rt.probe.reflect((0, -1, 0), (FRAC_1_SQRT_2, FRAC_1_SQRT_2, 0)), (1, 0, 0))


# ------------------------------------------------------------------ shape.rs:202-285
def test_intersect_scaled_and_translated_shape(rt):
    s = rt.TestShape()
    s.set_transformation(rt.scaling(2, 2, 2))
    rt.probe.intersect(s, (0, 0, -5), (0, 0, 1))
    o, d = rt.probe.saved_ray(s)
    assert_eq(o, (0, 0, -2.5))
    assert_eq(d, (0, 0, 0.5))
    s = rt.TestShape()
    s.set_transformation(rt.translation(5, 0, 0))
    rt.probe.intersect(s, (0, 0, -5), (0, 0, 1))
    o, d = rt.probe.saved_ray(s)
    assert_eq(o, (-5, 0, -5))
    assert_eq(d, (0, 0, 1))


def test_normal_on_translated_shape(rt):
    s = rt.TestShape()
    s.set_transformation(rt.translation(0, 1, 0))
    assert_abs_diff_eq(rt.probe.normal_at(s, (0, 1.70711, -0.70711)), (0.0, 0.6000001, -0.79999995))


def test_normal_on_transformed_shape(rt):
    s = rt.TestShape()
    s.set_transformation(rt.scaling(1.0, 0.5, 1.0) * rt.rotation_z(PI / 5.0))
    assert_abs_diff_eq(rt.probe.normal_at(s, (0, FRAC_1_SQRT_2, -FRAC_1_SQRT_2)), (-0.08352663, 0.9325296, -0.3513003))


def test_normal_is_normalized(rt):
    n = rt.probe.normal_at(rt.TestShape(), (1, 5, 10))
    assert_abs_diff_eq(n, norm3(n))


def test_normal_object_to_world_through_groups(rt):
    k = float(np.float32(1.0) / np.sqrt(np.float32(3.0)))
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    g2 = rt.GroupShape()
    g2.set_transformation(rt.scaling(1, 2, 3))
    g1 = rt.GroupShape()
    g1.set_transformation(rt.rotation_y(PI / 2.0))
    g2.add_child(s)
    g1.add_child(g2)
    inner = g1.get_children()[0].get_children()[0]
    assert_abs_diff_eq(rt.probe.normal_to_world(inner, (k, k, k)), (0.28571427, 0.42857143, -0.85714287))


def test_parent_space_bounding_box(rt):
    s = rt.Sphere()
    s.set_transformation(rt.translation(1, -3, 5) * rt.scaling(0.5, 2, 4))
    mn, mx = s.parent_space_bounding_box()
    assert_eq(mn, (0.5, -5, 1))
    assert_eq(mx, (1.5, -1, 9))


# ------------------------------------------------------------------ sphere.rs:94-145
def test_sphere_local_intersect_and_normals(rt):
    s = rt.Sphere()
    s.set_transformation(rt.scaling(2, 2, 2))
    ts, _, _ = rt.probe.local_intersect(s, (0, 0, -2.5), (0, 0, 0.5))
    assert_eq(ts, [3.0, 7.0])
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    ts, _, _ = rt.probe.local_intersect(s, (-5, 0, -5), (0, 0, 1))
    assert len(ts) == 0
    s = rt.Sphere()
    for p in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
        assert_eq(rt.probe.normal_at(s, p, local=True), p)
    k = float(np.float32(1.0) / np.sqrt(np.float32(3.0)))
    assert_abs_diff_eq(rt.probe.normal_at(s, (k, k, k), local=True), (k, k, k))


# ------------------------------------------------------------------ plane.rs:74-116
def test_plane(rt):
    p = rt.Plane()
    for q in ((0, 0, 0), (10, 0, -10), (-5, 0, 150)):
        assert_eq(rt.probe.normal_at(p, q, local=True), (0, 1, 0))
    assert len(rt.probe.local_intersect(p, (0, 10, 0), (0, 0, 1))[0]) == 0
    assert len(rt.probe.local_intersect(p, (0, 0, 0), (0, 0, 1))[0]) == 0
    assert_eq(rt.probe.local_intersect(p, (0, 1, 0), (0, -1, 0))[0], [1.0])
    assert_eq(rt.probe.local_intersect(p, (0, -1, 0), (0, 1, 0))[0], [1.0])


# ------------------------------------------------------------------ cube.rs:136-230
def test_ray_intersects_cube(rt):
    c = rt.Cube()
    cases = [((5, 0.5, 0), (-1, 0, 0), 4.0, 6.0), ((-5, 0.5, 0), (1, 0, 0), 4.0, 6.0),
             ((0.5, 5, 0), (0, -1, 0), 4.0, 6.0), ((0.5, -5, 0), (0, 1, 0), 4.0, 6.0),
             ((0.5, 0, 5), (0, 0, -1), 4.0, 6.0), ((0.5, 0.5, -5), (0, 0, 1), 4.0, 6.0),
             ((0, 0.5, 0), (0, 0, 1), -1.0, 1.0)]
    for o, d, t1, t2 in cases:
        assert_eq(rt.probe.local_intersect(c, o, d)[0], [t1, t2], msg=str(o))


def test_ray_misses_cube(rt):
    c = rt.Cube()
    cases = [((-2, 0, 0), (0.2673, 0.5345, 0.8018)), ((0, -2, 0), (0.8018, 0.2673, 0.5345)),
             ((0, 0, -2), (0.5345, 0.8018, 0.2673)), ((0, 0, 2), (0, 0, 1)), ((2, 0, 2), (0, 0, -1)),
             ((0, 2, 2), (0, -1, 0)), ((2, 2, 0), (-1, 0, 0))]
    for o, d in cases:
        assert len(rt.probe.local_intersect(c, o, d)[0]) == 0, str(o)


def test_cube_surface_normal(rt):
    c = rt.Cube()
    cases = [((1, 0.5, -0.8), (1, 0, 0)), ((-1, -0.2, 0.9), (-1, 0, 0)), ((-0.4, 1, -0.1), (0, 1, 0)),
             ((0.3, -1, -0.7), (0, -1, 0)), ((-0.6, 0.3, 1), (0, 0, 1)), ((0.4, 0.4, -1), (0, 0, -1)),
             ((1, 1, 1), (1, 0, 0)), ((-1, -1, -1), (-1, 0, 0))]
    for p, n in cases:
        assert_eq(rt.probe.normal_at(c, p, local=True), n, msg=str(p))


# ------------------------------------------------------------------ cylinder.rs:160-363
def test_ray_misses_cylinder(rt):
    c = rt.Cylinder()
    for o, d in [((1, 0, 0), (0, 1, 0)), ((0, 0, 0), (0, 1, 0)), ((0, 0, -5), (1, 1, 1))]:
        assert len(rt.probe.local_intersect(c, o, norm3(d))[0]) == 0, str(o)


def test_ray_intersects_cylinder_sides(rt):
    c = rt.Cylinder()
    cases = [((1, 0, -5), (0, 0, 1), 5.0, 5.0), ((0, 0, -5), (0, 0, 1), 4.0, 6.0),
             ((0.5, 0, -5), (0.1, 1, 1), 6.808006, 7.0886984)]
    for o, d, t1, t2 in cases:
        ts = rt.probe.local_intersect(c, o, norm3(d))[0]
        assert len(ts) == 2
        assert_abs_diff_eq(ts, [t1, t2], msg=str(o))


def test_ray_intersects_constrained_cylinder(rt):
    c = rt.Cylinder()
    c.minimum_y, c.maximum_y = 1.0, 2.0
    cases = [((0, 1.5, 0), (0.1, 1, 0), 0), ((0, 3, -5), (0, 0, 1), 0), ((0, 0, -5), (0, 0, 1), 0),
             ((0, 2, -5), (0, 0, 1), 0), ((0, 1, -5), (0, 0, 1), 0), ((0, 1.5, -2), (0, 0, 1), 2)]
    for o, d, n in cases:
        assert len(rt.probe.local_intersect(c, o, norm3(d))[0]) == n, str(o)


def test_ray_intersects_caps_of_closed_cylinder(rt):
    c = rt.Cylinder()
    c.minimum_y, c.maximum_y, c.closed = 1.0, 2.0, True
    cases = [((0, 3, 0), (0, -1, 0)), ((0, 3, -2), (0, -1, 2)), ((0, 4, -2), (0, -1, 1)), ((0, 0, -2), (0, 1, 2)),
             ((0, -1, -2), (0, 1, 1))]
    for o, d in cases:
        assert len(rt.probe.local_intersect(c, o, norm3(d))[0]) == 2, str(o)


def test_cylinder_normals(rt):
    c = rt.Cylinder()
    for p, n in [((1, 0, 0), (1, 0, 0)), ((0, 5, -1), (0, 0, -1)), ((0, -2, 1), (0, 0, 1)), ((-1, 1, 0), (-1, 0, 0))]:
        assert_eq(rt.probe.normal_at(c, p, local=True), n)
    c = rt.Cylinder()
    c.minimum_y, c.maximum_y, c.closed = 1.0, 2.0, True
    for p, n in [((0, 1, 0), (0, -1, 0)), ((0.5, 1, 0), (0, -1, 0)), ((0, 1, 0.5), (0, -1, 0)),
                 ((0, 2, 0), (0, 1, 0)), ((0.5, 2, 0), (0, 1, 0)), ((0, 2, 0.5), (0, 1, 0))]:
        assert_eq(rt.probe.normal_at(c, p, local=True), n)


# ------------------------------------------------------------------ cone.rs:189-298
def test_ray_intersects_cone_sides(rt):
    c = rt.Cone()
    cases = [((0, 0, -5), (0, 0, 1), 5.0, 5.0), ((0, 0, -4.999999), (1, 1, 1), 8.660253, 8.660253),
             ((1, 1, -5), (-0.5, -1, 1), 4.5500546, 49.449955)]
    for o, d, t1, t2 in cases:
        ts = rt.probe.local_intersect(c, o, norm3(d))[0]
        assert len(ts) == 2, str(o)
        # the reference checks these with debug_assert!(abs_diff_eq(.., f32 epsilon)) (cone.rs:215-232)
        assert_abs_diff_eq(ts, [t1, t2], msg=str(o))


def test_cone_ray_parallel_to_one_half(rt):
    ts = rt.probe.local_intersect(rt.Cone(), (0, 0, -1), norm3((0, 1, 1)))[0]
    assert len(ts) == 1
    assert_abs_diff_eq(ts[0], 0.35355338)


def test_ray_intersects_caps_of_closed_cone(rt):
    c = rt.Cone()
    c.minimum_y, c.maximum_y, c.closed = -0.5, 0.5, True
    for o, d, n in [((0, 0, -5), (0, 1, 0), 0), ((0, 0, -0.25), (0, 1, 1), 2), ((0, 0, -0.25), (0, 1, 0), 4)]:
        assert len(rt.probe.local_intersect(c, o, norm3(d))[0]) == n, str(o)


def test_cone_normal_and_bbox(rt):
    c = rt.Cone()
    for p, n in [((0, 0, 0), (0, 0, 0)), ((1, 1, 1), (1, -SQRT_2, 1)), ((-1, -1, 0), (-1, 1, 0))]:
        assert_eq(rt.probe.normal_at(c, p, local=True), n)
    mn, mx = c.bounding_box()
    assert_eq(mn, (-INF, -INF, -INF))
    assert_eq(mx, (INF, INF, INF))
    c = rt.Cone()
    c.minimum_y, c.maximum_y = -5.0, 3.0
    mn, mx = c.bounding_box()
    assert_eq(mn, (-5, -5, -5))
    assert_eq(mx, (5, 3, 5))


# ------------------------------------------------------------------ triangle.rs:100-176
def default_triangle(rt):
    return rt.Triangle((0, 1, 0), (-1, 0, 0), (1, 0, 0))


def test_triangle_construction_and_normal(rt):
    t = default_triangle(rt)
    p1, e1, e2, n = t.geometry()
    assert_eq(p1, (0, 1, 0))
    assert_eq(e1, (-1, -1, 0))
    assert_eq(e2, (1, -1, 0))
    assert_eq(n, (0, 0, -1))
    for p in ((0, 0.5, 0), (-0.5, 0.75, 0), (0.5, 0.25, 0)):
        assert_eq(rt.probe.normal_at(t, p, local=True), n)


def test_triangle_intersections(rt):
    t = default_triangle(rt)
    assert len(rt.probe.local_intersect(t, (0, -1, -2), (0, 1, 0))[0]) == 0
    assert len(rt.probe.local_intersect(t, (1, 1, -2), (0, 0, 1))[0]) == 0
    assert len(rt.probe.local_intersect(t, (-1, 1, -2), (0, 0, 1))[0]) == 0
    assert len(rt.probe.local_intersect(t, (0, -1, -2), (0, 0, 1))[0]) == 0
    assert_eq(rt.probe.local_intersect(t, (0, 0.5, -2), (0, 0, 1))[0], [2.0])


def test_triangle_bounding_box(rt):
    mn, mx = rt.Triangle((-3, 7, 2), (6, 2, -4), (2, -1, -1)).bounding_box()
    assert_eq(mn, (-3, -1, -4))
    assert_eq(mx, (6, 7, 2))


# ------------------------------------------------------------------ smooth_triangle.rs:69-108
def default_smooth_triangle(rt):
    return rt.SmoothTriangle((0, 1, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0))


def test_smooth_triangle_uv_and_normal(rt):
    t = default_smooth_triangle(rt)
    ts, objs, uvs = rt.probe.local_intersect(t, (-0.2, 0.3, -2), (0, 0, 1))
    assert_eq(uvs[0], (0.45, 0.25))
    assert_abs_diff_eq(rt.probe.normal_at(t, (0, 0, 0), u=0.45, v=0.25), (-0.5547002, 0.8320504, 0.0))
    comps = rt.probe.precompute((-0.2, 0.3, -2), (0, 0, 1), [(1.0, t)], uvs=[(0.45, 0.25)])
    assert_abs_diff_eq(comps["surface_normal"], (-0.5547002, 0.8320504, 0.0))


def test_smooth_triangle_renders_flat(rt):
    """local_intersect delegates to the inner flat Triangle (smooth_triangle.rs:39-41), so a
    world-level hit reports the flat triangle whose normal is constant."""
    t = default_smooth_triangle(rt)
    w = rt.World([t], rt.PointLight((0, 0, -10), (1, 1, 1)))
    ts, objs = rt.probe.world_intersect(w, (-0.2, 0.3, -2), (0, 0, 1))
    assert_eq(ts, [2.0])
    c1 = rt.probe.color_at(w, (-0.2, 0.3, -2), (0, 0, 1))
    c2 = rt.probe.color_at(w, (0.3, 0.2, -2), (0, 0, 1))
    assert np.all(c1 > 0)
    # flat normal (0,0,-1) everywhere: only the light direction varies between the two hits
    comps = rt.probe.precompute((-0.2, 0.3, -2), (0, 0, 1), [(2.0, rt.wrap_shape(int(objs[0])))])
    assert_eq(comps["surface_normal"], (0, 0, -1))
    assert c1.shape == c2.shape
