"""Golden vectors of shape/group.rs, shape/csg.rs and bounding_box.rs replayed against the CPU oracle."""
import math

import numpy as np

from tests.helpers import assert_abs_diff_eq, assert_eq
from tests.test_oracle_golden_shapes import norm3

PI = float(np.float32(math.pi))
UNION, INTERSECTION, DIFFERENCE = 0, 1, 2


# ------------------------------------------------------------------ group.rs:199-645
def test_add_child_and_material_propagation(rt):
    g = rt.GroupShape()
    s = rt.Sphere()
    g.add_child(s)
    kids = g.get_children()
    assert len(kids) == 1 and kids[0].handle == s.handle
    g = rt.GroupShape()
    for _ in range(3):
        g.add_child(rt.Sphere())
    g.set_material(rt.Material(shininess=123.456))
    # group.rs:96-100 pushes the material into every child: a unit-test level check through Phong's powf
    w = rt.World([g], rt.PointLight((0, 0, -10), (1, 1, 1)))
    eye = (0, math.sqrt(0.5), -math.sqrt(0.5))
    a = rt.probe.phong(g.get_children()[0], None, w.light, (0, 0, 0), (0, 0, -1), (0, 0, -1), 1.0)
    assert_eq(a, (1.9, 1.9, 1.9))


def test_intersect_empty_and_nonempty_group(rt):
    g = rt.GroupShape()
    assert len(rt.probe.local_intersect(g, (0, 0, 0), (0, 0, 1))[0]) == 0
    s1, s2, s3 = rt.Sphere(), rt.Sphere(), rt.Sphere()
    s2.set_transformation(rt.translation(0, 0, -3))
    s3.set_transformation(rt.translation(5, 0, 0))
    g = rt.GroupShape()
    for s in (s1, s2, s3):
        g.add_child(s)
    ts, objs, _ = rt.probe.local_intersect(g, (0, 0, -5), (0, 0, 1))
    order = np.argsort(ts, kind="stable")
    assert len(ts) == 4
    assert [int(objs[i]) for i in order] == [s2.handle, s2.handle, s1.handle, s1.handle]


def test_group_transform_baking(rt):
    expected = [[2, 0, 0, 10], [0, 2, 0, 0], [0, 0, 2, 0], [0, 0, 0, 1]]
    # before adding (group.rs:263-286)
    g = rt.GroupShape()
    g.set_transformation(rt.scaling(2, 2, 2))
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    g.add_child(s)
    assert_eq(g.get_children()[0].transformation().m, expected)
    assert len(rt.probe.intersect(g, (10, 0, -10), (0, 0, 1))[0]) == 2
    # after adding (group.rs:288-311)
    g = rt.GroupShape()
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    g.add_child(s)
    g.set_transformation(rt.scaling(2, 2, 2))
    assert_eq(g.get_children()[0].transformation().m, expected)
    assert len(rt.probe.intersect(g, (10, 0, -10), (0, 0, 1))[0]) == 2
    # before and after (group.rs:313-338)
    g = rt.GroupShape()
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    g.set_transformation(rt.scaling(3, 4, 8))
    g.add_child(s)
    g.set_transformation(rt.scaling(2, 2, 2))
    assert_eq(g.get_children()[0].transformation().m, expected)
    assert len(rt.probe.intersect(g, (10, 0, -10), (0, 0, 1))[0]) == 2


def nested(rt):
    g1 = rt.GroupShape()
    g1.set_transformation(rt.rotation_y(PI / 2.0))
    g2 = rt.GroupShape()
    g2.set_transformation(rt.scaling(1, 2, 3))
    s = rt.Sphere()
    s.set_transformation(rt.translation(5, 0, 0))
    g2.add_child(s)
    g1.add_child(g2)
    return g1.get_children()[0].get_children()[0]


def test_world_to_object_point_in_nested_child(rt):  # group.rs:340-360
    assert_abs_diff_eq(rt.probe.world_to_object(nested(rt), (-2, 0, -10)), (5.0, 0.0, -0.66666657))


def test_normal_on_nested_child(rt):  # group.rs:362-388
    assert_abs_diff_eq(rt.probe.normal_at(nested(rt), (1.7321, 1.1547, -5.5774)), (0.2857036, 0.42854306, -0.8571606))


def test_group_bounding_boxes(rt):  # group.rs:390-428
    s = rt.Sphere()
    s.set_transformation(rt.translation(2, 5, -3) * rt.scaling(2, 2, 2))
    c = rt.Cylinder()
    c.minimum_y, c.maximum_y = -2.0, 2.0
    c.set_transformation(rt.translation(-4, -1, 4) * rt.scaling(0.5, 1, 0.5))
    g = rt.GroupShape()
    g.add_child(s)
    g.add_child(c)
    mn, mx = g.bounding_box()
    assert_eq(mn, (-4.5, -3, -5))
    assert_eq(mx, (4, 7, 4.5))
    s = rt.Sphere()
    s.set_transformation(rt.scaling(2, 2, 2))
    c = rt.Cylinder()
    c.minimum_y, c.maximum_y = -1.0, 1.0
    c.set_transformation(rt.scaling(2, 2, 2))
    g = rt.GroupShape()
    g.add_child(s)
    g.add_child(c)
    g.set_transformation(rt.scaling(0.5, 0.5, 0.5))
    b1, b2 = g.bounding_box(), g.parent_space_bounding_box()
    assert_eq(b1[0], b2[0])
    assert_eq(b1[1], b2[1])


def test_group_bbox_culls_children(rt):  # group.rs:430-454
    import pytest

    from ray_tracer_challenge_b200.api import RtcError
    child = rt.TestShape()
    g = rt.GroupShape()
    g.add_child(child)
    rt.probe.intersect(g, (0, 0, -5), (0, 1, 0))
    with pytest.raises(RtcError):
        rt.probe.saved_ray(child)
    rt.probe.intersect(g, (0, 0, -5), (0, 0, 1))
    rt.probe.saved_ray(child)


def test_divide_partitions_children(rt):  # group.rs:456-531 (partition + make_subgroup through divide)
    s1, s2, s3 = rt.Sphere(), rt.Sphere(), rt.Sphere()
    s1.set_transformation(rt.translation(-2, -2, 0))
    s2.set_transformation(rt.translation(-2, 2, 0))
    s3.set_transformation(rt.scaling(4, 4, 4))
    g = rt.GroupShape()
    for s in (s1, s2, s3):
        g.add_child(s)
    g.divide(1)
    kids = g.get_children()
    assert kids[0].handle == s3.handle
    assert [k.handle for k in kids[1].get_children()] == [s1.handle, s2.handle]


def test_partition_left_right_single_children_stay_bare(rt):  # group.rs:456-490 + make_subgroup :70-77
    s1, s2, s3 = rt.Sphere(), rt.Sphere(), rt.Sphere()
    s1.set_transformation(rt.translation(-2, 0, 0))
    s2.set_transformation(rt.translation(2, 0, 0))
    g = rt.GroupShape()
    for s in (s1, s2, s3):
        g.add_child(s)
    g.divide(3)
    # straddler first, then the left partition, then the right; 1-element partitions are not wrapped
    assert [k.handle for k in g.get_children()] == [s3.handle, s1.handle, s2.handle]


def test_divide_with_too_few_children(rt):  # group.rs:533-591
    s1, s2, s3, s4 = rt.Sphere(), rt.Sphere(), rt.Sphere(), rt.Sphere()
    s1.set_transformation(rt.translation(-2, 0, 0))
    s2.set_transformation(rt.translation(2, 1, 0))
    s3.set_transformation(rt.translation(2, -1, 0))
    sub = rt.GroupShape()
    for s in (s1, s2, s3):
        sub.add_child(s)
    g = rt.GroupShape()
    g.add_child(sub)
    g.add_child(s4)
    g.divide(3)
    kids = g.get_children()
    assert kids[0].handle == sub.handle and kids[1].handle == s4.handle
    sub_kids = kids[0].get_children()
    assert sub_kids[0].handle == s1.handle
    assert [k.handle for k in sub_kids[1].get_children()] == [s2.handle, s3.handle]


def test_divide_preserves_pushed_down_transformation(rt):  # group.rs:593-645
    s1, s2, s3 = rt.Sphere(), rt.Sphere(), rt.Sphere()
    s1.set_transformation(rt.translation(-2, 0, 0))
    s2.set_transformation(rt.translation(2, -1, 0))
    s3.set_transformation(rt.translation(2, 1, 0))
    g = rt.GroupShape()
    g.set_transformation(rt.translation(1, 1, 0))
    for s in (s1, s2, s3):
        g.add_child(s)
    g.divide(2)
    kids = g.get_children()
    assert_eq(kids[0].transformation().m, rt.translation(-1, 1, 0).m)
    sub = kids[1].get_children()
    assert_eq(sub[0].transformation().m, rt.translation(3, 0, 0).m)
    assert_eq(sub[1].transformation().m, rt.translation(3, 2, 0).m)


# ------------------------------------------------------------------ csg.rs:177-393
def test_csg_operation_rules(rt):
    T, F = True, False
    table = [
        (UNION, T, T, T, F), (UNION, T, T, F, T), (UNION, T, F, T, F), (UNION, T, F, F, T),
        (UNION, F, T, T, F), (UNION, F, T, F, F), (UNION, F, F, T, T), (UNION, F, F, F, T),
        (INTERSECTION, T, T, T, T), (INTERSECTION, T, T, F, F), (INTERSECTION, T, F, T, T),
        (INTERSECTION, T, F, F, F), (INTERSECTION, F, T, T, T), (INTERSECTION, F, T, F, T),
        (INTERSECTION, F, F, T, F), (INTERSECTION, F, F, F, F),
        (DIFFERENCE, T, T, T, F), (DIFFERENCE, T, T, F, T), (DIFFERENCE, T, F, T, F), (DIFFERENCE, T, F, F, T),
        (DIFFERENCE, F, T, T, T), (DIFFERENCE, F, T, F, T), (DIFFERENCE, F, F, T, F), (DIFFERENCE, F, F, F, F),
    ]
    for op, hit_s1, in_s1, in_s2, expected in table:
        assert rt.probe.csg_allowed(op, hit_s1, in_s1, in_s2) == expected, (op, hit_s1, in_s1, in_s2)


def test_csg_filter_intersections(rt):
    for op, x0, x1 in ((UNION, 0, 3), (INTERSECTION, 1, 2), (DIFFERENCE, 0, 1)):
        s1, s2 = rt.Sphere(), rt.Cube()
        c = rt.CSG(op, s1, s2)
        xs = [(1.0, s1), (2.0, s2), (3.0, s1), (4.0, s2)]
        assert_eq(rt.probe.csg_filter(c, xs), [xs[x0][0], xs[x1][0]])


def test_ray_vs_csg(rt):
    c = rt.CSG(UNION, rt.Sphere(), rt.Cube())
    assert len(rt.probe.local_intersect(c, (0, 2, -5), (0, 0, 1))[0]) == 0
    s1 = rt.Sphere()
    s2 = rt.Sphere.build(rt.translation(0, 0, 0.5), rt.Material())
    c = rt.CSG(UNION, s1, s2)
    ts, objs, _ = rt.probe.local_intersect(c, (0, 0, -5), (0, 0, 1))
    assert_eq(ts, [4.0, 6.5])
    assert [int(o) for o in objs] == [s1.handle, s2.handle]


def test_csg_bounding_box_and_cull(rt):
    import pytest

    from ray_tracer_challenge_b200.api import RtcError
    right = rt.Sphere()
    right.set_transformation(rt.translation(2, 3, 4))
    shape = rt.CSG(DIFFERENCE, rt.Sphere(), right)
    mn, mx = shape.bounding_box()
    assert_eq(mn, (-1, -1, -1))
    assert_eq(mx, (3, 4, 5))
    left, right = rt.TestShape(), rt.TestShape()
    shape = rt.CSG(DIFFERENCE, left, right)
    rt.probe.intersect(shape, (0, 0, -5), (0, 1, 0))
    for t in (left, right):
        with pytest.raises(RtcError):
            rt.probe.saved_ray(t)
    rt.probe.intersect(shape, (0, 0, -5), (0, 0, 1))
    rt.probe.saved_ray(left)
    rt.probe.saved_ray(right)


def test_csg_includes(rt):  # csg.rs:111-117, group.rs:87-93
    a, b, c = rt.Sphere(), rt.Cube(), rt.Sphere()
    g = rt.GroupShape()
    g.add_child(a)
    csg = rt.CSG(UNION, g, b)
    assert rt.probe.includes(csg, a) and rt.probe.includes(csg, b) and rt.probe.includes(csg, g)
    assert not rt.probe.includes(csg, c)


# ------------------------------------------------------------------ bounding_box.rs:135-287
def test_bbox_contains(rt):
    mn, mx = (5, -2, 0), (11, 4, 7)
    cases = [((5, -2, 0), True), ((11, 4, 7), True), ((8, 1, 3), True), ((3, 0, 3), False), ((8, -4, 3), False),
             ((8, 1, -1), False), ((13, 1, 3), False), ((8, 5, 3), False), ((8, 1, 8), False)]
    for p, expected in cases:
        assert rt.probe.bbox_contains_point(mn, mx, p) == expected
    boxes = [((5, -2, 0), (11, 4, 7), True), ((6, -1, 1), (10, 3, 6), True), ((4, -3, -1), (10, 3, 6), False),
             ((6, -1, 1), (12, 5, 8), False)]
    for a, b, expected in boxes:
        assert rt.probe.bbox_contains_box(mn, mx, a, b) == expected


def test_bbox_transform(rt):
    m = rt.rotation_x(PI / 4.0) * rt.rotation_y(PI / 4.0)
    mn, mx = rt.probe.bbox_transform((-1, -1, -1), (1, 1, 1), m)
    assert_abs_diff_eq(mn, (-1.4142135, -1.7071067, -1.7071067))
    assert_abs_diff_eq(mx, (1.4142135, 1.7071067, 1.7071067))


def test_bbox_ray_tables(rt):
    table1 = [((5, 0.5, 0), (-1, 0, 0), True), ((-5, 0.5, 0), (1, 0, 0), True), ((0.5, 5, 0), (0, -1, 0), True),
              ((0.5, -5, 0), (0, 1, 0), True), ((0.5, 0, 5), (0, 0, -1), True), ((0.5, 0, -5), (0, 0, 1), True),
              ((0, 0.5, 0), (0, 0, 1), True), ((-2, 0, 0), (2, 4, 6), False), ((0, -2, 0), (6, 2, 4), False),
              ((0, 0, -2), (4, 6, 2), False), ((2, 0, 2), (0, 0, -1), False), ((0, 2, 2), (0, -1, 0), False),
              ((2, 2, 0), (-1, 0, 0), False)]
    for o, d, expected in table1:
        assert rt.probe.bbox_intersects((-1, -1, -1), (1, 1, 1), o, norm3(d)) == expected, o
    table2 = [((15, 1, 2), (-1, 0, 0), True), ((-5, -1, 4), (1, 0, 0), True), ((7, 6, 5), (0, -1, 0), True),
              ((9, -5, 6), (0, 1, 0), True), ((8, 2, 12), (0, 0, -1), True), ((6, 0, -5), (0, 0, 1), True),
              ((8, 1, 3.5), (0, 0, 1), True), ((9, -1, -8), (2, 4, 6), False), ((8, 3, -4), (6, 2, 4), False),
              ((9, -1, -2), (4, 6, 2), False), ((4, 0, 9), (0, 0, -1), False), ((8, 6, -1), (0, -1, 0), False),
              ((12, 5, 4), (-1, 0, 0), False)]
    for o, d, expected in table2:
        assert rt.probe.bbox_intersects((5, -2, 0), (11, 4, 7), o, norm3(d)) == expected, o


def test_bbox_splits(rt):
    cases = [((-1, -4, -5), (9, 6, 5), (4, 6, 5), (4, -4, -5)), ((-1, -2, -3), (9, 5.5, 3), (4, 5.5, 3), (4, -2, -3)),
             ((-1, -2, -3), (5, 8, 3), (5, 3, 3), (-1, 3, -3)), ((-1, -2, -3), (5, 3, 7), (5, 3, 2), (-1, -2, 2))]
    for mn, mx, left_max, right_min in cases:
        lmin, lmax, rmin, rmax = rt.probe.bbox_split(mn, mx)
        assert_eq(lmin, mn)
        assert_eq(lmax, left_max)
        assert_eq(rmin, right_min)
        assert_eq(rmax, mx)
