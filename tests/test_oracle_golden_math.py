"""Golden vectors of matrix.rs, transformations.rs and obj_parser.rs replayed against the CPU oracle.

These pin the host-side scene construction (cofactor inverse, transform products, OBJ normalisation and fan
triangulation) that feeds the hot path; tests/golden/triangles.obj is the reference's only on-disk fixture
(lib/resources/test/triangles.obj)."""
import math
import os

import numpy as np

from tests.helpers import F32_EPSILON, assert_abs_diff_eq, assert_eq

PI = float(np.float32(math.pi))
FRAC_1_SQRT_2 = float(np.float32(0.70710678118654752440))
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def mat(rt, rows):
    return rt.Matrix(np.asarray(rows, np.float32))


def mul_point(rt, m, p):
    return rt.probe.mat_mul_tuple(m, (*p, 1.0))[:3]


def mul_vector(rt, m, v):
    return rt.probe.mat_mul_tuple(m, (*v, 0.0))[:3]


# ------------------------------------------------------------------ matrix.rs:218-398
def test_determinants(rt):
    assert mat(rt, [[-2, -8, 3, 5], [-3, 1, 7, 3], [1, 2, -9, 6], [-6, 7, 7, -9]]).determinant() == -4071.0
    assert mat(rt, [[6, 4, 4, 4], [5, 5, 7, 6], [4, -9, 3, -7], [9, 1, 7, -6]]).determinant() == -2120.0
    assert mat(rt, [[-4, 2, -2, -3], [9, 6, 2, 6], [0, -5, 1, -5], [0, 0, 0, 0]]).determinant() == 0.0


def test_matrix_inversions(rt):
    cases = [
        ([[-5, 2, 6, -8], [1, -5, 1, 8], [7, 7, -6, -7], [1, -3, 7, 4]], 532.0,
         [[116, 240, 128, -24], [-430, -775, -236, 277], [-42, -119, -28, 105], [-278, -433, -160, 163]]),
        ([[8, -5, 9, 2], [7, 5, 6, 1], [-6, 0, 9, 6], [-3, 0, -9, -4]], -585.0,
         [[90, 90, 165, 315], [45, -72, -15, -18], [-210, -210, -255, -540], [405, 405, 450, 1125]]),
        ([[9, 3, 0, 9], [-5, -2, -6, -3], [-4, 9, 6, 4], [-7, 6, 6, 2]], 1620.0,
         [[-66, -126, 234, -360], [-126, 54, 594, -540], [-47, -237, -177, 210], [288, 108, -432, 540]]),
    ]
    for a, det, adj in cases:
        expected = np.asarray(adj, np.float32) * (np.float32(1.0) / np.float32(det))
        assert_abs_diff_eq(mat(rt, a).inverse().m, expected)


def test_inverse_undoes_multiplication(rt):
    a = mat(rt, [[3, -9, 7, 3], [3, -8, 2, -9], [-4, 4, 4, 1], [-6, 5, -1, 1]])
    b = mat(rt, [[8, 2, 2, 2], [3, -1, 7, 0], [7, 0, 5, 4], [6, -2, 0, 5]])
    c = a * b
    assert_abs_diff_eq((c * b.inverse()).m, a.m, epsilon=10.0 * F32_EPSILON)


def test_matrix_multiplication_and_transpose(rt):  # matrix.rs:233-268
    a = mat(rt, [[1, 2, 3, 4], [5, 6, 7, 8], [9, 8, 7, 6], [5, 4, 3, 2]])
    b = mat(rt, [[-2, 1, 2, 3], [3, 2, 1, -1], [4, 3, 6, 5], [1, 2, 7, 8]])
    assert_eq((a * b).m, [[20, 22, 50, 48], [44, 54, 114, 108], [40, 58, 110, 102], [16, 26, 46, 42]])
    assert_eq(a.transpose().m, np.asarray(a.m).T)


# ------------------------------------------------------------------ transformations.rs:76-256
def test_translation_scaling(rt):
    t = rt.translation(5, -3, 2)
    assert_eq(mul_point(rt, t, (-3, 4, 5)), (2, 1, 7))
    assert_eq(mul_point(rt, t.inverse(), (-3, 4, 5)), (-8, 7, 3))
    assert_eq(mul_vector(rt, t, (-3, 4, 5)), (-3, 4, 5))
    s = rt.scaling(2, 3, 4)
    assert_eq(mul_point(rt, s, (-4, 6, 8)), (-8, 18, 32))
    assert_eq(mul_vector(rt, s.inverse(), (-4, 6, 8)), (-2, 2, 2))


def test_rotations(rt):
    assert_abs_diff_eq(mul_point(rt, rt.rotation_x(PI / 4), (0, 1, 0)), (0, FRAC_1_SQRT_2, FRAC_1_SQRT_2))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_x(PI / 2), (0, 1, 0)), (0, 0, 1))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_x(PI / 4).inverse(), (0, 1, 0)), (0, FRAC_1_SQRT_2, -FRAC_1_SQRT_2))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_y(PI / 4), (0, 0, 1)), (FRAC_1_SQRT_2, 0, FRAC_1_SQRT_2))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_y(PI / 2), (0, 0, 1)), (1, 0, 0))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_z(PI / 4), (0, 1, 0)), (-FRAC_1_SQRT_2, FRAC_1_SQRT_2, 0))
    assert_abs_diff_eq(mul_point(rt, rt.rotation_z(PI / 2), (0, 1, 0)), (-1, 0, 0))


def test_shearing_and_chaining(rt):
    cases = [((1, 0, 0, 0, 0, 0), (5, 3, 4)), ((0, 1, 0, 0, 0, 0), (6, 3, 4)), ((0, 0, 1, 0, 0, 0), (2, 5, 4)),
             ((0, 0, 0, 1, 0, 0), (2, 7, 4)), ((0, 0, 0, 0, 1, 0), (2, 3, 6)), ((0, 0, 0, 0, 0, 1), (2, 3, 7))]
    for args, expected in cases:
        assert_eq(mul_point(rt, rt.shearing(*args), (2, 3, 4)), expected)
    chained = rt.translation(10, 5, 7) * rt.scaling(5, 5, 5) * rt.rotation_x(PI / 2)
    assert_eq(mul_point(rt, chained, (1, 0, 1)), (15, 0, 7))


def test_view_transforms(rt):
    assert_eq(rt.view_transform((0, 0, 0), (0, 0, -1), (0, 1, 0)).m, np.eye(4))
    assert_eq(rt.view_transform((0, 0, 0), (0, 0, 1), (0, 1, 0)).m, rt.scaling(-1, 1, -1).m)
    assert_eq(rt.view_transform((0, 0, 8), (0, 0, 0), (0, 1, 0)).m, rt.translation(0, 0, -8).m)
    t = rt.view_transform((1, 3, 2), (4, -2, 8), (1, 1, 0))
    expected = [[-0.50709254, 0.50709254, 0.6761234, -2.366432], [0.76771593, 0.6060915, 0.12121832, -2.828427],
                [-0.35856858, 0.59761435, -0.71713716, -0.00000023841858], [0, 0, 0, 1]]
    assert_abs_diff_eq(t.m, expected)


# ------------------------------------------------------------------ obj_parser.rs:301-535
def tri_points(t):
    p1, e1, e2, _ = t.geometry()
    return p1, p1 + e1, p1 + e2


def test_obj_normalises_vertices(rt):  # obj_parser.rs:329-343 — observed through a face
    text = "v -50 10 20\nv 30 -40 0\nv 10 -20 50\nv -10 30 10\nf 1 2 3\nf 1 3 4\n"
    g = rt.parse_obj(text)
    t1, t2 = g.get_children()
    a, b, c = tri_points(t1)
    assert_eq(a, (-1.0, 0.375, -0.125))
    assert_eq(b, (1.0, -0.875, -0.625))
    assert_eq(c, (0.5, -0.375, 0.625))
    assert_eq(tri_points(t2)[2], (0.0, 0.875, -0.375))


def test_obj_fan_triangulation(rt):  # obj_parser.rs:345-397
    text = "\nv -1 1 0\nv -1 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 1\n\nf 1 2 3 4 5\n"
    kids = rt.parse_obj(text).get_children()
    assert len(kids) == 3
    # one common scale (largest half-span = 1); every axis centred on its own mid-point (obj_parser.rs:250-264)
    v = {1: (-1, 0.5, -0.5), 2: (-1, -0.5, -0.5), 3: (1, -0.5, -0.5), 4: (1, 0.5, -0.5), 5: (0, 0.5, 0.5)}
    for t, idx in zip(kids, ((1, 2, 3), (1, 3, 4), (1, 4, 5))):
        for got, i in zip(tri_points(t), idx):
            assert_eq(got, v[i])


def test_obj_groups_fixture(rt):  # obj_parser.rs:407-452 with lib/resources/test/triangles.obj
    with open(os.path.join(GOLDEN, "triangles.obj")) as fh:
        g = rt.parse_obj(fh.read())
    g1, g2 = g.get_children()
    t1, t2 = g1.get_children()[0], g2.get_children()[0]
    assert_eq(tri_points(t1)[0], (-1, 1, 0))
    assert_eq(tri_points(t1)[1], (-1, 0, 0))
    assert_eq(tri_points(t1)[2], (1, 0, 0))
    assert_eq(tri_points(t2)[0], (-1, 1, 0))
    assert_eq(tri_points(t2)[1], (1, 0, 0))
    assert_eq(tri_points(t2)[2], (1, -1, 0))


def test_obj_single_group_is_returned_bare(rt):  # obj_parser.rs:454-489
    verts = "v .7 0 1\nv .5 -1 1\nv .5 0 1\nv -1 1 0\nv .6 .6 .6\nv 1 .7 -1\n"
    assert len(rt.parse_obj(verts + "f 1 2 3\nf 4 5 6\n").get_children()) == 2
    assert len(rt.parse_obj(verts + "g TestGroup\nf 1 2 3\nf 4 5 6\n").get_children()) == 2


def test_obj_faces_with_normals_make_smooth_triangles(rt):  # obj_parser.rs:503-534
    from ray_tracer_challenge_b200.api import SG_SMOOTH_TRIANGLE
    text = "v 0 1 0\nv -1 0 0\nv 1 0 0\nvn -1 0 0\nvn 1 0 0\nvn 0 1 0\nf 1//3 2//1 3//2\nf 1/0/3 2/102/1 3/14/2\n"
    kids = rt.parse_obj(text).get_children()
    assert [k.kind() for k in kids] == [SG_SMOOTH_TRIANGLE, SG_SMOOTH_TRIANGLE]
    for t in kids:
        a, b, c = tri_points(t)
        assert_eq(a, (0, 0.5, 0))  # scale = x half-span = 1; y centred on its mid-point 0.5
        assert_eq(b, (-1, -0.5, 0))
        assert_eq(c, (1, -0.5, 0))


def test_obj_errors(rt):
    import pytest

    from ray_tracer_challenge_b200.api import RtcError
    with pytest.raises(RtcError):
        rt.parse_obj("v 1 2\nf 1 2 3\n")
    with pytest.raises(RtcError):
        rt.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\nv 2 2 2\n")  # vertex after first face (obj_parser.rs:114-118)
    with pytest.raises(RtcError):
        rt.parse_obj("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n")
