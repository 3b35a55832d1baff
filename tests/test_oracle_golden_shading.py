"""Golden vectors of light/{phong_lighting,rectangle_light}.rs, pattern/*.rs and canvas.rs replayed against
the CPU oracle."""
import dataclasses
import math

import numpy as np

from tests.helpers import assert_abs_diff_eq, assert_eq

FRAC_1_SQRT_2 = float(np.float32(0.70710678118654752440))
PI = float(np.float32(math.pi))
WHITE, BLACK = (1, 1, 1), (0, 0, 0)
RED, YELLOW, GREEN, CYAN, BLUE, PURPLE, BROWN = (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1), (0, 0, 1), (1, 0, 1), (1, 0.5, 0)


# ------------------------------------------------------------------ phong_lighting.rs:77-271
def test_phong_scenarios(rt):
    s = rt.Sphere()
    m = rt.Material()
    n = (0, 0, -1)
    cases = [
        ((0, 0, -1), (0, 0, -10), 1.0, (1.9, 1.9, 1.9), True),
        ((0, FRAC_1_SQRT_2, FRAC_1_SQRT_2), (0, 0, -10), 1.0, (1.0, 1.0, 1.0), True),
        ((0, -FRAC_1_SQRT_2, -FRAC_1_SQRT_2), (0, 10, -10), 1.0, (1.6363853,) * 3, False),
        ((0, 0, -1), (0, 0, 10), 1.0, (0.1, 0.1, 0.1), False),
        ((0, 0, -1), (0, 0, -10), 0.0, (0.1, 0.1, 0.1), True),
    ]
    for eye, light_pos, li, expected, exact in cases:
        got = rt.probe.phong(s, m, rt.PointLight(light_pos, WHITE), (0, 0, 0), eye, n, li)
        (assert_eq if exact else assert_abs_diff_eq)(got, expected, msg=str((eye, light_pos)))
    # light offset 45 degrees: 0.1 + 0.9 * FRAC_1_SQRT_2 evaluated in f32 (phong_lighting.rs:117-137)
    e = float(np.float32(0.1) + np.float32(0.9) * np.float32(FRAC_1_SQRT_2))
    got = rt.probe.phong(s, m, rt.PointLight((0, 10, -10), WHITE), (0, 0, 0), (0, 0, -1), n, 1.0)
    assert_eq(got, (e, e, e))


def test_phong_with_pattern(rt):  # phong_lighting.rs:197-237
    m = rt.Material(ambient=1.0, diffuse=0.0, specular=0.0, color=(0.5, 0.5, 0.5), pattern=rt.Stripes(WHITE, BLACK))
    s = rt.Sphere()
    light = rt.PointLight((0, 0, -10), WHITE)
    assert_eq(rt.probe.phong(s, m, light, (0.9, 0, 0), (0, 0, -1), (0, 0, -1), 1.0), WHITE)
    assert_eq(rt.probe.phong(s, m, light, (1.1, 0, 0), (0, 0, -1), (0, 0, -1), 1.0), BLACK)


def test_phong_attenuates_by_light_intensity(rt):  # phong_lighting.rs:239-271
    s = rt.Sphere()
    m = rt.Material(ambient=0.1, diffuse=0.9, specular=0.0, color=WHITE)
    light = rt.PointLight((0, 0, -10), WHITE)
    for li, expected in ((1.0, WHITE), (0.5, (0.55, 0.55, 0.55)), (0.0, (0.1, 0.1, 0.1))):
        assert_abs_diff_eq(rt.probe.phong(s, m, light, (0, 0, -1), (0, 0, -1), (0, 0, -1), li), expected)


# ------------------------------------------------------------------ rectangle_light.rs:98-166
def test_rectangle_light_construction(rt):
    w = rt.World([], rt.RectangleLight(WHITE, (0, 0, 0), (2, 0, 0), 4, (0, 0, 1), 2, [0.5]))
    info = rt.probe.light_info(w)
    assert_eq(info["u_vec"], (0.5, 0, 0))
    assert_eq(info["v_vec"], (0, 0, 0.5))
    assert info["cells"] == 8
    assert_eq(info["position"], (1, 0, 0.5))


def test_point_on_rectangle_light(rt):
    w = rt.World([], rt.RectangleLight(WHITE, (0, 0, 0), (2, 0, 0), 4, (0, 0, 1), 2, [0.3, 0.7]))
    cases = [(0, 0, (0.15, 0, 0.35)), (1, 0, (0.65, 0, 0.35)), (0, 1, (0.15, 0, 0.85)), (2, 0, (1.15, 0, 0.35)),
             (3, 1, (1.65, 0, 0.85))]
    for u, v, expected in cases:
        assert_eq(rt.probe.point_on_light(w, u, v), expected)


def test_rectangle_light_intensity_at(rt):
    cases = [((0, 0, 2), 0.0), ((1, -1, 2), 0.5), ((1.5, 0, 2), 0.75), ((1.25, 1.25, 3), 0.75), ((0, 0, -2), 1.0)]
    for p, expected in cases:
        w = rt.World.default()
        w.set_light(rt.RectangleLight(WHITE, (-0.5, -0.5, -5), (1, 0, 0), 2, (0, 1, 0), 2, [0.7, 0.3, 0.9, 0.1, 0.5]))
        assert rt.probe.intensity_at(w, p) == expected, p


def test_rectangle_light_shading_uses_centre_position(rt):
    """phong_lighting.rs:36 uses light.position() (the rectangle centre) for diffuse/specular, and only scales by
    the sampled intensity."""
    w = rt.World.default()
    w.set_light(rt.RectangleLight(WHITE, (-0.5, -0.5, -5), (1, 0, 0), 2, (0, 1, 0), 2, [0.5]))
    a = rt.probe.color_at(w, (0, 0, -5), (0, 0, 1), 1)
    w2 = rt.World.default()
    w2.set_light(rt.PointLight((0, 0, -5), WHITE))
    b = rt.probe.color_at(w2, (0, 0, -5), (0, 0, 1), 1)
    assert_eq(a, b)


# ------------------------------------------------------------------ pattern/*.rs
def test_stripes(rt):  # stripes.rs:53-84
    p = rt.Stripes()
    for q in ((0, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 0, 2), (0.9, 0, 0), (-1.1, 0, 0)):
        assert_eq(rt.probe.pattern_color_at(p, q), WHITE, msg=str(q))
    for q in ((1, 0, 0), (-0.1, 0, 0), (-1, 0, 0)):
        assert_eq(rt.probe.pattern_color_at(p, q), BLACK, msg=str(q))


def test_gradient(rt):  # gradient.rs:50-65
    p = rt.Gradient()
    for x, e in ((0, 1.0), (0.25, 0.75), (0.5, 0.5), (0.75, 0.25)):
        assert_eq(rt.probe.pattern_color_at(p, (x, 0, 0)), (e, e, e))


def test_rings(rt):  # rings.rs:58-65
    p = rt.Rings(WHITE, BLACK)
    assert_eq(rt.probe.pattern_color_at(p, (0, 0, 0)), WHITE)
    for q in ((1, 0, 0), (0, 0, 1), (0.708, 0, 0.708)):
        assert_eq(rt.probe.pattern_color_at(p, q), BLACK)


def test_checkers(rt):  # checkers.rs:54-75
    p = rt.Checkers()
    for axis in range(3):
        for v, e in ((0.0, WHITE), (0.99, WHITE), (1.01, BLACK)):
            q = [0.0, 0.0, 0.0]
            q[axis] = v
            assert_eq(rt.probe.pattern_color_at(p, q), e)
    # |x|+|y|+|z| rather than the book's sum of floors
    assert_eq(rt.probe.pattern_color_at(p, (0.6, 0.6, 0)), BLACK)


def test_sine_2d(rt):  # sine_2d.rs:52-73
    p = rt.Sine2D()
    for q in ((0, 0, 0), (0, 1, 0), (0, 2, 0)):
        assert_eq(rt.probe.pattern_color_at(p, q), WHITE)
    assert_abs_diff_eq(rt.probe.pattern_color_at(p, (0, 0, 1)), (0.77015114,) * 3)
    assert_eq(rt.probe.pattern_color_at(p, (0, 0, 2)), (0.29192656,) * 3)
    assert_eq(rt.probe.pattern_color_at(p, (0, 0, PI)), BLACK)


def test_pattern_transformations(rt):  # pattern.rs:98-122
    obj = rt.Sphere.build(rt.scaling(2, 2, 2), rt.Material())
    assert_eq(rt.probe.pattern_color_at(rt.TestPattern(), (2, 3, 4), obj), (1, 1.5, 2))
    tp = rt.TestPattern()
    tp.set_transformation(rt.scaling(2, 2, 2))
    assert_eq(rt.probe.pattern_color_at(tp, (2, 3, 4), rt.Sphere()), (1, 1.5, 2))
    tp = rt.TestPattern()
    tp.set_transformation(rt.translation(0.5, 1.0, 1.5))
    assert_eq(rt.probe.pattern_color_at(tp, (2.5, 3, 3.5), obj), (0.75, 0.5, 0.25))


# ------------------------------------------------------------------ pattern/uv.rs:386-639
def test_uv_checkers(rt):
    p = rt.UVCheckers(2.0, 2.0, BLACK, WHITE)
    for u, v, e in ((0.0, 0.0, BLACK), (0.5, 0.0, WHITE), (0.0, 0.5, WHITE), (0.5, 0.5, BLACK), (1.0, 1.0, BLACK)):
        assert_eq(rt.probe.uv_color_at(p, u, v), e)


def test_spherical_map(rt):
    cases = [((0, 0, -1), 0.0, 0.5), ((1, 0, 0), 0.25, 0.5), ((0, 0, 1), 0.5, 0.5), ((-1, 0, 0), 0.75, 0.5),
             ((0, 1, 0), 0.5, 1.0), ((0, -1, 0), 0.5, 0.0), ((FRAC_1_SQRT_2, FRAC_1_SQRT_2, 0), 0.25, 0.75)]
    for p, u, v in cases:
        assert_abs_diff_eq(rt.probe.uv_map(0, p), (u, v), msg=str(p))


def test_texture_map_spherical_checkers(rt):
    tm = rt.TextureMap(rt.UVCheckers(16.0, 8.0, BLACK, WHITE), 0)
    cases = [((0.4315, 0.4670, 0.7719), WHITE), ((-0.9654, 0.2552, -0.0534), BLACK), ((0.1039, 0.7090, 0.6975), WHITE),
             ((-0.4986, -0.7856, -0.3663), BLACK), ((-0.0317, -0.9395, 0.3411), BLACK),
             ((0.4809, -0.7721, 0.4154), BLACK), ((0.0285, -0.9612, -0.2745), BLACK),
             ((-0.5734, -0.2162, -0.7903), WHITE), ((0.7688, -0.1470, 0.6223), BLACK),
             ((-0.7652, 0.2175, 0.6060), BLACK)]
    for p, e in cases:
        assert_eq(rt.probe.pattern_color_at(tm, p), e, msg=str(p))


def test_planar_map(rt):
    cases = [((0.25, 0, 0.5), 0.25, 0.5), ((0.25, 0, -0.25), 0.25, 0.75), ((0.25, 0.5, -0.25), 0.25, 0.75),
             ((1.25, 0, 0.5), 0.25, 0.5), ((0.25, 0, -1.75), 0.25, 0.25), ((1, 0, -1), 0.0, 0.0), ((0, 0, 0), 0.0, 0.0)]
    for p, u, v in cases:
        assert_eq(rt.probe.uv_map(1, p), (u, v), msg=str(p))


def test_cylindrical_map(rt):
    cases = [((0, 0, -1), 0.0, 0.0), ((0, 0.5, -1), 0.0, 0.07957747), ((0, 1, -1), 0.0, 0.15915494),
             ((0.70711, 0.5, -0.70711), 0.125, 0.07957747), ((1, 0.5, 0), 0.25, 0.07957747),
             ((0.70711, 0.5, 0.70711), 0.375, 0.07957747), ((0, -0.25, 1), 0.5, 0.9602113),
             ((-0.70711, 0.5, 0.70711), 0.625, 0.07957747), ((-1, 1.25, 0), 0.75, 0.19894367),
             ((-0.70711, 0.5, -0.70711), 0.875, 0.07957747)]
    for p, u, v in cases:
        assert_abs_diff_eq(rt.probe.uv_map(2, p), (u, v), msg=str(p))


def test_align_check(rt):
    ac = rt.AlignCheck((1, 1, 1), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1))
    cases = [(0.5, 0.5, (1, 1, 1)), (0.1, 0.9, (1, 0, 0)), (0.9, 0.9, (1, 1, 0)), (0.1, 0.1, (0, 1, 0)),
             (0.9, 0.1, (0, 1, 1))]
    for u, v, e in cases:
        assert_eq(rt.probe.uv_color_at(ac, u, v), e)


FRONT, BACK, LEFT, RIGHT, UP, DOWN = range(6)


def test_face_from_point(rt):
    cases = [((-1, 0.5, -0.25), LEFT), ((1.1, -0.75, 0.8), RIGHT), ((0.1, 0.6, 0.9), FRONT), ((-0.7, 0, -2), BACK),
             ((0.5, 1, 0.9), UP), ((-0.2, -1.3, 1.1), DOWN)]
    for p, face in cases:
        assert rt.probe.face_from_point(p) == face, p


def test_cube_uv_faces(rt):
    cases = {
        FRONT: [((-0.5, 0.5, 1), 0.25, 0.75), ((0.5, -0.5, 1), 0.75, 0.25)],
        BACK: [((0.5, 0.5, -1), 0.25, 0.75), ((-0.5, -0.5, -1), 0.75, 0.25)],
        LEFT: [((-1, 0.5, -0.5), 0.25, 0.75), ((-1, -0.5, 0.5), 0.75, 0.25)],
        RIGHT: [((1, 0.5, 0.5), 0.25, 0.75), ((1, -0.5, -0.5), 0.75, 0.25)],
        UP: [((-0.5, 1, -0.5), 0.25, 0.75), ((0.5, 1, 0.5), 0.75, 0.25)],
        DOWN: [((-0.5, -1, 0.5), 0.25, 0.75), ((0.5, -1, -0.5), 0.75, 0.25)],
    }
    for face, rows in cases.items():
        for p, u, v in rows:
            assert_eq(rt.probe.cube_uv(face, p), (u, v), msg=str((face, p)))


def align_check_cubic_map(rt):  # uv.rs:328-344
    left = rt.AlignCheck(YELLOW, CYAN, RED, BLUE, BROWN)
    front = rt.AlignCheck(CYAN, RED, YELLOW, BROWN, GREEN)
    right = rt.AlignCheck(RED, YELLOW, PURPLE, GREEN, WHITE)
    back = rt.AlignCheck(GREEN, PURPLE, CYAN, WHITE, BLUE)
    up = rt.AlignCheck(BROWN, CYAN, PURPLE, RED, YELLOW)
    down = rt.AlignCheck(PURPLE, BROWN, GREEN, BLUE, WHITE)
    return rt.CubicMap(front, back, left, right, up, down)


CUBE_MAP_TABLE = [
    ((-1, 0, 0), YELLOW), ((-1, 0.9, -0.9), CYAN), ((-1, 0.9, 0.9), RED), ((-1, -0.9, -0.9), BLUE),
    ((-1, -0.9, 0.9), BROWN), ((0, 0, 1), CYAN), ((-0.9, 0.9, 1), RED), ((0.9, 0.9, 1), YELLOW),
    ((-0.9, -0.9, 1), BROWN), ((0.9, -0.9, 1), GREEN), ((1, 0, 0), RED), ((1, 0.9, 0.9), YELLOW),
    ((1, 0.9, -0.9), PURPLE), ((1, -0.9, 0.9), GREEN), ((1, -0.9, -0.9), WHITE), ((0, 0, -1), GREEN),
    ((0.9, 0.9, -1), PURPLE), ((-0.9, 0.9, -1), CYAN), ((0.9, -0.9, -1), WHITE), ((-0.9, -0.9, -1), BLUE),
    ((0, 1, 0), BROWN), ((-0.9, 1, -0.9), CYAN), ((0.9, 1, -0.9), PURPLE), ((-0.9, 1, 0.9), RED),
    ((0.9, 1, 0.9), YELLOW), ((0, -1, 0), PURPLE), ((-0.9, -1, 0.9), BROWN), ((0.9, -1, 0.9), GREEN),
    ((-0.9, -1, -0.9), BLUE), ((0.9, -1, -0.9), WHITE),
]


def test_colors_on_mapped_cube(rt):
    pattern = align_check_cubic_map(rt)
    for p, e in CUBE_MAP_TABLE:
        assert_eq(rt.probe.pattern_color_at(pattern, p), e, msg=str(p))


# ------------------------------------------------------------------ canvas.rs:227-278
def test_scale_color_truncates(rt):
    assert rt.probe.scale_color(1.5) == 255
    assert rt.probe.scale_color(0.5) == 127  # canvas.rs:242 — "book says 128, but I'll trust Rust's rounding"
    assert rt.probe.scale_color(-0.5) == 0
    assert rt.probe.scale_color(1.0) == 255
    assert rt.probe.scale_color(0.8) == 204
    assert rt.probe.scale_color(0.6) == 153
    assert rt.probe.scale_color(float("nan")) == 255  # NaN.min(255) = 255 in Rust


def test_ppm_writer(rt):
    c = rt.Canvas(5, 3, u8=np.zeros((3, 5, 3), np.uint8))
    c._u8[0, 0] = (255, 0, 0)
    c._u8[1, 2] = (0, 127, 0)
    c._u8[2, 4] = (0, 0, 255)
    lines = c.to_ppm().splitlines()
    assert lines[:3] == ["P3", "5 3", "255"]
    assert lines[3] == "255 0 0 0 0 0 0 0 0 0 0 0 0 0 0"
    assert lines[4] == "0 0 0 0 0 0 0 127 0 0 0 0 0 0 0"
    assert lines[5] == "0 0 0 0 0 0 0 0 0 0 0 0 0 0 255"
    c = rt.Canvas(10, 2, u8=np.tile(np.array([255, 204, 153], np.uint8), (2, 10, 1)))
    lines = c.to_ppm().splitlines()[3:]
    assert lines[0] == "255 204 153 255 204 153 255 204 153 255 204 153 255 204 153 255 204"
    assert lines[1] == "153 255 204 153 255 204 153 255 204 153 255 204 153"
    assert lines[2] == lines[0] and lines[3] == lines[1]
