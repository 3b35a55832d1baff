"""Golden vectors of lib/src/world.rs (tests at world.rs:304-845) replayed against the CPU oracle."""
import dataclasses
import math

import numpy as np

from tests.helpers import F32_EPSILON, assert_abs_diff_eq, assert_eq

FRAC_1_SQRT_2 = float(np.float32(0.70710678118654752440))
SQRT_2 = float(np.float32(1.41421356237309504880))
ACNE = float(np.float32(F32_EPSILON * np.float32(10000.0)))  # world.rs:210


def glass_sphere(rt, **kw):  # world.rs:383-393
    return rt.Sphere.build(rt.identity_4x4(), rt.Material(transparency=1.0, refractive_index=1.5, **kw))


def test_intersect_world_with_ray(rt):  # world.rs:322-332
    w = rt.World.default()
    ts, _ = rt.probe.world_intersect(w, (0, 0, -5), (0, 0, 1))
    assert_eq(ts, [4.0, 4.5, 5.5, 6.0])


def test_precompute_intersection_state(rt):  # world.rs:334-344
    s = rt.Sphere()
    c = rt.probe.precompute((0, 0, -5), (0, 0, 1), [(4.0, s)])
    assert c["distance"] == 4.0
    assert_eq(c["point"], (0, 0, -1))
    assert_eq(c["eye_vector"], (0, 0, -1))
    assert_eq(c["surface_normal"], (0, 0, -1))
    assert not c["inside"]  # world.rs:346-353


def test_precompute_hit_occurs_inside(rt):  # world.rs:355-369
    s = rt.Sphere()
    c = rt.probe.precompute((0, 0, 0), (0, 0, 1), [(1.0, s)])
    assert_eq(c["point"], (0, 0, 1))
    assert_eq(c["eye_vector"], (0, 0, -1))
    assert c["inside"]
    assert_eq(c["surface_normal"], (0, 0, -1))


def test_precompute_reflection_vector(rt):  # world.rs:371-381
    c = rt.probe.precompute((0, 1, -1), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, rt.Plane())])
    assert_eq(c["reflection_vector"], (0, FRAC_1_SQRT_2, FRAC_1_SQRT_2))


def test_find_n1_and_n2(rt):  # world.rs:395-451
    a = glass_sphere(rt)
    a.set_transformation(rt.scaling(2, 2, 2))
    b = glass_sphere(rt)
    b.set_transformation(rt.translation(0, 0, -0.25))
    b.set_material(dataclasses.replace(b.material(), refractive_index=2.0))
    c = glass_sphere(rt)
    c.set_transformation(rt.translation(0, 0, 0.25))
    c.set_material(dataclasses.replace(c.material(), refractive_index=2.5))
    xs = [(2.0, a), (2.75, b), (3.25, c), (4.75, b), (5.25, c), (6.0, a)]
    expected = [(1.0, 1.5), (1.5, 2.0), (2.0, 2.5), (2.5, 2.5), (2.5, 1.5), (1.5, 1.0)]
    for i, (n1, n2) in enumerate(expected):
        comps = rt.probe.precompute((0, 0, -4), (0, 0, 1), xs, hit_index=i)
        assert (comps["n1"], comps["n2"]) == (n1, n2), f"intersection {i}"


def test_under_point_is_offset_below_surface(rt):  # world.rs:452-465
    s = glass_sphere(rt)
    s.set_transformation(rt.translation(0, 0, 1))
    c = rt.probe.precompute((0, 0, -5), (0, 0, 1), [(5.0, s)])
    assert c["under_point"][2] > ACNE / 2
    assert c["point"][2] < c["under_point"][2]


def world_with_reflective_plane(rt):
    w = rt.World.default()
    plane = rt.Plane.build(rt.translation(0, -1, 0), rt.Material(reflective=0.5))
    w.add_object(plane)
    return w, plane


def test_reflected_color_for_nonreflective_material(rt):  # world.rs:467-479
    w = rt.World.default()
    w.objects[1].set_material(dataclasses.replace(w.objects[1].material(), ambient=1.0))
    c = rt.probe.reflected_color(w, (0, 0, 0), (0, 0, 1), [(1.0, w.objects[1])], remaining=1)
    assert_eq(c, (0, 0, 0))


def test_reflected_color_for_reflective_material(rt):  # world.rs:481-494
    w, plane = world_with_reflective_plane(rt)
    c = rt.probe.reflected_color(w, (0, 0, -3), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, plane)], remaining=1)
    assert_abs_diff_eq(c, (0.19052197, 0.23815246, 0.14289148))


def test_shade_hit_with_reflective_material(rt):  # world.rs:496-508
    w, plane = world_with_reflective_plane(rt)
    c = rt.probe.shade_hit(w, (0, 0, -3), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, plane)], remaining=1)
    assert_abs_diff_eq(c, (0.8769108, 0.9245413, 0.8292803))


def test_mutually_reflective_surfaces_terminate(rt):  # world.rs:510-523
    m = rt.Material(reflective=1.0)
    lower = rt.Plane.build(rt.translation(0, -1, 0), m)
    upper = rt.Plane.build(rt.translation(0, 1, 0), m)
    w = rt.World([lower, upper], rt.PointLight((0, 0, 0), (0, 0, 0)))
    rt.probe.color_at(w, (0, 0, 0), (0, 1, 0), remaining=1)


def test_reflected_color_at_max_recursive_depth(rt):  # world.rs:525-537
    w, plane = world_with_reflective_plane(rt)
    c = rt.probe.reflected_color(w, (0, 0, -3), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, plane)], remaining=0)
    assert_abs_diff_eq(c, (0, 0, 0))


def test_shade_intersection(rt):  # world.rs:539-548
    w = rt.World.default()
    c = rt.probe.shade_hit(w, (0, 0, -5), (0, 0, 1), [(4.0, w.objects[0])], remaining=1)
    assert_abs_diff_eq(c, (0.38063288, 0.47579104, 0.28547466))


def test_shade_intersection_from_inside(rt):  # world.rs:550-560
    w = rt.World.default()
    w.set_light(rt.PointLight((0, 0.25, 0), (1, 1, 1)))
    c = rt.probe.shade_hit(w, (0, 0, 0), (0, 0, 1), [(0.5, w.objects[1])], remaining=1)
    assert_abs_diff_eq(c, (0.9045995, 0.9045995, 0.9045995))


def test_color_when_ray_misses(rt):  # world.rs:562-568
    assert_eq(rt.probe.color_at(rt.World.default(), (0, 0, -5), (0, 1, 0), 1), (0, 0, 0))


def test_color_when_ray_hits(rt):  # world.rs:570-576
    c = rt.probe.color_at(rt.World.default(), (0, 0, -5), (0, 0, 1), 1)
    assert_abs_diff_eq(c, (0.38063288, 0.47579104, 0.28547466))


def test_color_when_intersection_behind_ray(rt):  # world.rs:578-590
    w = rt.World.default()
    m = rt.Material(ambient=1.0)
    w.objects[0].set_material(m)
    w.objects[1].set_material(m)
    assert_eq(rt.probe.color_at(w, (0, 0, 0.75), (0, 0, -1), 1), m.color)


def test_is_shadowed(rt):  # world.rs:592-610
    w = rt.World.default()
    light = (-10, -10, -10)
    for p, expected in [((-10, -10, 10), False), ((10, 10, 10), True), ((-20, -20, -20), False), ((-5, -5, -5), False)]:
        assert rt.probe.is_shadowed(w, light, p) == expected, p


def test_point_light_intensity_at(rt):  # world.rs:612-630
    w = rt.World.default()
    cases = [((0, 1.0001, 0), 1.0), ((-1.0001, 0, 0), 1.0), ((0, 0, -1.0001), 1.0), ((0, 0, 1.0001), 0.0),
             ((1.0001, 0, 0), 0.0), ((0, -1.0001, 0), 0.0), ((0, 0, 0), 0.0)]
    for p, expected in cases:
        assert_abs_diff_eq(rt.probe.intensity_at(w, p), expected, msg=str(p))


def test_hit_should_offset_point(rt):  # world.rs:632-643
    s = rt.Sphere.build(rt.translation(0, 0, 1), rt.Material())
    c = rt.probe.precompute((0, 0, -5), (0, 0, 1), [(5.0, s)])
    assert c["over_point"][2] < -ACNE / 2
    assert c["over_point"][2] > -ACNE * 2
    assert c["point"][2] > c["over_point"][2]


def test_shade_hit_in_shadow(rt):  # world.rs:645-658
    s1 = rt.Sphere()
    s2 = rt.Sphere.build(rt.translation(0, 0, 10), rt.Material())
    w = rt.World([s1, s2], rt.PointLight((0, 0, -10), (1, 1, 1)))
    c = rt.probe.shade_hit(w, (0, 0, 5), (0, 0, 1), [(4.0, s2)], remaining=1)
    assert_eq(c, (0.1, 0.1, 0.1))


def test_refracted_color_of_opaque_surface(rt):  # world.rs:660-672
    w = rt.World.default()
    s = w.objects[0]
    c = rt.probe.refracted_color(w, (0, 0, -5), (0, 0, 1), [(4.0, s), (6.0, s)], 0, remaining=5)
    assert_abs_diff_eq(c, (0, 0, 0))


def glassy_default_world(rt):
    w = rt.World.default()
    w.objects[0].set_material(dataclasses.replace(w.objects[0].material(), transparency=1.0, refractive_index=1.5))
    return w


def test_refracted_color_at_max_depth(rt):  # world.rs:674-692
    w = glassy_default_world(rt)
    s = w.objects[0]
    c = rt.probe.refracted_color(w, (0, 0, -5), (0, 0, 1), [(4.0, s), (6.0, s)], 0, remaining=0)
    assert_abs_diff_eq(c, (0, 0, 0))


def test_refracted_color_under_total_internal_reflection(rt):  # world.rs:694-713
    w = glassy_default_world(rt)
    s = w.objects[0]
    xs = [(-FRAC_1_SQRT_2, s), (FRAC_1_SQRT_2, s)]
    c = rt.probe.refracted_color(w, (0, 0, FRAC_1_SQRT_2), (0, 1, 0), xs, 1, remaining=5)
    assert_abs_diff_eq(c, (0, 0, 0))


def test_refracted_color_with_refracted_ray(rt):  # world.rs:715-744
    w = rt.World.default()
    a, b = w.objects
    a.set_material(dataclasses.replace(a.material(), ambient=1.0, pattern=rt.TestPattern()))
    b.set_material(dataclasses.replace(b.material(), transparency=1.0, refractive_index=1.5))
    xs = [(-0.9899, a), (-0.4899, b), (0.4899, b), (0.9899, a)]
    c = rt.probe.refracted_color(w, (0, 0, 0.1), (0, 1, 0), xs, 2, remaining=5)
    assert_abs_diff_eq(c, (0, 0.9976768, 0.047521036))


def world_with_floor_and_ball(rt, **floor_kw):
    w = rt.World.default()
    floor = rt.Plane.build(rt.translation(0, -1, 0), rt.Material(**floor_kw))
    w.add_object(floor)
    ball = rt.Sphere.build(rt.translation(0, -3.5, -0.5), rt.Material(color=(1, 0, 0), ambient=0.5))
    w.add_object(ball)
    return w, floor


def test_shade_hit_with_transparent_material(rt):  # world.rs:746-777
    w, floor = world_with_floor_and_ball(rt, transparency=0.5, refractive_index=1.5)
    c = rt.probe.shade_hit(w, (0, 0, -3), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, floor)], remaining=5)
    assert_abs_diff_eq(c, (0.93638885, 0.68638885, 0.68638885))


def test_schlick_under_total_internal_reflection(rt):  # world.rs:779-790
    s = glass_sphere(rt)
    w = rt.World.default()
    xs = [(-FRAC_1_SQRT_2, s), (FRAC_1_SQRT_2, s)]
    assert rt.probe.schlick(w, (0, 0, FRAC_1_SQRT_2), (0, 1, 0), xs, 1) == 1.0


def test_schlick_perpendicular(rt):  # world.rs:791-802
    s = glass_sphere(rt)
    w = rt.World.default()
    assert_abs_diff_eq(rt.probe.schlick(w, (0, 0, 0), (0, 1, 0), [(-1.0, s), (1.0, s)], 1), 0.04)


def test_schlick_small_angle(rt):  # world.rs:804-812
    s = glass_sphere(rt)
    w = rt.World.default()
    assert_abs_diff_eq(rt.probe.schlick(w, (0, 0.99, -2.0), (0, 0, 1), [(1.8589, s)], 0), 0.48873067)


def test_shade_hit_with_reflective_transparent_material(rt):  # world.rs:814-844
    w, floor = world_with_floor_and_ball(rt, reflective=0.5, transparency=0.5, refractive_index=1.5)
    c = rt.probe.shade_hit(w, (0, 0, -3), (0, -FRAC_1_SQRT_2, FRAC_1_SQRT_2), [(SQRT_2, floor)], remaining=5)
    assert_abs_diff_eq(c, (0.93388665, 0.69640774, 0.6924002))
