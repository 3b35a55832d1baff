"""Python mirror of the reference's `lib` crate API, bound to a native library over include/rtc_scene.h.

The names follow the Rust reference (garfieldnate/ray_tracer_challenge, paths under lib/src/):
`World`, `Camera`, `Canvas`, `Sphere` / `Plane` / `Cube` / `Cylinder` / `Cone` / `Triangle` /
`SmoothTriangle` / `GroupShape` / `CSG`, `Material`, `Stripes` / `Gradient` / `Rings` / `Checkers` /
`Sine2D` / `TextureMap` / `CubicMap`, `PointLight` / `RectangleLight`, and the `transformations`
functions, so that scene code reads like the reference's `demos/src/bin/*.rs`.

This module contains NO rendering arithmetic: every matrix, bounding box, flattening step and pixel is
computed by the native library it is bound to (`load_api(path)`).  The product binding is created in
`ray_tracer_challenge_b200/__init__.py` from `librtc_host.so`; the test-suite binds the very same classes
to the CPU oracle to build identical scenes on both sides.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

F = C.c_float
FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)
U8P = C.POINTER(C.c_uint8)

SG_SPHERE, SG_PLANE, SG_CUBE, SG_CYLINDER, SG_CONE, SG_TRIANGLE, SG_SMOOTH_TRIANGLE, SG_GROUP, SG_CSG, SG_TEST_SHAPE = range(10)
SG_PAT_STRIPES, SG_PAT_GRADIENT, SG_PAT_RINGS, SG_PAT_CHECKERS, SG_PAT_SINE2D, SG_PAT_TEST = range(6)
SG_UV_CHECKERS, SG_UV_ALIGN_CHECK, SG_UV_IMAGE = 0, 1, 2
SG_MAP_SPHERICAL, SG_MAP_PLANAR, SG_MAP_CYLINDRICAL = 0, 1, 2


class SgStats(C.Structure):
    _fields_ = [
        ("primary_rays", C.c_uint64),
        ("secondary_rays", C.c_uint64),
        ("shadow_rays", C.c_uint64),
        ("shades", C.c_uint64),
        ("flops", C.c_double),
        ("ms", C.c_double),
        ("ms_total", C.c_double),
    ]

    @property
    def rays(self) -> int:
        return self.primary_rays + self.secondary_rays + self.shadow_rays

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_} | {"rays": self.rays}


_SIGNATURES = {
    "sg_last_error": (C.c_char_p, []),
    "sg_translation": (None, [F, F, F, FP]),
    "sg_scaling": (None, [F, F, F, FP]),
    "sg_rotation_x": (None, [F, FP]),
    "sg_rotation_y": (None, [F, FP]),
    "sg_rotation_z": (None, [F, FP]),
    "sg_shearing": (None, [F, F, F, F, F, F, FP]),
    "sg_view_transform": (None, [FP, FP, FP, FP]),
    "sg_matmul": (None, [FP, FP, FP]),
    "sg_inverse": (None, [FP, FP]),
    "sg_transpose": (None, [FP, FP]),
    "sg_determinant": (F, [FP]),
    "sg_create": (C.c_void_p, []),
    "sg_destroy": (None, [C.c_void_p]),
    "sg_pattern_new": (C.c_int, [C.c_void_p, C.c_int, FP, FP]),
    "sg_pattern_set_transform": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_uv_pattern_new": (C.c_int, [C.c_void_p, C.c_int, FP, C.c_int]),
    "sg_texture_map_new": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_canvas_new": (C.c_int, [C.c_void_p, C.c_int, C.c_int, FP]),
    "sg_canvas_from_ppm": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sg_canvas_size": (C.c_int, [C.c_void_p, C.c_int, IP, IP]),
    "sg_canvas_pixels": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_canvas_to_ppm": (C.c_int64, [C.c_void_p, C.c_int, C.c_char_p, C.c_int64]),
    "sg_uv_image_new": (C.c_int, [C.c_void_p, C.c_int]),
    "sg_cubic_map_new": (C.c_int, [C.c_void_p, IP]),
    "sg_material_new": (C.c_int, [C.c_void_p, FP, C.c_int]),
    "sg_shape_new": (C.c_int, [C.c_void_p, C.c_int]),
    "sg_triangle_new": (C.c_int, [C.c_void_p, FP, FP, FP]),
    "sg_smooth_triangle_new": (C.c_int, [C.c_void_p, FP, FP]),
    "sg_csg_new": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "sg_shape_clone": (C.c_int, [C.c_void_p, C.c_int]),
    "sg_shape_set_transform": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_shape_set_material": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_shape_set_casts_shadow": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_shape_set_bounds": (C.c_int, [C.c_void_p, C.c_int, F, F, C.c_int]),
    "sg_group_add_child": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_shape_divide": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_parse_obj": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sg_shape_kind": (C.c_int, [C.c_void_p, C.c_int]),
    "sg_shape_get_transform": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_shape_get_inverse": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_shape_get_inverse_transpose": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_shape_bounding_box": (C.c_int, [C.c_void_p, C.c_int, FP, FP]),
    "sg_shape_parent_space_bounding_box": (C.c_int, [C.c_void_p, C.c_int, FP, FP]),
    "sg_group_child_count": (C.c_int, [C.c_void_p, C.c_int]),
    "sg_group_child": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_triangle_get": (C.c_int, [C.c_void_p, C.c_int, FP]),
    "sg_world_new": (C.c_int, [C.c_void_p]),
    "sg_world_add_object": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "sg_world_set_point_light": (C.c_int, [C.c_void_p, C.c_int, FP, FP]),
    "sg_world_set_rect_light": (C.c_int, [C.c_void_p, C.c_int, FP, FP, FP, C.c_int, FP, C.c_int, FP, C.c_int, C.c_uint64]),
    "sg_camera_new": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, F, FP]),
    "sg_camera_render": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, FP, U8P, C.POINTER(SgStats)]),
}


def f32(values) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(values, dtype=np.float32))


def fptr(a: np.ndarray):
    return a.ctypes.data_as(FP)


class RtcError(RuntimeError):
    pass


@dataclass
class Material:
    """material.rs:19-51 (defaults are the builder's)."""

    color: Sequence[float] = (1.0, 1.0, 1.0)
    ambient: float = 0.1
    diffuse: float = 0.9
    specular: float = 0.9
    shininess: float = 200.0
    reflective: float = 0.0
    transparency: float = 0.0
    refractive_index: float = 1.0
    pattern: Optional["Pattern"] = None

    def params(self) -> np.ndarray:
        return f32([*self.color, self.ambient, self.diffuse, self.specular, self.shininess, self.reflective,
                    self.transparency, self.refractive_index])


# constants.rs:6-62
REFRACTION_VACCUM = 1.0
REFRACTION_AIR = 1.00029
REFRACTION_WATER = 1.333
REFRACTION_GLASS = 1.52
REFRACTION_DIAMOND = 2.417
DEFAULT_RAY_RECURSION_DEPTH = 5


def glass() -> Material:
    return Material(transparency=1.0, refractive_index=REFRACTION_GLASS)


def metal() -> Material:
    return Material(color=(0.5, 0.5, 0.5), ambient=1.0, diffuse=0.6, reflective=0.1, specular=0.4, shininess=10.0,
                    transparency=0.0, refractive_index=1.0)


def color_from_hex(s: str):
    """color.rs:93-108 — '#RRGGBB' -> components / 255 in f32."""
    s = s.lstrip("#")
    return tuple(float(np.float32(int(s[i:i + 2], 16)) / np.float32(255.0)) for i in (0, 2, 4))


def constant_jitter():
    """test/utils.rs:15-17"""
    return [0.5]


def hardcoded_jitter(sequence):
    """test/utils.rs:19-24"""
    return list(sequence)


@dataclass
class PointLight:
    position: Sequence[float]
    intensity: Sequence[float]


@dataclass
class RectangleLight:
    """rectangle_light.rs:34-59. `jitter`: cyclic table, or None for the counter-based generator (`seed`)."""

    intensity: Sequence[float]
    corner: Sequence[float]
    u_vec: Sequence[float]
    u_steps: int
    v_vec: Sequence[float]
    v_steps: int
    jitter: Optional[Sequence[float]] = None
    seed: int = 0


class Canvas:
    """canvas.rs:6-43 — f32 RGB framebuffer, row-major [y][x]."""

    def __init__(self, width: int, height: int, data: Optional[np.ndarray] = None, u8: Optional[np.ndarray] = None, api=None):
        self.width, self.height = width, height
        self.data = data if data is not None else np.zeros((height, width, 3), np.float32)
        self._u8 = u8
        self.api = api  # the native library that serialises this canvas (to_ppm) and binds it to UVImage

    def write_pixel(self, x: int, y: int, color):
        self.data[y, x] = color
        self._u8 = None

    def pixel_at(self, x: int, y: int):
        return tuple(float(v) for v in self.data[y, x])

    def to_u8(self) -> np.ndarray:
        """scale_color (canvas.rs:39-43) of every channel, computed by the library that rendered the frame."""
        if self._u8 is None:
            raise RtcError("canvas has no 8-bit plane")
        return self._u8

    def to_ppm(self) -> str:
        """canvas.rs:58-96 (P3, 70-column wrap): the native writer of the library this canvas came from, or, for a
        canvas with no library behind it, the pure-Python spelling below."""
        if self.api is not None:
            return self.api.canvas_to_ppm(self)
        u8 = self.to_u8()
        out = [f"P3\n{self.width} {self.height}\n255\n"]
        for row in u8.reshape(self.height, self.width * 3):
            line = ""
            n = len(row)
            for i, v in enumerate(row):
                line += str(int(v))
                if i != n - 1:
                    if len(line) < 70 - 3:
                        line += " "
                    else:
                        out.append(line + "\n")
                        line = ""
            if line:
                out.append(line + "\n")
        return "".join(out)


class Api:
    """One native library + one scene context.  Attribute access gives classes bound to this context."""

    def __init__(self, lib_path: str, extra_signatures: Optional[dict] = None):
        self.lib_path = lib_path
        self.lib = C.CDLL(lib_path, mode=C.RTLD_LOCAL)
        sigs = dict(_SIGNATURES)
        if extra_signatures:
            sigs.update(extra_signatures)
        for name, (res, args) in sigs.items():
            fn = getattr(self.lib, name)
            fn.restype = res
            fn.argtypes = args
        self.ctx = self.lib.sg_create()
        if not self.ctx:
            raise RtcError("sg_create failed")
        api = self

        class Matrix:
            """matrix.rs — 4x4 f32, row-major."""

            def __init__(self, data=None):
                self.m = f32(np.eye(4) if data is None else data).reshape(4, 4)

            def __mul__(self, other):
                if isinstance(other, Matrix):
                    out = np.empty(16, np.float32)
                    api.lib.sg_matmul(fptr(self.m), fptr(other.m), fptr(out))
                    return Matrix(out)
                return NotImplemented

            def inverse(self):
                out = np.empty(16, np.float32)
                api.lib.sg_inverse(fptr(self.m), fptr(out))
                return Matrix(out)

            def transpose(self):
                out = np.empty(16, np.float32)
                api.lib.sg_transpose(fptr(self.m), fptr(out))
                return Matrix(out)

            def determinant(self) -> float:
                return float(api.lib.sg_determinant(fptr(self.m)))

            def __repr__(self):
                return f"Matrix({self.m.tolist()})"

        self.Matrix = Matrix

        def _mk(fn, *args):
            out = np.empty(16, np.float32)
            fn(*args, fptr(out))
            return Matrix(out)

        self.identity_4x4 = lambda: Matrix()
        self.translation = lambda x, y, z: _mk(api.lib.sg_translation, x, y, z)
        self.scaling = lambda x, y, z: _mk(api.lib.sg_scaling, x, y, z)
        self.rotation_x = lambda r: _mk(api.lib.sg_rotation_x, r)
        self.rotation_y = lambda r: _mk(api.lib.sg_rotation_y, r)
        self.rotation_z = lambda r: _mk(api.lib.sg_rotation_z, r)
        self.shearing = lambda a, b, c, d, e, f: _mk(api.lib.sg_shearing, a, b, c, d, e, f)

        def view_transform(frm, to, up):
            a, b, c = f32(frm), f32(to), f32(up)
            return _mk(api.lib.sg_view_transform, fptr(a), fptr(b), fptr(c))

        self.view_transform = view_transform

        # ---------------------------------------------------------------- patterns
        class Pattern:
            _kind = -1

            def __init__(self, a=(1.0, 1.0, 1.0), b=(0.0, 0.0, 0.0)):
                self.a, self.b = tuple(a), tuple(b)
                ca, cb = f32(self.a), f32(self.b)
                self.handle = api.check(api.lib.sg_pattern_new(api.ctx, self._kind, fptr(ca), fptr(cb)))

            def set_transformation(self, t):
                api.check(api.lib.sg_pattern_set_transform(api.ctx, self.handle, fptr(t.m)))
                self.transform = t
                return self

        def pat(name, kind):
            return type(name, (Pattern,), {"_kind": kind})

        self.Pattern = Pattern
        self.Stripes = pat("Stripes", SG_PAT_STRIPES)
        self.Gradient = pat("Gradient", SG_PAT_GRADIENT)
        self.Rings = pat("Rings", SG_PAT_RINGS)
        self.Checkers = pat("Checkers", SG_PAT_CHECKERS)
        self.Sine2D = pat("Sine2D", SG_PAT_SINE2D)
        self.TestPattern = pat("TestPattern", SG_PAT_TEST)

        class UVCheckers:
            def __init__(self, width, height, a, b):
                p = f32([width, height, *a, *b])
                self.handle = api.check(api.lib.sg_uv_pattern_new(api.ctx, SG_UV_CHECKERS, fptr(p), 8))

        class AlignCheck:
            def __init__(self, main, ul, ur, bl, br):
                p = f32([*main, *ul, *ur, *bl, *br])
                self.handle = api.check(api.lib.sg_uv_pattern_new(api.ctx, SG_UV_ALIGN_CHECK, fptr(p), 15))

        class TextureMap(Pattern):
            def __init__(self, uv_pattern, mapping: int):
                self.handle = api.check(api.lib.sg_texture_map_new(api.ctx, uv_pattern.handle, mapping))

        class CubicMap(Pattern):
            def __init__(self, front, back, left, right, up, down):
                ids = (C.c_int * 6)(*[p.handle for p in (front, back, left, right, up, down)])
                self.handle = api.check(api.lib.sg_cubic_map_new(api.ctx, ids))

        class UVImage:
            """uv.rs:346-377 — `canvas` is an api.Canvas (its f32 pixels are handed to the library)."""

            def __init__(self, canvas):
                self.canvas_handle = api.native_canvas(canvas)
                self.handle = api.check(api.lib.sg_uv_image_new(api.ctx, self.canvas_handle))

        self.UVCheckers, self.AlignCheck, self.TextureMap, self.CubicMap = UVCheckers, AlignCheck, TextureMap, CubicMap
        self.UVImage = UVImage

        # ---------------------------------------------------------------- shapes
        class Shape:
            _kind = -1

            def __init__(self, handle=None):
                self.handle = api.check(api.lib.sg_shape_new(api.ctx, self._kind)) if handle is None else handle
                self._material = Material()

            @classmethod
            def new(cls):
                return cls()

            @classmethod
            def build(cls, transform, material):
                s = cls()
                s.set_transformation(transform)
                s.set_material(material)
                return s

            def set_transformation(self, t):
                api.check(api.lib.sg_shape_set_transform(api.ctx, self.handle, fptr(t.m)))
                return self

            def set_material(self, m: Material):
                self._material = m
                api.check(api.lib.sg_shape_set_material(api.ctx, self.handle, api.material_handle(m)))
                return self

            def material(self) -> Material:
                return self._material

            def set_casts_shadow(self, v: bool):
                api.check(api.lib.sg_shape_set_casts_shadow(api.ctx, self.handle, int(bool(v))))
                return self

            def divide(self, threshold: int):
                api.check(api.lib.sg_shape_divide(api.ctx, self.handle, int(threshold)))
                return self

            def clone(self):
                h = api.check(api.lib.sg_shape_clone(api.ctx, self.handle))
                c = api.wrap_shape(h)
                c._material = self._material
                return c

            def _mat(self, fn):
                out = np.empty(16, np.float32)
                api.check(fn(api.ctx, self.handle, fptr(out)))
                return Matrix(out)

            def transformation(self):
                return self._mat(api.lib.sg_shape_get_transform)

            def transformation_inverse(self):
                return self._mat(api.lib.sg_shape_get_inverse)

            def transformation_inverse_transpose(self):
                return self._mat(api.lib.sg_shape_get_inverse_transpose)

            def _bbox(self, fn):
                mn, mx = np.empty(3, np.float32), np.empty(3, np.float32)
                api.check(fn(api.ctx, self.handle, fptr(mn), fptr(mx)))
                return mn, mx

            def bounding_box(self):
                return self._bbox(api.lib.sg_shape_bounding_box)

            def parent_space_bounding_box(self):
                return self._bbox(api.lib.sg_shape_parent_space_bounding_box)

            def kind(self) -> int:
                return api.check(api.lib.sg_shape_kind(api.ctx, self.handle))

        class _Bounded(Shape):
            def __init__(self, handle=None):
                super().__init__(handle)
                self._lo, self._hi, self._closed = -math.inf, math.inf, False

            def _push(self):
                api.check(api.lib.sg_shape_set_bounds(api.ctx, self.handle, self._lo, self._hi, int(self._closed)))

            minimum_y = property(lambda s: s._lo, lambda s, v: (setattr(s, "_lo", float(v)), s._push())[0])
            maximum_y = property(lambda s: s._hi, lambda s, v: (setattr(s, "_hi", float(v)), s._push())[0])
            closed = property(lambda s: s._closed, lambda s, v: (setattr(s, "_closed", bool(v)), s._push())[0])

        class Triangle(Shape):
            def __init__(self, p1=None, p2=None, p3=None, handle=None):
                if handle is None:
                    a, b, c = f32(p1), f32(p2), f32(p3)
                    handle = api.check(api.lib.sg_triangle_new(api.ctx, fptr(a), fptr(b), fptr(c)))
                Shape.__init__(self, handle)

            def geometry(self):
                out = np.empty(12, np.float32)
                api.check(api.lib.sg_triangle_get(api.ctx, self.handle, fptr(out)))
                return out.reshape(4, 3)  # p1, e1, e2, normal

        class SmoothTriangle(Triangle):
            def __init__(self, p1=None, p2=None, p3=None, n1=None, n2=None, n3=None, handle=None):
                if handle is None:
                    p = f32([*p1, *p2, *p3])
                    n = f32([*n1, *n2, *n3])
                    handle = api.check(api.lib.sg_smooth_triangle_new(api.ctx, fptr(p), fptr(n)))
                Shape.__init__(self, handle)

        class GroupShape(Shape):
            _kind = SG_GROUP

            def add_child(self, child):
                api.check(api.lib.sg_group_add_child(api.ctx, self.handle, child.handle))
                return self

            def get_children(self):
                n = api.check(api.lib.sg_group_child_count(api.ctx, self.handle))
                return [api.wrap_shape(api.check(api.lib.sg_group_child(api.ctx, self.handle, i))) for i in range(n)]

            def set_material(self, m: Material):
                # group.rs:96-100 — pushed to the children by the library
                api.check(api.lib.sg_shape_set_material(api.ctx, self.handle, api.material_handle(m)))
                return self

        class CSG(GroupShape):
            _kind = SG_CSG

            def __init__(self, op=None, s1=None, s2=None, handle=None):
                if handle is None:
                    handle = api.check(api.lib.sg_csg_new(api.ctx, int(op), s1.handle, s2.handle))
                Shape.__init__(self, handle)

        self.Shape = Shape
        self.Sphere = type("Sphere", (Shape,), {"_kind": SG_SPHERE})
        self.Plane = type("Plane", (Shape,), {"_kind": SG_PLANE})
        self.Cube = type("Cube", (Shape,), {"_kind": SG_CUBE})
        self.Cylinder = type("Cylinder", (_Bounded,), {"_kind": SG_CYLINDER})
        self.Cone = type("Cone", (_Bounded,), {"_kind": SG_CONE})
        self.TestShape = type("TestShape", (Shape,), {"_kind": SG_TEST_SHAPE})
        self.Triangle, self.SmoothTriangle, self.GroupShape, self.CSG = Triangle, SmoothTriangle, GroupShape, CSG
        self._by_kind = {
            SG_SPHERE: self.Sphere, SG_PLANE: self.Plane, SG_CUBE: self.Cube, SG_CYLINDER: self.Cylinder,
            SG_CONE: self.Cone, SG_TRIANGLE: Triangle, SG_SMOOTH_TRIANGLE: SmoothTriangle, SG_GROUP: GroupShape,
            SG_CSG: CSG, SG_TEST_SHAPE: self.TestShape,
        }

        # ---------------------------------------------------------------- world / camera
        class World:
            """world.rs:18-21"""

            def __init__(self, objects=(), light=None):
                self.handle = api.check(api.lib.sg_world_new(api.ctx))
                self.objects = []
                self.light = None
                for o in objects:
                    self.add_object(o)
                if light is not None:
                    self.set_light(light)

            def add_object(self, shape):
                api.check(api.lib.sg_world_add_object(api.ctx, self.handle, shape.handle))
                self.objects.append(shape)
                return self

            def set_light(self, light):
                self.light = light
                i = f32(light.intensity)
                if isinstance(light, PointLight):
                    p = f32(light.position)
                    api.check(api.lib.sg_world_set_point_light(api.ctx, self.handle, fptr(p), fptr(i)))
                else:
                    c, u, v = f32(light.corner), f32(light.u_vec), f32(light.v_vec)
                    j = f32(light.jitter if light.jitter is not None else [])
                    api.check(api.lib.sg_world_set_rect_light(
                        api.ctx, self.handle, fptr(i), fptr(c), fptr(u), int(light.u_steps), fptr(v),
                        int(light.v_steps), fptr(j), int(j.size), int(light.seed)))
                return self

            @classmethod
            def default(cls):
                """world.rs:31-49 — the book's two-sphere world."""
                s1 = api.Sphere.build(api.identity_4x4(), Material(color=(0.8, 1.0, 0.6), diffuse=0.7, specular=0.2))
                s2 = api.Sphere.build(api.scaling(0.5, 0.5, 0.5), Material())
                return cls([s1, s2], PointLight((-10.0, 10.0, -10.0), (1.0, 1.0, 1.0)))

        class Camera:
            """camera.rs:8-91"""

            def __init__(self, width_pixels: int, height_pixels: int, field_of_view: float, transform):
                self.width_pixels, self.height_pixels = int(width_pixels), int(height_pixels)
                self.field_of_view = field_of_view
                self.handle = api.check(api.lib.sg_camera_new(api.ctx, self.width_pixels, self.height_pixels,
                                                              field_of_view, fptr(transform.m)))
                self.last_stats: Optional[SgStats] = None

            def render(self, world, reflection_recursion_depth: int = DEFAULT_RAY_RECURSION_DEPTH, want_u8=True):
                w, h = self.width_pixels, self.height_pixels
                rgb = np.zeros((h, w, 3), np.float32)
                u8 = np.zeros((h, w, 3), np.uint8) if want_u8 else None
                stats = SgStats()
                api.check(api.lib.sg_camera_render(
                    api.ctx, self.handle, world.handle, int(reflection_recursion_depth), fptr(rgb),
                    u8.ctypes.data_as(U8P) if u8 is not None else None, C.byref(stats)))
                self.last_stats = stats
                return Canvas(w, h, rgb, u8, api=api)

        self.World, self.Camera = World, Camera
        self.Material, self.PointLight, self.RectangleLight, self.Canvas = Material, PointLight, RectangleLight, Canvas

    # ------------------------------------------------------------------ helpers
    def check(self, rc: int) -> int:
        if rc < 0:
            raise RtcError(self.lib.sg_last_error().decode())
        return rc

    # ---- canvases as data: canvas.rs:58-200
    def native_canvas(self, canvas) -> int:
        """Hand an api.Canvas's f32 pixels to the library (sg_canvas_new); returns the native handle."""
        data = np.ascontiguousarray(canvas.data, dtype=np.float32)
        return self.check(self.lib.sg_canvas_new(self.ctx, int(canvas.width), int(canvas.height), fptr(data)))

    def canvas_from_ppm(self, text) -> "Canvas":
        """canvas_from_ppm (canvas.rs:119-182).  Raises RtcError carrying the reference's ParseError kind."""
        raw = text.encode() if isinstance(text, str) else bytes(text)
        h = self.check(self.lib.sg_canvas_from_ppm(self.ctx, raw, len(raw)))
        w, ht = C.c_int(), C.c_int()
        self.check(self.lib.sg_canvas_size(self.ctx, h, C.byref(w), C.byref(ht)))
        data = np.zeros((ht.value, w.value, 3), np.float32)
        if data.size:
            self.check(self.lib.sg_canvas_pixels(self.ctx, h, fptr(data)))
        return Canvas(w.value, ht.value, data, api=self)

    def canvas_to_ppm(self, canvas) -> str:
        """Canvas::to_ppm (canvas.rs:58-96) by the library's native writer."""
        h = self.native_canvas(canvas)
        n = self.check(self.lib.sg_canvas_to_ppm(self.ctx, h, None, 0))
        buf = C.create_string_buffer(int(n))
        self.check(self.lib.sg_canvas_to_ppm(self.ctx, h, buf, n))
        return buf.raw[:n].decode("ascii")

    def new_canvas(self, width: int, height: int) -> "Canvas":
        """Canvas::new (canvas.rs:19-25): black."""
        return Canvas(width, height, api=self)

    def material_handle(self, m: Material) -> int:
        p = m.params()
        ph = m.pattern.handle if m.pattern is not None else -1
        key = (p.tobytes(), ph)
        cache = self.__dict__.setdefault("_material_cache", {})
        h = cache.get(key)
        if h is None:
            h = cache[key] = self.check(self.lib.sg_material_new(self.ctx, fptr(p), ph))
        return h

    def wrap_shape(self, handle: int):
        kind = self.check(self.lib.sg_shape_kind(self.ctx, handle))
        cls = self._by_kind[kind]
        if kind in (SG_TRIANGLE, SG_SMOOTH_TRIANGLE, SG_CSG):
            return cls(handle=handle)
        return cls(handle)

    def parse_obj(self, text: str):
        """obj_parser.rs:100-216 + take_all_as_group (33-55)."""
        data = text.encode()
        return self.wrap_shape(self.check(self.lib.sg_parse_obj(self.ctx, data, len(data))))

    def close(self):
        if self.ctx:
            self.lib.sg_destroy(self.ctx)
            self.ctx = None


def load_api(lib_path: str, extra_signatures: Optional[dict] = None) -> Api:
    return Api(lib_path, extra_signatures)
