"""ctypes binding of the CPU oracle for the test-suite (test infrastructure only).

`load_oracle()` returns the same reference-shaped API object as the product (`ray_tracer_challenge_b200`),
bound to oracle/librtc_oracle.so, plus the `orc_*` probes that expose the sub-functions the reference's
unit tests pin.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ray_tracer_challenge_b200.api import FP, IP, U8P, F, SgStats, f32, fptr, load_api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "librtc_oracle.so")

V = C.c_void_p
I = C.c_int
_PROBES = {
    "orc_set_threads": (None, [V, I]),
    "orc_max_threads": (I, []),
    "orc_camera_render_rows": (I, [V, I, I, I, C.c_uint32, C.c_uint32, C.c_uint32, FP, U8P, C.POINTER(SgStats)]),
    "orc_shape_intersect": (I, [V, I, FP, FP, I, FP, IP, FP, I]),
    "orc_shape_normal_at": (I, [V, I, FP, I, F, F, FP]),
    "orc_shape_world_to_object": (I, [V, I, FP, FP]),
    "orc_shape_normal_to_world": (I, [V, I, FP, FP]),
    "orc_shape_includes": (I, [V, I, I]),
    "orc_test_shape_saved_ray": (I, [V, I, FP]),
    "orc_world_intersect": (I, [V, I, FP, FP, FP, IP, I]),
    "orc_hit_index": (I, [FP, I]),
    "orc_color_at": (I, [V, I, FP, FP, I, FP]),
    "orc_precompute": (I, [V, FP, FP, I, FP, IP, FP, I, FP]),
    "orc_comps_eval": (I, [V, I, FP, FP, I, FP, IP, I, I, I, FP]),
    "orc_is_shadowed": (I, [V, I, FP, FP]),
    "orc_intensity_at": (F, [V, I, FP]),
    "orc_light_info": (I, [V, I, FP]),
    "orc_point_on_light": (I, [V, I, I, I, I, FP]),
    "orc_phong": (I, [V, I, I, FP, FP, FP, FP, FP, F, FP]),
    "orc_pattern_color_at": (I, [V, I, I, FP, FP]),
    "orc_uv_pattern_color_at": (I, [V, I, F, F, FP]),
    "orc_uv_map": (None, [I, FP, FP]),
    "orc_face_from_point": (I, [FP]),
    "orc_cube_uv": (None, [I, FP, FP]),
    "orc_camera_ray": (I, [V, I, C.c_uint32, C.c_uint32, FP]),
    "orc_camera_info": (I, [V, I, FP]),
    "orc_scale_color": (I, [F]),
    "orc_bbox_transform": (None, [FP, FP, FP, FP, FP]),
    "orc_bbox_intersects": (I, [FP, FP, FP, FP]),
    "orc_bbox_split": (None, [FP, FP, FP]),
    "orc_bbox_contains_box": (I, [FP, FP, FP, FP]),
    "orc_bbox_contains_point": (I, [FP, FP, FP]),
    "orc_reflect": (None, [FP, FP, FP]),
    "orc_mat_mul_tuple": (None, [FP, FP, FP]),
    "orc_csg_allowed": (I, [I, I, I, I]),
    "orc_csg_filter": (I, [V, I, I, FP, IP, FP, I]),
}


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def _ints(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


class Probes:
    """Thin wrappers over orc_* returning numpy / python values."""

    def __init__(self, api):
        self.api, self.lib, self.ctx = api, api.lib, api.ctx

    def _ck(self, rc):
        return self.api.check(rc)

    def intersect(self, shape, origin, direction, local=False, cap=16):
        o, d = f32(origin), f32(direction)
        ts, objs, uvs = np.zeros(cap, np.float32), np.zeros(cap, np.int32), np.zeros(2 * cap, np.float32)
        n = self._ck(self.lib.orc_shape_intersect(self.ctx, shape.handle, fptr(o), fptr(d), int(local), fptr(ts),
                                                  objs.ctypes.data_as(IP), fptr(uvs), cap))
        return ts[:n].copy(), objs[:n].copy(), uvs[:2 * n].reshape(-1, 2).copy()

    def local_intersect(self, shape, origin, direction):
        return self.intersect(shape, origin, direction, local=True)

    def normal_at(self, shape, p, local=False, u=0.0, v=0.0):
        pp, out = f32(p), np.zeros(3, np.float32)
        self._ck(self.lib.orc_shape_normal_at(self.ctx, shape.handle, fptr(pp), int(local), u, v, fptr(out)))
        return out

    def world_to_object(self, shape, p):
        pp, out = f32(p), np.zeros(3, np.float32)
        self._ck(self.lib.orc_shape_world_to_object(self.ctx, shape.handle, fptr(pp), fptr(out)))
        return out

    def normal_to_world(self, shape, n):
        nn, out = f32(n), np.zeros(3, np.float32)
        self._ck(self.lib.orc_shape_normal_to_world(self.ctx, shape.handle, fptr(nn), fptr(out)))
        return out

    def saved_ray(self, test_shape):
        out = np.zeros(6, np.float32)
        self._ck(self.lib.orc_test_shape_saved_ray(self.ctx, test_shape.handle, fptr(out)))
        return out[:3], out[3:]

    def includes(self, a, b):
        return bool(self._ck(self.lib.orc_shape_includes(self.ctx, a.handle, b.handle)))

    def world_intersect(self, world, origin, direction, cap=64):
        o, d = f32(origin), f32(direction)
        ts, objs = np.zeros(cap, np.float32), np.zeros(cap, np.int32)
        n = self._ck(self.lib.orc_world_intersect(self.ctx, world.handle, fptr(o), fptr(d), fptr(ts),
                                                  objs.ctypes.data_as(IP), cap))
        return ts[:n].copy(), objs[:n].copy()

    def hit_index(self, ts):
        t = f32(ts)
        return self.lib.orc_hit_index(fptr(t), int(t.size))

    def color_at(self, world, origin, direction, remaining=5):
        o, d, out = f32(origin), f32(direction), np.zeros(3, np.float32)
        self._ck(self.lib.orc_color_at(self.ctx, world.handle, fptr(o), fptr(d), remaining, fptr(out)))
        return out

    def precompute(self, origin, direction, xs, hit_index=0, uvs=None):
        """xs: list of (t, shape).  Returns a dict of the PrecomputedValues (world.rs:165-182)."""
        o, d = f32(origin), f32(direction)
        ts, hs = f32([t for t, _ in xs]), _ints([s.handle for _, s in xs])
        uv = f32(uvs).reshape(-1) if uvs is not None else None
        out = np.zeros(24, np.float32)
        self._ck(self.lib.orc_precompute(self.ctx, fptr(o), fptr(d), len(xs), fptr(ts), hs.ctypes.data_as(IP),
                                         fptr(uv) if uv is not None else None, hit_index, fptr(out)))
        return dict(point=out[0:3], eye_vector=out[3:6], surface_normal=out[6:9], reflection_vector=out[9:12],
                    over_point=out[12:15], under_point=out[15:18], inside=bool(out[18]), n1=float(out[19]),
                    n2=float(out[20]), distance=float(out[21]))

    def _eval(self, what, world, origin, direction, xs, hit_index, remaining):
        o, d = f32(origin), f32(direction)
        ts, hs = f32([t for t, _ in xs]), _ints([s.handle for _, s in xs])
        out = np.zeros(3, np.float32)
        self._ck(self.lib.orc_comps_eval(self.ctx, world.handle, fptr(o), fptr(d), len(xs), fptr(ts),
                                         hs.ctypes.data_as(IP), hit_index, remaining, what, fptr(out)))
        return out

    def shade_hit(self, world, origin, direction, xs, hit_index=0, remaining=5):
        return self._eval(0, world, origin, direction, xs, hit_index, remaining)

    def reflected_color(self, world, origin, direction, xs, hit_index=0, remaining=5):
        return self._eval(1, world, origin, direction, xs, hit_index, remaining)

    def refracted_color(self, world, origin, direction, xs, hit_index=0, remaining=5):
        return self._eval(2, world, origin, direction, xs, hit_index, remaining)

    def schlick(self, world, origin, direction, xs, hit_index=0):
        return float(self._eval(3, world, origin, direction, xs, hit_index, 0)[0])

    def is_shadowed(self, world, light_position, p):
        lp, pp = f32(light_position), f32(p)
        return bool(self._ck(self.lib.orc_is_shadowed(self.ctx, world.handle, fptr(lp), fptr(pp))))

    def intensity_at(self, world, p):
        pp = f32(p)
        return float(self.lib.orc_intensity_at(self.ctx, world.handle, fptr(pp)))

    def light_info(self, world):
        out = np.zeros(10, np.float32)
        self._ck(self.lib.orc_light_info(self.ctx, world.handle, fptr(out)))
        return dict(u_vec=out[0:3], v_vec=out[3:6], position=out[6:9], cells=int(out[9]))

    def point_on_light(self, world, u, v, cursor=0):
        out = np.zeros(3, np.float32)
        self._ck(self.lib.orc_point_on_light(self.ctx, world.handle, u, v, cursor, fptr(out)))
        return out

    def phong(self, shape, material, light, p, eye, normal, light_intensity):
        mh = self.api.material_handle(material) if material is not None else -1
        lp, li, pp, e, n = f32(light.position), f32(light.intensity), f32(p), f32(eye), f32(normal)
        out = np.zeros(3, np.float32)
        self._ck(self.lib.orc_phong(self.ctx, shape.handle, mh, fptr(lp), fptr(li), fptr(pp), fptr(e), fptr(n),
                                    light_intensity, fptr(out)))
        return out

    def pattern_color_at(self, pattern, p, shape=None):
        pp, out = f32(p), np.zeros(3, np.float32)
        self._ck(self.lib.orc_pattern_color_at(self.ctx, pattern.handle, shape.handle if shape is not None else -1,
                                               fptr(pp), fptr(out)))
        return out

    def uv_color_at(self, uv, u, v):
        out = np.zeros(3, np.float32)
        self._ck(self.lib.orc_uv_pattern_color_at(self.ctx, uv.handle, u, v, fptr(out)))
        return out

    def uv_map(self, mapping, p):
        pp, out = f32(p), np.zeros(2, np.float32)
        self.lib.orc_uv_map(mapping, fptr(pp), fptr(out))
        return out

    def face_from_point(self, p):
        pp = f32(p)
        return self.lib.orc_face_from_point(fptr(pp))

    def cube_uv(self, face, p):
        pp, out = f32(p), np.zeros(2, np.float32)
        self.lib.orc_cube_uv(face, fptr(pp), fptr(out))
        return out

    def camera_ray(self, camera, x, y):
        out = np.zeros(6, np.float32)
        self._ck(self.lib.orc_camera_ray(self.ctx, camera.handle, x, y, fptr(out)))
        return out[:3], out[3:]

    def camera_info(self, camera):
        out = np.zeros(3, np.float32)
        self._ck(self.lib.orc_camera_info(self.ctx, camera.handle, fptr(out)))
        return dict(pixel_size=float(out[0]), half_width=float(out[1]), half_height=float(out[2]))

    def scale_color(self, v):
        return self.lib.orc_scale_color(v)

    def bbox_transform(self, mn, mx, m):
        a, b = f32(mn), f32(mx)
        omn, omx = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.lib.orc_bbox_transform(fptr(a), fptr(b), fptr(m.m), fptr(omn), fptr(omx))
        return omn, omx

    def bbox_intersects(self, mn, mx, origin, direction):
        a, b, o, d = f32(mn), f32(mx), f32(origin), f32(direction)
        return bool(self.lib.orc_bbox_intersects(fptr(a), fptr(b), fptr(o), fptr(d)))

    def bbox_split(self, mn, mx):
        a, b, out = f32(mn), f32(mx), np.zeros(12, np.float32)
        self.lib.orc_bbox_split(fptr(a), fptr(b), fptr(out))
        return out[0:3], out[3:6], out[6:9], out[9:12]

    def bbox_contains_box(self, mn, mx, omn, omx):
        a, b, c, d = f32(mn), f32(mx), f32(omn), f32(omx)
        return bool(self.lib.orc_bbox_contains_box(fptr(a), fptr(b), fptr(c), fptr(d)))

    def bbox_contains_point(self, mn, mx, p):
        a, b, c = f32(mn), f32(mx), f32(p)
        return bool(self.lib.orc_bbox_contains_point(fptr(a), fptr(b), fptr(c)))

    def reflect(self, v, n):
        a, b, out = f32(v), f32(n), np.zeros(3, np.float32)
        self.lib.orc_reflect(fptr(a), fptr(b), fptr(out))
        return out

    def mat_mul_tuple(self, m, t):
        tt, out = f32(t), np.zeros(4, np.float32)
        self.lib.orc_mat_mul_tuple(fptr(m.m), fptr(tt), fptr(out))
        return out

    def csg_allowed(self, op, hit_s1, in_s1, in_s2):
        return bool(self.lib.orc_csg_allowed(op, int(hit_s1), int(in_s1), int(in_s2)))

    def csg_filter(self, csg, xs):
        ts, hs = f32([t for t, _ in xs]), _ints([s.handle for _, s in xs])
        out = np.zeros(len(xs), np.float32)
        n = self._ck(self.lib.orc_csg_filter(self.ctx, csg.handle, len(xs), fptr(ts), hs.ctypes.data_as(IP), fptr(out),
                                             len(xs)))
        return out[:n]

    def set_threads(self, n):
        self.lib.orc_set_threads(self.ctx, int(n))

    def max_threads(self):
        return self.lib.orc_max_threads()

    def render_rows(self, camera, world, depth, y0, y1, ystep, want_u8=False):
        w, h = camera.width_pixels, camera.height_pixels
        rgb = np.zeros((h, w, 3), np.float32)
        u8 = np.zeros((h, w, 3), np.uint8) if want_u8 else None
        stats = SgStats()
        self._ck(self.lib.orc_camera_render_rows(self.ctx, camera.handle, world.handle, depth, y0, y1, ystep, fptr(rgb),
                                                 u8.ctypes.data_as(U8P) if u8 is not None else None, C.byref(stats)))
        return rgb, u8, stats


def load_oracle():
    if not os.path.exists(ORACLE_SO) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
            for f in ("oracle_capi.cpp", "rtc_oracle.hpp")):
        build_oracle()
    api = load_api(ORACLE_SO, _PROBES)
    api.probe = Probes(api)
    return api
