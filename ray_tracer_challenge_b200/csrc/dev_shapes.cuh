// dev_shapes.cuh — Shape::local_intersect / local_norm_at of every leaf kind and the reference's slab test.
// Part of rtc_device.cuh (include that, not this): compiled once per kernel build inside namespace rtc::RTC_NS.
#pragma once

namespace rtc {
namespace RTC_NS {

// ---------------------------------------------------------------------------------------------------
// cube.rs:90-129 — the reference's slab test, used for Cube::local_intersect and for every group / CSG
// bounding-box cull that has to be reproduced exactly.  fminf/fmaxf return the non-NaN operand like Rust's
// f32::min/max.
__device__ __forceinline__ bool aabb_ref(V3 o, V3 d, V3 mn, V3 mx, float& lo, float& hi) {
    V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // ray.rs:16
    float a = (mn.x - o.x) * inv.x, b = (mx.x - o.x) * inv.x;
    lo = fminf(a, b);
    hi = fmaxf(a, b);
    a = (mn.y - o.y) * inv.y, b = (mx.y - o.y) * inv.y;
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    a = (mn.z - o.z) * inv.z, b = (mx.z - o.z) * inv.z;
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    return hi >= fmaxf(0.0f, lo);
}

// Shape::local_intersect of every leaf kind.  Writes the distances in the reference's emission order and
// returns how many there are (0..4).
__device__ __forceinline__ int local_intersect(const DevScene& S, int type, int aux, float4 bd, V3 o, V3 d, float t[4],
                                               const float4* tri = nullptr) {
    switch (type) {
        case T_SPHERE: {  // sphere.rs:47-70 (centre is the origin)
            float a = dot(d, d);
            float b = 2.0f * dot(d, o);
            float c = dot(o, o) - 1.0f;
            float disc = b * b - 4.0f * a * c;
            if (disc < 0.0f) return 0;
            float two_a = 2.0f * a;
            float ds = sqrtf(disc);
            t[0] = (-b - ds) / two_a;
            t[1] = (-b + ds) / two_a;
            return 2;
        }
        case T_PLANE: {  // plane.rs:45-56
            if (fabsf(d.y) < kAcne) return 0;
            t[0] = -o.y / d.y;
            return 1;
        }
        case T_CUBE: {  // cube.rs:55-63
            float lo, hi;
            if (!aabb_ref(o, d, mk(-1.f, -1.f, -1.f), mk(1.f, 1.f, 1.f), lo, hi)) return 0;
            t[0] = lo;
            t[1] = hi;
            return 2;
        }
        case T_CYLINDER: {  // cylinder.rs:52-59, 84-151
            int n = 0;
            float two_a = 2.0f * (d.x * d.x + d.z * d.z);
            if (!(fabsf(two_a) < kCloseToZero)) {
                float b = 2.0f * (o.x * d.x + o.z * d.z);
                float c = o.x * o.x + o.z * o.z - 1.0f;
                float disc = b * b - 2.0f * two_a * c;
                if (!(disc < 0.0f)) {
                    float ds = sqrtf(disc);
                    float d1 = (-b - ds) / two_a;
                    float d2 = (-b + ds) / two_a;
                    if (d1 > d2) {
                        float tmp = d1;
                        d1 = d2;
                        d2 = tmp;
                    }
                    float y1 = o.y + d1 * d.y;
                    if (bd.x < y1 && y1 < bd.y) t[n++] = d1;
                    float y2 = o.y + d2 * d.y;
                    if (bd.x < y2 && y2 < bd.y) t[n++] = d2;
                }
            }
            if (n < 2 && bd.z != 0.0f) {  // caps only when the walls gave fewer than two hits
                float tc = (bd.x - o.y) / d.y;
                float x = o.x + tc * d.x, z = o.z + tc * d.z;
                if ((x * x + z * z) <= 1.0f + kCloseToZero) t[n++] = tc;
                tc = (bd.y - o.y) / d.y;
                x = o.x + tc * d.x, z = o.z + tc * d.z;
                if ((x * x + z * z) <= 1.0f + kCloseToZero) t[n++] = tc;
            }
            return n;
        }
        case T_CONE: {  // cone.rs:52-57, 89-174
            int n = 0;
            float two_a = 2.0f * (d.x * d.x - d.y * d.y + d.z * d.z);
            float b = 2.0f * (o.x * d.x - o.y * d.y + o.z * d.z);
            if (fabsf(two_a) < kCloseToZero) {
                if (!(fabsf(b) < kCloseToZero)) {
                    float c = o.x * o.x - o.y * o.y + o.z * o.z;
                    t[n++] = -c / (2.0f * b);
                }
            } else {
                float c = o.x * o.x - o.y * o.y + o.z * o.z;
                float disc = b * b - 2.0f * two_a * c;
                if (!(disc < 0.0f)) {
                    float ds = sqrtf(disc);
                    float d1 = (-b - ds) / two_a;
                    float d2 = (-b + ds) / two_a;
                    if (d1 > d2) {
                        float tmp = d1;
                        d1 = d2;
                        d2 = tmp;
                    }
                    float y1 = o.y + d1 * d.y;
                    if (bd.x < y1 && y1 < bd.y) t[n++] = d1;
                    float y2 = o.y + d2 * d.y;
                    if (bd.x < y2 && y2 < bd.y) t[n++] = d2;
                }
            }
            if (bd.z != 0.0f) {  // caps are always tested; the radius is |y|, not y^2
                float tc = (bd.x - o.y) / d.y;
                float x = o.x + tc * d.x, z = o.z + tc * d.z;
                if ((x * x + z * z) <= fabsf(bd.x) + kCloseToZero) t[n++] = tc;
                tc = (bd.y - o.y) / d.y;
                x = o.x + tc * d.x, z = o.z + tc * d.z;
                if ((x * x + z * z) <= fabsf(bd.y) + kCloseToZero) t[n++] = tc;
            }
            return n;
        }
        default: {  // T_TRIANGLE — triangle.rs:45-76
            const float4* tp = tri ? tri : S.tri + 3 * (size_t)aux;
            float4 q0 = __ldg(tp), q1 = __ldg(tp + 1), q2 = __ldg(tp + 2);
            V3 p1 = mk(q0.x, q0.y, q0.z), e1 = mk(q0.w, q1.x, q1.y), e2 = mk(q1.z, q1.w, q2.x);
            V3 dce2 = mk(d.y * e2.z - d.z * e2.y, d.z * e2.x - d.x * e2.z, d.x * e2.y - d.y * e2.x);
            float det = dot(e1, dce2);
            if (fabsf(det) < 0.0000001f) return 0;
            float f = 1.0f / det;
            V3 p1o = o - p1;
            float u = f * dot(p1o, dce2);
            if (u < 0.0f || u > 1.0f) return 0;
            V3 oce1 = mk(p1o.y * e1.z - p1o.z * e1.y, p1o.z * e1.x - p1o.x * e1.z, p1o.x * e1.y - p1o.y * e1.x);
            float v = f * dot(d, oce1);
            if (v < 0.0f || (u + v) > 1.0f) return 0;
            t[0] = f * dot(e2, oce1);
            return 1;
        }
    }
}

__device__ __forceinline__ float4 load_bound(const DevScene& S, int type, int aux) {
    return (type == T_CYLINDER || type == T_CONE) ? __ldg(&S.bound[aux]) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// The smallest non-negative distance the primitive reports for this object-space ray (what
// Intersection::hit would pick among its intersections), or a negative / NaN value when there is none.
// Same arithmetic as local_intersect, minus the work whose result cannot be the answer: for a sphere the
// far root is only divided out when the near root is negative (the sign of a quotient by 2a > 0 is the sign
// of its numerator).
__device__ __forceinline__ float nearest_t(const DevScene& S, int type, int aux, float4 bd, V3 o, V3 d,
                                           const float4* tri = nullptr) {
    {
        if (type == T_SPHERE) {  // sphere.rs:47-70
            float a = dot(d, d);
            float b = 2.0f * dot(d, o);
            float c = dot(o, o) - 1.0f;
            float disc = b * b - 4.0f * a * c;
            if (disc < 0.0f) return -1.0f;
            float two_a = 2.0f * a;
            float ds = sqrtf(disc);
            float n0 = -b - ds, n1 = -b + ds;
            if (two_a > 0.0f) {
                if (n0 >= 0.0f) return n0 / two_a;
                if (n1 >= 0.0f) return n1 / two_a;
                return -1.0f;
            }
            float t0 = n0 / two_a, t1 = n1 / two_a;
            if (t0 >= 0.0f && !(t1 < t0)) return t0;
            return t1 >= 0.0f ? t1 : t0;
        }
        if (type == T_PLANE) {  // plane.rs:45-56
            if (fabsf(d.y) < kAcne) return -1.0f;
            return -o.y / d.y;
        }
        if (type == T_CUBE) {  // cube.rs:55-63
            float lo, hi;
            if (!aabb_ref(o, d, mk(-1.f, -1.f, -1.f), mk(1.f, 1.f, 1.f), lo, hi)) return -1.0f;
            return lo >= 0.0f ? lo : hi;
        }
        {
            float t[4];
            int n = local_intersect(S, type, aux, bd, o, d, t, tri);
            float tn = -1.0f;
            for (int i = 0; i < n; i++)
                if (t[i] >= 0.0f && (!(tn >= 0.0f) || t[i] < tn)) tn = t[i];
            return tn;
        }
    }
}

// Shape::local_norm_at of every leaf kind.
__device__ __forceinline__ V3 local_normal(const DevScene& S, int type, int aux, V3 p) {
    switch (type) {
        case T_SPHERE: return p;                      // sphere.rs:71-73
        case T_PLANE: return mk(0.f, 1.f, 0.f);       // plane.rs:57-59
        case T_CUBE: {                                // cube.rs:66-80
            float xa = fabsf(p.x), ya = fabsf(p.y), za = fabsf(p.z);
            float mc = fmaxf(xa, fmaxf(ya, za));
            if (xa == mc) return mk(p.x, 0.f, 0.f);
            if (ya == mc) return mk(0.f, p.y, 0.f);
            return mk(0.f, 0.f, p.z);
        }
        case T_CYLINDER: {  // cylinder.rs:62-72
            float4 bd = __ldg(&S.bound[aux]);
            float dist2 = p.x * p.x + p.z * p.z;
            if (dist2 < 1.0f) {
                if (p.y >= bd.y - kCloseToZero) return mk(0.f, 1.f, 0.f);
                if (p.y <= bd.x + kCloseToZero) return mk(0.f, -1.f, 0.f);
            }
            return mk(p.x, 0.f, p.z);
        }
        case T_CONE: {  // cone.rs:60-73
            float4 bd = __ldg(&S.bound[aux]);
            float dist2 = p.x * p.x + p.z * p.z;
            if (dist2 < 1.0f) {
                if (p.y >= bd.y - kCloseToZero) return mk(0.f, 1.f, 0.f);
                if (p.y <= bd.x + kCloseToZero) return mk(0.f, -1.f, 0.f);
            }
            float y = sqrtf(p.x * p.x + p.z * p.z);
            y = (p.y > 0.0f) ? -y : y;
            return mk(p.x, y, p.z);
        }
        default: {  // triangle.rs:78-81 — the flat normal precomputed at construction (also for smooth triangles, Q5)
            float4 q2 = __ldg(S.tri + 3 * (size_t)aux + 2);
            return mk(q2.y, q2.z, q2.w);
        }
    }
}

}  // namespace RTC_NS
}  // namespace rtc
