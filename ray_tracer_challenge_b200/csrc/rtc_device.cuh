// rtc_device.cuh — device functions of the B200 path for the reference's Camera::render hot loop.
//
// One thread owns one pixel sample end to end: ray generation (camera.rs:60-74), world intersection
// (world.rs:52-60 — here a BVH walk with a short per-thread stack instead of the reference's linear scan
// and sort), the hit record (world.rs:212-283), Phong + patterns (phong_lighting.rs:12-63), shadow rays
// and the area-light cell loop (world.rs:104-119, rectangle_light.rs:76-88), and the reflect / refract
// recursion (world.rs:121-162) unrolled into an explicit bounded stack that keeps the reference's
// post-order arithmetic, so every colour is combined in exactly the order the Rust code combines it.
//
// Arithmetic is spelled left to right exactly as the reference spells it.  This file is compiled twice
// (rtc_kernels.cu): once with FMA contraction (the fast build) and once with -fmad=false (RTC_STRICT), the
// latter reproducing the Rust/IEEE evaluation bit for bit apart from libm (powf, cosf, atan2f, acosf).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "rtc_types.h"

#ifndef RTC_NS
#error "define RTC_NS (fast | strict) before including rtc_device.cuh"
#endif

namespace rtc {
namespace RTC_NS {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
// tuple.rs:44-46 (the w lanes are 0 for every hot-path call)
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// tuple.rs:29-43
__device__ __forceinline__ float magnitude(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 norm(V3 a) {
    float m = magnitude(a);
    return mk(a.x / m, a.y / m, a.z / m);
}
__device__ __forceinline__ V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
// ray.rs:42-44
__device__ __forceinline__ V3 reflect(V3 in, V3 n) { return -(n * 2.0f * dot(in, n) - in); }

struct Xf {
    float4 r0, r1, r2;
};
__device__ __forceinline__ Xf load_xf(const float4* p) { return Xf{__ldg(p), __ldg(p + 1), __ldg(p + 2)}; }
// matrix.rs:73-84 with w = 1 / w = 0 (the products with an exact 0 or 1 are exact)
__device__ __forceinline__ V3 xf_point(const Xf& m, V3 p) {
    return mk(m.r0.x * p.x + m.r0.y * p.y + m.r0.z * p.z + m.r0.w, m.r1.x * p.x + m.r1.y * p.y + m.r1.z * p.z + m.r1.w,
              m.r2.x * p.x + m.r2.y * p.y + m.r2.z * p.z + m.r2.w);
}
__device__ __forceinline__ V3 xf_vec(const Xf& m, V3 v) {
    return mk(m.r0.x * v.x + m.r0.y * v.y + m.r0.z * v.z, m.r1.x * v.x + m.r1.y * v.y + m.r1.z * v.z,
              m.r2.x * v.x + m.r2.y * v.y + m.r2.z * v.z);
}
// shape.rs:130 — inverse-transpose times the object normal = transpose of the stored inverse
__device__ __forceinline__ V3 xf_normal(const Xf& m, V3 n) {
    return mk(m.r0.x * n.x + m.r1.x * n.y + m.r2.x * n.z, m.r0.y * n.x + m.r1.y * n.y + m.r2.y * n.z,
              m.r0.z * n.x + m.r1.z * n.y + m.r2.z * n.z);
}

// explicitly fused / approximate arithmetic for the conservative pre-tests and the shadow filter (never for values
// that reach a pixel): the same instructions in the IEEE and the FMA-contracting build
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float rcp_(float a) { return __fdividef(1.0f, a); }

constexpr float kInfF = __builtin_huge_valf();
constexpr float kAcne = 1.1920929e-7f * 10000.0f;  // world.rs:210
constexpr float kCloseToZero = 0.000001f;           // cylinder.rs:82, cone.rs:87

// Work counters.  Every kernel keeps the four ray / shade counts (Rays: one register each, only touched by
// inlined code so they never leave the register file); the detailed build (STATS) also counts every unit of
// SURVEY.md Appendix E in Ctr<true>, which is what the out-of-line helpers receive (Ctr<false> is empty).
struct Rays {
    unsigned primary = 0, secondary = 0, shadow = 0, shades = 0;
};
template <bool STATS>
struct Ctr;
template <>
struct Ctr<false> {
    __device__ __forceinline__ void node() {}
    __device__ __forceinline__ void prim(int) {}
    __device__ __forceinline__ void xform() {}
    __device__ __forceinline__ void pattern() {}
    __device__ __forceinline__ void cell() {}
    __device__ __forceinline__ void schlick() {}
    __device__ __forceinline__ void refr_dir() {}
    __device__ __forceinline__ void overflow() {}
    __device__ __forceinline__ void refiltered() {}
};
template <>
struct Ctr<true> {
    unsigned nodes = 0, prims[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xforms = 0, patterns = 0, cells = 0, schlicks = 0, refr_dirs = 0,
             overflows = 0, refilters = 0;
    __device__ __forceinline__ void node() { nodes++; }
    __device__ __forceinline__ void prim(int t) { prims[t]++; }
    __device__ __forceinline__ void xform() { xforms++; }
    __device__ __forceinline__ void pattern() { patterns++; }
    __device__ __forceinline__ void cell() { cells++; }
    __device__ __forceinline__ void schlick() { schlicks++; }
    __device__ __forceinline__ void refr_dir() { refr_dirs++; }
    __device__ __forceinline__ void overflow() { overflows++; }
    __device__ __forceinline__ void refiltered() { refilters++; }
};

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

// ---------------------------------------------------------------------------------------------------
// Nearest hit so far.  `order` is the primitive's depth-first index: the reference's tie-break (Q8).
struct Hit {
    float t;
    int pos;
    int order;
};
__device__ __forceinline__ void consider(Hit& best, float t, int pos, int order) {
    const bool better = t >= 0.0f && (t < best.t || (t == best.t && order < best.order));
    best.t = better ? t : best.t;
    best.pos = better ? pos : best.pos;
    best.order = better ? order : best.order;
}

// The world ray as seen by the primitive being tested; cached by transform id so that a mesh whose
// triangles share one transform pays for one ray transform per traversal, not one per triangle.
struct ObjRay {
    int xf_id = -1;
    V3 o, d;
};

// Every enclosing GroupShape culls with the forward ray (group.rs:119-125); the walk stops at a CSG because
// CSG subtrees are evaluated whole by csg_eval.
__device__ __noinline__ bool ancestors_pass(const DevScene& S, int node, V3 o, V3 d) {
    while (node >= 0) {
        const DevNode& n = S.nodes[node];
        float lo, hi;
        if (!aabb_ref(o, d, ld3(n.bmin), ld3(n.bmax), lo, hi)) return false;
        node = n.parent;
    }
    return true;
}

// CSG::local_intersect (csg.rs:87-104) for a whole CSG subtree, flattened at commit time into a post-order
// instruction list.  ht/hp receive the filtered hits (distance, primitive position) in the reference's
// sorted order; returns their count.
template <bool STATS>
__device__ __noinline__ int csg_eval(const DevScene& S, int pc, V3 wo, V3 wd, float* ht, int* hp, Ctr<STATS>& k) {
    V3 ro[kCsgRayDepth], rd[kCsgRayDepth];
    int s_mark[kCsgRayDepth], m_mark[kCsgRayDepth];
    int rsp = 0, n = 0;
    ro[0] = wo;
    rd[0] = wd;
    k.prim(T_CSG);
    const int end = S.csg_ops[pc].skip;  // the root's ENTER skips to one past its EXIT
    while (pc < end) {
        DevCsgOp op = S.csg_ops[pc];
        switch (op.op) {
            case OP_CSG_ENTER: {
                const DevNode& nd = S.nodes[op.arg];
                Xf m{nd.inv[0], nd.inv[1], nd.inv[2]};
                V3 o2 = xf_point(m, ro[rsp]), d2 = xf_vec(m, rd[rsp]);  // shape.rs:60-70 on the CSG itself
                k.xform();
                float lo, hi;
                k.node();
                if (!aabb_ref(o2, d2, ld3(nd.bmin), ld3(nd.bmax), lo, hi)) {  // csg.rs:90-93
                    pc = op.skip;
                    break;
                }
                rsp++;
                ro[rsp] = o2;
                rd[rsp] = d2;
                s_mark[rsp] = n;
                pc++;
                break;
            }
            case OP_CSG_MID:
                m_mark[rsp] = n;
                pc++;
                break;
            case OP_CSG_EXIT: {
                const int s = s_mark[rsp], m = m_mark[rsp], e = n;
                const int csg_op = S.nodes[op.arg].op;
                for (int i = s; i < m; i++) hp[i] |= 0x40000000;  // came from s1 (csg.rs:46 `s1.includes`)
                for (int i = s + 1; i < e; i++) {                 // stable insertion sort by distance (csg.rs:101)
                    float ti = ht[i];
                    int pi = hp[i];
                    int j = i - 1;
                    while (j >= s && ht[j] > ti) {
                        ht[j + 1] = ht[j];
                        hp[j + 1] = hp[j];
                        j--;
                    }
                    ht[j + 1] = ti;
                    hp[j + 1] = pi;
                }
                bool in1 = false, in2 = false;  // csg.rs:37-58
                int w = s;
                for (int i = s; i < e; i++) {
                    bool hit1 = (hp[i] & 0x40000000) != 0;
                    bool allowed;
                    if (csg_op == 0)
                        allowed = (hit1 && !in2) || (!hit1 && !in1);
                    else if (csg_op == 1)
                        allowed = (hit1 && in2) || (!hit1 && in1);
                    else
                        allowed = (hit1 && !in2) || (!hit1 && in1);
                    if (allowed) {
                        ht[w] = ht[i];
                        hp[w] = hp[i] & 0x3fffffff;
                        w++;
                    }
                    if (hit1)
                        in1 = !in1;
                    else
                        in2 = !in2;
                }
                n = w;
                rsp--;
                pc++;
                break;
            }
            case OP_GROUP: {
                const DevNode& nd = S.nodes[op.arg];
                float lo, hi;
                k.node();
                if (!aabb_ref(ro[rsp], rd[rsp], ld3(nd.bmin), ld3(nd.bmax), lo, hi))  // group.rs:122-125
                    pc = op.skip;
                else
                    pc++;
                break;
            }
            default: {  // OP_PRIM
                int4 h = __ldg(&S.head[op.arg]);
                Xf m = load_xf(S.xform + 3 * (size_t)h.y);
                V3 o2 = xf_point(m, ro[rsp]), d2 = xf_vec(m, rd[rsp]);
                k.xform();
                float t[4];
                int type = h.x & 15;
                k.prim(type);
                int c = local_intersect(S, type, h.z, load_bound(S, type, h.z), o2, d2, t);
                for (int i = 0; i < c; i++) {
                    if (n < kCsgHitCap) {
                        ht[n] = t[i];
                        hp[n] = op.arg;
                        n++;
                    } else {
                        k.overflow();
                    }
                }
                pc++;
                break;
            }
        }
    }
    return n;
}

// Test one stored primitive against the world ray for the nearest-hit search.
template <bool STATS>
__device__ __forceinline__ void test_prim(const DevScene& S, int pos, V3 o, V3 d, ObjRay& cache, Hit& best, Ctr<STATS>& k) {
    const float4* rec = S.rec + 4 * (size_t)pos;
    int4 h = __ldg(reinterpret_cast<const int4*>(rec));
    int type = h.x & 15;
    if (type == T_CSG) {
        float ht[kCsgHitCap];
        int hp[kCsgHitCap];
        int n = csg_eval<STATS>(S, h.z, o, d, ht, hp, k);
        for (int i = 0; i < n; i++) {
            if (ht[i] >= 0.0f) {
                consider(best, ht[i], hp[i], __ldg(&S.head[hp[i]]).w);
                break;  // the list is sorted: the first non-negative entry is this CSG's nearest
            }
        }
        return;
    }
    float tn;
    k.prim(type);
    if (type == T_TRIANGLE) {
        // a mesh's triangles share one transform: the object-space ray is kept across primitives (shape.rs:60-70)
        if (h.y != cache.xf_id) {
            Xf m = load_xf(S.xform + 3 * (size_t)h.y);
            cache.o = xf_point(m, o);
            cache.d = xf_vec(m, d);
            cache.xf_id = h.y;
            k.xform();
        }
        tn = nearest_t(S, T_TRIANGLE, h.z, make_float4(0.f, 0.f, 0.f, 0.f), cache.o, cache.d, rec + 1);
    } else {
        Xf m = load_xf(rec + 1);  // the primitive's own inverse transform travels in its record
        k.xform();
        tn = nearest_t(S, type, h.z, load_bound(S, type, h.z), xf_point(m, o), xf_vec(m, d));
    }
    if (!(tn >= 0.0f)) return;
    if (((h.x >> 4) & kFlagHasParent) && !ancestors_pass(S, __ldg(&S.head[pos + S.n_prims]).x, o, d)) return;
    consider(best, tn, pos, h.w);
}

// Conservative slab test for the acceleration structure (boxes are padded at build time, so this never
// rejects a primitive the reference would have hit; it is not part of the reference's semantics).
__device__ __forceinline__ bool slab(V3 o, V3 inv, float lx, float ly, float lz, float hx, float hy, float hz, float tmax,
                                     float& tnear) {
    float a = (lx - o.x) * inv.x, b = (hx - o.x) * inv.x;
    float lo = fminf(a, b), hi = fmaxf(a, b);
    a = (ly - o.y) * inv.y, b = (hy - o.y) * inv.y;
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    a = (lz - o.z) * inv.z, b = (hz - o.z) * inv.z;
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    tnear = lo;
    return hi >= fmaxf(lo, 0.0f) && lo <= tmax;
}

// World::intersect + Intersection::hit (world.rs:52-60, intersection.rs:30-35): nearest t >= 0 with the
// depth-first tie-break, searched through the BVH.  `best.t` on entry is the search limit (exclusive).
// ANY: stop at the first hit (shadow rays when every primitive casts a shadow).
template <bool STATS, bool ANY>
__device__ __noinline__ void nearest_hit(const DevScene& S, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    ObjRay cache;
    for (int i = 0; i < S.n_linear; i++) {
        test_prim<STATS>(S, __ldg(&S.linear[i]), o, d, cache, best, k);
        if (ANY && best.pos >= 0) return;
    }
    if (S.bvh_root < 0) return;
    V3 inv = mk(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[kBvhStack];
    int sp = 0;
    int node = S.bvh_root;
    // "while-while" traversal: every lane first descends to its next leaf (lanes that are there already wait), then
    // all lanes test their leaf's primitives together — instead of serialising an inner-node step for some lanes with
    // a leaf for the others in every iteration (ncu: 4.5 of 32 lanes active in the primitive tests before).
    constexpr int kDone = -0x7fffffff - 1;  // never a leaf code: ~((first << 4) | (count - 1)) > INT_MIN
    for (;;) {
        while (node >= 0) {
            const float4* np = reinterpret_cast<const float4*>(S.bvh + node);
            float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            int4 link = __ldg(reinterpret_cast<const int4*>(np + 3));
            float t0, t1;
            k.node();
            k.node();
            bool h0 = slab(o, inv, a.x, a.y, a.z, a.w, b.x, b.y, best.t, t0);
            bool h1 = slab(o, inv, b.z, b.w, c.x, c.y, c.z, c.w, best.t, t1);
            if (h0 && h1) {
                int near = link.x, far = link.y;
                if (t1 < t0) {
                    near = link.y;
                    far = link.x;
                }
                if (sp < kBvhStack) stack[sp++] = far;
                node = near;
            } else if (h0) {
                node = link.x;
            } else if (h1) {
                node = link.y;
            } else {
                node = sp == 0 ? kDone : stack[--sp];
            }
        }
        if (node == kDone) return;
        {
            int code = ~node;
            int first = code >> 4, count = (code & 15) + 1;
            for (int i = 0; i < count; i++) {
                test_prim<STATS>(S, first + i, o, d, cache, best, k);
                if (ANY && best.pos >= 0) return;
            }
        }
        if (sp == 0) return;
        node = stack[--sp];
    }
}

// ---------------------------------------------------------------------------------------------------
// n1 / n2 (world.rs:235-263) without materialising the sorted list.  In the render
// path the hit is the first t >= 0 entry, so the containers are decided by the NEGATIVE-t intersections:
// an object is open when it has an odd number of them, and the open object whose last negative hit is
// latest in the sorted order (t, then depth-first order) is the innermost one.
struct Containers {
    float best_t = -kInfF;  // innermost open object other than the hit object
    int best_order = -1;
    int best_pos = -1;
    bool hit_open = false;  // the hit object itself is open (we are leaving it)
    float hit_t = -kInfF;
    int hit_order = -1;
};
__device__ __forceinline__ void container_add(Containers& c, int hit_pos, float t_last, int pos, int order) {
    if (pos == hit_pos) {
        c.hit_open = true;
        c.hit_t = t_last;
        c.hit_order = order;
    } else if (t_last > c.best_t || (t_last == c.best_t && order > c.best_order)) {
        c.best_t = t_last;
        c.best_order = order;
        c.best_pos = pos;
    }
}
template <bool STATS>
__device__ __forceinline__ void container_prim(const DevScene& S, int pos, V3 o, V3 d, ObjRay& cache, int hit_pos, Containers& c,
                                               Ctr<STATS>& k) {
    int4 h = __ldg(&S.head[pos]);
    int type = h.x & 15;
    if (type == T_CSG) {
        float ht[kCsgHitCap];
        int hp[kCsgHitCap];
        int n = csg_eval<STATS>(S, h.z, o, d, ht, hp, k);
        // per leaf: parity and last negative hit among the filtered hits
        for (int i = 0; i < n; i++) {
            if (!(ht[i] < 0.0f)) break;
            int p = hp[i];
            bool seen = false;
            for (int j = 0; j < i; j++) seen |= (hp[j] == p);
            if (seen) continue;
            int cnt = 0;
            float last = 0.f;
            for (int j = i; j < n && ht[j] < 0.0f; j++)
                if (hp[j] == p) {
                    cnt++;
                    last = ht[j];
                }
            if (cnt & 1) container_add(c, hit_pos, last, p, __ldg(&S.head[p]).w);
        }
        return;
    }
    if (h.y != cache.xf_id) {
        Xf m = load_xf(S.xform + 3 * (size_t)h.y);
        cache.o = xf_point(m, o);
        cache.d = xf_vec(m, d);
        cache.xf_id = h.y;
        k.xform();
    }
    float t[4];
    k.prim(type);
    int n = local_intersect(S, type, h.z, load_bound(S, type, h.z), cache.o, cache.d, t);
    int cnt = 0;
    float last = -kInfF;
    for (int i = 0; i < n; i++)
        if (t[i] < 0.0f) {
            cnt++;
            last = fmaxf(last, t[i]);
        }
    if (!(cnt & 1)) return;
    // a grouped primitive only reports hits when every enclosing group's box passes the forward ray (Q6)
    int parent = __ldg(&S.head[pos + S.n_prims]).x;
    if (parent >= 0 && !ancestors_pass(S, parent, o, d)) return;
    container_add(c, hit_pos, last, pos, h.w);
}
template <bool STATS>
__device__ __noinline__ void find_containers(const DevScene& S, V3 o, V3 d, int hit_pos, float& n1, float& n2, Ctr<STATS>& k) {
    Containers c;
    ObjRay cache;
    for (int i = 0; i < S.n_linear; i++) container_prim<STATS>(S, __ldg(&S.linear[i]), o, d, cache, hit_pos, c, k);
    if (S.bvh_root >= 0) {
        // walk the backward half-line: the forward half-line of the reversed ray
        V3 inv = mk(-1.0f / d.x, -1.0f / d.y, -1.0f / d.z);
        int stack[kBvhStack];
        int sp = 0;
        int node = S.bvh_root;
        constexpr int kDone = -0x7fffffff - 1;
        for (;;) {  // while-while, as in nearest_hit
            while (node >= 0) {
                const float4* np = reinterpret_cast<const float4*>(S.bvh + node);
                float4 a = __ldg(np), b = __ldg(np + 1), cc = __ldg(np + 2);
                int4 link = __ldg(reinterpret_cast<const int4*>(np + 3));
                float t0, t1;
                k.node();
                k.node();
                // A sphere or a cube has an odd number of hits behind the origin only when the origin is inside it (both
                // roots of an outside origin have one sign, cube.rs:124-128 rejects a box that is not straddled), and
                // the tree's boxes are padded far beyond f32 rounding: subtrees of such primitives are culled with a
                // point-in-box test instead of the unbounded backward ray (a refraction hit deep in a 100 k-sphere
                // field no longer walks every node along the whole half-line).
                bool h0, h1;
                if (link.z & 1)
                    h0 = o.x >= a.x && o.x <= a.w && o.y >= a.y && o.y <= b.x && o.z >= a.z && o.z <= b.y;
                else
                    h0 = slab(o, inv, a.x, a.y, a.z, a.w, b.x, b.y, kInfF, t0);
                if (link.z & 2)
                    h1 = o.x >= b.z && o.x <= cc.y && o.y >= b.w && o.y <= cc.z && o.z >= cc.x && o.z <= cc.w;
                else
                    h1 = slab(o, inv, b.z, b.w, cc.x, cc.y, cc.z, cc.w, kInfF, t1);
                if (h0 && h1) {
                    if (sp < kBvhStack) stack[sp++] = link.y;
                    node = link.x;
                } else if (h0) {
                    node = link.x;
                } else if (h1) {
                    node = link.y;
                } else {
                    node = sp == 0 ? kDone : stack[--sp];
                }
            }
            if (node == kDone) break;
            {
                int code = ~node;
                int first = code >> 4, count = (code & 15) + 1;
                for (int i = 0; i < count; i++) container_prim<STATS>(S, first + i, o, d, cache, hit_pos, c, k);
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    auto index_of = [&](int pos) { return S.materials[(__ldg(&S.head[pos]).x >> 8)].refractive_index; };
    float other = (c.best_pos >= 0) ? index_of(c.best_pos) : 1.0f;  // REFRACTION_VACCUM, world.rs:242
    if (c.hit_open) {
        // leaving the hit object: n1 is the innermost open object (possibly the hit object itself)
        bool hit_is_inner = c.best_pos < 0 || c.hit_t > c.best_t || (c.hit_t == c.best_t && c.hit_order > c.best_order);
        n1 = hit_is_inner ? index_of(hit_pos) : other;
        n2 = other;
    } else {
        n1 = other;
        n2 = index_of(hit_pos);
    }
}

// ---------------------------------------------------------------------------------------------------
// Patterns (pattern/*.rs).  floor(..) as i32 % 2 with Rust semantics: the cast saturates and NaN maps to 0
// (__float2int_rz does both), and % keeps the dividend's sign so negative odd numbers give -1 != 0 (Q16).
__device__ __forceinline__ bool even_floor(float v) { return (__float2int_rz(floorf(v)) % 2) == 0; }
__device__ __forceinline__ float rem_euclid(float a, float b) {
    float r = fmodf(a, b);
    return (r < 0.0f) ? r + fabsf(b) : r;
}
__device__ __forceinline__ V3 uv_color(const DevScene& S, int id, float u, float v) {
    const DevUvPattern& p = S.uvs[id];
    if (p.kind == 0) {  // UVCheckers, uv.rs:45-56
        int u2 = __float2int_rz(floorf(u * p.p[0]));
        int v2 = __float2int_rz(floorf(v * p.p[1]));
        int s = (int)((unsigned)u2 + (unsigned)v2);
        return (s % 2 == 0) ? ld3(p.p + 2) : ld3(p.p + 5);
    }
    if (p.kind == 2) {  // UVImage, uv.rs:366-377: nearest pixel, v flipped (row 0 is the top of the image)
        const int w = __float_as_int(p.p[1]), h = __float_as_int(p.p[2]);
        const float x = u * (float)(w - 1);
        const float y = (1.f - v) * (float)(h - 1);
        // `x.round() as usize`: half away from zero; the cast saturates and maps NaN to 0.  Beyond the canvas the
        // reference panics (canvas.rs:35); the device clamps to the edge.
        const int xi = min(max(__float2int_rz(roundf(x)), 0), w - 1);
        const int yi = min(max(__float2int_rz(roundf(y)), 0), h - 1);
        const float4 t = __ldg(&S.texels[(size_t)__float_as_int(p.p[0]) + (size_t)yi * w + xi]);
        return mk(t.x, t.y, t.z);
    }
    // AlignCheck, uv.rs:155-176
    if (v > 0.8f) {
        if (u < 0.2f) return ld3(p.p + 3);
        if (u > 0.8f) return ld3(p.p + 6);
    } else if (v < 0.2f) {
        if (u < 0.2f) return ld3(p.p + 9);
        if (u > 0.8f) return ld3(p.p + 12);
    }
    return ld3(p.p);
}
__device__ __forceinline__ float u_from_azimuth(V3 p) {  // uv.rs:117-132
    const float frac_1_2pi = 1.0f / (2.0f * 3.14159265358979323846f);
    float theta = atan2f(p.x, p.z);
    float raw_u = theta * frac_1_2pi;
    return 1.f - (raw_u + 0.5f);
}
__device__ __noinline__ V3 pattern_color(const DevScene& S, int pid, V3 object_point) {
    const DevPattern& P = S.patterns[pid];
    Xf m{P.inv[0], P.inv[1], P.inv[2]};
    V3 p = xf_point(m, object_point);  // pattern.rs:17
    V3 a = ld3(P.a), b = ld3(P.b);
    switch (P.kind) {
        case 0: return even_floor(p.x) ? a : b;                                       // stripes.rs:39-45
        case 1: return a + (b * (p.x - floorf(p.x)));                                 // gradient.rs:33-36 (b holds `distance`)
        case 2: return even_floor(sqrtf(p.x * p.x + p.z * p.z)) ? a : b;              // rings.rs:38-50
        case 3: return even_floor(fabsf(p.x) + fabsf(p.y) + fabsf(p.z)) ? a : b;      // checkers.rs:38-46
        case 4: {                                                                     // sine_2d.rs:39-44
            float cosine = cosf(p.x + p.z);
            float fraction = (-cosine + 1.0f) / 2.0f;
            return a + (b * fraction);
        }
        case 5: return p;  // TestPattern, pattern.rs:84-86
        case 6: {          // TextureMap, uv.rs:89-132,196-212
            const float pi = 3.14159265358979323846f;
            float u, v;
            if (P.mapping == 0) {
                u = u_from_azimuth(p);
                float radius = magnitude(p);
                float phi = acosf(p.y / radius);
                v = 1.f - phi * 0.318309886183790671538f;
            } else if (P.mapping == 1) {
                u = rem_euclid(p.x, 1.f);
                v = rem_euclid(p.z, 1.f);
            } else {
                u = u_from_azimuth(p);
                v = rem_euclid(p.y, 2.f * pi) * (1.0f / (2.0f * pi));
            }
            return uv_color(S, P.uv[0], u, v);
        }
        default: {  // CubicMap, uv.rs:256-326 (Face: Front 0, Back 1, Left 2, Right 3, Up 4, Down 5)
            float coord = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z));
            int face;
            if (coord == p.x)
                face = 3;
            else if (coord == -p.x)
                face = 2;
            else if (coord == p.y)
                face = 4;
            else if (coord == -p.y)
                face = 5;
            else if (coord == p.z)
                face = 0;
            else
                face = 1;
            float u, v;
            switch (face) {
                case 0: u = fmodf(p.x + 1.f, 2.f) / 2.f, v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
                case 1: u = fmodf(1.f - p.x, 2.f) / 2.f, v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
                case 2: u = fmodf(p.z + 1.f, 2.f) / 2.f, v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
                case 3: u = fmodf(1.f - p.z, 2.f) / 2.f, v = fmodf(p.y + 1.f, 2.f) / 2.f; break;
                case 4: u = fmodf(p.x + 1.f, 2.f) / 2.f, v = fmodf(1.f - p.z, 2.f) / 2.f; break;
                default: u = fmodf(p.x + 1.f, 2.f) / 2.f, v = fmodf(p.z + 1.f, 2.f) / 2.f; break;
            }
            return uv_color(S, P.uv[face], u, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Counter-based stand-in for thread_rng().sample(OpenClosed01) (rectangle_light.rs:46): identical to the
// oracle's jitter_hash / jitter_open_closed01.
__device__ __forceinline__ float jitter_value(unsigned long long seed, unsigned pixel, unsigned path, unsigned index) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(pixel + 1u);
    z ^= ((unsigned long long)path << 32) | index;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    unsigned bits = (unsigned)(z >> 32);
    return (float)((bits >> 8) + 1u) * 5.9604644775390625e-08f;
}

// ---------------------------------------------------------------------------------------------------
// What a thread needs to trace: the scene and the small-scene table (both in the kernel parameter block).
struct Env {
    const DevScene& S;
    const SmallScene& SS;
};

// Small scenes keep these in (dynamic) shared memory, kSmallSmemBytes in all:
//   tab     : the primitive table, kSmallStride x float4 per primitive {head, row0, row1, row2, bound, ball}, copied
//             from the parameter block by stage_small_scene — every thread of a warp reads the same entry, so a row
//             is one broadcast LDS.128 with an immediate offset;
//   org     : per thread, the object-space origin of the current shade's shadow rays for the first kOrgCache
//             primitives (element (i, c) of thread t at org[(i * 3 + c) * 128 + t]) — the per-cell shadow loop;
//   samples : table-mode area light: the `cells` sample points (cell-mask loops, intensity_cells);
//   plane cells : per (caster plane, cell) the constants of filter_plane_cell.
__device__ __forceinline__ const float4* small_tab() {
    extern __shared__ float4 rtc_smem[];
    return rtc_smem;
}
__device__ __forceinline__ float* small_org() {
    extern __shared__ float4 rtc_smem[];
    return reinterpret_cast<float*>(rtc_smem + kSmallCap * kSmallStride) + threadIdx.x;
}
// Shadow filter, plane test with the light sample folded in (SmallScene::plane_cells): for light point L and shading
// point p the object-space direction's y is r1.L - r1.p, so everything that depends on L alone is staged once per
// block: {r1.L, tol * sum|r1_k L_k|, EPSILON' * |L|_1}.  See filter_plane_cell.
constexpr float kTolP = 3.814697265625e-06f;  // 2^-18 = 64 ulp: planes and cubes (bounds are term-wise, no conditioning)
__device__ __forceinline__ float4 plane_cell_constants(float4 r1, float4 L) {
    const float px = r1.x * L.x, py = r1.y * L.y, pz = r1.z * L.z;
    return make_float4(px + py + pz, kTolP * (fabsf(px) + fabsf(py) + fabsf(pz)),
                       (1.1920929e-3f * (1.0f + 2.0f * kTolP)) * (fabsf(L.x) + fabsf(L.y) + fabsf(L.z)), 0.0f);
}
__device__ __forceinline__ const float4* small_plane_cells() {
    extern __shared__ float4 rtc_smem[];
    return rtc_smem + kSmallCap * kSmallStride + kOrgCache * 3 * 128 / 4 + kSampleCap;
}
__device__ __forceinline__ const float4* small_samples() {  // table-mode light samples (SmallScene::cell_masks)
    extern __shared__ float4 rtc_smem[];
    return rtc_smem + kSmallCap * kSmallStride + kOrgCache * 3 * 128 / 4;
}
__device__ __forceinline__ void stage_small_scene(const DevScene& S, const SmallScene& SS) {
    extern __shared__ float4 rtc_smem[];
    const float4* src = reinterpret_cast<const float4*>(SS.p);
    for (int i = threadIdx.x; i < SS.n * kSmallStride; i += blockDim.x) rtc_smem[i] = src[i];
    if (SS.cell_masks && S.jitter_len > 0) {
        float4* dst = rtc_smem + kSmallCap * kSmallStride + kOrgCache * 3 * 128 / 4;
        for (int i = threadIdx.x; i < S.cells; i += blockDim.x) dst[i] = __ldg(&S.samples[i]);
        if (SS.plane_cells) {  // see plane_cell_constants
            const int n_planes = SS.caster_end.y - SS.caster_end.x;
            for (int i = threadIdx.x; i < n_planes * S.cells; i += blockDim.x) {
                const float4 r1 = SS.p[SS.caster_end.x + i / S.cells].r1;
                const float4 L = __ldg(&S.samples[i % S.cells]);
                dst[kSampleCap + i] = plane_cell_constants(r1, L);
            }
        }
    }
    __syncthreads();
}

// Object-space origin of primitive i: from the per-shade cache (all shadow rays of one shade share their
// origin, so `inverse * origin`, shape.rs:60-70, is evaluated once per primitive instead of once per light
// cell — the same arithmetic, hoisted) or computed.
__device__ __forceinline__ V3 small_origin(bool cached, int i, const Xf& m, V3 o) {
    if (cached && i < kOrgCache) {
        const float* org = small_org();
        return mk(org[(i * 3 + 0) * 128], org[(i * 3 + 1) * 128], org[(i * 3 + 2) * 128]);
    }
    return xf_point(m, o);
}

__device__ __forceinline__ void cache_origins(const Env& E, V3 o) {
    const int n = E.SS.n < kOrgCache ? E.SS.n : kOrgCache;
    const float4* tab = small_tab();
    float* org = small_org();
    for (int i = 0; i < n; i++) {
        Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
        V3 o2 = xf_point(m, o);
        org[(i * 3 + 0) * 128] = o2.x;
        org[(i * 3 + 1) * 128] = o2.y;
        org[(i * 3 + 2) * 128] = o2.z;
    }
}

// Ray-vs-bounding-ball pre-test of the small-scene loops: true when the LINE o + t d stays clear of the primitive's
// world-space ball {centre, radius} grown by `pad_rate * |centre - o|^2` (SmallPrim::ball / bound.w) — then the
// reference's intersection test reports no hit either, with any sign of t (its f32 error, which for a far, small
// object grows with the squared distance over the radius, is inside the padding), and the exact test is skipped.
// ~14 instructions against ~90 for a sphere and ~250 for a cylinder or cone.  NaN ball (planes, CSG): never true.
__device__ __forceinline__ bool ball_missed(float4 ball, float pad_rate, V3 o, V3 d) {
    const float wx = ball.x - o.x, wy = ball.y - o.y, wz = ball.z - o.z;
    const float ww = fma_(wx, wx, fma_(wy, wy, wz * wz));
    const float wd = fma_(wx, d.x, fma_(wy, d.y, wz * d.z));
    const float dd = fma_(d.x, d.x, fma_(d.y, d.y, d.z * d.z));
    const float R = fma_(pad_rate, ww, ball.w);
    // squared distance of the centre from the line, times |d|^2, against the squared radius times |d|^2 (0.1 % margin)
    return fma_(ww, dd, -(wd * wd)) > R * R * dd * 1.001f;
}

// One item of a small scene of any kind (primitive or CSG root), honouring the cull chain: the general form
// (out of line: cylinders, cones, triangles and CSG roots are the rare members of small scenes).
template <bool STATS>
__device__ __noinline__ void test_small(const Env& E, int i, bool cached, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    const float4* tab = small_tab();
    const int4 head = *reinterpret_cast<const int4*>(tab + i * kSmallStride);
    const int type = head.x & 15;
    if (type == T_CSG) {
        float ht[kCsgHitCap];
        int hp[kCsgHitCap];
        int n = csg_eval<STATS>(E.S, head.z, o, d, ht, hp, k);
        for (int j = 0; j < n; j++) {
            if (ht[j] >= 0.0f) {
                consider(best, ht[j], hp[j], __ldg(&E.S.head[hp[j]]).w);
                break;
            }
        }
        return;
    }
    Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
    V3 o2 = cached ? small_origin(true, i, m, o) : xf_point(m, o);
    V3 d2 = xf_vec(m, d);
    k.xform();
    k.prim(type);
    float tn = nearest_t(E.S, type, head.z, tab[i * kSmallStride + 4], o2, d2);
    if (((head.x >> 4) & kFlagHasParent) && tn >= 0.0f && !ancestors_pass(E.S, head.y, o, d)) return;
    consider(best, tn, i, head.w);
}

// Nearest hit among the items [begin, ends.w) of a small scene, which are runs of spheres, planes, cubes and
// "everything else" ending at ends.x / .y / .z / .w: one tight loop per kind, no per-item dispatch.
// ANY: return true as soon as some item is hit in [0, best.t) (shadow rays when every object casts a shadow).
template <bool STATS, bool ANY>
__device__ __forceinline__ bool scan_small(const Env& E, bool cached, int begin, int4 ends, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    constexpr bool any = ANY;
    const float4* tab = small_tab();
    int i = begin;
    for (; i < ends.x; i++) {  // spheres — sphere.rs:47-70
        if (ball_missed(tab[i * kSmallStride + 5], tab[i * kSmallStride + 4].w, o, d)) continue;
        Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
        V3 o2 = small_origin(cached, i, m, o);
        V3 d2 = xf_vec(m, d);
        k.xform();
        k.prim(T_SPHERE);
        float t = nearest_t(E.S, T_SPHERE, 0, make_float4(0.f, 0.f, 0.f, 0.f), o2, d2);
        if (any && t >= 0.0f && t < best.t) return true;
        consider(best, t, i, __float_as_int(tab[i * kSmallStride].w));
    }
    for (; i < ends.y; i++) {  // planes — plane.rs:45-56 only reads the y components of the object-space ray
        float4 r1 = tab[i * kSmallStride + 2];
        float oy;
        if (cached && i < kOrgCache)
            oy = small_org()[(i * 3 + 1) * 128];
        else
            oy = r1.x * o.x + r1.y * o.y + r1.z * o.z + r1.w;
        float dy = r1.x * d.x + r1.y * d.y + r1.z * d.z;
        k.xform();
        k.prim(T_PLANE);
        float t = (fabsf(dy) < kAcne) ? -1.0f : -oy / dy;
        if (any && t >= 0.0f && t < best.t) return true;
        consider(best, t, i, __float_as_int(tab[i * kSmallStride].w));
    }
    for (; i < ends.z; i++) {  // cubes — cube.rs:55-63
        if (ball_missed(tab[i * kSmallStride + 5], tab[i * kSmallStride + 4].w, o, d)) continue;
        Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
        V3 o2 = small_origin(cached, i, m, o);
        V3 d2 = xf_vec(m, d);
        k.xform();
        k.prim(T_CUBE);
        float t = nearest_t(E.S, T_CUBE, 0, make_float4(0.f, 0.f, 0.f, 0.f), o2, d2);
        if (any && t >= 0.0f && t < best.t) return true;
        consider(best, t, i, __float_as_int(tab[i * kSmallStride].w));
    }
    for (; i < ends.w; i++) {  // cylinders, cones, triangles, CSG roots
        if (ball_missed(tab[i * kSmallStride + 5], tab[i * kSmallStride + 4].w, o, d)) continue;
        const int before = best.pos;
        test_small<STATS>(E, i, cached, o, d, best, k);
        if (any && best.pos != before) return true;
    }
    return false;
}

// World::intersect + Intersection::hit for the nearest hit, general (BVH) or small-scene form.
template <bool STATS, bool SMALL>
__device__ __forceinline__ void find_hit(const Env& E, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    if (SMALL) {
        const SmallScene& SS = E.SS;
        if (SS.has_cull_chain) {
            for (int i = 0; i < SS.n; i++) test_small<STATS>(E, i, false, o, d, best, k);
        } else {
            int begin = 0;
            int4 ends = SS.caster_end;
#pragma unroll 1
            for (int seg = 0; seg < 2; seg++) {  // casters, then non-casters: one copy of the loops
                scan_small<STATS, false>(E, false, begin, ends, o, d, best, k);
                begin = ends.w;
                ends = SS.other_end;
            }
        }
    } else {
        nearest_hit<STATS, false>(E.S, o, d, best, k);
    }
}

// ---------------------------------------------------------------------------------------------------
// Shadow filter (small scenes made of spheres, planes and axis-aligned cubes only — SmallScene::filter_ok).
//
// World::is_shadowed only needs a BOOLEAN per light sample: "is the nearest hit in [0, distance) a shadow caster".
// The reference gets it the expensive way: normalise the direction (sqrt + 3 divisions), intersect everything,
// divide out every root.  Away from the decision boundaries the boolean does not depend on any of that rounding, so
// the filter evaluates the same predicate on the UN-normalised segment point -> light (parameter s in [0, 1),
// t = s * distance) with fused multiply-adds and approximate reciprocals, carries a forward error bound that covers
// both its own rounding and the reference's (object-space origins are the reference's own values, bit for bit; only
// the direction differs), and answers only when every comparison it needs is decided by a margin larger than that
// bound.  Anything closer than the margin — tangent rays, roots at the light, ties between objects, NaN / inf —
// returns F_UNSURE and the caller runs the exact test.  Frames are therefore bit-identical with the filter on or off
// (tests/test_gpu_parity.py::test_shadow_filter_changes_no_pixel).
enum : int { F_MISS = 0, F_HIT = 1, F_UNSURE = 2 };
struct FRes {
    int code;
    float s, e;  // F_HIT: segment parameter of the nearest hit and its error bound
};

// sphere.rs:47-70 on the segment o + s * d, d = M * (light - point).  `tol` = SmallScene::tol_sphere, which scales
// with the worst condition number of the spheres' transforms (set at commit).
// The part that runs once the discriminant is clearly positive: which root is Intersection::hit's, is it in [0, 1).
__device__ __forceinline__ FRes sphere_roots(float a, float b, float oo, float spread, float disc, float tol) {
    const float rs = rsqrtf(disc), ia = rcp_(a);
    const float sq = disc * rs;
    const float s0 = (-b - sq) * ia, s1 = (-b + sq) * ia;
    const float x = oo * ia;
    // |error of a root| <= tol * (sqrt(oo / a) + spread / sqrt(disc) + |root|); + tol for the comparison with 1
    const float e = tol * (x * rsqrtf(x + 1e-30f) + spread * rs + fmaxf(fabsf(s0), fabsf(s1)) + 1.0f);
    const bool p0 = s0 > e, n0 = s0 < -e, p1 = s1 > e, n1 = s1 < -e;
    if (n0 && n1) return FRes{F_MISS, 0.f, 0.f};
    if (!(p0 || (n0 && p1))) return FRes{F_UNSURE, 0.f, 0.f};
    const float cand = p0 ? s0 : s1;  // Intersection::hit: the smallest non-negative root
    if (cand < 1.0f - e) return FRes{F_HIT, cand, e};
    if (cand > 1.0f + e) return FRes{F_MISS, 0.f, 0.f};
    return FRes{F_UNSURE, 0.f, 0.f};
}
__device__ __forceinline__ FRes filter_sphere(const Xf& m, V3 o, V3 v, float tol) {
    const float dx = fma_(m.r0.x, v.x, fma_(m.r0.y, v.y, m.r0.z * v.z));
    const float dy = fma_(m.r1.x, v.x, fma_(m.r1.y, v.y, m.r1.z * v.z));
    const float dz = fma_(m.r2.x, v.x, fma_(m.r2.y, v.y, m.r2.z * v.z));
    const float a = fma_(dx, dx, fma_(dy, dy, dz * dz));
    const float b = fma_(dx, o.x, fma_(dy, o.y, dz * o.z));  // half the reference's b
    const float oo = fma_(o.x, o.x, fma_(o.y, o.y, o.z * o.z));
    const float c = oo - 1.0f;
    const float disc = fma_(b, b, -(a * c));                 // a quarter of the reference's discriminant (times |v|^2)
    const float spread = oo + fabsf(c);
    const float td = tol * (a * spread);                     // a * spread >= b^2 + |a c|
    if (disc < -td) return FRes{F_MISS, 0.f, 0.f};
    if (!(disc > td)) return FRes{F_UNSURE, 0.f, 0.f};
    return sphere_roots(a, b, oo, spread, disc, tol);
}

// plane.rs:45-56: only the y row of the inverse is needed.  `len` ~ |light - point| (the reference compares the
// NORMALISED direction's y with EPSILON).
// `upper`: len is only an upper bound of the length (the cell-mask path passes the 1-norm): a direction that is not
// clearly steeper than EPSILON against the bound is undecided rather than a miss.
template <bool NEED_S>
__device__ __forceinline__ FRes filter_plane(float4 r1, float oy, V3 v, float len, bool upper = false) {
    const float px = r1.x * v.x, py = r1.y * v.y, pz = r1.z * v.z;
    const float dy = px + py + pz;
    const float edy = kTolP * (fabsf(px) + fabsf(py) + fabsf(pz));
    const float mag = fabsf(dy), thr = kAcne * len;
    if (!upper && mag + edy < thr * (1.0f - kTolP)) return FRes{F_MISS, 0.f, 0.f};  // plane.rs:49
    if (!(mag - edy > thr * (1.0f + kTolP)) || oy == 0.0f) return FRes{F_UNSURE, 0.f, 0.f};
    if ((oy < 0.0f) == (dy < 0.0f)) return FRes{F_MISS, 0.f, 0.f};  // t = -oy / dy < 0
    const float aoy = fabsf(oy);
    if (aoy < (mag - edy) * (1.0f - kTolP)) {
        if (!NEED_S) return FRes{F_HIT, 0.f, 0.f};
        const float im = rcp_(mag), s = aoy * im;
        return FRes{F_HIT, s, s * (edy * im + 4.0f * kTolP)};
    }
    if (aoy > (mag + edy) * (1.0f + kTolP)) return FRes{F_MISS, 0.f, 0.f};
    return FRes{F_UNSURE, 0.f, 0.f};
}

// cube.rs:55-63 + 90-129 for a cube whose inverse has a diagonal 3x3 part (checked at commit): every direction
// component is ONE product, so each slab distance differs from the reference's by a few ulp, never by cancellation.
__device__ __forceinline__ FRes filter_cube(const Xf& m, V3 o, V3 v) {
    const float dx = m.r0.x * v.x, dy = m.r1.y * v.y, dz = m.r2.z * v.z;
    if (dx == 0.0f || dy == 0.0f || dz == 0.0f) return FRes{F_UNSURE, 0.f, 0.f};
    const float ix = rcp_(dx), iy = rcp_(dy), iz = rcp_(dz);
    float p = (-1.0f - o.x) * ix, q = (1.0f - o.x) * ix;
    float lo = fminf(p, q), hi = fmaxf(p, q);
    p = (-1.0f - o.y) * iy, q = (1.0f - o.y) * iy;
    lo = fmaxf(lo, fminf(p, q)), hi = fminf(hi, fmaxf(p, q));
    p = (-1.0f - o.z) * iz, q = (1.0f - o.z) * iz;
    lo = fmaxf(lo, fminf(p, q)), hi = fminf(hi, fmaxf(p, q));
    const float e = kTolP * (fabsf(lo) + fabsf(hi) + 1.0f);
    const float g = hi - fmaxf(lo, 0.0f);
    if (g < -e) return FRes{F_MISS, 0.f, 0.f};
    if (!(g > e) || !(fabsf(lo) > e)) return FRes{F_UNSURE, 0.f, 0.f};
    const float cand = lo > 0.0f ? lo : hi;
    if (cand < 1.0f - e) return FRes{F_HIT, cand, e};
    if (cand > 1.0f + e) return FRes{F_MISS, 0.f, 0.f};
    return FRes{F_UNSURE, 0.f, 0.f};
}

// The items [begin, ends.z) of the small-scene table against the segment.  MODE 0: casters, any hit decides
// (every object casts); MODE 1: casters, keep the nearest hit (s_c, e_c); MODE 2: non-casters against the nearest
// caster hit.  Returns F_UNSURE as soon as some test is undecided; otherwise F_HIT / F_MISS, meaning
//   MODE 0/1: some / no caster is hit in [0, 1);  MODE 2: F_HIT = a non-caster is clearly nearer than every caster.
template <bool STATS, int MODE>
__device__ __forceinline__ int filter_scan(const Env& E, bool cached, int begin, int4 ends, V3 p, V3 v, float len, float& s_c,
                                           float& e_c, Ctr<STATS>& k) {
    const float4* tab = small_tab();
    const float tol = E.SS.tol_sphere;
    int result = F_MISS;
    auto take = [&](const FRes& r) -> bool {  // true: the scan is decided
        if (r.code == F_UNSURE) {
            result = F_UNSURE;
            return true;
        }
        if (r.code == F_MISS) return false;
        if (MODE == 0) {
            result = F_HIT;
            return true;
        }
        if (MODE == 1) {
            result = F_HIT;
            e_c = r.s < s_c ? r.e : e_c;
            s_c = fminf(s_c, r.s);
            return false;
        }
        if (r.s + r.e < s_c - e_c) {  // clearly in front of the nearest caster: the point is lit (world.rs:113-118)
            result = F_HIT;
            return true;
        }
        if (r.s - r.e > s_c + e_c) return false;  // clearly behind it
        result = F_UNSURE;
        return true;
    };
    int i = begin;
    for (; i < ends.x; i++) {
        Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
        k.xform();
        k.prim(T_SPHERE);
        if (take(filter_sphere(m, small_origin(cached, i, m, p), v, tol))) return result;
    }
    for (; i < ends.y; i++) {
        float4 r1 = tab[i * kSmallStride + 2];
        float oy;
        if (cached && i < kOrgCache)
            oy = small_org()[(i * 3 + 1) * 128];
        else
            oy = r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w;
        k.xform();
        k.prim(T_PLANE);
        if (take(filter_plane<MODE != 0>(r1, oy, v, len))) return result;
    }
    for (; i < ends.z; i++) {
        Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
        k.xform();
        k.prim(T_CUBE);
        if (take(filter_cube(m, small_origin(cached, i, m, p), v))) return result;
    }
    return result;
}

// 0: lit, 1: shadowed, 2: undecided (run the exact test)
template <bool STATS>
__device__ __forceinline__ int shadow_filter(const Env& E, bool cached, V3 light_position, V3 p, Ctr<STATS>& k) {
    const SmallScene& SS = E.SS;
    const V3 v = light_position - p;
    const float vv = fma_(v.x, v.x, fma_(v.y, v.y, v.z * v.z));
    const float len = vv * rsqrtf(vv);
    float s_c = kInfF, e_c = 0.0f;
    if (E.S.all_cast_shadow) return filter_scan<STATS, 0>(E, cached, 0, SS.caster_end, p, v, len, s_c, e_c, k);
    int r = filter_scan<STATS, 1>(E, cached, 0, SS.caster_end, p, v, len, s_c, e_c, k);
    if (r != F_HIT) return r == F_MISS ? 0 : 2;
    r = filter_scan<STATS, 2>(E, cached, SS.caster_end.w, SS.other_end, p, v, len, s_c, e_c, k);
    return r == F_UNSURE ? 2 : (r == F_HIT ? 0 : 1);
}

// World::is_shadowed (world.rs:104-119): nearest hit on the point->light ray; shadowed iff that object
// casts a shadow and is nearer than the light (Q9).
//
// Small scenes test the shadow casters first: with no caster in [0, distance) the answer is "lit" whatever
// the non-casting objects do, and otherwise only a non-casting object NEARER than the nearest caster (same
// (t, depth-first order) comparison as Intersection::hit) can un-shadow the point.  Same predicate as the
// reference's, evaluated with fewer intersection tests.
template <bool STATS>
__device__ __forceinline__ bool shadow_exact_small(const Env& E, bool cached, V3 light_position, V3 p, Ctr<STATS>& k) {
    const DevScene& S = E.S;
    const SmallScene& SS = E.SS;
    V3 v = light_position - p;
    float distance = magnitude(v);
    V3 direction = mk(v.x / distance, v.y / distance, v.z / distance);
    // order -1: a hit AT the light distance is never accepted (`<`, world.rs:116)
    Hit best{distance, -1, -1};
    if (!SS.two_pass_shadows || SS.has_cull_chain) {  // a CSG root or a cull chain: plain nearest-hit search
        for (int i = 0; i < SS.n; i++) test_small<STATS>(E, i, cached, p, direction, best, k);
        return best.pos >= 0 && ((__ldg(&S.head[best.pos]).x >> 4) & kFlagCastsShadow);
    }
    if (S.all_cast_shadow)  // every object casts: any hit in [0, distance) shadows the point
        return scan_small<STATS, true>(E, cached, 0, SS.caster_end, p, direction, best, k);
    scan_small<STATS, false>(E, cached, 0, SS.caster_end, p, direction, best, k);
    if (best.pos < 0) return false;
    const int caster = best.pos;
    scan_small<STATS, false>(E, cached, SS.caster_end.w, SS.other_end, p, direction, best, k);
    return best.pos == caster;
}
// One shadow ray of a small scene, out of line (ONE copy of the filter and of the exact test in the kernel: the
// kernel's instruction footprint, not its arithmetic, limits the issue rate).  skip_filter: the caller already
// knows the filter cannot decide this ray.
template <bool STATS>
__device__ __noinline__ bool shadow_query_small(const Env& E, bool cached, bool skip_filter, V3 light_position, V3 p, Ctr<STATS>& k) {
    if (E.SS.filter_ok) {
        const int f = skip_filter ? 2 : shadow_filter<STATS>(E, cached, light_position, p, k);
        if (f != 2) return f == 1;
        k.refiltered();
    }
    return shadow_exact_small<STATS>(E, cached, light_position, p, k);
}
template <bool STATS, bool SMALL, bool CACHED>
__device__ __forceinline__ bool is_shadowed(const Env& E, V3 light_position, V3 p, Rays& r, Ctr<STATS>& k) {
    r.shadow++;
    if (SMALL) return shadow_query_small<STATS>(E, CACHED, false, light_position, p, k);
    const DevScene& S = E.S;
    V3 v = light_position - p;
    float distance = magnitude(v);
    V3 direction = mk(v.x / distance, v.y / distance, v.z / distance);
    Hit best{distance, -1, -1};  // order -1: a hit AT the light distance is never accepted (`<`, world.rs:116)
    if (S.all_cast_shadow) {
        nearest_hit<STATS, true>(S, p, direction, best, k);
        return best.pos >= 0;
    }
    nearest_hit<STATS, false>(S, p, direction, best, k);
    return best.pos >= 0 && ((__ldg(&S.head[best.pos]).x >> 4) & kFlagCastsShadow);
}

// RectangleLight::intensity_at (rectangle_light.rs:76-88) for filter_ok scenes with a table-mode light: the light
// samples are the same for every shade (staged in shared memory), so the loops are turned inside out — primitive
// outside, light cell inside — and the per-shade part of every test (object-space origin, |o|^2, ...) leaves the
// cell loop.  Pass 1 runs the shadow filter of every CASTER against up to 32 cells at a time and keeps two bit
// masks: cells where some caster is clearly hit, cells where some test was undecided.  Cells in neither mask are
// lit.  Pass 2 revisits the others one by one: hit cells need the nearest-caster / non-caster comparison
// (shadow_filter) when the scene has non-casting objects, undecided cells run the exact test.
// Distances from p to a bounding ball {centre, radius}: no point of the ball is farther than ball_reach, none is
// nearer than ball_gap (negative inside).  Approximate square roots: the callers compare with a 0.1 % margin.
__device__ __forceinline__ float ball_reach(float4 ball, V3 p) {
    const float dx = p.x - ball.x, dy = p.y - ball.y, dz = p.z - ball.z;
    const float ww = fma_(dx, dx, fma_(dy, dy, dz * dz));
    return ww * rsqrtf(ww + 1e-30f) + ball.w;
}
__device__ __forceinline__ float ball_gap(float4 ball, V3 p) {
    const float dx = p.x - ball.x, dy = p.y - ball.y, dz = p.z - ball.z;
    const float ww = fma_(dx, dx, fma_(dy, dy, dz * dz));
    return ww * rsqrtf(ww + 1e-30f) - ball.w;
}

// filter_plane with the per-(plane, cell) constants of plane_cell_constants: q = {r1.L, tol*|r1||L|, eps'*|L|_1},
// rp = r1.p, erp = tol * sum|r1_k p_k|, p1 = eps' * |p|_1 (per shade).  |L - p| <= |L|_1 + |p|_1 stands in for the length
// in the reference's `direction.y.abs() < EPSILON` test (plane.rs:49), so a direction that is not clearly steeper than
// that is undecided rather than a miss.
__device__ __forceinline__ int filter_plane_cell(float4 q, float oy, float rp, float erp, float p1) {
    const float dy = q.x - rp;
    const float edy = q.y + erp;
    const float mag = fabsf(dy), lo = mag - edy;
    if (!(lo > q.z + p1) || oy == 0.0f) return F_UNSURE;
    if ((oy < 0.0f) == (dy < 0.0f)) return F_MISS;  // t = -oy / dy < 0
    const float aoy = fabsf(oy);
    if (aoy < lo * (1.0f - kTolP)) return F_HIT;
    if (aoy > (mag + edy) * (1.0f + kTolP)) return F_MISS;
    return F_UNSURE;
}

// Bundle reject: every shadow segment of a shade runs from p to a light sample inside the ball (Lc, Rl), so all of
// them lie in the cone-like solid { x : |x - (p + s (Lc - p))| <= s Rl, 0 <= s <= 1 }.  A primitive whose bounding ball
// (C, R) stays outside it — f(s) = |w + s u|^2 - (R + s Rl)^2 > 0 on [0, 1], w = p - C, u = Lc - p — cannot be hit by
// any of them.  R is padded by 0.1 % plus 2^-17 * |w|^2 / R: beyond that clearance neither this filter nor the
// reference's f32 discriminant (whose rounding error grows with |w|^2 / R^2) can report a hit.
__device__ __forceinline__ bool bundle_misses(float4 ball, float pad_over_r, float4 light_ball, V3 p) {
    const V3 w = mk(p.x - ball.x, p.y - ball.y, p.z - ball.z);
    const V3 u = mk(light_ball.x - p.x, light_ball.y - p.y, light_ball.z - p.z);
    const float ww = fma_(w.x, w.x, fma_(w.y, w.y, w.z * w.z));
    const float R = fma_(pad_over_r, ww, ball.w), Rl = light_ball.w;
    const float A = fma_(u.x, u.x, fma_(u.y, u.y, u.z * u.z)) - Rl * Rl;
    const float B = fma_(w.x, u.x, fma_(w.y, u.y, w.z * u.z)) - R * Rl;
    const float C = ww - R * R;
    if (!(C > 0.0f) || !(A + 2.0f * B + C > 0.0f)) return false;  // an end of the bundle touches the ball
    if (!(A > 0.0f)) return A <= 0.0f;      // concave or linear: the minimum over [0, 1] is at an end (NaN: no reject)
    if (B >= 0.0f || -B >= A) return true;  // convex, vertex outside (0, 1)
    return C * A > B * B * 1.0001f;
}

// TABLE: the light samples are the staged table.  Otherwise (jitter `None`, rectangle_light.rs:46: the counter-based
// generator) a chunk's sample points are drawn first — two jitter values per cell in the reference's order
// `for v { for u { j_u, j_v } }`, point_on_light's arithmetic (rectangle_light.rs:60-66) — into a per-thread array,
// and the same loops read them from there.
template <bool STATS, bool TABLE>
__device__ __forceinline__ float intensity_cells(const Env& E, V3 p, unsigned pixel, unsigned path, Rays& r, Ctr<STATS>& k) {
    const DevScene& S = E.S;
    const SmallScene& SS = E.SS;
    const float4* tab = small_tab();
    const float4* table = small_samples();
    const int cells = S.cells;
    const float tol = SS.tol_sphere;
    const int4 ends = SS.caster_end;
    r.shadow += cells;
    int lit = 0;
    float4 drawn[TABLE ? 1 : 32];
    for (int c0 = 0; c0 < cells; c0 += 32) {
        const int nc = min(32, cells - c0);
        const unsigned full = nc == 32 ? 0xffffffffu : ((1u << nc) - 1u);
        if (!TABLE) {
            const V3 corner = ld3(S.corner), u_vec = ld3(S.u_vec), v_vec = ld3(S.v_vec);
            for (int j = 0; j < nc; j++) {
                const unsigned cell = (unsigned)(c0 + j);
                const int v = (int)cell / S.u_steps, u = (int)cell - v * S.u_steps;
                const float j1 = jitter_value(S.seed, pixel, path, 2u * cell);
                const float j2 = jitter_value(S.seed, pixel, path, 2u * cell + 1u);
                const V3 lp = corner + u_vec * ((float)u + j1) + v_vec * ((float)v + j2);
                drawn[j] = make_float4(lp.x, lp.y, lp.z, 0.f);
            }
        }
        const float4* smp = TABLE ? table + c0 : drawn;
        unsigned hit = 0u, unsure = 0u;
        float far_hit = 0.0f;  // no caster hit of this chunk is farther from p than this (bounding balls)
        int i = 0;
        for (; i < ends.x && (hit | unsure) != full; i++) {  // caster spheres
            if (bundle_misses(tab[i * kSmallStride + 5], tab[i * kSmallStride + 4].w, SS.light_ball, p)) continue;
            const unsigned hit_before = hit;
            const Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
            const V3 o = xf_point(m, p);  // the reference's object-space origin (shape.rs:60-70), once per shade
            const float oo = fma_(o.x, o.x, fma_(o.y, o.y, o.z * o.z));
            const float c = oo - 1.0f;
            const float spread = oo + fabsf(c);
            const float ts = tol * spread;
#pragma unroll 4
            for (int j = 0; j < nc; j++) {
                const float4 L = smp[j];
                const float vx = L.x - p.x, vy = L.y - p.y, vz = L.z - p.z;
                const float dx = fma_(m.r0.x, vx, fma_(m.r0.y, vy, m.r0.z * vz));
                const float dy = fma_(m.r1.x, vx, fma_(m.r1.y, vy, m.r1.z * vz));
                const float dz = fma_(m.r2.x, vx, fma_(m.r2.y, vy, m.r2.z * vz));
                const float a = fma_(dx, dx, fma_(dy, dy, dz * dz));
                const float b = fma_(dx, o.x, fma_(dy, o.y, dz * o.z));
                const float disc = fma_(b, b, -(a * c));
                const float td = a * ts;
                if (!(disc < -td)) {  // not a clear miss
                    const int code = disc > td ? sphere_roots(a, b, oo, spread, disc, tol).code : F_UNSURE;
                    hit |= (unsigned)(code == F_HIT) << j;
                    unsure |= (unsigned)(code == F_UNSURE) << j;
                }
            }
            for (int j = 0; j < nc; j++) {
                k.xform();
                k.prim(T_SPHERE);
            }
            if (hit != hit_before) far_hit = fmaxf(far_hit, ball_reach(tab[i * kSmallStride + 5], p));
        }
        for (i = ends.x; i < ends.y && (hit | unsure) != full; i++) {  // caster planes
            const unsigned hit_before = hit;
            const float4 r1 = tab[i * kSmallStride + 2];
            const float tx = r1.x * p.x, ty = r1.y * p.y, tz = r1.z * p.z;
            const float rp = tx + ty + tz;
            const float oy = rp + r1.w;  // the reference's object-space origin.y (shape.rs:60-70)
            if (SS.plane_cells) {
                const float4* pc = small_plane_cells() + (i - ends.x) * cells + c0;
                const float erp = kTolP * (fabsf(tx) + fabsf(ty) + fabsf(tz));
                const float p1 = (kAcne * (1.0f + 2.0f * kTolP)) * (fabsf(p.x) + fabsf(p.y) + fabsf(p.z));
                if (i - ends.x < 2) {
                    // the whole bundle at once: with the bounds of the cell constants, every cell's direction is clearly
                    // steep and points away from the plane's side the point is on (a floor under a light above it):
                    // filter_plane_cell would answer F_MISS for each of them
                    const float4 pb = SS.plane_bundle[i - ends.x];
                    const float slack = pb.z + erp + pb.w + p1;
                    if ((oy > 0.0f && (pb.x - rp) > slack) || (oy < 0.0f && (rp - pb.y) > slack)) continue;
                }
#pragma unroll 4
                for (int j = 0; j < nc; j++) {
                    const int code = filter_plane_cell(pc[j], oy, rp, erp, p1);
                    hit |= (unsigned)(code == F_HIT) << j;
                    unsure |= (unsigned)(code == F_UNSURE) << j;
                }
                for (int j = 0; j < nc; j++) {
                    k.xform();
                    k.prim(T_PLANE);
                }
                if (hit != hit_before) far_hit = kInfF;  // a plane has no bounding ball
                continue;
            }
#pragma unroll 4
            for (int j = 0; j < nc; j++) {
                const float4 L = smp[j];
                const V3 v = mk(L.x - p.x, L.y - p.y, L.z - p.z);
                // |v| <= |v|_1: a conservative stand-in for the length in the `direction.y.abs() < EPSILON` test
                const int code = filter_plane<false>(r1, oy, v, fabsf(v.x) + fabsf(v.y) + fabsf(v.z), true).code;
                hit |= (unsigned)(code == F_HIT) << j;
                unsure |= (unsigned)(code == F_UNSURE) << j;
            }
            for (int j = 0; j < nc; j++) {
                k.xform();
                k.prim(T_PLANE);
            }
            if (hit != hit_before) far_hit = kInfF;
        }
        for (i = ends.y; i < ends.z && (hit | unsure) != full; i++) {  // caster cubes
            if (bundle_misses(tab[i * kSmallStride + 5], tab[i * kSmallStride + 4].w, SS.light_ball, p)) continue;
            const unsigned hit_before = hit;
            const Xf m{tab[i * kSmallStride + 1], tab[i * kSmallStride + 2], tab[i * kSmallStride + 3]};
            const V3 o = xf_point(m, p);
#pragma unroll 2
            for (int j = 0; j < nc; j++) {
                const float4 L = smp[j];
                const int code = filter_cube(m, o, mk(L.x - p.x, L.y - p.y, L.z - p.z)).code;
                hit |= (unsigned)(code == F_HIT) << j;
                unsure |= (unsigned)(code == F_UNSURE) << j;
            }
            for (int j = 0; j < nc; j++) {
                k.xform();
                k.prim(T_CUBE);
            }
            if (hit != hit_before) far_hit = fmaxf(far_hit, ball_reach(tab[i * kSmallStride + 5], p));
        }
        for (int j = 0; j < nc; j++) k.cell();
        // pass 2.  A non-caster only matters where it is NEARER than the nearest caster hit (world.rs:113-118): when
        // every non-caster's bounding ball begins beyond the reach of every caster that was hit, the hit cells are
        // shadowed as they stand.
        bool hits_final = S.all_cast_shadow != 0;
        if (!hits_final && (hit & ~unsure) != 0u) {
            float near_other = kInfF;
            for (int q = ends.w; q < SS.other_end.z; q++) {
                const float4 ball = tab[q * kSmallStride + 5];
                const bool has_ball = q < SS.other_end.x || q >= SS.other_end.y;  // spheres and cubes; planes have none
                near_other = fminf(near_other, has_ball ? ball_gap(ball, p) : 0.0f);
            }
            hits_final = far_hit * 1.001f < near_other;  // false for NaN
        }
        unsigned todo = hits_final ? unsure : (hit | unsure);
        lit += __popc(full & ~(hit | unsure));
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1u;
            const float4 L = smp[j];
            lit += !shadow_query_small<STATS>(E, false, (unsure >> j) & 1u, mk(L.x, L.y, L.z), p, k);
        }
    }
    return (float)lit / (float)cells;  // `total += 1.0` per lit cell is exact in f32
}

// Light::intensity_at (point_light.rs:28-34, rectangle_light.rs:76-88).  DRAWN: the kernel build for small scenes whose
// area light draws its jitter from the counter-based generator (only that build carries intensity_cells<.., false>).
template <bool STATS, bool SMALL, bool DRAWN>
__device__ __forceinline__ float intensity_at(const Env& E, V3 p, unsigned pixel, unsigned path, Rays& r, Ctr<STATS>& k) {
    const DevScene& S = E.S;
    if (!S.light_is_rect) return is_shadowed<STATS, SMALL, false>(E, ld3(S.light_pos), p, r, k) ? 0.f : 1.f;
    if (SMALL && E.SS.cell_masks)
        return DRAWN ? intensity_cells<STATS, false>(E, p, pixel, path, r, k) : intensity_cells<STATS, true>(E, p, pixel, path, r, k);
    if (SMALL) cache_origins(E, p);
    float total = 0.f;
    int cell = 0;
    for (int v = 0; v < S.v_steps; v++) {
        for (int u = 0; u < S.u_steps; u++, cell++) {
            V3 lp;
            k.cell();
            if (S.jitter_len > 0) {
                float4 s = __ldg(&S.samples[cell]);  // table mode: point_on_light is the same for every shade
                lp = mk(s.x, s.y, s.z);
            } else {
                float j1 = jitter_value(S.seed, pixel, path, 2u * cell);
                float j2 = jitter_value(S.seed, pixel, path, 2u * cell + 1u);
                // rectangle_light.rs:60-66
                lp = ld3(S.corner) + ld3(S.u_vec) * ((float)u + j1) + ld3(S.v_vec) * ((float)v + j2);
            }
            if (!is_shadowed<STATS, SMALL, SMALL>(E, lp, p, r, k)) total += 1.0f;
        }
    }
    return total / (float)S.cells;
}

__device__ __forceinline__ float powi5(float x) {  // llvm.powi with a constant 5: x * (x^2)^2
    float x2 = x * x;
    return x * (x2 * x2);
}

// One pending shade_hit whose children are still being traced (world.rs:62-86).
struct Frame {
    V3 surface;
    V3 refl;          // reflected_color once known
    V3 refr_o, refr_d;
    float reflective, transparency;
    float reflectance;  // Schlick R, or < 0 when the plain sum applies (world.rs:80-85)
    int remaining;
    unsigned path;
    int stage;  // 1: waiting for the reflection subtree, 2: waiting for the refraction subtree
    int has_refr;
};

__device__ __forceinline__ V3 combine(V3 surface, V3 reflected, V3 refracted, float reflectance) {
    if (reflectance >= 0.0f) return surface + reflected * reflectance + refracted * (1.0f - reflectance);
    return surface + reflected + refracted;
}

// World::color_at (world.rs:88-101) with the recursion of reflected_color / refracted_color replaced by an
// explicit stack of at most depth+1 frames, evaluated in the reference's order (surface, then the whole
// reflection subtree, then the whole refraction subtree) and combined bottom-up with the same arithmetic.
// out_t / out_pos (optional) receive the primary hit.
//
// The loop runs one ray per iteration.  With CONVERGE (the host picks that build for scenes with reflective AND
// transparent materials, whose ray trees branch) EVERY lane of the warp calls this (active = false: no ray) and all
// lanes meet at a warp vote at the top: lanes whose tree is finished wait there, the others start their next ray —
// whichever branch produced it (first child, refraction sibling after a finished reflection subtree) — TOGETHER.
// Without the vote the compiler's reconvergence points leave lanes that took different exits of the body running
// their iterations one after the other (c5: 5 of 32 lanes active in the traversal code; 126 -> 64 ms with it).  Scenes
// whose trees are chains run ~5 % faster without it.
template <bool STATS, bool SMALL, bool CONVERGE, bool DRAWN = false>
__device__ __forceinline__ V3 color_at(const Env& E, bool active, V3 ro, V3 rd, int depth, unsigned pixel, Rays& r, Ctr<STATS>& k,
                                       float* out_t, int* out_pos) {
    const DevScene& S = E.S;
    Frame stack[kMaxFrames];
    int sp = 0;
    int remaining = depth;
    unsigned path = 1u;
    bool primary = true, running = active;
    V3 result = mk(0.f, 0.f, 0.f);
    for (;;) {
        if (CONVERGE) {
            if (!__any_sync(0xffffffffu, running)) break;
            if (!running) continue;
        }
        Hit best{kInfF, -1, 0x7fffffff};
        find_hit<STATS, SMALL>(E, ro, rd, best, k);
        if (primary) {
            if (out_t) *out_t = best.pos >= 0 ? best.t : -1.0f;
            if (out_pos) *out_pos = best.pos;
            primary = false;
        }
        V3 c = mk(0.f, 0.f, 0.f);
        if (best.pos >= 0) {
            // ---- precompute_values (world.rs:212-283)
            r.shades++;
            int4 h = __ldg(&S.head[best.pos]);
            int type = h.x & 15;
            const DevMaterial& mat = S.materials[h.x >> 8];
            Xf m = load_xf(S.xform + 3 * (size_t)h.y);
            V3 point = ro + rd * best.t;
            V3 object_point = xf_point(m, point);
            V3 n = norm(xf_normal(m, local_normal(S, type, h.z, object_point)));  // shape.rs:148-154,130-145
            V3 eye = -rd;
            V3 reflectv = reflect(rd, n);
            if (dot(n, eye) < 0.0f) n = -n;
            V3 over_point = point + n * kAcne;
            // ---- shade_hit (world.rs:62-86): light intensity first, then Phong (phong_lighting.rs:12-63)
            float li = intensity_at<STATS, SMALL, DRAWN>(E, over_point, pixel, path, r, k);
            V3 material_color = ld3(mat.color);
            if (mat.pattern >= 0) {
                k.pattern();
                material_color = pattern_color(S, mat.pattern, xf_point(m, over_point));  // pattern.rs:15-19 at over_point (Q4)
            }
            V3 light_rgb = ld3(S.light_rgb);
            V3 effective = material_color * light_rgb;
            V3 ambient = effective * mat.ambient;
            V3 surface = ambient;
            if (li != 0.f) {
                V3 to_light = norm(ld3(S.light_pos) - over_point);
                float lnc = dot(to_light, n);
                V3 diffuse = mk(0.f, 0.f, 0.f), specular = mk(0.f, 0.f, 0.f);
                if (!(lnc < 0.0f)) {
                    diffuse = effective * mat.diffuse * lnc;
                    V3 sr = reflect(-to_light, n);
                    float rec = dot(sr, eye);
                    if (!(rec <= 0.0f)) {
                        // `intensity * specular * factor` (phong_lighting.rs:56-57): with specular == 0 the product
                        // is 0 for every finite factor, so powf is only evaluated when it can matter
                        float factor = (mat.specular == 0.0f && mat.shininess <= 1.0e4f) ? 1.0f : powf(rec, mat.shininess);
                        specular = light_rgb * mat.specular * factor;
                    }
                }
                surface = ambient + (diffuse + specular) * li;
            }
            // ---- children (world.rs:121-162) with the reference's asymmetric guards (Q11)
            bool want_refl = mat.reflective != 0.0f && remaining >= 1;
            bool want_refr = false;
            float reflectance = -1.0f;
            V3 refr_d = mk(0.f, 0.f, 0.f);
            if (mat.transparency != 0.0f) {
                float n1, n2;
                find_containers<STATS>(S, ro, rd, best.pos, n1, n2, k);
                float cos_i = dot(eye, n);
                if (remaining != 0) {
                    float n_ratio = n1 / n2;  // world.rs:196-207
                    float sin2 = (n_ratio * n_ratio) * (1.0f - cos_i * cos_i);
                    if (!(sin2 > 1.0f)) {
                        k.refr_dir();
                        float cos_t = sqrtf(1.0f - sin2);
                        refr_d = n * (n_ratio * cos_i - cos_t) - (eye * n_ratio);
                        want_refr = true;
                    }
                }
                if (mat.reflective > 0.0f && mat.transparency > 0.0f) {  // schlick_reflectance, world.rs:285-303
                    k.schlick();
                    float cosine = cos_i;
                    bool tir = false;
                    if (n1 > n2) {
                        float nn = n1 / n2;
                        float sin2_t = (nn * nn) * (1.0f - cosine * cosine);
                        if (sin2_t > 1.0f)
                            tir = true;
                        else
                            cosine = sqrtf(1.0f - sin2_t);
                    }
                    if (tir) {
                        reflectance = 1.0f;
                    } else {
                        float q = (n1 - n2) / (n1 + n2);
                        float r0 = q * q;
                        reflectance = r0 + (1.0f - r0) * powi5(1.0f - cosine);
                    }
                }
            }
            if ((want_refl || want_refr) && sp < kMaxFrames) {
                Frame& f = stack[sp++];
                f.surface = surface;
                f.refl = mk(0.f, 0.f, 0.f);
                f.refr_o = point - n * kAcne;  // under_point
                f.refr_d = refr_d;
                f.reflective = mat.reflective;
                f.transparency = mat.transparency;
                f.reflectance = reflectance;
                f.remaining = remaining;
                f.path = path;
                f.has_refr = want_refr;
                r.secondary++;
                remaining = remaining - 1;
                if (want_refl) {
                    f.stage = 1;
                    ro = over_point;
                    rd = reflectv;
                    path = path * 3u + 1u;
                } else {
                    f.stage = 2;
                    ro = f.refr_o;
                    rd = refr_d;
                    path = path * 3u + 2u;
                }
                continue;  // to the vote: trace the child
            }
            c = combine(surface, mk(0.f, 0.f, 0.f), mk(0.f, 0.f, 0.f), reflectance);
        }
        // ---- hand the colour to the waiting frames
        for (;;) {
            if (sp == 0) {
                if (!CONVERGE) return c;
                result = c;
                running = false;
                break;
            }
            Frame& f = stack[sp - 1];
            if (f.stage == 1) {
                f.refl = c * f.reflective;  // world.rs:131
                if (f.has_refr) {
                    f.stage = 2;
                    ro = f.refr_o;
                    rd = f.refr_d;
                    remaining = f.remaining - 1;
                    path = f.path * 3u + 2u;
                    r.secondary++;
                    break;
                }
                c = combine(f.surface, f.refl, mk(0.f, 0.f, 0.f), f.reflectance);
            } else {
                c = combine(f.surface, f.refl, c * f.transparency, f.reflectance);  // world.rs:159-160
            }
            sp--;
        }
    }
    return result;
}

// Camera::ray_for_pixel (camera.rs:60-74)
__device__ __forceinline__ void ray_for_pixel(const DevScene& S, int x, int y, V3& o, V3& d) {
    float x_offset = ((float)x + 0.5f) * S.pixel_size;
    float y_offset = ((float)y + 0.5f) * S.pixel_size;
    float world_x = S.half_w - x_offset;
    float world_y = S.half_h - y_offset;
    Xf m{S.cam_inv[0], S.cam_inv[1], S.cam_inv[2]};
    V3 pixel = xf_point(m, mk(world_x, world_y, -1.0f));
    o = xf_point(m, mk(0.f, 0.f, 0.f));
    d = norm(pixel - o);
}

// Canvas::scale_color (canvas.rs:39-43): clamp, then truncate; NaN -> 255 because f32::min returns the
// non-NaN operand (fminf does the same) and `as u8` saturates.
__device__ __forceinline__ unsigned char scale_color(float c) {
    float s = fmaxf(fminf(c * 255.0f, 255.0f), 0.0f);
    return (unsigned char)__float2uint_rz(s);
}

}  // namespace RTC_NS
}  // namespace rtc
