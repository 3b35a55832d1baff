// dev_bvh.cuh — nearest hit through the BVH (while-while), CSG programs, cull chains, the n1 / n2 container walk.
// Part of rtc_device.cuh (include that, not this): compiled once per kernel build inside namespace rtc::RTC_NS.
#pragma once

namespace rtc {
namespace RTC_NS {

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

// Test one stored primitive against the world ray for the nearest-hit search: test_prim, below, does a mesh triangle
// itself and sends the rest here.
// Everything but a mesh triangle: CSG trees and the analytic shapes.  Out of line: the traversal loop then holds only
// the slab tests and the triangle test, which is what keeps it inside the instruction cache (profiles/README.md).
template <bool STATS>
__device__ __noinline__ void test_prim_other(const DevScene& S, int pos, int4 h, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    const float4* rec = S.rec + 4 * (size_t)pos;
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
    k.prim(type);
    Xf m = load_xf(rec + 1);  // the primitive's own inverse transform travels in its record
    k.xform();
    float tn = nearest_t(S, type, h.z, load_bound(S, type, h.z), xf_point(m, o), xf_vec(m, d));
    if (!(tn >= 0.0f)) return;
    if (((h.x >> 4) & kFlagHasParent) && !ancestors_pass(S, __ldg(&S.head[pos + S.n_prims]).x, o, d)) return;
    consider(best, tn, pos, h.w);
}

template <bool STATS>
__device__ __forceinline__ void test_prim(const DevScene& S, int pos, V3 o, V3 d, ObjRay& cache, Hit& best, Ctr<STATS>& k) {
    const float4* rec = S.rec + 4 * (size_t)pos;
    int4 h = __ldg(reinterpret_cast<const int4*>(rec));
    if ((h.x & 15) != T_TRIANGLE) {
        test_prim_other<STATS>(S, pos, h, o, d, best, k);
        return;
    }
    k.prim(T_TRIANGLE);
    // a mesh's triangles share one transform: the object-space ray is kept across primitives (shape.rs:60-70)
    if (h.y != cache.xf_id) {
        Xf m = load_xf(S.xform + 3 * (size_t)h.y);
        cache.o = xf_point(m, o);
        cache.d = xf_vec(m, d);
        cache.xf_id = h.y;
        k.xform();
    }
    float tn = nearest_t(S, T_TRIANGLE, h.z, make_float4(0.f, 0.f, 0.f, 0.f), cache.o, cache.d, rec + 1);
    if (!(tn >= 0.0f)) return;
    if (((h.x >> 4) & kFlagHasParent) && !ancestors_pass(S, __ldg(&S.head[pos + S.n_prims]).x, o, d)) return;
    consider(best, tn, pos, h.w);
}

// Conservative slab test for the acceleration structure (boxes are padded at build time, so this never rejects a
// primitive the reference would have hit; it is not part of the reference's semantics).  The distances are formed as
// fma(plane, 1 / d, -(o / d)) — one FFMA per plane instead of a subtraction and a product, in both kernel builds; the two
// product roundings are at most 2^-22 of the scene's extent in length, which the builder's padding includes
// (rtc_commit.cu).  A direction component below 2^-60 in magnitude is replaced by +-2^-60 first (tree_inverse): its
// reciprocal stays finite, so no inf - inf = NaN can appear — a ray parallel to a slab gets distances of +-1e18 with the
// right signs: inside the slab no constraint, outside a miss (and it is the padded box the ray has to clear).
__device__ __forceinline__ bool slab(V3 noi, V3 inv, float lx, float ly, float lz, float hx, float hy, float hz, float tmax,
                                     float& tnear) {
    float a = fma_(lx, inv.x, noi.x), b = fma_(hx, inv.x, noi.x);
    float lo = fminf(a, b), hi = fmaxf(a, b);
    a = fma_(ly, inv.y, noi.y), b = fma_(hy, inv.y, noi.y);
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    a = fma_(lz, inv.z, noi.z), b = fma_(hz, inv.z, noi.z);
    lo = fmaxf(lo, fminf(a, b));
    hi = fminf(hi, fmaxf(a, b));
    tnear = lo;
    return hi >= fmaxf(lo, 0.0f) && lo <= tmax;
}

__device__ __forceinline__ float tree_inverse(float d) {
    const float tiny = 8.673617379884035e-19f;  // 2^-60
    return 1.0f / (fabsf(d) < tiny ? copysignf(tiny, d) : d);
}

// World::intersect + Intersection::hit (world.rs:52-60, intersection.rs:30-35): nearest t >= 0 with the
// depth-first tie-break, searched through the BVH.  `best.t` on entry is the search limit (exclusive).
// ANY: stop at the first hit (shadow rays when every primitive casts a shadow).
template <bool STATS, bool ANY>
__device__ __noinline__ void nearest_hit(const DevScene& S, V3 o, V3 d, Hit& best, Ctr<STATS>& k) {
    ObjRay cache;
#pragma unroll 1
    for (int i = 0; i < S.n_linear; i++) {
        test_prim<STATS>(S, __ldg(&S.linear[i]), o, d, cache, best, k);
        if (ANY && best.pos >= 0) return;
    }
    if (S.bvh_root < 0) return;
    const V3 inv = mk(tree_inverse(d.x), tree_inverse(d.y), tree_inverse(d.z));
    const V3 noi = mk(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
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
            bool h0 = slab(noi, inv, a.x, a.y, a.z, a.w, b.x, b.y, best.t, t0);
            bool h1 = slab(noi, inv, b.z, b.w, c.x, c.y, c.z, c.w, best.t, t1);
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
// The CSG case of container_prim, out of line (one copy for both of its call sites; rarely taken).
template <bool STATS>
__device__ __noinline__ void container_csg(const DevScene& S, int pc, V3 o, V3 d, int hit_pos, Containers& c, Ctr<STATS>& k) {
    float ht[kCsgHitCap];
    int hp[kCsgHitCap];
    int n = csg_eval<STATS>(S, pc, o, d, ht, hp, k);
    // per leaf: parity and last negative hit among the filtered hits
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        if (!(ht[i] < 0.0f)) break;
        int p = hp[i];
        bool seen = false;
#pragma unroll 1
        for (int j = 0; j < i; j++) seen |= (hp[j] == p);
        if (seen) continue;
        int cnt = 0;
        float last = 0.f;
#pragma unroll 1
        for (int j = i; j < n && ht[j] < 0.0f; j++)
            if (hp[j] == p) {
                cnt++;
                last = ht[j];
            }
        if (cnt & 1) container_add(c, hit_pos, last, p, __ldg(&S.head[p]).w);
    }
}
template <bool STATS>
__device__ __forceinline__ void container_prim(const DevScene& S, int pos, V3 o, V3 d, ObjRay& cache, int hit_pos, Containers& c,
                                               Ctr<STATS>& k) {
    int4 h = __ldg(&S.head[pos]);
    int type = h.x & 15;
    if (type == T_CSG) {
        container_csg<STATS>(S, h.z, o, d, hit_pos, c, k);
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
#pragma unroll 1
    for (int i = 0; i < S.n_linear; i++) container_prim<STATS>(S, __ldg(&S.linear[i]), o, d, cache, hit_pos, c, k);
    if (S.bvh_root >= 0) {
        // walk the backward half-line: the forward half-line of the reversed ray
        const V3 inv = mk(tree_inverse(-d.x), tree_inverse(-d.y), tree_inverse(-d.z));
        const V3 noi = mk(-(o.x * inv.x), -(o.y * inv.y), -(o.z * inv.z));
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
                    h0 = slab(noi, inv, a.x, a.y, a.z, a.w, b.x, b.y, kInfF, t0);
                if (link.z & 2)
                    h1 = o.x >= b.z && o.x <= cc.y && o.y >= b.w && o.y <= cc.z && o.z >= cc.x && o.z <= cc.w;
                else
                    h1 = slab(noi, inv, b.z, b.w, cc.x, cc.y, cc.z, cc.w, kInfF, t1);
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
#pragma unroll 1
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

}  // namespace RTC_NS
}  // namespace rtc
