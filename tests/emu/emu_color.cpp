// emu_color.cpp — TEST INFRASTRUCTURE ONLY (built and loaded by tests/test_device_code_on_host.py, never by the
// product): the WHOLE device code of the render path (rtc_device.cuh: shapes, tree, CSG, patterns, small-scene loops,
// shadow filter and cell masks, color_at) compiled for the HOST behind a shim of the CUDA intrinsics it uses, one
// "thread" per "block".  It runs World::color_at for caller-supplied rays over the arrays rtc::flatten() — the host half
// of rtc_scene_commit — would upload, so kernel logic can be compared with the oracle on a machine without a GPU.
// With -ffp-contract=off the host evaluates the strict build's IEEE expression order; libm is glibc's on both sides.
#include <math.h>
#include <string.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "rtc_internal.h"  // Flattened, RtcScene, rtc::flatten (exported by librtc_b200.so)

#undef __device__
#undef __forceinline__
#undef __noinline__
#undef __shared__
#define __device__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__

namespace {
struct EmuDim {
    unsigned x = 0, y = 0, z = 0;
};
}  // namespace
static EmuDim threadIdx, blockIdx;
static EmuDim blockDim{1, 1, 1}, gridDim{1, 1, 1};

template <typename T>
static inline T __ldg(const T* p) {
    return *p;
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline int __float_as_int(float f) {
    int i;
    memcpy(&i, &f, 4);
    return i;
}
static inline float __int_as_float(int i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}
static inline unsigned __float_as_uint(float f) {
    unsigned i;
    memcpy(&i, &f, 4);
    return i;
}
static inline float __uint_as_float(unsigned i) {
    float f;
    memcpy(&f, &i, 4);
    return f;
}
static inline int __float2int_rz(float f) { return (int)f; }
static inline unsigned __float2uint_rz(float f) { return f > 0.0f ? (unsigned)f : 0u; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
[[noreturn]] static inline void __trap() { abort(); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __any_sync(unsigned, int p) { return p; }
// PixelStream (dev_shade.cuh) is compiled but never run here: one lane per "warp"
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    unsigned old = *p;
    *p += v;
    return old;
}
static inline void __syncthreads() {}

namespace rtc {
namespace emu {
float4 rtc_smem[kSmallSmemBytes / sizeof(float4) + 64];  // the block's dynamic shared memory
}
}  // namespace rtc

#define RTC_NS emu
#include "dev_math.cuh"
#include "dev_shapes.cuh"
#include "dev_bvh.cuh"
#include "dev_patterns.cuh"
#include "dev_small.cuh"
#include "dev_shadow.cuh"
#include "dev_shade.cuh"

namespace {
// upload_replica (rtc_api.cu) with host pointers
void host_scene(const RtcScene* s, const rtc::Flattened& f, rtc::DevScene& d) {
    memset(&d, 0, sizeof(d));
    rtc::rows3(s->cam_inv, d.cam_inv);
    d.half_w = s->half_w, d.half_h = s->half_h, d.pixel_size = s->pixel_size;
    d.width = (int)s->width, d.height = (int)s->height;
    d.light_is_rect = s->light_is_rect;
    memcpy(d.light_pos, s->light_pos, 12);
    memcpy(d.light_rgb, s->light_rgb, 12);
    memcpy(d.corner, s->corner, 12);
    memcpy(d.u_vec, s->u_cell, 12);
    memcpy(d.v_vec, s->v_cell, 12);
    d.u_steps = s->u_steps, d.v_steps = s->v_steps, d.cells = s->u_steps * s->v_steps;
    d.jitter_len = (int)s->jitter.size();
    d.seed = s->seed;
    auto at = [](const auto& v) { return v.empty() ? nullptr : v.data(); };
    d.jitter = at(s->jitter), d.samples = at(f.samples), d.head = at(f.head), d.rec = at(f.rec), d.xform = at(f.xform);
    d.tri = at(f.tri), d.bound = at(f.bound), d.bvh = at(f.bvh), d.linear = at(f.linear), d.nodes = at(f.nodes);
    d.csg_ops = at(f.ops), d.materials = at(f.materials), d.patterns = at(f.patterns), d.uvs = at(f.uvs), d.texels = at(f.texels);
    d.small_image = at(f.small_image);
    d.n_linear = (int)f.linear.size(), d.bvh_root = f.bvh_root, d.n_prims = f.n_pos, d.all_cast_shadow = f.all_cast_shadow;
}
}  // namespace

static unsigned long long g_rays[4];  // primary, secondary, shadow rays and shades of the last emu_color_at call
extern "C" void emu_last_rays(unsigned long long out[4]) { memcpy(out, g_rays, sizeof(g_rays)); }
static int g_converge = 0;
// 1: the converging build of color_at (a warp vote before every ray; here a warp of one lane)
extern "C" void emu_set_converge(int on) { g_converge = on; }

// World::color_at for n rays.  use_small: take the small-scene path when the scene qualifies (as the kernels do);
// use_filter: with the shadow filter and the cell-mask loops where eligible.  out_path: 0 tree / linear path, 1 small.
extern "C" int emu_color_at(RtcScene* s, uint32_t n, const float* origins, const float* directions, int depth, int use_small,
                            int use_filter, float* out_rgb, float* out_t, int32_t* out_path) {
    using namespace rtc;
    Flattened f;
    if (int rc = flatten(s, f)) return rc;
    DevScene S;
    host_scene(s, f, S);
    SmallScene SS = f.small;
    if (!use_small) SS.n = 0;
    if (!use_filter) SS.filter_ok = SS.cell_masks = SS.plane_cells = 0;
    const bool small = SS.n > 0;
    const bool drawn = small && SS.cell_masks && S.jitter_len == 0;
    if (out_path) *out_path = small ? 1 : 0;
    const emu::Env E{S, SS};
    memset(g_rays, 0, sizeof(g_rays));
    for (uint32_t i = 0; i < n; i++) {
        if (small) emu::stage_small_scene(S, SS);  // one thread per block: it stages the whole table
        emu::V3 o = emu::ld3(origins + 3 * (size_t)i), d = emu::ld3(directions + 3 * (size_t)i);
        emu::Ctr<false> k;
        emu::Rays r;
        float t = -1.0f;
        int pos = -1;
        emu::V3 c;
        if (drawn)
            c = g_converge ? emu::color_at<false, true, true, true>(E, true, o, d, depth, i, r, k, &t, &pos)
                           : emu::color_at<false, true, false, true>(E, true, o, d, depth, i, r, k, &t, &pos);
        else if (small)
            c = g_converge ? emu::color_at<false, true, true, false>(E, true, o, d, depth, i, r, k, &t, &pos)
                           : emu::color_at<false, true, false, false>(E, true, o, d, depth, i, r, k, &t, &pos);
        else
            c = g_converge ? emu::color_at<false, false, true, false>(E, true, o, d, depth, i, r, k, &t, &pos)
                           : emu::color_at<false, false, false, false>(E, true, o, d, depth, i, r, k, &t, &pos);
        emu::finish_rays(S, r);
        g_rays[0] += r.primary, g_rays[1] += r.secondary, g_rays[2] += r.shadow, g_rays[3] += r.shades;
        out_rgb[3 * (size_t)i] = c.x, out_rgb[3 * (size_t)i + 1] = c.y, out_rgb[3 * (size_t)i + 2] = c.z;
        if (out_t) out_t[i] = t;
    }
    return 0;
}
