// dev_patterns.cuh — patterns, UV maps, image textures; the counter-based jitter generator.
// Part of rtc_device.cuh (include that, not this): compiled once per kernel build inside namespace rtc::RTC_NS.
#pragma once

namespace rtc {
namespace RTC_NS {

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
// Counter-based stand-in for thread_rng().sample(OpenClosed01) (rectangle_light.rs:46): identical to the oracle's
// jitter_key / jitter_hash / jitter_open_closed01.  32-bit arithmetic only: one key per intensity_at call (seed, pixel,
// path), then two IMADs and three shift-xors per drawn value.
__device__ __forceinline__ unsigned jitter_mix32(unsigned x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ unsigned jitter_key(unsigned long long seed, unsigned pixel, unsigned path) {
    const unsigned k = jitter_mix32((unsigned)seed ^ pixel);
    return jitter_mix32(k ^ (unsigned)(seed >> 32) ^ (path * 0x9E3779B9u));
}
__device__ __forceinline__ float jitter_value(unsigned key, unsigned index) {
    const unsigned bits = jitter_mix32(key + index * 0x9E3779B9u);
    return (float)((bits >> 8) + 1u) * 5.9604644775390625e-08f;
}

}  // namespace RTC_NS
}  // namespace rtc
