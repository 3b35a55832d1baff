// dev_shade.cuh — World::color_at with the bounded reflect / refract stack, Camera::ray_for_pixel, Canvas::scale_color.
// Part of rtc_device.cuh (include that, not this): compiled once per kernel build inside namespace rtc::RTC_NS.
#pragma once

namespace rtc {
namespace RTC_NS {

__device__ __forceinline__ float powi5(float x) {  // llvm.powi with a constant 5: x * (x^2)^2
    float x2 = x * x;
    return x * (x2 * x2);
}

// One pending shade_hit whose children are still being traced (world.rs:62-86).  The reference adds its three terms
// left to right, `(surface + reflected * R) + refracted * (1 - R)` (world.rs:80-85), so a frame holds ONE colour: the
// surface colour while the reflection subtree is traced, then the first partial sum while the refraction subtree is.
// Only the fields a frame's children need are written (a mirror's frame: acc, reflective, reflectance, path, flags —
// the refraction ray and the transparency are stored for frames that refract); `remaining` is not stored at all: the
// frame at stack index i was pushed by a ray with remaining = depth - i.
struct Frame {
    V3 acc;             // stage 1: surface; stage 2: surface + reflected * R (or surface + reflected)
    float reflective;
    float reflectance;  // Schlick R, or < 0 when the plain sum applies (world.rs:80-85)
    unsigned path;
    int flags;          // bit 0: waiting for the refraction subtree (stage 2); bit 1: a refraction ray is pending
    float transparency;
    V3 refr_o, refr_d;
};

// world.rs:80-85, first and second addition
__device__ __forceinline__ V3 add_reflected(V3 surface, V3 reflected, float reflectance) {
    return reflectance >= 0.0f ? surface + reflected * reflectance : surface + reflected;
}
__device__ __forceinline__ V3 add_refracted(V3 partial, V3 refracted, float reflectance) {
    return reflectance >= 0.0f ? partial + refracted * (1.0f - reflectance) : partial + refracted;
}
__device__ __forceinline__ V3 combine(V3 surface, V3 reflected, V3 refracted, float reflectance) {
    return add_refracted(add_reflected(surface, reflected, reflectance), refracted, reflectance);
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
            const int4 h = __ldg(&S.head[best.pos]);
            const DevMaterial& mat = S.materials[h.x >> 8];
            V3 n, over_point, material_color;
            {
                // everything that needs the hit primitive's transform happens here, BEFORE the light loop, so the
                // twelve registers of `m` (and the hit point) are dead while intensity_at runs; the pattern colour
                // (pattern.rs:15-19, evaluated at over_point, Q4) does not depend on the light
                const Xf m = load_xf(S.xform + 3 * (size_t)h.y);
                const V3 point = ro + rd * best.t;
                const V3 object_point = xf_point(m, point);
                n = norm(xf_normal(m, local_normal(S, h.x & 15, h.z, object_point)));  // shape.rs:148-154,130-145
                if (dot(n, -rd) < 0.0f) n = -n;
                over_point = point + n * kAcne;
                material_color = ld3(mat.color);
                if (mat.pattern >= 0) {
                    k.pattern();
                    material_color = pattern_color(S, mat.pattern, xf_point(m, over_point));
                }
            }
            // ---- shade_hit (world.rs:62-86): light intensity first, then Phong (phong_lighting.rs:12-63)
            const float li = intensity_at<STATS, SMALL, DRAWN>(E, over_point, pixel, path, r, k);
            const V3 eye = -rd;
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
                f.reflectance = reflectance;
                f.path = path;
                r.secondary++;
                remaining = remaining - 1;
                V3 under_point = over_point;
                if (want_refr) {
                    under_point = (ro + rd * best.t) - n * kAcne;  // `point` again: the same expression, the same bits
                    f.transparency = mat.transparency;
                    f.refr_o = under_point;
                    f.refr_d = refr_d;
                }
                if (want_refl) {
                    f.acc = surface;
                    f.reflective = mat.reflective;
                    f.flags = want_refr ? 2 : 0;
                    // reflect(rd, n) (world.rs:225) is bit-for-bit the same for n and -n — every product and every sum
                    // changes sign twice — so it is evaluated here from the flipped normal instead of being kept live
                    // across the light loop
                    ro = over_point;
                    rd = reflect(rd, n);
                    path = path * 3u + 1u;
                } else {  // no reflection: reflected_color is black (world.rs:126-128)
                    f.acc = add_reflected(surface, mk(0.f, 0.f, 0.f), reflectance);
                    f.flags = 1;
                    ro = under_point;
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
            const int flags = f.flags;
            const float reflectance = f.reflectance;
            if (!(flags & 1)) {                        // the reflection subtree is done
                const V3 reflected = c * f.reflective;  // world.rs:131
                if (flags & 2) {                        // its refraction sibling is next
                    f.acc = add_reflected(f.acc, reflected, reflectance);
                    f.flags = 1;
                    ro = f.refr_o;
                    rd = f.refr_d;
                    remaining = depth - sp;  // the frame's ray had depth - (sp - 1)
                    path = f.path * 3u + 2u;
                    r.secondary++;
                    break;
                }
                c = combine(f.acc, reflected, mk(0.f, 0.f, 0.f), reflectance);
            } else {
                c = add_refracted(f.acc, c * f.transparency, reflectance);  // world.rs:159-160
            }
            sp--;
        }
    }
    return result;
}

// is_shadowed calls of a thread's shades (see Rays)
__device__ __forceinline__ void finish_rays(const DevScene& S, Rays& r) { r.shadow = r.shades * (unsigned)(S.light_is_rect ? S.cells : 1); }

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
