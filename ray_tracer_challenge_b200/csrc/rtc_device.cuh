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

// The device code, in dependency order (each part opens namespace rtc::RTC_NS itself):
#include "dev_math.cuh"      // vectors, affine transforms, constants, work counters
#include "dev_shapes.cuh"    // local_intersect / local_norm_at of every leaf kind, the reference's slab test
#include "dev_bvh.cuh"       // nearest hit through the BVH, CSG programs, cull chains, n1 / n2 container walk
#include "dev_patterns.cuh"  // patterns, UV maps, image textures, the counter-based jitter generator
#include "dev_small.cuh"     // small-scene path: shared-memory table, typed loops, ray-vs-ball pre-test
#include "dev_shadow.cuh"    // is_shadowed / intensity_at: shadow filter, exact test, cell-mask loops
#include "dev_shade.cuh"     // color_at (bounded reflect / refract stack), ray_for_pixel, scale_color
#if defined(__CUDACC__)
#include "dev_wave.cuh"      // the wavefront renderer of tree scenes: rays as work items, refilled tree walks
#endif
