// Checks dev_math.cuh's div3 (shared-divisor IEEE division) against the compiler's own `/` on the device, bit for bit:
// random operands over the whole exponent range, exact zeros, denormals, infinities, NaNs, and the shapes norm() feeds it
// (components of a vector divided by its magnitude).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define RTC_NS chk
#include "../../ray_tracer_challenge_b200/csrc/rtc_types.h"
#include "../../ray_tracer_challenge_b200/csrc/dev_math.cuh"
using namespace rtc::chk;

__device__ unsigned rnd(unsigned long long& s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(s >> 32);
}
__device__ float pick(unsigned long long& s, int mode) {
    unsigned b = rnd(s);
    if (mode == 0) return __uint_as_float(b);                                   // any bit pattern
    if (mode == 1) return __uint_as_float((b & 0x807fffffu) | ((100u + (rnd(s) % 56u)) << 23));  // moderate exponents
    unsigned k = rnd(s) % 16u;
    if (k == 0) return 0.0f;
    if (k == 1) return -0.0f;
    if (k == 2) return __uint_as_float(b & 0x807fffffu);                        // denormal
    return __uint_as_float((b & 0x807fffffu) | ((120u + (rnd(s) % 16u)) << 23));
}
__global__ void check(unsigned long long seed, int iters, unsigned long long* bad, unsigned long long* fast_taken) {
    unsigned long long s = seed + 977ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long nbad = 0;
    for (int it = 0; it < iters; it++) {
        const int mode = it % 4;
        V3 a = mk(pick(s, mode), pick(s, mode), pick(s, mode));
        float m = mode == 3 ? magnitude(a) : fabsf(pick(s, mode));
        V3 got = div3(a, m);
        V3 want = mk(a.x / m, a.y / m, a.z / m);
        const unsigned g[3] = {__float_as_uint(got.x), __float_as_uint(got.y), __float_as_uint(got.z)};
        const unsigned w[3] = {__float_as_uint(want.x), __float_as_uint(want.y), __float_as_uint(want.z)};
        for (int c = 0; c < 3; c++) {
            const bool both_nan = (g[c] & 0x7fffffffu) > 0x7f800000u && (w[c] & 0x7fffffffu) > 0x7f800000u;
            if (g[c] != w[c] && !both_nan) nbad++;
        }
    }
    if (nbad) atomicAdd(bad, nbad);
}
int main() {
    unsigned long long *bad, *fast;
    cudaMallocManaged(&bad, 8);
    cudaMallocManaged(&fast, 8);
    *bad = 0, *fast = 0;
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    check<<<blocks, threads>>>(12345, iters, bad, fast);
    cudaError_t e = cudaDeviceSynchronize();
    printf("div3_check: %s, %llu vectors, %llu mismatching components\n", cudaGetErrorString(e),
           (unsigned long long)blocks * threads * iters, *bad);
    return *bad == 0 && e == cudaSuccess ? 0 : 1;
}
