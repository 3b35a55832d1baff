// rtc_lbvh.cu — K3: the tree of a big scene built ON THE DEVICE (linear BVH over Morton codes).
//
// The host builder (rtc_commit.cu: binned SAH) costs 25-30 ms for 10^5 primitives; Camera::render consumes its World
// (camera.rs:76), so a one-shot render pays that for a kernel of well under a millisecond.  This builder takes the
// padded boxes of the bounded items and returns the same two things the host builder produces — the items' leaf order
// and the array of binary nodes (rtc_types.h: DevBvhNode) — in four small launches:
//
//   lbvh_codes    Morton code of every box centre inside the scene's bounds, `bits` bits per axis with
//                 3 * bits + ceil(log2 n) <= kBvhStack - 4: the tree's depth is at most the number of code bits plus the
//                 depth of the balanced position-split subtrees among equal codes, so it always fits the traversal stack
//   lbvh_sort     bitonic sort of (code, item) pairs, padded to a power of two (hand-written: log^2 n passes of n / 2
//                 compare-exchanges; 10^5 items: 153 launches, ~0.5 ms)
//   lbvh_topology Karras 2012: node i's range, split and children from the common prefixes of neighbouring codes
//                 (equal codes are told apart by their sorted position, so coincident centres still give a tree)
//   lbvh_refit    leaves walk up; the second child to arrive at a node writes the node's two child boxes, links
//                 (subtrees of at most `leaf_size` items become one leaf) and "closed primitives only" bits
//   lbvh_depth    longest root-to-leaf path: a tree deeper than the traversal stack allows is REFUSED (the caller falls
//                 back to the host builder, which balances) — never silently truncated
//
// The tree only decides which primitives get the reference's exact test (its boxes are the host builder's padded
// boxes), so any correct tree renders the same pixels; the SAH tree is better (fewer visits per ray), this one is
// there when the build is the frame.  Group boxes and divide() of the reference (group.rs:48-77, 115-133) play no part
// in either: primitives arrive here already flattened in depth-first order.
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "rtc_internal.h"

namespace rtc {
namespace {

struct LbvhBox {
    float lo[3], hi[3];
};

__device__ __forceinline__ unsigned long long spread21(unsigned v) {  // 21 bits -> every third bit
    unsigned long long x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void lbvh_codes(const LbvhBox* boxes, int n, int padded, float3 lo, float3 scale, float qmax, unsigned long long* keys,
                           int* items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= padded) return;
    if (i >= n) {  // padding sorts last
        keys[i] = ~0ull;
        items[i] = -1;
        return;
    }
    const LbvhBox b = boxes[i];
    const float cx = 0.5f * (b.lo[0] + b.hi[0]), cy = 0.5f * (b.lo[1] + b.hi[1]), cz = 0.5f * (b.lo[2] + b.hi[2]);
    const unsigned qx = (unsigned)fminf(fmaxf((cx - lo.x) * scale.x, 0.0f), qmax);
    const unsigned qy = (unsigned)fminf(fmaxf((cy - lo.y) * scale.y, 0.0f), qmax);
    const unsigned qz = (unsigned)fminf(fmaxf((cz - lo.z) * scale.z, 0.0f), qmax);
    keys[i] = spread21(qx) << 2 | spread21(qy) << 1 | spread21(qz);
    items[i] = i;
}

// one pass of the bitonic network: partner = i ^ j, ascending blocks of size k; ties broken by the item index so the
// order is a total one (deterministic whatever the thread scheduling)
__global__ void lbvh_sort(unsigned long long* keys, int* items, int padded, int j, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = i ^ j;
    if (i >= padded || p <= i) return;
    const unsigned long long ka = keys[i], kb = keys[p];
    const int ia = items[i], ib = items[p];
    const bool a_after_b = ka > kb || (ka == kb && (unsigned)ia > (unsigned)ib);
    const bool ascending = (i & k) == 0;
    if (a_after_b == ascending) {
        keys[i] = kb, keys[p] = ka;
        items[i] = ib, items[p] = ia;
    }
}

// The passes whose partner distance j stays inside a block's 2048 elements, in shared memory: for every k in
// [k_first, k_last] (doubling), the steps j = min(k / 2, 1024) .. 1.  With k_first = 2 it sorts every 2048-element
// block outright (66 passes of lbvh_sort in one launch); with k_first = k_last = k > 2048 it finishes the merge of
// size k after lbvh_sort has done its steps j >= 2048.  Same comparisons, same total order, as lbvh_sort.
constexpr int kSortBlock = 2048;
__global__ void __launch_bounds__(kSortBlock / 2) lbvh_sort_shared(unsigned long long* keys, int* items, int k_first, int k_last) {
    __shared__ unsigned long long s_key[kSortBlock];
    __shared__ int s_item[kSortBlock];
    const int base = blockIdx.x * kSortBlock;
    for (int t = threadIdx.x; t < kSortBlock; t += blockDim.x) s_key[t] = keys[base + t], s_item[t] = items[base + t];
    __syncthreads();
    for (int k = k_first; k <= k_last; k <<= 1)
        for (int j = min(k >> 1, kSortBlock / 2); j > 0; j >>= 1) {
            // thread t owns the pair (i, i ^ j) with bit j of i clear
            const int t = threadIdx.x;
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
            const int p = i | j;
            const unsigned long long ka = s_key[i], kb = s_key[p];
            const int ia = s_item[i], ib = s_item[p];
            const bool a_after_b = ka > kb || (ka == kb && (unsigned)ia > (unsigned)ib);
            const bool ascending = ((base + i) & k) == 0;
            if (a_after_b == ascending) {
                s_key[i] = kb, s_key[p] = ka;
                s_item[i] = ib, s_item[p] = ia;
            }
            __syncthreads();
        }
    for (int t = threadIdx.x; t < kSortBlock; t += blockDim.x) keys[base + t] = s_key[t], items[base + t] = s_item[t];
}

// length of the common prefix of the keys at sorted positions i and j (-1 outside the array); equal keys: 64 + the
// common prefix of the positions (Karras 2012, section 4)
__device__ __forceinline__ int prefix(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz(i ^ j);
}

struct LbvhNode {
    int left, right;    // >= 0: internal node; < 0: ~(sorted position of a leaf)
    int first, last;    // sorted positions covered
    int parent;
};

__global__ void lbvh_topology(const unsigned long long* keys, int n, LbvhNode* nodes, int* leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = prefix(keys, n, i, i + 1) - prefix(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = prefix(keys, n, i, i - d);
    int lmax = 2;
    while (prefix(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (prefix(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = prefix(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (prefix(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = first == gamma ? ~gamma : gamma, right = last == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    // a node's own thread writes its children and range; its `parent` field is written by its parent's thread
    nodes[i].left = left, nodes[i].right = right, nodes[i].first = first, nodes[i].last = last;
    if (left >= 0) nodes[left].parent = i; else leaf_parent[~left] = i;
    if (right >= 0) nodes[right].parent = i; else leaf_parent[~right] = i;
    if (i == 0) nodes[0].parent = -1;
}

struct LbvhBounds {
    float lo[3], hi[3];
    int closed;
};

// The link the traversal follows for a child: an internal node, or a leaf code ~((first << 4) | (count - 1)) when the
// child's whole subtree holds at most leaf_size items (they are contiguous in the sorted order).
__device__ __forceinline__ int link_of(const LbvhNode* nodes, int child, int leaf_size) {
    if (child < 0) return ~((~child << 4) | 0);
    const int count = nodes[child].last - nodes[child].first + 1;
    return count <= leaf_size ? ~((nodes[child].first << 4) | (count - 1)) : child;
}

__global__ void lbvh_refit(const LbvhBox* boxes, const unsigned char* closed, const int* items, int n, const LbvhNode* nodes,
                           const int* leaf_parent, int leaf_size, LbvhBounds* bounds, int* arrivals, DevBvhNode* out) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int node = leaf_parent[leaf];
    while (node >= 0) {
        __threadfence();  // this thread's bounds of the level below are visible before it announces itself
        if (atomicAdd(&arrivals[node], 1) == 0) return;  // the first child to arrive leaves the node to the second
        __threadfence();
        const LbvhNode nd = nodes[node];
        LbvhBounds cb[2];
        const int child[2] = {nd.left, nd.right};
        for (int c = 0; c < 2; c++) {
            if (child[c] < 0) {
                const int item = items[~child[c]];
                const LbvhBox b = boxes[item];
                for (int a = 0; a < 3; a++) cb[c].lo[a] = b.lo[a], cb[c].hi[a] = b.hi[a];
                cb[c].closed = closed[item];
            } else {  // written by another thread, maybe on another SM, during this launch: read past the (incoherent) L1
                const float* src = reinterpret_cast<const float*>(&bounds[child[c]]);
                for (int a = 0; a < 3; a++) cb[c].lo[a] = __ldcg(src + a), cb[c].hi[a] = __ldcg(src + 3 + a);
                cb[c].closed = __ldcg(reinterpret_cast<const int*>(src + 6));
            }
        }
        LbvhBounds me;
        for (int a = 0; a < 3; a++) me.lo[a] = fminf(cb[0].lo[a], cb[1].lo[a]), me.hi[a] = fmaxf(cb[0].hi[a], cb[1].hi[a]);
        me.closed = cb[0].closed & cb[1].closed;
        bounds[node] = me;
        DevBvhNode o;
        o.a = make_float4(cb[0].lo[0], cb[0].lo[1], cb[0].lo[2], cb[0].hi[0]);
        o.b = make_float4(cb[0].hi[1], cb[0].hi[2], cb[1].lo[0], cb[1].lo[1]);
        o.c = make_float4(cb[1].lo[2], cb[1].hi[0], cb[1].hi[1], cb[1].hi[2]);
        o.d = make_int4(link_of(nodes, nd.left, leaf_size), link_of(nodes, nd.right, leaf_size), (cb[0].closed ? 1 : 0) | (cb[1].closed ? 2 : 0), 0);
        out[node] = o;
        node = nd.parent;
    }
}

// levels of inner nodes on the way from the root to each leaf that the traversal actually descends through
__global__ void lbvh_depth(const LbvhNode* nodes, const int* leaf_parent, int n, int leaf_size, int* max_depth) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int depth = 0;
    for (int node = leaf_parent[leaf]; node >= 0; node = nodes[node].parent)
        if (nodes[node].last - nodes[node].first + 1 > leaf_size) depth++;
    atomicMax(max_depth, depth);
}

#define LBVH_TRY(expr)                                                         \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            fail(RTC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

// carves 256-byte aligned pieces out of one buffer
struct Carver {
    char* base;
    size_t used = 0;
    template <class T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + used);
        used = (used + std::max<size_t>(count, 1) * sizeof(T) + 255) & ~size_t(255);
        return p;
    }
};

}  // namespace

// TreeBuilderFn (rtc_internal.h).  ctx: an LbvhContext (the current device is the scene's first replica).
// Returns 0 and fills `out`, or non-zero: the caller builds on the host instead (too few items, CUDA error).
int lbvh_build(void* ctx_, const TreeBuildInput& in, TreeBuildOutput& out) {
    const int n = in.n;
    if (n < 2) return 1;
    LbvhContext& ctx = *static_cast<LbvhContext*>(ctx_);
    cudaStream_t stream = ctx.stream;
    const bool timing = getenv("RTC_TIMING") != nullptr;  // tuning aid
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rtc lbvh]   %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    int padded = 1, log2n = 0;
    while (padded < n) padded <<= 1, log2n++;
    // depth <= 3 * bits (code bits) + log2n (position bits among equal codes): keep it inside the traversal stack
    const int bits = std::max(1, std::min(21, (kBvhStack - 4 - log2n) / 3));
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    {
        struct Range {
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        };
        std::vector<Range> part(parallel_chunks((size_t)n, 16384));
        parallel_for((size_t)n, 16384, [&](size_t b, size_t e, int k) {
            Range r;
            for (size_t i = b; i < e; i++)
                for (int a = 0; a < 3; a++) {
                    const float c = 0.5f * (in.boxes[6 * i + a] + in.boxes[6 * i + 3 + a]);
                    r.lo[a] = std::min(r.lo[a], c), r.hi[a] = std::max(r.hi[a], c);
                }
            part[k] = r;
        });
        for (const Range& r : part)
            for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], r.lo[a]), hi[a] = std::max(hi[a], r.hi[a]);
    }
    const float cells = (float)(1u << bits);
    float scale[3];
    for (int a = 0; a < 3; a++) scale[a] = hi[a] > lo[a] ? cells / (hi[a] - lo[a]) : 0.0f;

    // ---- buffers: one device block and one pinned block, kept with the device slot between builds
    static_assert(sizeof(LbvhBox) == 6 * sizeof(float), "boxes arrive as 6 floats per item");
    auto layout = [&](Carver& c, LbvhBox*& boxes, unsigned char*& closed, unsigned long long*& keys, int*& items, int*& leaf_parent,
                      int*& arrivals, int*& depth, LbvhNode*& nodes, LbvhBounds*& bounds, DevBvhNode*& outn) {
        boxes = c.take<LbvhBox>(n), closed = c.take<unsigned char>(n), keys = c.take<unsigned long long>(padded);
        items = c.take<int>(padded), leaf_parent = c.take<int>(n), arrivals = c.take<int>(n), depth = c.take<int>(1);
        nodes = c.take<LbvhNode>(n), bounds = c.take<LbvhBounds>(n), outn = c.take<DevBvhNode>(n);
    };
    LbvhBox* d_boxes;
    unsigned char* d_closed;
    unsigned long long* d_keys;
    int *d_items, *d_leaf_parent, *d_arrivals, *d_depth;
    LbvhNode* d_nodes;
    LbvhBounds* d_bounds;
    DevBvhNode* d_out;
    Carver sizing{nullptr};
    layout(sizing, d_boxes, d_closed, d_keys, d_items, d_leaf_parent, d_arrivals, d_depth, d_nodes, d_bounds, d_out);
    if (sizing.used > *ctx.scratch_bytes) {
        if (*ctx.scratch) cudaFree(*ctx.scratch);
        *ctx.scratch = nullptr, *ctx.scratch_bytes = 0;
        LBVH_TRY(cudaMalloc(ctx.scratch, sizing.used + sizing.used / 4));
        *ctx.scratch_bytes = sizing.used + sizing.used / 4;
    }
    Carver dev{*ctx.scratch};
    layout(dev, d_boxes, d_closed, d_keys, d_items, d_leaf_parent, d_arrivals, d_depth, d_nodes, d_bounds, d_out);
    // pinned staging: [boxes | closed] up, [order | nodes | depth] down
    const size_t up_bytes = ((size_t)n * sizeof(LbvhBox) + n + 255) & ~size_t(255);
    const size_t down_bytes = (size_t)n * sizeof(int) + (size_t)n * sizeof(DevBvhNode) + 256;
    if (up_bytes + down_bytes > *ctx.pinned_bytes) {
        if (*ctx.pinned) cudaFreeHost(*ctx.pinned);
        *ctx.pinned = nullptr, *ctx.pinned_bytes = 0;
        LBVH_TRY(cudaHostAlloc(ctx.pinned, (up_bytes + down_bytes) * 5 / 4, cudaHostAllocDefault));
        *ctx.pinned_bytes = (up_bytes + down_bytes) * 5 / 4;
    }
    char* h_up = *ctx.pinned;
    char* h_down = *ctx.pinned + up_bytes;
    parallel_stream_copy(h_up, in.boxes, (size_t)n * sizeof(LbvhBox));
    parallel_stream_copy(h_up + (size_t)n * sizeof(LbvhBox), in.closed, n);
    lap("centroid bounds + staging");
    LBVH_TRY(cudaMemcpyAsync(d_boxes, h_up, (size_t)n * sizeof(LbvhBox), cudaMemcpyHostToDevice, stream));
    LBVH_TRY(cudaMemcpyAsync(d_closed, h_up + (size_t)n * sizeof(LbvhBox), (size_t)n, cudaMemcpyHostToDevice, stream));
    LBVH_TRY(cudaMemsetAsync(d_arrivals, 0, (size_t)n * sizeof(int), stream));
    LBVH_TRY(cudaMemsetAsync(d_depth, 0, sizeof(int), stream));
    LBVH_TRY(cudaMemsetAsync(d_out, 0, (size_t)n * sizeof(DevBvhNode), stream));
    const int threads = 256;
    lbvh_codes<<<(padded + threads - 1) / threads, threads, 0, stream>>>(d_boxes, n, padded, make_float3(lo[0], lo[1], lo[2]),
                                                                         make_float3(scale[0], scale[1], scale[2]), cells - 1.0f, d_keys,
                                                                         d_items);
    if (padded >= kSortBlock) {  // 2^17 items: 28 launches instead of 153
        lbvh_sort_shared<<<padded / kSortBlock, kSortBlock / 2, 0, stream>>>(d_keys, d_items, 2, kSortBlock);
        for (int k = 2 * kSortBlock; k <= padded; k <<= 1) {
            for (int j = k >> 1; j >= kSortBlock; j >>= 1)
                lbvh_sort<<<(padded + threads - 1) / threads, threads, 0, stream>>>(d_keys, d_items, padded, j, k);
            lbvh_sort_shared<<<padded / kSortBlock, kSortBlock / 2, 0, stream>>>(d_keys, d_items, k, k);
        }
    } else {
        for (int k = 2; k <= padded; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) lbvh_sort<<<(padded + threads - 1) / threads, threads, 0, stream>>>(d_keys, d_items, padded, j, k);
    }
    lbvh_topology<<<(n + threads - 1) / threads, threads, 0, stream>>>(d_keys, n, d_nodes, d_leaf_parent);
    lbvh_refit<<<(n + threads - 1) / threads, threads, 0, stream>>>(d_boxes, d_closed, d_items, n, d_nodes, d_leaf_parent, in.leaf_size,
                                                                   d_bounds, d_arrivals, d_out);
    lbvh_depth<<<(n + threads - 1) / threads, threads, 0, stream>>>(d_nodes, d_leaf_parent, n, in.leaf_size, d_depth);
    LBVH_TRY(cudaGetLastError());
    lap("launches queued");
    int* h_order = reinterpret_cast<int*>(h_down);
    DevBvhNode* h_nodes = reinterpret_cast<DevBvhNode*>(h_down + (((size_t)n * sizeof(int) + 63) & ~size_t(63)));
    int* h_depth = reinterpret_cast<int*>(reinterpret_cast<char*>(h_nodes) + (size_t)n * sizeof(DevBvhNode));
    LBVH_TRY(cudaMemcpyAsync(h_depth, d_depth, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LBVH_TRY(cudaMemcpyAsync(h_order, d_items, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    LBVH_TRY(cudaMemcpyAsync(h_nodes, d_out, (size_t)(n - 1) * sizeof(DevBvhNode), cudaMemcpyDeviceToHost, stream));
    LBVH_TRY(cudaStreamSynchronize(stream));
    lap("device done, tree on the host");
    if (*h_depth > kBvhStack - 2) return 1;  // cannot happen (the bit budget above); refused rather than truncated if it does
    out.depth = *h_depth;
    out.order.resize(n);
    out.nodes.resize(n - 1);
    parallel_copy(out.order.data(), h_order, (size_t)n * sizeof(int));
    parallel_copy(out.nodes.data(), h_nodes, (size_t)(n - 1) * sizeof(DevBvhNode));
    // the whole tree fits one leaf: the caller wraps it (as it does for the host builder)
    out.root = n <= in.leaf_size ? ~((0 << 4) | (n - 1)) : 0;
    lap("copies out of staging");
    return 0;
}

}  // namespace rtc
