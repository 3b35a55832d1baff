// rtc_api.cu — the C ABI of include/rtc_b200.h: scene intake, commit (validation, transform
// de-duplication, binned-SAH BVH over the book's bounding boxes, CSG flattening, upload of one replica per
// device) and the render entry points that drive the kernels of rtc_kernels.cu.
//
// There is no CPU rendering path in this library: every pixel comes from the sm_100a kernels, and every
// entry point that needs a device fails with RTC_ERR_NO_DEVICE when there is none.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <future>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rtc_b200.h"
#include "rtc_launch.h"
#include "rtc_types.h"

using namespace rtc;

namespace {

thread_local std::string g_error;
int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            return fail(RTC_ERR_NO_DEVICE, std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
    } while (0)

struct Box {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const Box& o) {
        for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], o.lo[a]), hi[a] = std::max(hi[a], o.hi[a]);
    }
    float area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
    bool finite() const {
        for (int a = 0; a < 3; a++)
            if (!std::isfinite(lo[a]) || !std::isfinite(hi[a]) || lo[a] > hi[a]) return false;
        return true;
    }
};

// Device-side resources that outlive a scene: creating streams / events and allocating the frame and scene
// buffers costs more than rendering a small frame, so they are pooled per device and leased to a scene at
// commit (Camera::render_b200 commits a fresh scene on every call, like the reference's render takes its
// World by value).
struct DeviceSlot {
    int device = -1;
    cudaStream_t stream = nullptr;       // kernels
    cudaStream_t copy_stream = nullptr;  // device-to-host copies, overlapped with the kernels of later slices
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> slice_done;
    float* d_rgb = nullptr;
    unsigned char* d_u8 = nullptr;
    size_t frame_px = 0;
    DevCounters* d_counters = nullptr;
    unsigned* d_tile_cost = nullptr;  // clock cycles per frame tile, recorded by the render that learns the order
    int* d_tile_order = nullptr;      // launch order of this shard's tiles
    int tile_capacity = 0;
    int sm_count = 148;
    char* arena = nullptr;  // scene arrays, one allocation
    size_t arena_bytes = 0;
    char* staging = nullptr;  // pinned host mirror of the arena for one asynchronous upload
    size_t staging_bytes = 0;
    void* d_flush = nullptr;
    size_t flush_bytes = 0;
};

std::mutex g_pool_mutex;
std::vector<DeviceSlot*> g_pool;  // idle slots

int lease_slot(int device, DeviceSlot** out) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i]->device == device) {
                *out = g_pool[i];
                g_pool.erase(g_pool.begin() + i);
                return 0;
            }
    }
    CUDA_TRY(cudaSetDevice(device));
    DeviceSlot* d = new DeviceSlot();
    d->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&d->ev0));
    CUDA_TRY(cudaEventCreate(&d->ev1));
    CUDA_TRY(cudaMalloc(&d->d_counters, sizeof(DevCounters)));
    CUDA_TRY(cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device));
    // the kernels keep their bounce and traversal stacks in local memory
    size_t have = 0;
    cudaDeviceGetLimit(&have, cudaLimitStackSize);
    if (have < 8192) CUDA_TRY(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
    *out = d;
    return 0;
}
void return_slot(DeviceSlot* d) {
    if (!d) return;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool.push_back(d);
}
int ensure_frame(DeviceSlot* d, size_t px) {
    if (px <= d->frame_px) return 0;
    CUDA_TRY(cudaSetDevice(d->device));
    if (d->d_rgb) cudaFree(d->d_rgb);
    if (d->d_u8) cudaFree(d->d_u8);
    d->d_rgb = nullptr, d->d_u8 = nullptr, d->frame_px = 0;
    CUDA_TRY(cudaMalloc(&d->d_rgb, std::max<size_t>(px * 3 * sizeof(float), 16)));
    CUDA_TRY(cudaMalloc(&d->d_u8, std::max<size_t>(px * 3, 16)));
    d->frame_px = px;
    return 0;
}
int ensure_tiles(DeviceSlot* d, int total_tiles) {
    if (total_tiles <= d->tile_capacity) return 0;
    CUDA_TRY(cudaSetDevice(d->device));
    if (d->d_tile_cost) cudaFree(d->d_tile_cost);
    if (d->d_tile_order) cudaFree(d->d_tile_order);
    d->tile_capacity = 0;
    CUDA_TRY(cudaMalloc(&d->d_tile_cost, (size_t)total_tiles * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&d->d_tile_order, (size_t)total_tiles * sizeof(int)));
    d->tile_capacity = total_tiles;
    return 0;
}
int ensure_arena(DeviceSlot* d, size_t bytes) {
    CUDA_TRY(cudaSetDevice(d->device));
    if (bytes > d->arena_bytes) {
        if (d->arena) cudaFree(d->arena);
        d->arena = nullptr, d->arena_bytes = 0;
        size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 16);
        CUDA_TRY(cudaMalloc(&d->arena, cap));
        d->arena_bytes = cap;
    }
    if (bytes > d->staging_bytes) {
        if (d->staging) cudaFreeHost(d->staging);
        d->staging = nullptr, d->staging_bytes = 0;
        size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 16);
        CUDA_TRY(cudaHostAlloc(&d->staging, cap, cudaHostAllocDefault));
        d->staging_bytes = cap;
    }
    return 0;
}

// One committed replica of the scene on one device.
struct Replica {
    DeviceSlot* slot = nullptr;
    DevScene scene{};
    SmallScene small{};
    int cell_masks_eligible = 0, plane_cells_eligible = 0;
    int filter_eligible = 0;  // SmallScene::filter_ok as computed at commit (RTC_OPT_SHADOW_FILTER masks it per render)
    // Longest-first launch order learnt from the previous render of the same shard (see render_impl)
    int order_shard = -1, order_n_shards = -1, order_depth = -1, order_filter = -1;  // what d_tile_order was learnt for
    bool learning = false;  // this render records the tile costs
};

}  // namespace

struct RtcScene {
    bool have_camera = false, have_light = false, committed = false;
    uint32_t width = 0, height = 0;
    float half_w = 0, half_h = 0, pixel_size = 0;
    float cam_inv[16];
    std::vector<RtcPrim> prims;
    std::vector<RtcNode> nodes;
    std::vector<int32_t> refs;
    std::vector<RtcMaterial> materials;
    std::vector<RtcPattern> patterns;
    std::vector<RtcUvPattern> uvs;
    struct Texture {
        uint32_t width, height;
        std::vector<float> rgb;
    };
    std::vector<Texture> textures;
    bool light_is_rect = false;
    float light_pos[3], light_rgb[3], corner[3], u_cell[3], v_cell[3];
    int u_steps = 1, v_steps = 1;
    std::vector<float> jitter;
    uint64_t seed = 0;
    int strict_fp = 1, leaf_size = 0 /* automatic */, bvh_min_prims = kSmallCap + 1;
    int render_slices = 6;  // kernel / copy pipeline depth when rendering into host memory
    int adaptive_order = 1;  // launch a shard's bands longest-first, learnt from the previous render
    int shadow_filter = 1;   // RTC_OPT_SHADOW_FILTER
    int converge = -1;         // color_at warp vote: -1 automatic (branching ray trees), 0 / 1 forced (RTC_CONVERGE)
    bool has_branching_materials = false;  // some material is reflective AND transparent (set at commit)
    int order_max_waves = 24;  // longest-first order only for launches shorter than this many waves of blocks
    std::vector<Replica> replicas;
    std::vector<int> replica_devices;
    std::vector<int> pos_to_prim;  // device position -> API primitive index (-1 for CSG pseudo-primitives)
    // commit statistics
    int n_bvh_nodes = 0, n_linear = 0, n_xforms = 0;
};

namespace {

void release(RtcScene* s) {
    for (Replica& r : s->replicas) {
        if (!r.slot) continue;
        cudaSetDevice(r.slot->device);
        cudaStreamSynchronize(r.slot->stream);
        cudaStreamSynchronize(r.slot->copy_stream);
        return_slot(r.slot);
        r.slot = nullptr;
    }
    s->replicas.clear();
    s->committed = false;
}

// Packs the scene arrays into one staging buffer (256-byte aligned sections) for a single upload.
struct ArenaWriter {
    size_t bytes = 0;
    struct Part {
        const void* src;
        size_t n, off;
    };
    std::vector<Part> parts;
    template <class T>
    size_t add(const std::vector<T>& v) {
        size_t off = bytes;
        if (!v.empty()) {
            parts.push_back({v.data(), v.size() * sizeof(T), off});
            bytes = (bytes + v.size() * sizeof(T) + 255) & ~size_t(255);
        }
        return off;
    }
};

void rows3(const float m[16], float4 out[3]) {
    for (int r = 0; r < 3; r++) out[r] = make_float4(m[r * 4], m[r * 4 + 1], m[r * 4 + 2], m[r * 4 + 3]);
}

// ---- binned-SAH BVH over the top-level items -----------------------------------------------------------
struct BuildItem {
    Box box;
    float centroid[3];
    int item;     // index into the bounded item list
    bool closed;  // sphere or cube: it has an odd number of hits behind a ray origin only if the origin is inside it
};
struct Builder {
    std::vector<BuildItem>& items;
    std::vector<DevBvhNode>& nodes;
    int leaf_size;
    // returns the link for the range [b, e): >= 0 inner node, < 0 leaf code
    int build(int b, int e, int depth) {
        int n = e - b;
        if (n <= leaf_size || depth > 40) {
            if (n <= 16) return ~((b << 4) | (n - 1));
            // forced leaf too large for the 4-bit count: split in the middle regardless of cost
            return make_inner(b, (b + e) / 2, e, depth);
        }
        Box cb;
        for (int i = b; i < e; i++)
            for (int a = 0; a < 3; a++)
                cb.lo[a] = std::min(cb.lo[a], items[i].centroid[a]), cb.hi[a] = std::max(cb.hi[a], items[i].centroid[a]);
        int axis = 0;
        float ext = cb.hi[0] - cb.lo[0];
        for (int a = 1; a < 3; a++)
            if (cb.hi[a] - cb.lo[a] > ext) axis = a, ext = cb.hi[a] - cb.lo[a];
        int mid = -1;
        if (ext > 0.f) {
            constexpr int kBins = 16;
            Box bins[kBins];
            int counts[kBins] = {0};
            float scale = kBins / ext;
            auto bin_of = [&](const BuildItem& it) {
                int k = (int)((it.centroid[axis] - cb.lo[axis]) * scale);
                return std::min(std::max(k, 0), kBins - 1);
            };
            for (int i = b; i < e; i++) {
                int k = bin_of(items[i]);
                bins[k].grow(items[i].box);
                counts[k]++;
            }
            float right_area[kBins];
            int right_count[kBins];
            Box acc;
            int cnt = 0;
            for (int k = kBins - 1; k > 0; k--) {
                acc.grow(bins[k]);
                cnt += counts[k];
                right_area[k] = acc.area();
                right_count[k] = cnt;
            }
            Box left;
            int lcnt = 0, best_k = -1;
            float best_cost = INFINITY;
            for (int k = 0; k < kBins - 1; k++) {
                left.grow(bins[k]);
                lcnt += counts[k];
                if (lcnt == 0 || right_count[k + 1] == 0) continue;
                float cost = left.area() * lcnt + right_area[k + 1] * right_count[k + 1];
                if (cost < best_cost) best_cost = cost, best_k = k;
            }
            if (best_k >= 0) {
                auto it = std::partition(items.begin() + b, items.begin() + e,
                                         [&](const BuildItem& x) { return bin_of(x) <= best_k; });
                mid = (int)(it - items.begin());
            }
        }
        if (mid <= b || mid >= e) {  // degenerate: equal centroids — split by count
            mid = (b + e) / 2;
            std::nth_element(items.begin() + b, items.begin() + mid, items.begin() + e,
                             [&](const BuildItem& x, const BuildItem& y) { return x.centroid[axis] < y.centroid[axis]; });
        }
        return make_inner(b, mid, e, depth);
    }
    // Subtrees of the top `par_depth` levels are built by separate threads into their own node arrays (the ranges of
    // `items` they partition are disjoint) and appended afterwards with their inner links shifted.
    int par_depth = 0;
    static void append(std::vector<DevBvhNode>& dst, std::vector<DevBvhNode>& sub, int& link) {
        const int off = (int)dst.size();
        for (DevBvhNode& n : sub) {
            if (n.d.x >= 0) n.d.x += off;
            if (n.d.y >= 0) n.d.y += off;
        }
        if (link >= 0) link += off;
        dst.insert(dst.end(), sub.begin(), sub.end());
    }
    int make_inner(int b, int mid, int e, int depth) {
        int idx = (int)nodes.size();
        nodes.emplace_back();
        Box b0, b1;
        for (int i = b; i < mid; i++) b0.grow(items[i].box);
        for (int i = mid; i < e; i++) b1.grow(items[i].box);
        int c0, c1;
        if (depth < par_depth && e - b > 4096) {
            std::vector<DevBvhNode> left_nodes, right_nodes;
            Builder left{items, left_nodes, leaf_size}, right{items, right_nodes, leaf_size};
            left.par_depth = right.par_depth = par_depth;
            auto task = std::async(std::launch::async, [&] { return left.build(b, mid, depth + 1); });
            c1 = right.build(mid, e, depth + 1);
            c0 = task.get();
            append(nodes, left_nodes, c0);
            append(nodes, right_nodes, c1);
        } else {
            c0 = build(b, mid, depth + 1);
            c1 = build(mid, e, depth + 1);
        }
        DevBvhNode& n = nodes[idx];
        n.a = make_float4(b0.lo[0], b0.lo[1], b0.lo[2], b0.hi[0]);
        n.b = make_float4(b0.hi[1], b0.hi[2], b1.lo[0], b1.lo[1]);
        n.c = make_float4(b1.lo[2], b1.hi[0], b1.hi[1], b1.hi[2]);
        // d.z bit i: child i holds closed primitives only (find_containers may cull it with a point-in-box test)
        bool closed0 = true, closed1 = true;
        for (int i = b; i < mid; i++) closed0 = closed0 && items[i].closed;
        for (int i = mid; i < e; i++) closed1 = closed1 && items[i].closed;
        n.d = make_int4(c0, c1, (closed0 ? 1 : 0) | (closed1 ? 2 : 0), 0);
        return idx;
    }
};

// worst-case number of intersections a leaf kind can emit (cone: 2 walls + 2 caps)
int max_hits(int type) {
    switch (type) {
        case RTC_SPHERE: return 2;
        case RTC_PLANE: return 1;
        case RTC_CUBE: return 2;
        case RTC_CYLINDER: return 3;
        case RTC_CONE: return 4;
        default: return 1;
    }
}

struct Flattened {
    std::vector<int4> head;  // 2 * n_pos entries: [pos] main, [n_pos + pos] {cull-chain parent node, api prim, 0, 0}
    std::vector<float4> xform, tri, bound, rec;
    std::vector<DevBvhNode> bvh;
    std::vector<int> linear;
    std::vector<DevNode> nodes;
    std::vector<DevCsgOp> ops;
    std::vector<DevMaterial> materials;
    std::vector<DevPattern> patterns;
    std::vector<DevUvPattern> uvs;
    std::vector<float4> texels;
    int leaf_size = 0;  // the BVH leaf size used
    std::vector<float4> samples;
    SmallScene small{};
    int bvh_root = -1;
    int n_pos = 0;
    int all_cast_shadow = 1;
};

int flatten(RtcScene* s, Flattened& f) {
    const int np = (int)s->prims.size(), nn = (int)s->nodes.size();
    const bool timing = getenv("RTC_TIMING") != nullptr;  // tuning aid: phase times of the host half on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[rtc commit] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    // ---- validate references
    for (int i = 0; i < np; i++) {
        const RtcPrim& p = s->prims[i];
        if (p.type < RTC_SPHERE || p.type > RTC_TRIANGLE) return fail(RTC_ERR_INVALID, "primitive " + std::to_string(i) + ": bad type");
        if (p.material < 0 || p.material >= (int)s->materials.size())
            return fail(RTC_ERR_INVALID, "primitive " + std::to_string(i) + ": bad material index");
        if (p.parent < -1 || p.parent >= nn) return fail(RTC_ERR_INVALID, "primitive " + std::to_string(i) + ": bad parent");
    }
    for (int i = 0; i < nn; i++) {
        const RtcNode& n = s->nodes[i];
        if (n.kind != RTC_NODE_GROUP && n.kind != RTC_NODE_CSG) return fail(RTC_ERR_INVALID, "node: bad kind");
        if (n.parent < -1 || n.parent >= nn || n.parent == i) return fail(RTC_ERR_INVALID, "node: bad parent");
        if (n.child_begin < 0 || n.child_count < 0 || n.child_begin + n.child_count > (int)s->refs.size())
            return fail(RTC_ERR_INVALID, "node: bad child range");
        if (n.kind == RTC_NODE_CSG && (n.child_count != 2 || n.op < 0 || n.op > 2)) return fail(RTC_ERR_INVALID, "csg node: needs 2 children and a valid operator");
        for (int c = 0; c < n.child_count; c++) {
            int r = s->refs[n.child_begin + c];
            if (r >= np || (r < 0 && ~r >= nn)) return fail(RTC_ERR_INVALID, "node: bad child reference");
        }
    }
    for (const RtcMaterial& m : s->materials)
        if (m.pattern < -1 || m.pattern >= (int)s->patterns.size()) return fail(RTC_ERR_INVALID, "material: bad pattern index");
    for (const RtcPattern& p : s->patterns) {
        if (p.kind < RTC_PAT_STRIPES || p.kind > RTC_PAT_CUBIC_MAP) return fail(RTC_ERR_INVALID, "pattern: bad kind");
        int need = p.kind == RTC_PAT_TEXTURE_MAP ? 1 : (p.kind == RTC_PAT_CUBIC_MAP ? 6 : 0);
        for (int i = 0; i < need; i++)
            if (p.uv[i] < 0 || p.uv[i] >= (int)s->uvs.size()) return fail(RTC_ERR_INVALID, "pattern: bad uv pattern index");
    }

    // ---- which CSG (if any) is the outermost CSG ancestor of each node / primitive
    std::vector<int> top_csg_of_node(nn, -1);
    auto resolve = [&](int node) {
        int top = -1, guard = 0;
        for (int a = node; a >= 0; a = s->nodes[a].parent) {
            if (s->nodes[a].kind == RTC_NODE_CSG) top = a;
            if (++guard > nn) return -2;
        }
        return top;
    };
    for (int i = 0; i < nn; i++) {
        top_csg_of_node[i] = resolve(i);
        if (top_csg_of_node[i] == -2) return fail(RTC_ERR_INVALID, "node parents form a cycle");
    }
    std::vector<int> prim_top_csg(np, -1);
    for (int i = 0; i < np; i++) prim_top_csg[i] = s->prims[i].parent >= 0 ? top_csg_of_node[s->prims[i].parent] : -1;

    // ---- top-level items: free primitives and outermost CSG nodes
    struct Item {
        int prim;  // >= 0 primitive, else ~csg node
        Box box;
    };
    std::vector<Item> bounded, unbounded;
    auto add_item = [&](int ref, const float* lo, const float* hi) {
        Item it;
        it.prim = ref;
        for (int a = 0; a < 3; a++) it.box.lo[a] = lo[a], it.box.hi[a] = hi[a];
        (it.box.finite() ? bounded : unbounded).push_back(it);
    };
    for (int i = 0; i < np; i++)
        if (prim_top_csg[i] < 0) add_item(i, s->prims[i].bbox_min, s->prims[i].bbox_max);
    for (int i = 0; i < nn; i++)
        if (s->nodes[i].kind == RTC_NODE_CSG && top_csg_of_node[i] == i) add_item(~i, s->nodes[i].world_bbox_min, s->nodes[i].world_bbox_max);
    if ((int)bounded.size() < s->bvh_min_prims) {  // tiny scene: test everything for every ray, no tree
        unbounded.insert(unbounded.end(), bounded.begin(), bounded.end());
        bounded.clear();
        // group the linear list by (shadow casters first, kind) — the small-scene kernels loop over runs of one
        // kind; depth-first order inside a run (ties are broken by the stored order field, not by position)
        std::stable_sort(unbounded.begin(), unbounded.end(), [&](const Item& a, const Item& b) {
            auto key = [&](const Item& it) {
                int dfs = it.prim >= 0 ? it.prim : np + ~it.prim;
                int caster = it.prim >= 0 ? (s->prims[it.prim].casts_shadow ? 0 : 1) : 0;
                int type = it.prim >= 0 ? s->prims[it.prim].type : 99;
                int bucket = type == RTC_SPHERE ? 0 : type == RTC_PLANE ? 1 : type == RTC_CUBE ? 2 : 3;
                return std::make_tuple(caster, bucket, dfs);
            };
            return key(a) < key(b);
        });
    }

    lap("validate + item lists");
    // ---- BVH
    std::vector<BuildItem> build_items(bounded.size());
    for (size_t i = 0; i < bounded.size(); i++) {
        BuildItem& b = build_items[i];
        b.box = bounded[i].box;
        b.item = (int)i;
        b.closed = bounded[i].prim >= 0 && (s->prims[bounded[i].prim].type == RTC_SPHERE || s->prims[bounded[i].prim].type == RTC_CUBE);
        for (int a = 0; a < 3; a++) {
            // pad: the tree must never reject a hit the reference would report (it is only an accelerator)
            float ext = b.box.hi[a] - b.box.lo[a];
            float pad = 1e-4f * ext + 1e-5f * std::max(std::fabs(b.box.lo[a]), std::fabs(b.box.hi[a])) + 1e-6f;
            b.box.lo[a] -= pad;
            b.box.hi[a] += pad;
            b.centroid[a] = 0.5f * (b.box.lo[a] + b.box.hi[a]);
        }
    }
    if (!build_items.empty()) {
        // Leaf size: a mesh's triangles share one transform (the object-space ray is cached), so a few per leaf cost
        // less than the extra boxes; every other primitive pays its own ray transform, and one per leaf wins
        // (measured on B200: 102 k triangles 0.50 ms at 4 vs 0.58 at 1; 100 k spheres 38.7 ms at 1 vs 51.7 at 4).
        size_t n_triangles = 0;
        for (const Item& it : bounded) n_triangles += it.prim >= 0 && s->prims[it.prim].type == RTC_TRIANGLE;
        int leaf = s->leaf_size > 0 ? s->leaf_size : (2 * n_triangles > bounded.size() ? 4 : 1);
        if (const char* env = getenv("RTC_BVH_LEAF")) leaf = atoi(env);  // tuning aid
        f.leaf_size = std::min(std::max(leaf, 1), 16);
        Builder builder{build_items, f.bvh, f.leaf_size};
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        while ((1u << builder.par_depth) < hw && builder.par_depth < 5) builder.par_depth++;  // up to 32 subtree tasks
        f.bvh.reserve(build_items.size());
        int root = builder.build(0, (int)build_items.size(), 0);
        if (root < 0) {  // a single leaf: wrap it so the traversal always starts at an inner node
            DevBvhNode n;
            Box b;
            for (auto& it : build_items) b.grow(it.box);
            n.a = make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]);
            n.b = make_float4(b.hi[1], b.hi[2], NAN, NAN);  // second child: a NaN box never passes the slab test
            n.c = make_float4(NAN, NAN, NAN, NAN);
            n.d = make_int4(root, root, 0, 0);
            f.bvh.push_back(n);
            root = (int)f.bvh.size() - 1;
        }
        f.bvh_root = root;
    }

    lap("bvh build");
    // ---- device positions: BVH order, then the linear list, then CSG-internal primitives
    std::vector<int> item_refs;
    for (const BuildItem& b : build_items) item_refs.push_back(bounded[b.item].prim);
    const int n_tree = (int)item_refs.size();
    for (const Item& it : unbounded) item_refs.push_back(it.prim);
    const int n_items = (int)item_refs.size();
    std::vector<int> prim_pos(np, -1), csg_pos(nn, -1);
    for (int i = 0; i < n_items; i++) {
        if (item_refs[i] >= 0)
            prim_pos[item_refs[i]] = i;
        else
            csg_pos[~item_refs[i]] = i;
    }
    int next = n_items;
    for (int i = 0; i < np; i++)
        if (prim_top_csg[i] >= 0) prim_pos[i] = next++;
    f.n_pos = next;
    for (int i = n_tree; i < n_items; i++) f.linear.push_back(i);

    // ---- transforms (deduplicated bitwise), triangle and bound tables, heads
    struct XfKey {
        uint32_t w[12];
        bool operator==(const XfKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
    };
    struct XfHash {
        size_t operator()(const XfKey& k) const {  // FNV-1a over the 12 words
            uint64_t h = 1469598103934665603ull;
            for (uint32_t v : k.w) h = (h ^ v) * 1099511628211ull;
            return (size_t)h;
        }
    };
    std::unordered_map<XfKey, int, XfHash> xf_ids;
    f.xform.reserve(3 * (size_t)np);
    XfKey last_key{};
    int last_id = -1;
    // only a mesh's triangles share transforms in practice (and only they profit: the object-space ray is cached by
    // transform id), so every other primitive gets its own slot without a lookup
    auto xform_id = [&](const float m[16], bool dedup) {
        if (!dedup) {
            float4 r[3];
            rows3(m, r);
            f.xform.insert(f.xform.end(), r, r + 3);
            return (int)f.xform.size() / 3 - 1;
        }
        XfKey key;
        memcpy(key.w, m, sizeof(key.w));
        if (last_id >= 0 && key == last_key) return last_id;  // a mesh's triangles share one transform
        last_key = key;
        auto it = xf_ids.find(key);
        if (it != xf_ids.end()) return last_id = it->second;
        int id = (int)f.xform.size() / 3;
        float4 r[3];
        rows3(m, r);
        f.xform.insert(f.xform.end(), r, r + 3);
        xf_ids.emplace(key, id);
        return last_id = id;
    };
    f.head.assign(2 * (size_t)f.n_pos, make_int4(0, 0, 0, 0));
    s->pos_to_prim.assign(f.n_pos, -1);
    f.tri.reserve(3 * (size_t)np);
    for (int i = 0; i < np; i++) {
        const RtcPrim& p = s->prims[i];
        int pos = prim_pos[i];
        int aux = 0;
        if (p.type == RTC_TRIANGLE) {
            aux = (int)f.tri.size() / 3;
            const float* q = p.params;  // p1, e1, e2, normal
            f.tri.push_back(make_float4(q[0], q[1], q[2], q[3]));
            f.tri.push_back(make_float4(q[4], q[5], q[6], q[7]));
            f.tri.push_back(make_float4(q[8], q[9], q[10], q[11]));
        } else if (p.type == RTC_CYLINDER || p.type == RTC_CONE) {
            aux = (int)f.bound.size();
            f.bound.push_back(make_float4(p.params[0], p.params[1], p.params[2] != 0.f ? 1.f : 0.f, 0.f));
        }
        int flags = p.casts_shadow ? kFlagCastsShadow : 0;
        if (!p.casts_shadow) f.all_cast_shadow = 0;
        bool in_linear = pos >= n_tree && pos < n_items;
        if (in_linear && p.parent >= 0) flags |= kFlagHasParent;
        f.head[pos] = make_int4(p.type | (flags << 4) | (p.material << 8), xform_id(p.inv, p.type == RTC_TRIANGLE), aux, i);
        f.head[f.n_pos + pos] = make_int4(prim_top_csg[i] < 0 ? p.parent : -1, i, 0, 0);
        s->pos_to_prim[pos] = i;
    }

    lap("positions, transforms, heads");
    // ---- reference shape-tree nodes
    f.nodes.resize(nn);
    for (int i = 0; i < nn; i++) {
        const RtcNode& n = s->nodes[i];
        DevNode& d = f.nodes[i];
        rows3(n.inv, d.inv);
        for (int a = 0; a < 3; a++) d.bmin[a] = n.bbox_min[a], d.bmax[a] = n.bbox_max[a];
        d.kind = n.kind;
        d.parent = n.parent;
        d.op = n.op;
        d.pad = 0;
    }

    // ---- CSG programs
    struct Emit {
        RtcScene* s;
        Flattened& f;
        std::vector<int>& prim_pos;
        int worst_hits = 0, max_depth = 0;
        int emit(int ref, int csg_depth) {  // returns worst-case hit count of the subtree
            if (ref >= 0) {
                f.ops.push_back(DevCsgOp{OP_PRIM, prim_pos[ref], 0, 0});
                return max_hits(s->prims[ref].type);
            }
            int node = ~ref;
            const RtcNode& n = s->nodes[node];
            int total = 0;
            if (n.kind == RTC_NODE_GROUP) {
                size_t idx = f.ops.size();
                f.ops.push_back(DevCsgOp{OP_GROUP, node, 0, 0});
                for (int c = 0; c < n.child_count; c++) total += emit(s->refs[n.child_begin + c], csg_depth);
                f.ops[idx].skip = (int)f.ops.size();
            } else {
                max_depth = std::max(max_depth, csg_depth + 1);
                size_t idx = f.ops.size();
                f.ops.push_back(DevCsgOp{OP_CSG_ENTER, node, 0, 0});
                total += emit(s->refs[n.child_begin], csg_depth + 1);
                f.ops.push_back(DevCsgOp{OP_CSG_MID, node, 0, 0});
                total += emit(s->refs[n.child_begin + 1], csg_depth + 1);
                f.ops.push_back(DevCsgOp{OP_CSG_EXIT, node, 0, 0});
                f.ops[idx].skip = (int)f.ops.size();
            }
            worst_hits = std::max(worst_hits, total);
            return total;
        }
    } emitter{s, f, prim_pos};
    for (int i = 0; i < nn; i++) {
        if (csg_pos[i] < 0) continue;
        int start = (int)f.ops.size();
        emitter.emit(~i, 0);
        int pos = csg_pos[i];
        // order: depth-first index of the CSG's first leaf (ties are resolved with the leaves' own orders)
        f.head[pos] = make_int4(T_CSG | (kFlagCastsShadow << 4), 0, start, 0);
        f.head[f.n_pos + pos] = make_int4(s->nodes[i].parent, -1, 0, 0);
    }
    if (emitter.worst_hits > kCsgHitCap)
        return fail(RTC_ERR_CAPACITY, "a CSG subtree can produce " + std::to_string(emitter.worst_hits) +
                                          " intersections on one ray; the device hit buffer holds " + std::to_string(kCsgHitCap));
    if (emitter.max_depth > kCsgRayDepth - 1)
        return fail(RTC_ERR_CAPACITY, "CSG nesting depth " + std::to_string(emitter.max_depth) + " exceeds " + std::to_string(kCsgRayDepth - 1));

    lap("nodes + csg programs");
    // ---- traversal records: head + the rows the intersection test needs, one 64 B fetch per primitive
    f.rec.assign(4 * (size_t)f.n_pos, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int pos = 0; pos < f.n_pos; pos++) {
        int4 h = f.head[pos];
        memcpy(&f.rec[4 * (size_t)pos], &h, sizeof(h));
        int type = h.x & 15;
        if (type == T_CSG) continue;
        const float4* src = (type == T_TRIANGLE) ? &f.tri[3 * (size_t)h.z] : &f.xform[3 * (size_t)h.y];
        for (int r = 0; r < 3; r++) f.rec[4 * (size_t)pos + 1 + r] = src[r];
    }

    // ---- shading tables
    s->has_branching_materials = false;
    for (const RtcMaterial& m : s->materials) {
        if (m.reflective != 0.0f && m.transparency != 0.0f) s->has_branching_materials = true;
        DevMaterial d{};
        memcpy(d.color, m.color, sizeof(d.color));
        d.ambient = m.ambient, d.diffuse = m.diffuse, d.specular = m.specular, d.shininess = m.shininess;
        d.reflective = m.reflective, d.transparency = m.transparency, d.refractive_index = m.refractive_index;
        d.pattern = m.pattern;
        f.materials.push_back(d);
    }
    for (const RtcPattern& p : s->patterns) {
        DevPattern d{};
        rows3(p.inv, d.inv);
        memcpy(d.a, p.a, sizeof(d.a));
        memcpy(d.b, p.b, sizeof(d.b));
        d.kind = p.kind, d.mapping = p.mapping;
        memcpy(d.uv, p.uv, sizeof(d.uv));
        f.patterns.push_back(d);
    }
    std::vector<size_t> texel_base;
    for (const RtcScene::Texture& t : s->textures) {
        texel_base.push_back(f.texels.size());
        for (size_t i = 0; i < (size_t)t.width * t.height; i++)
            f.texels.push_back(make_float4(t.rgb[3 * i], t.rgb[3 * i + 1], t.rgb[3 * i + 2], 0.f));
    }
    if (f.texels.size() >= (1u << 30)) return fail(RTC_ERR_CAPACITY, "image textures exceed 2^30 pixels");
    for (const RtcUvPattern& u : s->uvs) {
        DevUvPattern d{};
        d.kind = u.kind;
        memcpy(d.p, u.params, sizeof(d.p));
        if (u.kind == RTC_UV_IMAGE) {  // {first texel, width, height} as integers
            const int t = (int)u.params[0];
            if (t < 0 || t >= (int)s->textures.size() || (float)t != u.params[0]) return fail(RTC_ERR_INVALID, "uv image: bad texture index");
            const int base = (int)texel_base[t], w = (int)s->textures[t].width, h = (int)s->textures[t].height;
            if (w < 1 || h < 1) return fail(RTC_ERR_INVALID, "uv image: empty canvas");
            memcpy(&d.p[0], &base, 4), memcpy(&d.p[1], &w, 4), memcpy(&d.p[2], &h, 4);
        }
        f.uvs.push_back(d);
    }
    // ---- table-mode light samples: point_on_light (rectangle_light.rs:60-66) is the same for every shade
    if (s->light_is_rect && !s->jitter.empty()) {
        size_t cursor = 0, L = s->jitter.size();
        for (int v = 0; v < s->v_steps; v++)
            for (int u = 0; u < s->u_steps; u++) {
                float j1 = s->jitter[cursor % L], j2 = s->jitter[(cursor + 1) % L];
                cursor += 2;
                float su = (float)u + j1, sv = (float)v + j2;
                float p[3];
                for (int a = 0; a < 3; a++) {
                    volatile float t1 = s->u_cell[a] * su;  // volatile: no host-side contraction / reassociation
                    volatile float t2 = s->corner[a] + t1;
                    volatile float t3 = s->v_cell[a] * sv;
                    p[a] = t2 + t3;
                }
                f.samples.push_back(make_float4(p[0], p[1], p[2], 0.f));
            }
    }
    // ---- small-scene table (kernel parameter block): no tree, every item in the linear list, few enough of them
    memset(&f.small, 0, sizeof(f.small));
    if (f.bvh_root < 0 && n_items > 0 && n_items <= kSmallCap) {
        f.small.n = n_items;
        f.small.two_pass_shadows = 1;
        int ends[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n_items; i++) {
            SmallPrim& sp = f.small.p[i];
            int4 h = f.head[i];
            sp.head = make_int4(h.x, f.head[f.n_pos + i].x, h.z, h.w);
            int type = h.x & 15;
            bool caster = type == T_CSG || ((h.x >> 4) & kFlagCastsShadow);
            int bucket = (caster ? 0 : 4) + (type == T_SPHERE ? 0 : type == T_PLANE ? 1 : type == T_CUBE ? 2 : 3);
            for (int b = bucket; b < 8; b++) ends[b] = i + 1;
            if ((h.x >> 4) & kFlagHasParent) f.small.has_cull_chain = 1;
            sp.ball = make_float4(NAN, NAN, NAN, NAN);  // a NaN ball never rejects
            if (type == T_CSG) {
                f.small.two_pass_shadows = 0;
                sp.r0 = sp.r1 = sp.r2 = sp.bound = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            sp.r0 = f.xform[3 * (size_t)h.y], sp.r1 = f.xform[3 * (size_t)h.y + 1], sp.r2 = f.xform[3 * (size_t)h.y + 2];
            sp.bound = (type == T_CYLINDER || type == T_CONE) ? f.bound[h.z] : make_float4(0.f, 0.f, 0.f, 0.f);
            // ---- world-space bounding ball (rtc_device.cuh: ball_missed, bundle_misses).  Sphere / cube: centre =
            // forward transform of the origin, radius = the largest stretch of the forward 3x3 (bounded by
            // sqrt(|T|_1 |T|_inf)), times sqrt(3) for a cube's corners; other bounded shapes: the ball around their
            // world box.  bound.w = 2^-17 cond^2 / radius is the rate at which the tested radius grows with the squared
            // distance of the ray origin: beyond that clearance no f32 intersection test of the reference reports a hit.
            {
                const double m[3][4] = {{sp.r0.x, sp.r0.y, sp.r0.z, sp.r0.w}, {sp.r1.x, sp.r1.y, sp.r1.z, sp.r1.w}, {sp.r2.x, sp.r2.y, sp.r2.z, sp.r2.w}};
                const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                                   m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
                double t[3][3], n1 = 0.0, ninf = 0.0, mi = 0.0;  // t = forward 3x3 = inverse of m's 3x3
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++) {
                        const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
                        t[a][b] = (m[b1][a1] * m[b2][a2] - m[b1][a2] * m[b2][a1]) / det;
                    }
                for (int a = 0; a < 3; a++) {
                    n1 = std::max(n1, std::fabs(t[0][a]) + std::fabs(t[1][a]) + std::fabs(t[2][a]));
                    ninf = std::max(ninf, std::fabs(t[a][0]) + std::fabs(t[a][1]) + std::fabs(t[a][2]));
                    mi = std::max(mi, std::fabs(m[a][0]) + std::fabs(m[a][1]) + std::fabs(m[a][2]));
                }
                const double cond = std::max(1.0, mi * ninf);
                double c[3], radius;
                if (type == T_SPHERE || type == T_CUBE) {
                    radius = std::sqrt(n1 * ninf) * (type == T_CUBE ? std::sqrt(3.0) : 1.0);
                    for (int a = 0; a < 3; a++) c[a] = -(t[a][0] * m[0][3] + t[a][1] * m[1][3] + t[a][2] * m[2][3]);
                } else {
                    const RtcPrim& pr = s->prims[h.w];
                    double d2 = 0.0;
                    for (int a = 0; a < 3; a++) {
                        c[a] = 0.5 * ((double)pr.bbox_min[a] + pr.bbox_max[a]);
                        d2 += 0.25 * ((double)pr.bbox_max[a] - pr.bbox_min[a]) * ((double)pr.bbox_max[a] - pr.bbox_min[a]);
                    }
                    radius = std::sqrt(d2);
                }
                if (std::isfinite(radius) && std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2]) && radius > 0.0 &&
                    std::isfinite(cond)) {
                    sp.ball = make_float4((float)c[0], (float)c[1], (float)c[2], (float)(radius * 1.001 + 1e-6));
                    sp.bound.w = (float)(std::ldexp(1.0, -17) * cond * cond / radius);
                }
            }
        }
        f.small.caster_end = make_int4(ends[0], ends[1], ends[2], ends[3]);
        f.small.other_end = make_int4(ends[4], ends[5], ends[6], ends[7]);
        // ---- shadow-filter eligibility (rtc_device.cuh: shadow_filter): spheres whose transform is well conditioned
        // (the filter's error bound scales with the condition number), planes (term-wise bound: any transform),
        // cubes whose inverse has a diagonal 3x3 part (every direction component is a single product)
        bool ok = f.small.two_pass_shadows && !f.small.has_cull_chain && ends[2] == ends[3] && ends[6] == ends[7];
        double worst = 1.0;
        for (int i = 0; i < n_items && ok; i++) {
            const SmallPrim& sp = f.small.p[i];
            const int type = sp.head.x & 15;
            const double m[3][3] = {{sp.r0.x, sp.r0.y, sp.r0.z}, {sp.r1.x, sp.r1.y, sp.r1.z}, {sp.r2.x, sp.r2.y, sp.r2.z}};
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) ok = ok && std::isfinite(m[a][b]);
            if (type == T_CUBE) {
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++) ok = ok && (a == b ? m[a][b] != 0.0 : m[a][b] == 0.0);
            } else if (type == T_SPHERE) {
                const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                                   m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
                double norm_m = 0.0, norm_i = 0.0;
                for (int a = 0; a < 3; a++) {
                    double row_m = 0.0, row_i = 0.0;
                    for (int b = 0; b < 3; b++) {
                        const int a1 = (a + 1) % 3, a2 = (a + 2) % 3, b1 = (b + 1) % 3, b2 = (b + 2) % 3;
                        row_m += std::fabs(m[a][b]);
                        row_i += std::fabs((m[b1][a1] * m[b2][a2] - m[b1][a2] * m[b2][a1]) / det);  // inverse = adjugate / det
                    }
                    norm_m = std::max(norm_m, row_m), norm_i = std::max(norm_i, row_i);
                }
                const double cond = norm_m * norm_i;
                ok = ok && std::isfinite(cond) && cond <= 64.0;
                if (ok) worst = std::max(worst, cond);
            }
        }
        f.small.filter_ok = ok ? 1 : 0;
        // cell-mask path: a table-mode light whose samples fit the staging area, or the counter-mode generator
        const bool table_light = s->light_is_rect && !f.samples.empty() && f.samples.size() <= (size_t)kSampleCap;
        const bool counter_light = s->light_is_rect && s->jitter.empty() && s->u_steps > 0 && s->v_steps > 0;
        f.small.cell_masks = ok && (table_light || counter_light);
        if (f.small.cell_masks) {
            // the bundle reject (rtc_device.cuh: bundle_misses): a ball around the light samples, and for every sphere /
            // cube its world-space bounding ball — centre = forward transform of the origin, radius = the largest
            // stretch of the forward 3x3 (bounded by sqrt(|T|_1 |T|_inf)), times sqrt(3) for a cube's corners
            // the light's possible sample points: the table's, or (counter mode: jitter in (0, 1]) the whole rectangle
            std::vector<float4> pts = f.samples;
            if (pts.empty())
                for (int cu = 0; cu < 2; cu++)
                    for (int cv = 0; cv < 2; cv++)
                        pts.push_back(make_float4(s->corner[0] + s->u_cell[0] * (cu * s->u_steps) + s->v_cell[0] * (cv * s->v_steps),
                                                  s->corner[1] + s->u_cell[1] * (cu * s->u_steps) + s->v_cell[1] * (cv * s->v_steps),
                                                  s->corner[2] + s->u_cell[2] * (cu * s->u_steps) + s->v_cell[2] * (cv * s->v_steps), 0.f));
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
            for (const float4& q : pts) {
                const double v[3] = {q.x, q.y, q.z};
                for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], v[a]), hi[a] = std::max(hi[a], v[a]);
            }
            const double lc[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
            double rl = 0.0;
            for (const float4& q : pts)
                rl = std::max(rl, std::sqrt((q.x - lc[0]) * (q.x - lc[0]) + (q.y - lc[1]) * (q.y - lc[1]) + (q.z - lc[2]) * (q.z - lc[2])));
            f.small.light_ball = make_float4((float)lc[0], (float)lc[1], (float)lc[2], (float)(rl * 1.001 + 1e-6));
            const int n_planes = ends[1] - ends[0];
            f.small.plane_cells = table_light && n_planes > 0 && (size_t)n_planes * f.samples.size() <= (size_t)kPlaneCellCap;
            // bounds of the per-(plane, cell) constants over all cells (rtc_device.cuh: plane_cell_constants), padded
            // beyond the f32 rounding of the device's own evaluation: lets a shade settle the plane for every cell at once
            for (int q = 0; q < 2; q++) {
                f.small.plane_bundle[q] = make_float4(NAN, NAN, NAN, NAN);  // NaN: the bundle test never decides
                if (!f.small.plane_cells || q >= n_planes) continue;
                const float4 r1 = f.small.p[ends[0] + q].r1;
                double lo = 1e300, hi = -1e300, e_max = 0.0, l1_max = 0.0;
                for (const float4& L : f.samples) {
                    const double px = (double)r1.x * L.x, py = (double)r1.y * L.y, pz = (double)r1.z * L.z;
                    const double a = std::fabs(px) + std::fabs(py) + std::fabs(pz);
                    lo = std::min(lo, px + py + pz - 1e-6 * a), hi = std::max(hi, px + py + pz + 1e-6 * a);
                    e_max = std::max(e_max, 3.814697265625e-06 * a * 1.000001);
                    l1_max = std::max(l1_max, 1.1920929e-3 * (1.0 + 7.7e-6) * (std::fabs(L.x) + std::fabs(L.y) + std::fabs(L.z)) * 1.000001);
                }
                f.small.plane_bundle[q] = make_float4((float)lo, (float)hi, (float)e_max, (float)l1_max);
                // round outwards
                f.small.plane_bundle[q].x = std::nextafter(f.small.plane_bundle[q].x, -INFINITY);
                f.small.plane_bundle[q].y = std::nextafter(f.small.plane_bundle[q].y, INFINITY);
                f.small.plane_bundle[q].z = std::nextafter(f.small.plane_bundle[q].z, INFINITY);
                f.small.plane_bundle[q].w = std::nextafter(f.small.plane_bundle[q].w, INFINITY);
            }
        }
        f.small.tol_sphere = (float)(std::ldexp(1.0, -24) * 64.0 * (worst + 1.0));
    }
    lap("records, tables, small scene");
    s->n_bvh_nodes = (int)f.bvh.size();
    s->n_linear = (int)f.linear.size();
    s->n_xforms = (int)f.xform.size() / 3;
    return 0;
}

int upload_replica(RtcScene* s, const Flattened& f, Replica& r, int device) {
    int rc;
    if ((rc = lease_slot(device, &r.slot))) return rc;
    DeviceSlot* slot = r.slot;
    CUDA_TRY(cudaSetDevice(device));
    DevScene& d = r.scene;
    memset(&d, 0, sizeof(d));
    rows3(s->cam_inv, d.cam_inv);
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
    ArenaWriter w;
    size_t o_jitter = w.add(s->jitter), o_samples = w.add(f.samples), o_head = w.add(f.head), o_rec = w.add(f.rec);
    size_t o_xform = w.add(f.xform), o_tri = w.add(f.tri), o_bound = w.add(f.bound), o_bvh = w.add(f.bvh);
    size_t o_linear = w.add(f.linear), o_nodes = w.add(f.nodes), o_ops = w.add(f.ops), o_mat = w.add(f.materials);
    size_t o_pat = w.add(f.patterns), o_uv = w.add(f.uvs), o_tex = w.add(f.texels);
    if ((rc = ensure_arena(slot, std::max<size_t>(w.bytes, 256)))) return rc;
    for (const auto& p : w.parts) memcpy(slot->staging + p.off, p.src, p.n);
    if (w.bytes) CUDA_TRY(cudaMemcpyAsync(slot->arena, slot->staging, w.bytes, cudaMemcpyHostToDevice, slot->stream));
    auto at = [&](size_t off, bool present) -> const void* { return present ? slot->arena + off : nullptr; };
    d.jitter = (const float*)at(o_jitter, !s->jitter.empty());
    d.samples = (const float4*)at(o_samples, !f.samples.empty());
    d.head = (const int4*)at(o_head, !f.head.empty());
    d.rec = (const float4*)at(o_rec, !f.rec.empty());
    d.xform = (const float4*)at(o_xform, !f.xform.empty());
    d.tri = (const float4*)at(o_tri, !f.tri.empty());
    d.bound = (const float4*)at(o_bound, !f.bound.empty());
    d.bvh = (const DevBvhNode*)at(o_bvh, !f.bvh.empty());
    d.linear = (const int*)at(o_linear, !f.linear.empty());
    d.nodes = (const DevNode*)at(o_nodes, !f.nodes.empty());
    d.csg_ops = (const DevCsgOp*)at(o_ops, !f.ops.empty());
    d.materials = (const DevMaterial*)at(o_mat, !f.materials.empty());
    d.patterns = (const DevPattern*)at(o_pat, !f.patterns.empty());
    d.uvs = (const DevUvPattern*)at(o_uv, !f.uvs.empty());
    d.texels = (const float4*)at(o_tex, !f.texels.empty());
    d.n_linear = (int)f.linear.size();
    d.bvh_root = f.bvh_root;
    d.n_prims = f.n_pos;
    d.all_cast_shadow = f.all_cast_shadow;
    r.small = f.small;
    r.filter_eligible = f.small.filter_ok;
    r.cell_masks_eligible = f.small.cell_masks;
    r.plane_cells_eligible = f.small.plane_cells;
    if ((rc = ensure_frame(slot, (size_t)s->width * s->height))) return rc;
    return 0;
}

double flops_of(const RtcStats& st, uint64_t pixels) {
    const double pf[8] = {30, 2, 29, 45, 50, 46, 0, 0};
    double f = 33.0 * st.xforms + 26.0 * st.node_visits;
    for (int i = 0; i < 8; i++) f += pf[i] * st.prim_tests[i];
    f += 36.0 * st.primary_rays + (85.0 + 75.0 + 9.0) * st.shades + 41.0 * st.patterns;
    f += 13.0 * st.shadow_rays + 14.0 * st.cells + 20.0 * st.schlicks + 25.0 * st.refr_dirs + 3.0 * pixels;
    return f;
}

void add_counters(RtcStats& st, const DevCounters& c) {
    st.primary_rays += c.primary, st.secondary_rays += c.secondary, st.shadow_rays += c.shadow, st.shades += c.shades;
    st.node_visits += c.node_visits;
    for (int i = 0; i < 8; i++) st.prim_tests[i] += c.prim_tests[i];
    st.xforms += c.xforms, st.patterns += c.patterns, st.cells += c.cells, st.schlicks += c.schlicks;
    st.refr_dirs += c.refr_dirs, st.capacity_overflows += c.overflows;
}

// Copy the rows of bands [b0, b1) of this shard's band list from the device frame into the caller's full-size
// canvas (band j of the list is frame band `shard + j * n_shards`).
int copy_bands(DeviceSlot* slot, const RtcScene* s, int shard, int n_shards, int b0, int b1, const void* src, void* dst,
               size_t px_bytes) {
    const size_t row = (size_t)s->width * px_bytes;
    const size_t band = row * kBandRows;
    if (b1 <= b0) return 0;
    const int first_band = shard + b0 * n_shards, last_band = shard + (b1 - 1) * n_shards;
    const int last_rows = std::min<int>(kBandRows, (int)s->height - last_band * kBandRows);
    if (n_shards == 1) {  // contiguous rows
        size_t off = (size_t)first_band * band;
        size_t bytes = (size_t)(b1 - b0 - 1) * band + row * last_rows;
        CUDA_TRY(cudaMemcpyAsync((char*)dst + off, (const char*)src + off, bytes, cudaMemcpyDeviceToHost, slot->copy_stream));
        return 0;
    }
    const int full = (last_rows == kBandRows) ? (b1 - b0) : (b1 - b0 - 1);  // every band but possibly the last is full
    size_t off = (size_t)first_band * band;
    if (full > 0)
        CUDA_TRY(cudaMemcpy2DAsync((char*)dst + off, band * n_shards, (const char*)src + off, band * n_shards, band, full,
                                   cudaMemcpyDeviceToHost, slot->copy_stream));
    if (full < b1 - b0) {
        size_t o2 = (size_t)last_band * band;
        CUDA_TRY(cudaMemcpyAsync((char*)dst + o2, (const char*)src + o2, row * last_rows, cudaMemcpyDeviceToHost,
                                 slot->copy_stream));
    }
    return 0;
}

int render_impl(RtcScene* s, int depth, int shard0, int n_shards_ext, float* rgb, uint8_t* u8, RtcStats* stats, bool detailed) {
    if (!s) return fail(RTC_ERR_INVALID, "null scene");
    if (!s->committed) return fail(RTC_ERR_STATE, "rtc_render before rtc_scene_commit");
    if (depth < 0 || depth > kMaxFrames - 1)
        return fail(RTC_ERR_CAPACITY, "reflection_recursion_depth must be in [0, " + std::to_string(kMaxFrames - 1) + "]");
    auto t0 = std::chrono::steady_clock::now();
    const int ndev = (int)s->replicas.size();
    const bool external = n_shards_ext > 0;
    if (external && ndev != 1) return fail(RTC_ERR_STATE, "rtc_render_shard needs a scene committed on exactly one device");
    if (external && (shard0 < 0 || shard0 >= n_shards_ext)) return fail(RTC_ERR_INVALID, "bad shard index");
    const int n_shards = external ? n_shards_ext : ndev;
    const int total_bands = ((int)s->height + kBandRows - 1) / kBandRows;
    const bool copy_out = rgb || u8;
    RtcStats st;
    memset(&st, 0, sizeof(st));
    st.n_devices = ndev;
    st.detailed = detailed;
    for (int i = 0; i < ndev; i++) {
        Replica& r = s->replicas[i];
        DeviceSlot* slot = r.slot;
        const int shard = external ? shard0 : i;
        r.small.filter_ok = r.filter_eligible && s->shadow_filter;
        r.small.cell_masks = r.cell_masks_eligible && s->shadow_filter;
        r.small.plane_cells = r.plane_cells_eligible && r.small.cell_masks;
        const int nb = shard < total_bands ? (total_bands - shard + n_shards - 1) / n_shards : 0;
        CUDA_TRY(cudaSetDevice(slot->device));
        CUDA_TRY(cudaMemsetAsync(slot->d_counters, 0, sizeof(DevCounters), slot->stream));
        // Longest-processing-time-first: a frame's cost is concentrated in a few tiles (deep reflection trees: one
        // block there runs ~10x longer than the average), and a long block that starts late is the tail of the
        // launch — the term that limits strong scaling over shards (c3, 1/8 of the frame: the SMs were busy for
        // only 53 % of the launch).  The first render of a (shard, depth) configuration runs in natural order and
        // records how many clock cycles every tile's block took; later renders launch the shard's tiles
        // most-expensive-first from a list kept on the device.  The order changes no pixel.  It needs one
        // copy-free launch, so it is used when the frame stays on the device or is copied in one piece — and only
        // for launches of fewer than `order_max_waves` waves of blocks (a whole 4K frame, 73 waves, has no tail to
        // speak of and runs a few percent faster in natural order: neighbouring tiles share their cache lines).
        const int n_slices = copy_out ? std::max(1, std::min(s->render_slices, nb)) : 1;
        const int tiles_x = ((int)s->width + kTileW - 1) / kTileW;
        int rc0;
        if ((rc0 = ensure_tiles(slot, total_bands * tiles_x))) return rc0;
        const long long launch_blocks = (long long)nb * tiles_x;
        const bool order_wanted = s->adaptive_order && n_slices == 1 && nb > 0 &&
                                  launch_blocks < (long long)s->order_max_waves * 5 * slot->sm_count;
        const bool learnt = r.order_shard == shard && r.order_n_shards == n_shards && r.order_depth == depth &&
                            r.order_filter == s->shadow_filter;
        const bool use_order = order_wanted && learnt;
        r.learning = order_wanted && !learnt;
        if (r.learning) CUDA_TRY(cudaMemsetAsync(slot->d_tile_cost, 0, (size_t)total_bands * tiles_x * sizeof(unsigned), slot->stream));
        // With a host destination the frame is rendered in a few slices so that the device-to-host copy of one
        // slice overlaps the kernel of the next (the 4K canvases are 124 MB: ~2.3 ms of PCIe against ~2 ms of
        // kernel); left on the device it is one launch.
        while ((int)slot->slice_done.size() < n_slices) {
            cudaEvent_t e;
            CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            slot->slice_done.push_back(e);
        }
        CUDA_TRY(cudaEventRecord(slot->ev0, slot->stream));
        for (int k = 0; k < n_slices; k++) {
            const int b0 = (int)((int64_t)nb * k / n_slices), b1 = (int)((int64_t)nb * (k + 1) / n_slices);
            if (b1 <= b0) continue;
            DevFrame F{slot->d_rgb, slot->d_u8, shard, n_shards, depth, b1 - b0, b0, use_order ? slot->d_tile_order : nullptr,
                       r.learning ? slot->d_tile_cost : nullptr, s->converge < 0 ? (int)s->has_branching_materials : s->converge};
            if (s->strict_fp)
                strict::launch_render(r.scene, r.small, F, slot->d_counters, detailed, slot->stream);
            else
                fast::launch_render(r.scene, r.small, F, slot->d_counters, detailed, slot->stream);
            CUDA_TRY(cudaGetLastError());
            st.launches++;
            if (copy_out) {
                CUDA_TRY(cudaEventRecord(slot->slice_done[k], slot->stream));
                CUDA_TRY(cudaStreamWaitEvent(slot->copy_stream, slot->slice_done[k], 0));
                int rc;
                if (rgb && (rc = copy_bands(slot, s, shard, n_shards, b0, b1, slot->d_rgb, rgb, 3 * sizeof(float)))) return rc;
                if (u8 && (rc = copy_bands(slot, s, shard, n_shards, b0, b1, slot->d_u8, u8, 3))) return rc;
            }
        }
        CUDA_TRY(cudaEventRecord(slot->ev1, slot->stream));
    }
    for (int i = 0; i < ndev; i++) {
        DeviceSlot* slot = s->replicas[i].slot;
        CUDA_TRY(cudaSetDevice(slot->device));
        CUDA_TRY(cudaStreamSynchronize(slot->stream));
        CUDA_TRY(cudaStreamSynchronize(slot->copy_stream));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, slot->ev0, slot->ev1));
        st.kernel_ms = std::max(st.kernel_ms, (double)ms);
        DevCounters c;
        CUDA_TRY(cudaMemcpy(&c, slot->d_counters, sizeof(c), cudaMemcpyDeviceToHost));
        add_counters(st, c);
        Replica& r = s->replicas[i];
        if (r.learning) {  // sort this shard's tiles by the cycles their blocks took: the launch order from now on
            const int shard = external ? shard0 : i;
            const int tiles_x = ((int)s->width + kTileW - 1) / kTileW;
            std::vector<unsigned> cost((size_t)total_bands * tiles_x);
            CUDA_TRY(cudaMemcpy(cost.data(), slot->d_tile_cost, cost.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
            std::vector<int> order;
            for (int b = shard; b < total_bands; b += n_shards)
                for (int x = 0; x < tiles_x; x++) order.push_back(b * tiles_x + x);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
            if (const char* dump = getenv("RTC_DUMP_TILE_COST")) {  // tuning aid: the learnt costs, in launch order
                if (FILE* fh = fopen(dump, "w")) {
                    for (int id : order) fprintf(fh, "%d %d %u\n", id / tiles_x, id % tiles_x, cost[id]);
                    fclose(fh);
                }
            }
            CUDA_TRY(cudaMemcpy(slot->d_tile_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
            r.order_shard = shard, r.order_n_shards = n_shards, r.order_depth = depth, r.order_filter = s->shadow_filter;
            r.learning = false;
        }
    }
    st.flops = detailed ? flops_of(st, st.primary_rays) : 0.0;
    st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = st;
    if (st.capacity_overflows)
        return fail(RTC_ERR_CAPACITY, "CSG hit buffer overflowed " + std::to_string(st.capacity_overflows) + " times");
    return 0;
}

}  // namespace

extern "C" {

const char* rtc_last_error(void) { return g_error.c_str(); }

int rtc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(RTC_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return n;
}

int rtc_scene_create(RtcScene** out) {
    if (!out) return fail(RTC_ERR_INVALID, "null out pointer");
    *out = new RtcScene();
    if (const char* env = getenv("RTC_ADAPTIVE_ORDER")) (*out)->adaptive_order = atoi(env) != 0;  // tuning aids
    if (const char* env = getenv("RTC_SHADOW_FILTER")) (*out)->shadow_filter = atoi(env) != 0;
    if (const char* env = getenv("RTC_ORDER_MAX_WAVES")) (*out)->order_max_waves = atoi(env);
    if (const char* env = getenv("RTC_CONVERGE")) (*out)->converge = atoi(env);
    return 0;
}
void rtc_scene_destroy(RtcScene* s) {
    if (!s) return;
    release(s);
    delete s;
}

int rtc_set_camera(RtcScene* s, uint32_t w, uint32_t h, float half_w, float half_h, float pixel_size, const float inv[16]) {
    if (!s || !inv) return fail(RTC_ERR_INVALID, "null argument");
    if (w == 0 || h == 0 || (uint64_t)w * h > (1ull << 31) / 3) return fail(RTC_ERR_INVALID, "bad canvas size");
    s->width = w, s->height = h, s->half_w = half_w, s->half_h = half_h, s->pixel_size = pixel_size;
    memcpy(s->cam_inv, inv, sizeof(s->cam_inv));
    s->have_camera = true;
    s->committed = false;
    return 0;
}
int rtc_set_primitives(RtcScene* s, uint32_t n, const RtcPrim* prims) {
    if (!s || (n && !prims)) return fail(RTC_ERR_INVALID, "null argument");
    if (n >= (1u << 27)) return fail(RTC_ERR_CAPACITY, "too many primitives");
    s->prims.assign(prims, prims + n);
    s->committed = false;
    return 0;
}
int rtc_set_nodes(RtcScene* s, uint32_t n_nodes, const RtcNode* nodes, uint32_t n_refs, const int32_t* refs) {
    if (!s || (n_nodes && !nodes) || (n_refs && !refs)) return fail(RTC_ERR_INVALID, "null argument");
    s->nodes.assign(nodes, nodes + n_nodes);
    s->refs.assign(refs, refs + n_refs);
    s->committed = false;
    return 0;
}
int rtc_set_materials(RtcScene* s, uint32_t n, const RtcMaterial* m) {
    if (!s || (n && !m)) return fail(RTC_ERR_INVALID, "null argument");
    if (n >= (1u << 23)) return fail(RTC_ERR_CAPACITY, "too many materials");
    s->materials.assign(m, m + n);
    s->committed = false;
    return 0;
}
int rtc_set_patterns(RtcScene* s, uint32_t n, const RtcPattern* p, uint32_t n_uv, const RtcUvPattern* uv) {
    if (!s || (n && !p) || (n_uv && !uv)) return fail(RTC_ERR_INVALID, "null argument");
    s->patterns.assign(p, p + n);
    s->uvs.assign(uv, uv + n_uv);
    s->committed = false;
    return 0;
}
int rtc_set_textures(RtcScene* s, uint32_t n, const RtcTexture* t) {
    if (!s || (n && !t)) return fail(RTC_ERR_INVALID, "null argument");
    s->textures.clear();
    for (uint32_t i = 0; i < n; i++) {
        if (!t[i].rgb && t[i].width * t[i].height) return fail(RTC_ERR_INVALID, "texture without pixels");
        RtcScene::Texture x;
        x.width = t[i].width, x.height = t[i].height;
        x.rgb.assign(t[i].rgb, t[i].rgb + (size_t)t[i].width * t[i].height * 3);
        s->textures.push_back(std::move(x));
    }
    s->committed = false;
    return 0;
}
int rtc_set_point_light(RtcScene* s, const float position[3], const float intensity[3]) {
    if (!s || !position || !intensity) return fail(RTC_ERR_INVALID, "null argument");
    s->light_is_rect = false;
    memcpy(s->light_pos, position, 12);
    memcpy(s->light_rgb, intensity, 12);
    memset(s->corner, 0, 12), memset(s->u_cell, 0, 12), memset(s->v_cell, 0, 12);
    s->u_steps = s->v_steps = 1;
    s->jitter.clear();
    s->have_light = true;
    s->committed = false;
    return 0;
}
int rtc_set_rect_light(RtcScene* s, const float intensity[3], const float corner[3], const float u_cell[3], int32_t u_steps,
                       const float v_cell[3], int32_t v_steps, const float position[3], const float* table, uint32_t table_len,
                       uint64_t seed) {
    if (!s || !intensity || !corner || !u_cell || !v_cell || !position || (table_len && !table))
        return fail(RTC_ERR_INVALID, "null argument");
    if (u_steps < 1 || v_steps < 1 || (int64_t)u_steps * v_steps > (1 << 20)) return fail(RTC_ERR_INVALID, "bad light steps");
    s->light_is_rect = true;
    memcpy(s->light_rgb, intensity, 12);
    memcpy(s->corner, corner, 12);
    memcpy(s->u_cell, u_cell, 12);
    memcpy(s->v_cell, v_cell, 12);
    memcpy(s->light_pos, position, 12);
    s->u_steps = u_steps, s->v_steps = v_steps;
    s->jitter.assign(table, table + table_len);
    s->seed = seed;
    s->have_light = true;
    s->committed = false;
    return 0;
}
int rtc_set_option(RtcScene* s, int32_t option, int64_t value) {
    if (!s) return fail(RTC_ERR_INVALID, "null scene");
    switch (option) {
        case RTC_OPT_FMA_CONTRACTION: s->strict_fp = value == 0; return 0;  // may change between renders
        case RTC_OPT_BVH_LEAF_SIZE:
            if (value < 0 || value > 16) return fail(RTC_ERR_INVALID, "leaf size must be in [1,16], or 0 = automatic");
            s->leaf_size = (int)value;
            s->committed = false;
            return 0;
        case RTC_OPT_ADAPTIVE_ORDER:
            s->adaptive_order = value != 0;
            return 0;
        case RTC_OPT_SHADOW_FILTER:  // may change between renders
            s->shadow_filter = value != 0;
            return 0;
        case RTC_OPT_RENDER_SLICES:
            if (value < 1 || value > 64) return fail(RTC_ERR_INVALID, "render slices must be in [1,64]");
            s->render_slices = (int)value;
            return 0;
        case RTC_OPT_BVH_MIN_PRIMS:
            s->bvh_min_prims = (int)std::max<int64_t>(0, value);
            s->committed = false;
            return 0;
    }
    return fail(RTC_ERR_INVALID, "unknown option");
}

int rtc_scene_inspect(RtcScene* s, RtcCommitInfo* out) {
    if (!s || !out) return fail(RTC_ERR_INVALID, "null argument");
    if (!s->have_camera) return fail(RTC_ERR_STATE, "camera not set");
    if (!s->have_light) return fail(RTC_ERR_STATE, "World light should be set");
    const auto t0 = std::chrono::steady_clock::now();
    Flattened f;
    int rc = flatten(s, f);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    out->host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    out->n_positions = f.n_pos, out->n_bvh_nodes = (int)f.bvh.size(), out->n_linear = (int)f.linear.size();
    out->n_xforms = (int)f.xform.size() / 3, out->bvh_leaf_size = f.leaf_size;
    out->small_n = f.small.n, out->filter_ok = f.small.filter_ok, out->cell_masks = f.small.cell_masks;
    out->plane_cells = f.small.plane_cells, out->converge = s->has_branching_materials ? 1 : 0;
    out->tol_sphere = f.small.tol_sphere;
    out->light_ball[0] = f.small.light_ball.x, out->light_ball[1] = f.small.light_ball.y;
    out->light_ball[2] = f.small.light_ball.z, out->light_ball[3] = f.small.light_ball.w;
    return 0;
}

int rtc_scene_commit(RtcScene* s, int32_t n_devices, const int32_t* device_ids) {
    if (!s) return fail(RTC_ERR_INVALID, "null scene");
    if (!s->have_camera) return fail(RTC_ERR_STATE, "camera not set");
    if (!s->have_light) return fail(RTC_ERR_STATE, "World light should be set");  // world.rs:66
    int visible = rtc_device_count();
    if (visible < 0) return visible;
    if (visible == 0) return fail(RTC_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
    if (n_devices == 0) n_devices = visible;
    if (n_devices < 0 || n_devices > visible) return fail(RTC_ERR_INVALID, "bad device count");
    release(s);
    Flattened f;
    int rc = flatten(s, f);
    if (rc) return rc;
    s->replicas.resize(n_devices);
    s->replica_devices.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        s->replica_devices[i] = device_ids ? device_ids[i] : i;
        if (s->replica_devices[i] < 0 || s->replica_devices[i] >= visible) {
            release(s);
            return fail(RTC_ERR_INVALID, "bad device id");
        }
        if ((rc = upload_replica(s, f, s->replicas[i], s->replica_devices[i]))) {
            release(s);
            return rc;
        }
    }
    s->committed = true;
    return 0;
}

int rtc_shard_bands(uint32_t height, int32_t shard, int32_t n_shards, uint32_t* first_rows) {
    static_assert(RTC_BAND_ROWS == kBandRows, "public band height must match the kernel tile height");
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(RTC_ERR_INVALID, "bad shard index");
    const int total = ((int)height + kBandRows - 1) / kBandRows;
    int n = 0;
    for (int b = shard; b < total; b += n_shards, n++)
        if (first_rows) first_rows[n] = (uint32_t)b * kBandRows;
    return n;
}

int rtc_render(RtcScene* s, int32_t depth, float* rgb, uint8_t* u8, RtcStats* stats) {
    return render_impl(s, depth, 0, 0, rgb, u8, stats, false);
}
int rtc_render_shard(RtcScene* s, int32_t depth, int32_t shard, int32_t n_shards, float* rgb, uint8_t* u8, RtcStats* stats) {
    if (n_shards < 1) return fail(RTC_ERR_INVALID, "n_shards must be >= 1");
    return render_impl(s, depth, shard, n_shards, rgb, u8, stats, false);
}
int rtc_render_detailed(RtcScene* s, int32_t depth, float* rgb, uint8_t* u8, RtcStats* stats) {
    return render_impl(s, depth, 0, 0, rgb, u8, stats, true);
}

int rtc_trace_rays(RtcScene* s, uint32_t n, const float* origins, const float* directions, int32_t depth, float* out_rgb,
                   float* out_t, int32_t* out_prim) {
    if (!s || !origins || !directions || !out_rgb) return fail(RTC_ERR_INVALID, "null argument");
    if (!s->committed) return fail(RTC_ERR_STATE, "rtc_trace_rays before rtc_scene_commit");
    if (depth < 0 || depth > kMaxFrames - 1) return fail(RTC_ERR_CAPACITY, "depth out of range");
    if (n == 0) return 0;
    Replica& rep = s->replicas[0];
    rep.small.filter_ok = rep.filter_eligible && s->shadow_filter;
    rep.small.cell_masks = rep.cell_masks_eligible && s->shadow_filter;
    rep.small.plane_cells = rep.plane_cells_eligible && rep.small.cell_masks;
    struct {
        int device;
        cudaStream_t stream;
        DevCounters* d_counters;
        DevScene& scene;
        SmallScene& small;
    } r{rep.slot->device, rep.slot->stream, rep.slot->d_counters, rep.scene, rep.small};
    CUDA_TRY(cudaSetDevice(r.device));
    float *d_o = nullptr, *d_d = nullptr, *d_rgb = nullptr, *d_t = nullptr;
    int* d_pos = nullptr;
    size_t b3 = (size_t)n * 3 * sizeof(float);
    CUDA_TRY(cudaMalloc(&d_o, b3));
    CUDA_TRY(cudaMalloc(&d_d, b3));
    CUDA_TRY(cudaMalloc(&d_rgb, b3));
    CUDA_TRY(cudaMalloc(&d_t, (size_t)n * sizeof(float)));
    CUDA_TRY(cudaMalloc(&d_pos, (size_t)n * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(d_o, origins, b3, cudaMemcpyHostToDevice, r.stream));
    CUDA_TRY(cudaMemcpyAsync(d_d, directions, b3, cudaMemcpyHostToDevice, r.stream));
    CUDA_TRY(cudaMemsetAsync(r.d_counters, 0, sizeof(DevCounters), r.stream));
    if (s->strict_fp)
        strict::launch_trace(r.scene, r.small, (int)n, d_o, d_d, depth, d_rgb, d_t, d_pos, r.d_counters, r.stream);
    else
        fast::launch_trace(r.scene, r.small, (int)n, d_o, d_d, depth, d_rgb, d_t, d_pos, r.d_counters, r.stream);
    CUDA_TRY(cudaGetLastError());
    std::vector<int> pos(n);
    CUDA_TRY(cudaMemcpyAsync(out_rgb, d_rgb, b3, cudaMemcpyDeviceToHost, r.stream));
    if (out_t) CUDA_TRY(cudaMemcpyAsync(out_t, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, r.stream));
    CUDA_TRY(cudaMemcpyAsync(pos.data(), d_pos, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, r.stream));
    CUDA_TRY(cudaStreamSynchronize(r.stream));
    if (out_prim)
        for (uint32_t i = 0; i < n; i++) out_prim[i] = pos[i] >= 0 ? s->pos_to_prim[pos[i]] : -1;
    cudaFree(d_o), cudaFree(d_d), cudaFree(d_rgb), cudaFree(d_t), cudaFree(d_pos);
    return 0;
}

void* rtc_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        g_error = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
void rtc_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int rtc_host_register(void* ptr, size_t bytes) {
    CUDA_TRY(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}
int rtc_host_unregister(void* ptr) {
    CUDA_TRY(cudaHostUnregister(ptr));
    return 0;
}

int rtc_flush_l2(RtcScene* s) {
    if (!s || !s->committed) return fail(RTC_ERR_STATE, "scene not committed");
    for (Replica& rep : s->replicas) {
        DeviceSlot* r = rep.slot;
        CUDA_TRY(cudaSetDevice(r->device));
        if (!r->d_flush) {
            r->flush_bytes = 256u << 20;  // > 126 MB L2
            CUDA_TRY(cudaMalloc(&r->d_flush, r->flush_bytes));
        }
        CUDA_TRY(cudaMemsetAsync(r->d_flush, 0x5a, r->flush_bytes, r->stream));
        CUDA_TRY(cudaStreamSynchronize(r->stream));
    }
    return 0;
}

int rtc_measure_fp32_peak(int32_t device, double* tflops, double* sm_clock_mhz) {
    if (!tflops) return fail(RTC_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0, khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device));
    if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms * 16, iters = 32768;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        fast::launch_fma_peak(d, blocks, iters, 0);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        double fl = (double)blocks * 256 * 8 * 2 * iters;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d);
    return 0;
}

}  // extern "C"
