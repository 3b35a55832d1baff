// ThreadSanitizer run of the host flattener (tests/test_host_threads.py builds this file together with
// csrc/host/rtc_host.cpp under -fsanitize=thread): a divided 20 000-sphere field and a divided 6 000-triangle mesh,
// flattened five times by the worker pool.  Any data race in the subtree walks, the material merge or the pool itself
// makes TSan print a report and exit non-zero.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../include/rtc_scene.h"
extern "C" int sg_flatten(sg_ctx* c, int w, int* counts, void* prims, void* nodes, int32_t* refs, int* prim_shapes);
int main() {
    sg_ctx* c = sg_create();
    const float mp[10] = {0.8f, 0.6f, 0.3f, 0.1f, 0.8f, 0.4f, 60.f, 0.f, 0.f, 1.f};
    const float mq[10] = {0.7f, 0.7f, 0.8f, 0.1f, 0.5f, 0.4f, 100.f, 0.5f, 0.f, 1.f};
    const int m0 = sg_material_new(c, mp, -1), m1 = sg_material_new(c, mq, -1);
    int kind_group = -1;
    for (int k = 0; k < 10; k++) {  // find the group kind: the one sg_group_add_child accepts
        int g = sg_shape_new(c, k);
        if (g < 0) continue;
        int s = sg_shape_new(c, 0);
        if (sg_group_add_child(c, g, s) == 0) { kind_group = k; break; }
    }
    if (kind_group < 0) { printf("no group kind\n"); return 1; }
    const int grp = sg_shape_new(c, kind_group);
    unsigned long long x = 88172645463325252ull;
    auto rnd = [&] { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (float)((x >> 11) * (1.0 / 9007199254740992.0)); };
    for (int i = 0; i < 20000; i++) {
        int s = sg_shape_new(c, 0);
        float t[16], sc[16], m[16];
        sg_translation((rnd() * 2 - 1) * 30, (rnd() * 2 - 1) * 30, (rnd() * 2 - 1) * 30, t);
        sg_scaling(0.3f, 0.3f, 0.3f, sc);
        sg_matmul(t, sc, m);
        sg_shape_set_transform(c, s, m);
        sg_shape_set_material(c, s, i % 8 == 1 ? m1 : m0);
        sg_group_add_child(c, grp, s);
    }
    sg_shape_divide(c, grp, 8);
    const int mesh = sg_shape_new(c, kind_group);
    for (int i = 0; i < 100; i++)
        for (int j = 0; j < 60; j++) {
            float p1[3] = {(float)i, 40.f, (float)j}, p2[3] = {i + 1.f, 40.f, (float)j}, p3[3] = {(float)i, 40.f, j + 1.f};
            sg_group_add_child(c, mesh, sg_triangle_new(c, p1, p2, p3));
        }
    sg_shape_divide(c, mesh, 4);
    const int w = sg_world_new(c);
    sg_world_add_object(c, w, grp);
    sg_world_add_object(c, w, mesh);
    const float lp[3] = {-10, 100, -10}, li[3] = {1, 1, 1};
    sg_world_set_point_light(c, w, lp, li);
    int counts[6];
    for (int it = 0; it < 5; it++) {
        if (sg_flatten(c, w, counts, nullptr, nullptr, nullptr, nullptr)) { printf("flatten failed: %s\n", sg_last_error()); return 1; }
    }
    printf("flattened: %d prims %d nodes %d refs %d materials\n", counts[0], counts[1], counts[2], counts[3]);
    sg_destroy(c);
    return 0;
}
