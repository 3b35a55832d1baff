"""Driver for compute-sanitizer: one 160x64 frame of EVERY render_tiles / trace_rays instantiation, both kernel builds.

    compute-sanitizer --tool memcheck  python tools/sanitize_frames.py
    compute-sanitizer --tool racecheck python tools/sanitize_frames.py --quick

The kernels carve four regions out of one dynamic shared-memory block by offset arithmetic (dev_small.cuh), index 1.3 KB
bounce stacks and traversal stacks dynamically, and stage the frame epilogue through static shared memory: memcheck sees
every out-of-bounds access of those, racecheck the missing barriers.  Frames are also rendered at 163x67 so the ragged
right / bottom tiles (per-pixel stores) run next to the whole tiles (staged 16-byte stores)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ray_tracer_challenge_b200 as rt  # noqa: E402
from ray_tracer_challenge_b200 import scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true", help="IEEE build only, 160x64 only")
a = ap.parse_args()
api = rt.new_session()
api.set_render_options(device_ids=[0])

CASES = [
    # (label, builder, kwargs)                                             render_tiles<STATS, SMALL, CONVERGE, DRAWN>
    ("soft_shadows table", scenes.soft_shadows, dict(u_steps=4, v_steps=4)),                   # <., 1, 0, 0> cell masks + plane cells
    ("soft_shadows drawn", scenes.soft_shadows, dict(u_steps=4, v_steps=4, jitter=None)),      # <., 1, 0, 1>
    ("soft_shadows 10x10 table", scenes.soft_shadows, dict(u_steps=10, v_steps=10)),           # 100 staged samples
    ("filter_zoo drawn", scenes.filter_zoo, dict()),                                           # non-casters, cubes
    ("reflect_refract", scenes.reflect_refract, dict()),                                       # <., 1, 1, 0> converge
    ("reflect_refract csg", scenes.reflect_refract, dict(with_csg=True)),                      # CSG roots in a small scene
    ("first_textures", scenes.first_textures, dict(u_steps=3, v_steps=3)),                     # org cache path, drawn, not filterable
    ("shapes_zoo", scenes.shapes_zoo, dict(area_light=True)),
    ("csg_gallery", scenes.csg_gallery, dict()),
    ("textured", scenes.textured, dict()),
    ("dragon_element", scenes.dragon_element, dict(n_u=24, n_v=12)),                           # <., 0, 0, 0> BVH, triangles
    ("stress", scenes.stress, dict(n_spheres=3000, n_each=8, n_csg=4)),                        # <., 0, 1, 0> BVH converge, CSG
]
sizes = [(160, 64)] if a.quick else [(160, 64), (163, 67)]
builds = [False] if a.quick else [False, True]
total = 0
for label, builder, kw in CASES:
    for (w, h) in sizes:
        cam, world = builder(api, width=w, height=h, **kw)
        p = cam.prepare(world)
        for fma in builds:
            for detailed in (False, True):
                img = p.render(5, fma=fma, detailed=detailed)
                total += 1
            for n_shards in (2,):
                for shard in range(n_shards):
                    p.render(5, shard=shard, n_shards=n_shards, fma=fma)
                    total += 1
            # trace_rays<SMALL, DRAWN>: 257 rays (a ragged last block)
            rng = np.random.default_rng(1)
            o = rng.uniform(-3, 3, size=(257, 3)).astype(np.float32)
            d = rng.normal(size=(257, 3)).astype(np.float32)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            p.trace_rays(o, d, 5, fma=fma)
            total += 1
        assert np.isfinite(img.data).all(), label
        p.release()
        print(f"ok {label} {w}x{h}", flush=True)
print(f"sanitize_frames: {total} launches", flush=True)
