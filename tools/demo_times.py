"""Kernel time of the reference's camera demos at their shipped sizes (device-resident frames).
    python tools/demo_times.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from ray_tracer_challenge_b200 import scenes  # noqa: E402

DEMOS = [
    ("soft_shadows (as shipped: jitter None -> counter RNG)", scenes.soft_shadows, dict(width=1000, height=400, jitter=None, seed=1)),
    ("soft_shadows (jitter table)", scenes.soft_shadows, dict(width=1000, height=400)),
    ("reflect_refract", scenes.reflect_refract, dict(width=1000, height=500)),
    ("hexagons", scenes.hexagons, dict(width=1000, height=500)),
    ("first_scene", scenes.first_scene, dict(width=1000, height=500)),
    ("first_plane", scenes.first_plane, dict(width=1000, height=500)),
    ("first_patterns", scenes.first_patterns, dict(width=1000, height=500)),
    ("first_textures (10x10 area light, jitter None)", scenes.first_textures, dict(width=1000, height=500)),
    ("skybox", scenes.skybox, dict(width=800, height=400)),
    ("here_be_dragons (6 x 13k-triangle synthetic mesh)", scenes.here_be_dragons, dict(width=500, height=200, n_u=80, n_v=40)),
]
api = rt.new_session()
api.set_render_options(device_ids=[0])
for name, build, kw in DEMOS:
    cam, world = build(api, **kw)
    p = cam.prepare(world)
    ms = []
    for _ in range(5):
        p.render(5, want_rgb=False, want_u8=False)
        ms.append(p.last_stats.kernel_ms)
    st = p.last_stats
    print(f"{name:55s} {kw['width']}x{kw['height']}  {min(ms[1:]):8.3f} ms  {st.rays:>10d} rays  {st.rays / min(ms[1:]) / 1e3:9.1f} Mrays/s", flush=True)
    p.release()
