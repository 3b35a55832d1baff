"""Shadow-filter statistics of a bench workload: shadow rays, fallbacks to the exact test, kernel time on / off.
    python tools/filter_stats.py c3 c1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import build_scene  # noqa: E402

for wl in sys.argv[1:] or ["c3"]:
    api = rt.new_session()
    api.set_render_options(device_ids=[0])
    cam, world, depth, _ = build_scene(api, wl)
    p = cam.prepare(world)
    for on in (1, 0):
        p.set_option(6, on)
        p.render(depth, want_rgb=False, want_u8=False, detailed=True)
        st = p.last_stats
        ms = []
        for _ in range(4):
            p.render(depth, want_rgb=False, want_u8=False)
            ms.append(p.last_stats.kernel_ms)
        print(f"{wl} filter={on}: shadow rays {st.shadow_rays}, exact fallbacks {st.prim_tests[7]} "
              f"({st.prim_tests[7] / max(st.shadow_rays, 1):.4%}), prim tests {list(st.prim_tests)[:3]}, kernel {min(ms):.3f} ms", flush=True)
    p.release()
