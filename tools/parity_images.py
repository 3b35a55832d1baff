"""Debug aid: render every parity scene on the GPU (fast + strict) and with the oracle, write the frames and the
diff masks as PNGs plus a JSON report into gpurun_out/."""
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from tests.oracle_binding import load_oracle  # noqa: E402
from tests.parity import compare_frames  # noqa: E402
from tests.test_gpu_parity import SCENES  # noqa: E402

out = os.path.join(ROOT, "gpurun_out", "parity")
os.makedirs(out, exist_ok=True)
oracle = load_oracle()
oracle.probe.set_threads(oracle.probe.max_threads())
gpu = rt.new_session()
names = sys.argv[1:] or sorted(SCENES)
report = {}
for name in names:
    build, kw = SCENES[name]
    ocam, ow = build(oracle, **kw)
    want = ocam.render(ow, 5)
    gcam, gw = build(gpu, **kw)
    p = gcam.prepare(gw)
    Image.fromarray(want.to_u8()).save(os.path.join(out, f"{name}_oracle.png"))
    for strict in (True, False):
        got = p.render(5, fma=not strict)
        rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
        rep["rays_gpu"], rep["rays_oracle"] = int(p.last_stats.rays), int(ocam.last_stats.rays)
        rep["kernel_ms"] = p.last_stats.kernel_ms
        tag = "ieee" if strict else "fma"
        report[f"{name}:{tag}"] = rep
        d = np.abs(got.to_u8().astype(np.int16) - want.to_u8().astype(np.int16)).max(axis=2)
        mask = np.zeros(d.shape + (3,), np.uint8)
        mask[d == 1] = (0, 90, 0)
        mask[(d > 1) & (d <= 8)] = (255, 255, 0)
        mask[d > 8] = (255, 0, 0)
        if d.max() > 0:
            Image.fromarray(mask).save(os.path.join(out, f"{name}_{tag}_diff.png"))
        if not strict:
            Image.fromarray(got.to_u8()).save(os.path.join(out, f"{name}_gpu.png"))
        print(f"{name:28s} {tag:6s} exact={rep['exact_u8']:.5f} within1={rep['within_1lsb']:.5f} gross={rep['gross']:.5f} "
              f"max={rep['max_u8_diff']:3d} f32exact={rep['bit_exact_f32']:.4f} rays {rep['rays_gpu']}/{rep['rays_oracle']} "
              f"{rep['kernel_ms']:.2f} ms", flush=True)
    p.release()
with open(os.path.join(out, "report.json"), "w") as fh:
    json.dump(report, fh, indent=1)
