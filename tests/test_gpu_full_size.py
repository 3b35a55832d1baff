"""BASELINE.json's configurations at FULL size on the GPU, checked against the oracle on a row sample (the oracle
renders every k-th row of the same frame in a few seconds), plus size-independent properties of the whole frame."""
import numpy as np
import pytest

from bench import WORKLOADS, build_scene
from tests.parity import assert_parity, compare_frames

pytestmark = pytest.mark.gpu

# workload -> (row stride of the oracle sample, minimum fraction of sampled pixels that must be bit-identical in 8 bits)
# c5 (100 k spheres): the reference's sphere test is noise-dominated at the silhouettes of far, tiny spheres, where its
# own group boxes decide whether a phantom hit is reported (see test_sphere_field_parity_at_scale): all but a few
# 1e-5 of the pixels are identical.
CASES = {"c1": (8, 1.0), "c2": (24, 1.0), "c3": (48, 1.0), "c4": (24, 1.0), "c5": (48, 0.9995)}


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


@pytest.mark.parametrize("workload", sorted(CASES))
def test_full_size_frame_matches_oracle_rows(workload, gpu, oracle):
    ystep, min_exact = CASES[workload]
    gcam, gworld, depth, _ = build_scene(gpu, workload)
    got = gcam.render_b200(gworld, depth)
    stats = gcam.last_rtc_stats
    w, h = gcam.width_pixels, gcam.height_pixels
    assert got.data.shape == (h, w, 3)
    # properties of the whole frame: never-rendered border, finite values, 8-bit plane == scale_color(f32 plane)
    assert not got.data[-1].any() and not got.data[:, -1].any()
    assert np.isfinite(got.data).all()
    expect_u8 = np.clip(got.data * np.float32(255.0), 0, 255).astype(np.uint8)  # truncation, canvas.rs:39-43
    assert np.array_equal(expect_u8, got.to_u8())
    assert stats.primary_rays == (w - 1) * (h - 1)
    # oracle on every ystep-th row
    oracle.probe.set_threads(oracle.probe.max_threads())
    ocam, oworld, _, _ = build_scene(oracle, workload)
    rgb, u8, ost = oracle.probe.render_rows(ocam, oworld, depth, ystep // 2, h, ystep, want_u8=True)
    rows = np.arange(ystep // 2, h - 1, ystep)
    rep = compare_frames(got.to_u8()[rows], u8[rows], got.data[rows], rgb[rows])
    loose = min_exact < 1.0
    assert_parity(rep, min_within=0.9995 if loose else 0.9999, max_gross=0.0005 if loose else 0.0001,
                  label=f"{workload} rows {ystep // 2}::{ystep}")
    print(workload, rep)
    assert rep["exact_u8"] >= min_exact - 1e-4, rep


def test_sphere_field_parity_at_scale(gpu, oracle):
    """A 20 k-sphere field (config 5's structure) against the oracle's full frame.  The reference's sphere test is
    noise-dominated at the silhouette of far, tiny spheres (b^2 - 4ac cancels in f32), and there the reference's own
    group boxes decide whether a phantom hit is reported; the BVH reproduces all but ~1e-6 of those rays."""
    from ray_tracer_challenge_b200 import scenes

    kw = dict(width=640, height=360, n_spheres=20_000, n_each=16, n_csg=8)
    oracle.probe.set_threads(oracle.probe.max_threads())
    ocam, ow = scenes.stress(oracle, **kw)
    want = ocam.render(ow, 5)
    gcam, gw = scenes.stress(gpu, **kw)
    got = gcam.render_b200(gw, 5)
    rep = compare_frames(got.to_u8(), want.to_u8(), got.data, want.data)
    assert_parity(rep, min_within=0.9995, max_gross=0.0005, label="20k-sphere field")
    rays_gpu, rays_cpu = gcam.last_rtc_stats.rays, ocam.last_stats.rays
    assert abs(rays_gpu - rays_cpu) <= 1e-4 * rays_cpu, (rays_gpu, rays_cpu)
