"""The wavefront renderer (csrc/dev_wave.cuh, RTC_OPT_WAVEFRONT = 8): rays as work items, one queue per bounce level, tree
walks / shading / post-order combine in kernels of their own.  It must render what the streaming kernel renders, bit
for bit, with the same ray counts — and a chunk whose ray pool overflows is rendered again by the streaming kernel, so
the frame is complete either way."""
import numpy as np
import pytest

from ray_tracer_challenge_b200 import scenes
from tests.parity import compare_frames

pytestmark = pytest.mark.gpu
RTC_OPT_WAVEFRONT = 8


@pytest.fixture(scope="module")
def gpu():
    import ray_tracer_challenge_b200 as rt

    return rt.new_session()


@pytest.mark.parametrize("kw", [dict(width=384, height=216, n_spheres=5000, n_each=8, n_csg=4),
                                dict(width=163, height=67, n_spheres=1500, n_each=4, n_csg=2)])  # ragged edges too
def test_wavefront_frame_equals_streaming_frame(gpu, oracle, kw):
    cam, world = scenes.stress(gpu, **kw)
    assert gpu.inspect(cam, world)["converge"] == 1  # branching ray trees: the scene class both renderers are for
    p = cam.prepare(world)
    try:
        for depth in (5, 1, 0):
            p.set_option(RTC_OPT_WAVEFRONT, 0)
            stream = p.render(depth)
            st_stream = p.last_stats
            p.set_option(RTC_OPT_WAVEFRONT, 1)
            wave = p.render(depth)
            st_wave = p.last_stats
            assert st_wave.launches >= 8 and st_wave.wave_overflows == 0, (st_wave.launches, st_wave.wave_overflows)
            assert (st_wave.primary_rays, st_wave.secondary_rays, st_wave.shadow_rays, st_wave.shades) == \
                   (st_stream.primary_rays, st_stream.secondary_rays, st_stream.shadow_rays, st_stream.shades), depth
            assert np.array_equal(wave.data.view(np.uint32), stream.data.view(np.uint32)), depth
            assert np.array_equal(wave.to_u8(), stream.to_u8())
    finally:
        p.release()
    ocam, oworld = scenes.stress(oracle, **kw)
    want = ocam.render(oworld, 0)
    rep = compare_frames(wave.to_u8(), want.to_u8(), wave.data, want.data)
    assert rep["exact_u8"] >= 0.9995, rep


def test_wavefront_shards_reassemble(gpu):
    kw = dict(width=200, height=120, n_spheres=2000, n_each=4, n_csg=2)
    cam, world = scenes.stress(gpu, **kw)
    p = cam.prepare(world)
    try:
        p.set_option(RTC_OPT_WAVEFRONT, 1)
        whole = p.render(5)
        parts = np.zeros_like(whole.data)
        for shard in range(3):
            p.render(5, out_rgb=parts, shard=shard, n_shards=3)
        assert np.array_equal(parts.view(np.uint32), whole.data.view(np.uint32))
    finally:
        p.release()
