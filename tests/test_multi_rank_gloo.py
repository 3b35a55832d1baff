"""The N > 1 host path on CPU: two `gloo` ranks shard one frame by interleaved bands, write them into one shared
canvas and reduce their counters — with the CPU oracle standing in for the device renderer (the same sharding
rule, shared-memory gather and reductions that bench.py uses with one GPU per rank)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, width, height, result_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    import torch.distributed as dist

    from ray_tracer_challenge_b200 import multi, scenes
    from ray_tracer_challenge_b200.sharding import rows_of
    from tests.oracle_binding import load_oracle

    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    oracle = load_oracle()
    cam, world = scenes.soft_shadows(oracle, width=width, height=height, u_steps=2, v_steps=2)
    canvas = multi.open_shared_canvas(dist, rank, f"rtc_test_{port}", width, height)
    rays = 0
    for rows in rows_of(height, rank, world_size):  # this rank's bands
        rgb, u8, st = oracle.probe.render_rows(cam, world, 5, rows.start, rows.stop, 1, want_u8=True)
        last = min(rows.stop, height - 1)  # camera.rs:80 — the last row is never rendered
        canvas.rgb[rows.start:last] = rgb[rows.start:last]
        canvas.u8[rows.start:last] = u8[rows.start:last]
        rays += st.rays
    dist.barrier()
    total_rays = multi.reduce_scalar(dist, rays, "sum")
    slowest = multi.reduce_scalar(dist, float(rank + 1), "max")
    if rank == 0:
        full = cam.render(world, 5)
        ok = (np.array_equal(canvas.rgb.view(np.uint32), full.data.view(np.uint32))
              and np.array_equal(canvas.u8, full.to_u8()) and total_rays == cam.last_stats.rays and slowest == world_size)
        with open(result_path, "w") as fh:
            fh.write("ok" if ok else f"mismatch rays={total_rays} vs {cam.last_stats.rays} slowest={slowest}")
    dist.barrier()
    canvas.close()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_reassemble_the_frame(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, port, 61, 37, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"
