"""Tile costs (clock cycles per 16x8-pixel block) of one shard of an N-way split, as the learning render records them:
    RTC_DUMP_TILE_COST=out.txt python tools/tile_dump.py --workload c3 --shards 8"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ray_tracer_challenge_b200 as rt  # noqa: E402
from bench import WORKLOADS, build_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
ap.add_argument("--shards", type=int, default=8)
a = ap.parse_args()
api = rt.new_session()
api.set_render_options(device_ids=[0])
cam, world, depth, desc = build_scene(api, a.workload)
p = cam.prepare(world)
for i in range(8):
    p.render(depth, want_rgb=False, want_u8=False, shard=0, n_shards=a.shards)
    print(i, p.last_stats.kernel_ms)
