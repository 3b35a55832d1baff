"""Band sharding of one frame over several GPUs / processes: pure host arithmetic.

The frame is cut into bands of BAND_ROWS rows (the kernel's tile height); band b belongs to shard b % n_shards.
Pixels are independent, so the shards need no exchange: every rank renders its bands and copies them into its
rows of one shared host canvas (`rtc_render_shard`)."""
from __future__ import annotations

BAND_ROWS = 8


def bands_of(height: int, shard: int, n_shards: int) -> list[int]:
    """Indices of the bands owned by `shard`."""
    if n_shards < 1 or not 0 <= shard < n_shards:
        raise ValueError("bad shard index")
    total = (height + BAND_ROWS - 1) // BAND_ROWS
    return list(range(shard, total, n_shards))


def rows_of(height: int, shard: int, n_shards: int) -> list[range]:
    """Row ranges [first, last) of the bands owned by `shard`."""
    return [range(b * BAND_ROWS, min((b + 1) * BAND_ROWS, height)) for b in bands_of(height, shard, n_shards)]
