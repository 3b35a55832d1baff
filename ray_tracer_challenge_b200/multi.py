"""Host plumbing for one frame rendered by several processes (one per GPU, `torchrun`): a canvas in POSIX shared
memory that every rank writes its bands into — the "simple host gather" of SURVEY.md §8e, with no copy and no
collective on the data path — plus the two scalar reductions the benchmark needs (torch.distributed is plumbing
here, not the product)."""
from __future__ import annotations

from multiprocessing import shared_memory

import numpy as np


class SharedCanvas:
    """width x height f32 RGB plane followed by the 8-bit plane, shared by all ranks of one node."""

    def __init__(self, name: str, width: int, height: int, create: bool):
        self.width, self.height = width, height
        size = width * height * 15
        if create:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
        else:
            self.shm = shared_memory.SharedMemory(name=name)
            # Python < 3.13 registers attached segments with the resource tracker, which then tries to unlink them a
            # second time at exit; only the creating rank owns the segment
            try:
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = create
        self.buf = np.ndarray((size,), np.uint8, buffer=self.shm.buf)
        self.rgb = self.buf[: width * height * 12].view(np.float32).reshape(height, width, 3)
        self.u8 = self.buf[width * height * 12:].reshape(height, width, 3)

    @property
    def address(self) -> int:
        return self.buf.ctypes.data

    @property
    def nbytes(self) -> int:
        return self.buf.nbytes

    def close(self):
        self.rgb = self.u8 = self.buf = None
        self.shm.close()
        if self.owner:
            self.shm.unlink()


def open_shared_canvas(dist, rank: int, name: str, width: int, height: int) -> SharedCanvas:
    """Rank 0 creates the canvas, everybody else attaches after a barrier."""
    canvas = SharedCanvas(name, width, height, create=True) if rank == 0 else None
    dist.barrier()
    if rank != 0:
        canvas = SharedCanvas(name, width, height, create=False)
    return canvas


def reduce_scalar(dist, value: float, op: str, device="cpu") -> float:
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())
