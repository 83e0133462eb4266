"""Shard-by-point multi-GPU plumbing: one process per GPU, `torch.distributed`
only carries the 128-byte NCCL id; the per-iteration exchange (reduced camera system, costs) is
NCCL inside the CUDA library, in-stream."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import _lib

ID_BYTES = _lib.BA_NCCL_ID_BYTES


def broadcast_unique_id(make_id: Callable[[], bytes], rank: int, group=None, device=None) -> bytes:
    """Rank 0 calls `make_id()` (128 bytes); every rank returns the same bytes."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = device if (backend == "nccl" and device is not None) else "cpu"
    buf = torch.zeros(ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        if len(raw) != ID_BYTES:
            raise ValueError("NCCL unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    raw = (C.c_ubyte * ID_BYTES)()
    _lib.check(lib, None, lib.ba_nccl_unique_id(raw), "ba_nccl_unique_id")
    return bytes(raw)


def init_solver(solver, rank: int, world: int, group=None, device=None) -> None:
    """Create the solver's NCCL communicator (no-op for world == 1)."""
    if world <= 1:
        return
    uid = broadcast_unique_id(nccl_unique_id, rank, group=group, device=device)
    solver.dist_init(uid, rank, world)


def shard_point_range(pt_offsets: np.ndarray, rank: int, world: int):
    """Point range [lo, hi) of `rank`: the CSR over observations cut at equal observation counts, so every
    shard carries the same linearisation work."""
    n_pt = len(pt_offsets) - 1
    n_obs = int(pt_offsets[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(pt_offsets, (n_obs * r) // world, side="left")))
    cuts.append(n_pt)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n_pt))
    return int(cuts[rank]), int(cuts[rank + 1])


def gather_flags(flags_local: np.ndarray, obs_range, n_obs: int, group=None) -> np.ndarray:
    """All-gather the per-rank outlier flags into the full (n_obs,) array (host side, any backend)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    objs = [None] * world
    dist.all_gather_object(objs, (int(obs_range[0]), int(obs_range[1]), np.asarray(flags_local)), group=group)
    out = np.zeros(n_obs, dtype=np.uint8)
    for lo, hi, f in objs:
        out[lo:hi] = f
    return out


def gather_points(X_local: np.ndarray, pt_range, n_pt: int, group=None) -> Optional[np.ndarray]:
    """All-gather the per-rank point blocks into the full (n_pt, 3) array (host side, any backend)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    objs = [None] * world
    dist.all_gather_object(objs, (int(pt_range[0]), int(pt_range[1]), np.asarray(X_local)), group=group)
    out = np.zeros((n_pt, 3), dtype=np.float64)
    for lo, hi, x in objs:
        out[lo:hi] = x
    return out
