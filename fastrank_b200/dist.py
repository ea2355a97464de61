"""Multi-GPU plumbing: one process per GPU, queries sharded across ranks, one NCCL
all-reduce of the fixed-point metric sums per evaluation.

The reference has no distributed path; queries are independent units whose only
cross-query operation is the sum in evaluate_mean (evaluators.rs:178-183).  Every rank runs
the same host control flow (coordinate ascent is deterministic given the all-reduced sums),
so no broadcast of weights is needed.

torch.distributed is used for rendezvous only (shipping the 128-byte NCCL unique id); the
data-path collective is issued by the native library on its own stream.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(query_sizes: np.ndarray, world: int) -> np.ndarray:
    """Split queries (in order) into `world` contiguous ranges balanced by document count.
    Returns world+1 boundaries into the query list."""
    sizes = np.asarray(query_sizes, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        k = min(max(k, bounds[-1]), len(sizes))
        bounds.append(k)
    bounds.append(len(sizes))
    return np.asarray(bounds, dtype=np.int64)


def shard_rows(qid: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Row indices of this rank's shard: whole queries only, first-appearance query order,
    contiguous ranges balanced by document count.  Rows keep their original relative order,
    so instance ids inside a shard are ascending in the original ids (the tie-break of
    evaluators.rs:46-48 is preserved)."""
    qid = np.asarray(qid)
    uniq, first, inverse, counts = np.unique(qid, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")          # queries by first appearance
    rank_of = np.empty(len(order), dtype=np.int64)
    rank_of[order] = np.arange(len(order))
    bounds = shard_bounds(counts[order], world)
    qrank = rank_of[inverse]
    mask = (qrank >= bounds[rank]) & (qrank < bounds[rank + 1])
    return np.nonzero(mask)[0]


class Communicator:
    """Owns one fr_dev_comm and (optionally) installs it as the process default."""

    def __init__(self, ptr, rank: int, world: int):
        self.ptr, self.rank, self.world = ptr, rank, world

    def install_default(self):
        from ._native import lib

        lib.fr_dev_set_default_comm(self.ptr)

    def allreduce_u64(self, values) -> np.ndarray:
        from ._native import ffi, lib

        buf = np.ascontiguousarray(values, dtype=np.uint64).copy()
        if lib.fr_dev_comm_allreduce_u64(self.ptr, ffi.cast("uint64_t*", buf.ctypes.data), len(buf)):
            raise RuntimeError(ffi.string(lib.fr_dev_last_error()).decode())
        return buf

    def close(self):
        from ._native import ffi, lib

        if self.ptr is not None:
            if lib.fr_dev_default_comm() == self.ptr:
                lib.fr_dev_set_default_comm(ffi.NULL)
            lib.fr_dev_comm_destroy(self.ptr)
            self.ptr = None


def exchange_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 produces the id (make_id()), every rank returns the same 128 bytes.  Uses the
    already-initialised torch.distributed process group (nccl on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist

    payload = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(payload, src=0)
    out = payload[0]
    assert isinstance(out, (bytes, bytearray)) and len(out) == 128
    del torch
    return bytes(out)


def init_communicator(rank: int, world: int, device: int, install_default: bool = True) -> Communicator:
    """Creates the NCCL communicator of the native library for this rank."""
    import torch  # noqa: F401  (loads the bundled libnccl.so.2 the library dlopens)

    from ._native import ffi, lib

    def make_id() -> bytes:
        buf = ffi.new("uint8_t[128]")
        if lib.fr_dev_comm_unique_id(buf):
            raise RuntimeError(ffi.string(lib.fr_dev_last_error()).decode())
        return bytes(ffi.buffer(buf, 128))

    uid = exchange_unique_id(make_id, rank, world)
    out = ffi.new("fr_dev_comm**")
    if lib.fr_dev_comm_create(device, rank, world, ffi.from_buffer("uint8_t[]", bytearray(uid)), out):
        raise RuntimeError(ffi.string(lib.fr_dev_last_error()).decode())
    comm = Communicator(out[0], rank, world)
    if install_default:
        comm.install_default()
    return comm


def combine_fixed_point(local_sums: np.ndarray, local_queries: int, group=None) -> Tuple[np.ndarray, int]:
    """Host-side statement of the reduction protocol (used by the gloo CPU tests and as
    documentation of what the NCCL path computes on the device): integer sum of the
    fixed-point metric sums and of the query counts, hence order- and world-size-independent."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.concatenate([np.asarray(local_sums, dtype=np.int64).ravel(), [local_queries]]).astype(np.int64))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.numpy()
    return out[:-1].reshape(np.shape(local_sums)), int(out[-1])
