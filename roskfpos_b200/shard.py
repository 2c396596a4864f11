"""Multi-GPU host logic: filters shard by index, no data-path collective; the only exchange is
the final sum of the per-rank statistics vectors.  One process per GPU,
`torch.distributed` (NCCL on the GPU box, gloo in the CPU tests) is only the plumbing."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_filters: int, rank: int, world: int):
    """Contiguous index range [lo, hi) of rank `rank`: GPU g owns filters [g*N/G, (g+1)*N/G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = (n_filters * rank) // world
    hi = (n_filters * (rank + 1)) // world
    return lo, hi


def reduce_stats(stats, device=None, group=None):
    """Sums the per-rank [sum |e|^2, sum |e_xy|^2, n, n_bad] (kfpos_batch_error_stats) over all ranks
    with one all_reduce and returns (summed vector, rmse, rmse_xy).  Without an initialised process
    group it is the identity (single GPU)."""
    import torch
    import torch.distributed as dist
    v = torch.as_tensor(np.asarray(stats, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if device is not None:
            v = v.to(device)
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        v = v.cpu()
    s = v.numpy()
    n = max(s[2], 1.0)
    return s, float(np.sqrt(s[0] / n)), float(np.sqrt(s[1] / n))


# ---- a raw NCCL communicator for the C ABI (kfpos_stats_allreduce takes an ncclComm_t).  torch.distributed
# does not hand its communicators out, so one is created beside it: rank 0 draws the unique id, the process
# group (NCCL or gloo) carries its 128 bytes to the other ranks, every rank joins with ncclCommInitRank.
_NCCL = None


def _nccl():
    global _NCCL
    if _NCCL is None:
        import ctypes as C
        import glob
        import os
        try:
            _NCCL = C.CDLL("libnccl.so.2")  # the copy PyTorch has already mapped, when there is one
        except OSError:
            import torch
            cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib",
                                           "libnccl.so*"))
            if not cands:
                raise
            _NCCL = C.CDLL(cands[0], mode=C.RTLD_GLOBAL)
    return _NCCL


class NcclComm:
    """Owns one ncclComm_t; int(comm) is the raw handle."""

    def __init__(self, handle, lib):
        self.handle, self._lib = handle, lib

    def __int__(self):
        return int(self.handle.value)

    def close(self):
        if self.handle is not None and self.handle.value:
            self._lib.ncclCommDestroy(self.handle)
        self.handle = None


def nccl_comm(rank: int = 0, world: int = 1, device: int = 0, group=None) -> NcclComm:
    """One communicator over `world` ranks (one per GPU).  With world == 1 no process group is needed."""
    import ctypes as C
    import torch

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]
    lib = _nccl()
    lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    uid = UniqueId()
    if rank == 0:
        if lib.ncclGetUniqueId(C.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    if world > 1:
        import torch.distributed as dist
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=0, group=group)
        C.memmove(C.byref(uid), box[0], 128)
    comm = C.c_void_p()
    with torch.cuda.device(device):
        rc = lib.ncclCommInitRank(C.byref(comm), world, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank failed ({rc})")
    return NcclComm(comm, lib)
