"""Rank discovery, row partitioning and the few collectives of the path.

Mirrors pyLOM/utils/parall.py:24-48 (`worksplit`), :87-99 (`pprint`) and the mpi4py facade in
pyLOM/utils/mpi.py:17-19,39-129.  One process per GPU; ranks come from torch.distributed (or the
RANK / WORLD_SIZE environment that torchrun sets).  The reference binds MPI_RANK/MPI_SIZE at
import time; here they are live properties of the process group, exposed as functions AND as
module attributes refreshed by `init_distributed`.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

MPI_RANK = int(os.environ.get("RANK", "0")) if "WORLD_SIZE" in os.environ else 0
MPI_SIZE = int(os.environ.get("WORLD_SIZE", "1"))


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def size():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def init_distributed(backend=None):
    """Join the process group described by the torchrun environment (no-op for one process)."""
    global MPI_RANK, MPI_SIZE
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend)
    MPI_RANK, MPI_SIZE = rank(), size()
    return MPI_RANK, MPI_SIZE


_c_comm = None


def c_comm():
    """The library's own communicator (pl_comm_t wrapping an NCCL communicator created INSIDE libpylom_b200): rank 0
    draws the unique id with pl_get_unique_id, the 128 bytes are broadcast over the existing process group, every rank
    calls pl_comm_init_rank.  This is the handle the collective C entry points (pl_tsqr_svd_dist_f64,
    pl_tsqr_svd_host_dist_f64) take -- the counterpart of MPI_COMM_WORLD in the reference's dtsqr_svd."""
    global _c_comm
    if _c_comm is None:
        import ctypes
        from .. import _lib
        L = _lib.lib()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank() == 0:
            _lib.check(L.pl_get_unique_id(ident.data_ptr()), "pl_get_unique_id")
        if is_distributed():
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = ident.to(dev)
            dist.broadcast(t, 0)
            ident = t.cpu()
        h = ctypes.c_void_p()
        _lib.check(L.pl_comm_init_rank(ctypes.byref(h), ident.data_ptr(), rank(), size()), "pl_comm_init_rank")
        _c_comm = h
    return _c_comm


def c_comm_destroy():
    global _c_comm
    if _c_comm is not None:
        from .. import _lib
        _lib.lib().pl_comm_destroy(_c_comm)
        _c_comm = None


def use_c_comm():
    """The collective C entry point is used whenever the process group runs on NCCL (PL_NO_C_COMM=1 keeps the
    torch.distributed composition, which is also what the gloo CPU tests exercise)."""
    return is_distributed() and dist.get_backend() == "nccl" and not os.environ.get("PL_NO_C_COMM")


def worksplit(istart, iend, whoAmI, nWorkers=None):
    """Contiguous range of worker `whoAmI`; the remainder goes one each to the lowest ranks
    (pyLOM/utils/parall.py:24-48)."""
    if nWorkers is None:
        nWorkers = size()
    istart_l, iend_l = istart, iend
    irange = iend - istart
    if nWorkers < irange:
        per = irange // nWorkers
        istart_l = istart + whoAmI * per
        iend_l = istart_l + per
        rem = irange - per * nWorkers
        if rem > whoAmI:
            istart_l += whoAmI
            iend_l += whoAmI + 1
        else:
            istart_l += rem
            iend_l += rem
    else:
        istart_l = whoAmI if whoAmI < iend else iend
        iend_l = whoAmI + 1 if whoAmI < iend else iend
    return istart_l, iend_l


def pprint(r, *args, **kwargs):
    """Print on rank `r` only (r < 0: every rank)  -- pyLOM/utils/parall.py:87-99."""
    if r < 0 or rank() == r:
        print(*args, **kwargs)
        sys.stdout.flush()


def mpi_barrier():
    if is_distributed():
        dist.barrier()


def mpi_reduce(x, root=0, op="sum", all=True):
    """Sum-reduce a scalar / tensor over ranks (pyLOM/utils/mpi.py:62-80).  Always an allreduce."""
    if not is_distributed():
        return x
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    if isinstance(x, torch.Tensor):
        y = x.clone()
        dist.all_reduce(y, op=ops[op])
        return y
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    y = torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    dist.all_reduce(y, op=ops[op])
    y = y.cpu().numpy()
    return float(y) if y.ndim == 0 else y


def mpi_allgather_rows(x):
    """Stack the same-shaped 2-D tensor of every rank along rows: ONE all-gather (NCCL over NVLink on
    GPUs).  Replaces the 2*log2(P) send/recv rounds of the reference butterfly
    (pyLOM/vmmath/svd.py:67-115, src/svd.c:602-669)."""
    if not is_distributed():
        return x
    P = dist.get_world_size()
    out = torch.empty((P * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out
