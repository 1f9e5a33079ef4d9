"""Error convention (pyLOM/utils/errors.py:22-28): the reference prints and aborts COMM_WORLD;
here the C ABI's non-zero return becomes a Python exception (and the NCCL group is torn down)."""
import sys

from .parall import rank


def raiseError(errmsg):
    print("%d - %s" % (rank(), errmsg), file=sys.stderr, flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:
            pass
    raise RuntimeError(errmsg)


def raiseWarning(warnmsg, allranks=False):
    if allranks or rank() == 0:
        print("Warning! %d - %s" % (rank(), warnmsg), file=sys.stderr, flush=True)
