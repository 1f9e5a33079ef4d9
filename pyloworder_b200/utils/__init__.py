"""Runtime services of the reference's `pyLOM.utils` that the hot path touches
(pyLOM/utils/{mpi,parall,cr,errors,gpu}.py), re-based on torch.distributed / NCCL."""
from .parall import MPI_RANK, MPI_SIZE, worksplit, pprint, mpi_barrier, mpi_reduce, mpi_allgather_rows, init_distributed, is_distributed
from .cr import cr, cr_nvtx, cr_start, cr_stop, cr_info, cr_reset, cr_time
from .errors import raiseError, raiseWarning
from .gpu import gpu_device, gpu_to_cpu, cpu_to_gpu
