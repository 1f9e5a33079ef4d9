"""pyloworder_b200 -- the POD / TSQR-SVD hot path of pyLOM (ArnauMiro/pyLowOrder 3.2.8) rebuilt for
NVIDIA B200 (sm_100a): hand-written CUDA behind a C ABI, same Python names and signatures as
`pyLOM.POD.{run,truncate,reconstruct}` and `pyLOM.math.{tsqr_svd,temporal_mean,subtract_mean,...}`.

    import pyloworder_b200 as pyLOM
    U, S, V = pyLOM.POD.run(X, remove_mean=True)
"""
__version__ = "0.1.0"

from . import utils, vmmath, POD, DMD
from . import vmmath as math
from .utils import pprint, cr_info, gpu_device, gpu_to_cpu, cpu_to_gpu
