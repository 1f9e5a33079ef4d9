"""ctypes binding of libpylom_b200.so (C ABI in include/pylom_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible every
entry point raises.  PyTorch is used for device buffers, streams and torch.distributed only.
"""
import ctypes, os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PL_LIBPATH: an instrumented build of the same sources (probes/build_variant.sh); still the CUDA library, never a CPU path
_LIBPATH = os.environ.get("PL_LIBPATH") or os.path.join(_HERE, "libpylom_b200.so")
_lib = None

_i64, _int, _vp, _sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t

_SIGS = {
    "pl_version": (_int, []),
    "pl_last_error": (ctypes.c_char_p, []),
    "pl_launch_count": (_i64, []),
    "pl_temporal_mean_f64": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "pl_subtract_mean_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_center_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_temporal_variance_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_norm_variance_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_qr_factor_var_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "pl_matmul_workspace_bytes": (_sz, [_i64, _i64]),
    "pl_matmul_f64": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "pl_matmul_tn_workspace_bytes": (_sz, [_i64, _i64]),
    "pl_matmul_tn_f64": (_int, [_vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "pl_vecmat_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_widen_f32_f64": (_int, [_vp, _vp, _i64, _vp]),
    "pl_narrow_f64_f32": (_int, [_vp, _vp, _i64, _vp]),
    "pl_complex_embed_f64": (_int, [_vp, _vp, _i64, _i64, _vp]),
    "pl_complex_pack_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pl_rmse_workspace_bytes": (_sz, []),
    "pl_rmse_sums_f64": (_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pl_qr_workspace_bytes": (_sz, [_i64, _i64]),
    "pl_qr_factor_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_qr_apply_q_f64": (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_qr_inplace_rows": (_i64, [_i64, _i64]),
    "pl_qr_workspace_bytes_inplace": (_sz, [_i64, _i64]),
    "pl_qr_factor_inplace_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_qr_apply_q_inplace_f64": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_pod_run_inplace_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_svd_workspace_bytes": (_sz, [_i64]),
    "pl_svd_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "pl_tsqr_svd_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
    "pl_pod_run_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp, _sz, _vp]),
    "pl_reconstruct_f64": (_int, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _sz, _vp]),
    "pl_tsqr_svd_host_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64]),
    "pl_tsqr_host_factor_f64": (_int, [_vp, _vp, _i64, _i64]),
    "pl_tsqr_host_stack_f64": (_int, [_vp, _vp, _vp, _vp, _i64, _i64]),
    "pl_tsqr_host_apply_f64": (_int, [_vp, _vp, _i64, _i64]),
    "pl_get_unique_id": (_int, [_vp]),
    "pl_comm_init_rank": (_int, [ctypes.POINTER(_vp), _vp, _int, _int]),
    "pl_comm_rank": (_int, [_vp]),
    "pl_comm_size": (_int, [_vp]),
    "pl_comm_destroy": (_int, [_vp]),
    "pl_tsqr_svd_dist_workspace_bytes": (_sz, [_vp, _i64, _i64, _int]),
    "pl_tsqr_svd_dist_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _vp, _sz, _vp]),
    "pl_tsqr_svd_host_dist_f64": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64]),
    "pl_host_chunk_rows": (_int, [_i64, _i64, _vp, _int]),
    "pl_host_cache_free": (None, []),
    "pl_profile_enable": (None, [_int]),
    "pl_profile_read": (_int, [_vp, _vp, _int]),
}
EXPORTS = tuple(_SIGS)


def libpath():
    return _LIBPATH


def lib():
    """Load (once) and return the CDLL; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIBPATH):
            raise RuntimeError(
                f"{_LIBPATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C pyloworder_b200/csrc). pyloworder_b200 has no CPU fallback.")
        L = ctypes.CDLL(_LIBPATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pl_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
