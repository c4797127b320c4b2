"""ctypes loader for libmlb.so, the C-ABI CUDA library (include/mlb.h).

There is NO CPU fallback: importing this module without the built library, or calling a
compute entry without a CUDA device, fails loudly.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MLB_LIB selects an alternative build of the same library (tuning experiments only)
LIB_PATH = os.environ.get("MLB_LIB") or os.path.join(_HERE, "libmlb.so")

# --- enums (keep in sync with include/mlb.h)
MODEL_TOY, MODEL_HIER, MODEL_FN, MODEL_HH = 0, 1, 2, 3
PARAM_UNBOUNDED, PARAM_NORMAL = 0, 1
(SOLVER_QUASI_NEWTON, SOLVER_NEWTON, SOLVER_NEWTON_LS, SOLVER_MICI_NEWTON,
 SOLVER_MICI_QUASI_NEWTON) = range(5)  # fmt: skip
FLAG_SINGLE_KERNEL = 1
FLAG_NUTS_NO_EXTRA_SUBTREE_CHECKS, FLAG_NUTS_TWO_SIDED_DELTA_H = 2, 4
NORM_MAX, NORM_EUCLIDEAN = 0, 1
ST_OK, ST_DIVERGED, ST_NOT_CONVERGED, ST_NON_REVERSIBLE, ST_H_DIVERGED, ST_LINALG = range(6)
ADAPT_ROWS, STAT_ROWS = 5, 8

EXPORTS = (
    "mlb_version", "mlb_last_error", "mlb_model_create", "mlb_model_destroy",
    "mlb_model_dim_u", "mlb_model_dim_y", "mlb_model_dim_q", "mlb_model_jac_rows",
    "mlb_model_gram_rows", "mlb_model_gram_dim", "mlb_constr", "mlb_jacob_constr_blocks",
    "mlb_decompose_gram", "mlb_log_det_sqrt_gram", "mlb_val_and_grad_log_det_sqrt_gram",
    "mlb_lmult_by_jacob_constr", "mlb_rmult_by_jacob_constr",
    "mlb_lmult_by_pinv_jacob_constr", "mlb_lmult_by_inv_jacob_product",
    "mlb_normal_space_component", "mlb_solve_projection", "mlb_neg_log_dens", "mlb_h1",
    "mlb_project_onto_cotangent_space", "mlb_leapfrog_steps", "mlb_hmc_sample",
    "mlb_philox_normals_host", "mlb_philox_uniform_host", "mlb_leapfrog_steps_host",
    "mlb_aos_to_soa", "mlb_soa_to_aos", "mlb_launch_count", "mlb_measure_fp64_tflops",
    "mlb_selftest_fast_math", "mlb_selftest_exp", "mlb_nuts_sample",
    "mlb_tridiagonal_solve", "mlb_tridiagonal_pos_def_log_det",
    "mlb_ssm_decompose_gram", "mlb_ssm_lmult_by_inv_gram",
    "mlb_ssm_lmult_by_inv_jacob_product", "mlb_ssm_log_det_sqrt_gram",
    "mlb_ssm_lmult_by_jacob_constr", "mlb_ssm_rmult_by_jacob_constr",
    "mlb_ssm_sv_constr_blocks", "mlb_ssm_projection_update", "mlb_ssm_band_inv_gram",
    "mlb_ssm_sv_grad_log_det_sqrt_gram",
    "mlb_ssm_project_onto_cotangent_space", "mlb_scratch_trim",
    "mlb_lu_triangular_solve", "mlb_jvp_cholesky_mtx_mult_by_vct", "mlb_dense_cholesky",
    "mlb_gp_decompose_gram", "mlb_gp_lmult_by_inv_gram", "mlb_gp_lmult_by_inv_jacob_product",
    "mlb_gp_jacob_constr_mult", "mlb_gp_log_det_sqrt_gram",
)  # fmt: skip


class ModelDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("dim_y", C.c_int32), ("parametrization", C.c_int32),
                ("n_data", C.c_int32), ("data", C.POINTER(C.c_double))]  # fmt: skip


class SolverOpts(C.Structure):
    _fields_ = [
        ("solver", C.c_int32), ("max_iters", C.c_int32),
        ("max_line_search_iters", C.c_int32), ("norm", C.c_int32),
        ("constraint_tol", C.c_double), ("position_tol", C.c_double),
        ("divergence_tol", C.c_double), ("reverse_check_tol", C.c_double),
        ("reverse_check_norm", C.c_int32), ("flags", C.c_int32),
    ]  # fmt: skip


class AdaptOpts(C.Structure):
    _fields_ = [("adapt_stat_target", C.c_double),
                ("log_step_size_reg_coefficient", C.c_double),
                ("iter_decay_coeff", C.c_double), ("iter_offset", C.c_double)]  # fmt: skip


class MlbError(RuntimeError):
    """Process-level failure reported by libmlb (bad argument, CUDA error, ...)."""


_lib = None


def lib():
    """Load libmlb.so once.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MlbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import "
                "__graft_entry__ as g; g.build()'` (there is no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        L.mlb_last_error.restype = C.c_char_p
        L.mlb_launch_count.restype = C.c_int64
        L.mlb_measure_fp64_tflops.restype = C.c_double
        L.mlb_measure_fp64_tflops.argtypes = [C.c_int32]
        L.mlb_scratch_trim.argtypes = [C.c_size_t]
        L.mlb_philox_uniform_host.restype = C.c_double
        L.mlb_philox_uniform_host.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.mlb_philox_normals_host.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int32,
                                              C.c_void_p]  # fmt: skip
        L.mlb_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
        for name in ("mlb_model_destroy", "mlb_model_dim_u", "mlb_model_dim_y",
                     "mlb_model_dim_q", "mlb_model_jac_rows", "mlb_model_gram_rows",
                     "mlb_model_gram_dim"):  # fmt: skip
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MlbError(f"libmlb error {rc}: {lib().mlb_last_error().decode()}")


def trim_scratch(keep_bytes=0):
    """Hand the cached scratch of the library's own memory pools (tree storage of the
    dynamic sampler, sweep scratch of the ODE / state-space kernels) back to the driver."""
    torch.cuda.synchronize()
    check(lib().mlb_scratch_trim(C.c_size_t(int(keep_bytes))))


def require_cuda():
    if not torch.cuda.is_available():
        raise MlbError("a CUDA device is required (there is no CPU fallback)")


# --- tensor plumbing: chain arrays are [n_chain, dim] VIEWS of contiguous [dim, n_chain]


def soa_empty(n, dim, device="cuda"):
    return torch.empty(dim, n, dtype=torch.float64, device=device).t()


def soa_zeros(n, dim, device="cuda"):
    return torch.zeros(dim, n, dtype=torch.float64, device=device).t()


def to_soa(x, device="cuda"):
    """Any [n, dim] array-like -> fp64 device tensor in the component-major layout."""
    t = torch.as_tensor(x, dtype=torch.float64)
    if t.ndim == 1:
        t = t[None]
    if t.device.type == "cuda" and is_soa(t):
        return t
    out = soa_empty(t.shape[0], t.shape[1], device)
    out.copy_(t)
    return out


def is_soa(t):
    return (t.ndim == 2 and t.dtype == torch.float64 and t.device.type == "cuda"
            and (t.shape[1] == 1 or t.stride(0) == 1) and t.stride(1) >= t.shape[0])  # fmt: skip


def soa_clone(t):
    out = soa_empty(t.shape[0], t.shape[1], t.device)
    out.copy_(t)
    return out


def p_soa(t, rows=None):
    """(pointer, ld) of a component-major chain array; validates layout."""
    if not is_soa(t):
        raise MlbError("expected an fp64 CUDA chain array in component-major layout "
                       f"(got shape {tuple(t.shape)}, strides {t.stride()}, {t.dtype}, "
                       f"{t.device})")  # fmt: skip
    if rows is not None and t.shape[1] != rows:
        raise MlbError(f"expected {rows} components, got {t.shape[1]}")
    ld = t.stride(1) if t.shape[1] > 1 else t.shape[0]
    return C.c_void_p(t.data_ptr()), ld


def p_vec(t, dtype=torch.float64):
    if t is None:
        return None
    if not (t.is_contiguous() and t.dtype == dtype and t.device.type == "cuda"):
        raise MlbError("expected a contiguous CUDA per-chain vector")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
