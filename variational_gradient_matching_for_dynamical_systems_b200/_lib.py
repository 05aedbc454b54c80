"""ctypes binding of ``libvgm_b200.so`` (the C ABI declared in ``include/vgm_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  The library itself links the CUDA runtime statically and does not need libcuda at
load time, so it can be *loaded* (symbols inspected) on a machine without a GPU; any compute
call there fails with a CUDA error.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvgm_b200.so")

VGM_OK, VGM_EINVAL, VGM_ECUDA, VGM_ENOTCONV, VGM_ENOTSPD, VGM_EDRIVER, VGM_ENCCL = 0, -1, -2, -3, -4, -5, -6
COMM_ID_BYTES = 128
KERNEL_RBF, KERNEL_RQ = 0, 1
(KERNEL_PERIODIC, KERNEL_LOCALLY_PERIODIC, KERNEL_RBF_LIN, KERNEL_SIGMOID, KERNEL_SIGMOID2, KERNEL_EXP_SIN_SQUARED,
 KERNEL_MATERN) = range(2, 9)


class VGMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvgm_b200 error {code}: {msg}")
        self.code = code


class NotConverged(VGMError):
    pass


c_dp = C.c_void_p  # device pointers travel as integers
c_ip = C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [("A", c_dp), ("B", c_dp), ("C", c_dp),
                ("lda", C.c_int), ("ldb", C.c_int), ("ldc", C.c_int),
                ("N", C.c_int), ("M", C.c_int), ("K", C.c_int),
                ("rowsA", c_ip), ("rowsC", c_ip), ("n_rows_dev", c_ip),
                ("alpha", C.c_double), ("beta", C.c_double), ("C0", c_dp),
                ("U1", c_dp), ("V1", c_dp), ("g1", C.c_double),
                ("U2", c_dp), ("V2", c_dp), ("g2", C.c_double),
                ("dotw", c_dp), ("dot_out", c_dp), ("dot_ld", C.c_int), ("band", C.c_int)]


class SweepPlan(C.Structure):
    _fields_ = [("T", C.c_int), ("ld", C.c_int), ("n_hidden", C.c_int),
                ("slot_state", c_ip), ("pair_ptr", c_ip), ("pair_ode", c_ip), ("pair_code", c_ip),
                ("pair_live", c_ip), ("pair_self", c_ip), ("ode_idx", c_ip), ("ode_state", c_ip),
                ("slot_has_self", c_ip),
                ("n_levels", C.c_int), ("level_ptr_host", C.POINTER(C.c_int)), ("level_slots", c_ip),
                ("n_live", C.c_int), ("live_state", c_ip), ("live_level_ptr_host", C.POINTER(C.c_int)),
                ("n_live_pairs", C.c_int), ("live_pair_slot", c_ip), ("live_pair_src", c_ip),
                ("live_pair_level_ptr_host", C.POINTER(C.c_int)),
                ("ruu_const", C.c_int), ("ruu_value", C.c_double),
                ("n_early", C.c_int), ("early_slots", c_ip), ("bulk_slots", c_ip)]


class Operators(C.Structure):
    _fields_ = [("D", c_dp), ("DT", c_dp), ("M", c_dp), ("G", c_dp), ("G_bf16", c_dp), ("M_bf16x2", c_dp),
                ("G_shift", C.c_double), ("precond_c", C.c_double),
                ("band_D", C.c_int), ("band_M", C.c_int), ("band_M_lp", C.c_int), ("band_G", C.c_int),
                ("DtD", c_dp), ("obs_rhs", c_dp)]


class SolverOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("check_every", C.c_int), ("inner_iters", C.c_int),
                ("optimistic_passes", C.c_int), ("small_level_solver", C.c_int), ("dense_short_grids", C.c_int)]


class SweepInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("unconverged", C.c_int), ("gemm_calls", C.c_int),
                ("kernel_launches", C.c_int)]


class ModelTables(C.Structure):
    _fields_ = [("n_states", C.c_int), ("n_odes", C.c_int), ("n_param", C.c_int),
                ("ode_class", c_ip), ("ode_idx", c_ip), ("ode_state", c_ip),
                ("theta_mode", C.c_int), ("theta_literal", C.c_double * 8),
                ("out_row0", C.c_int), ("out_row_stride", C.c_int), ("out_n_rows", C.c_int),
                ("pipe_chunks", C.c_int), ("pipe_halo_x", C.c_int), ("pipe_halo_dx", C.c_int),
                ("pipe_state_ptr_host", C.POINTER(C.c_int)), ("pipe_slot_ptr_host", C.POINTER(C.c_int)),
                ("pipe_row_ptr_host", C.POINTER(C.c_int)), ("pipe_out_slot_ptr_host", C.POINTER(C.c_int)),
                ("pipe_last_rows", c_ip)]


# name -> (restype, argtypes); every symbol declared in include/vgm_b200.h
_P = C.POINTER
SIGNATURES = {
    "vgm_last_error": (C.c_char_p, []),
    "vgm_version": (C.c_int, []),
    "vgm_shutdown": (C.c_int, []),
    "vgm_kernel_build": (C.c_int, [C.c_int, _P(C.c_double), C.c_int, c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int,
                                   c_dp, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "vgm_dgemm_nt": (C.c_int, [_P(GemmArgs), C.c_void_p]),
    "vgm_dgemm_dot_tiles": (C.c_int, [C.c_int, C.c_int]),
    "vgm_dgemm_set_variant": (C.c_int, [C.c_int]),
    "vgm_band_halfwidth": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, _P(C.c_int), C.c_void_p]),
    "vgm_profile_enable": (C.c_int, [C.c_int]),
    "vgm_profile_collect": (C.c_int, [C.c_int, _P(C.c_double), _P(C.c_double), _P(C.c_double), _P(C.c_longlong)]),
    "vgm_convert_bf16": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, C.c_void_p]),
    "vgm_split_bf16": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, C.c_void_p]),
    "vgm_tc_gemm_bf16": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                   C.c_int, C.c_void_p]),
    "vgm_tc_gemm_bf16_out": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                       C.c_int, C.c_void_p]),
    "vgm_tc_gemm_bf16_out_dots": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                            C.c_int, c_dp, C.c_int, c_dp, C.c_int, C.c_void_p]),
    "vgm_transpose": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vgm_gm_operators_ws_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vgm_gm_operators": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_size_t,
                                   C.c_void_p]),
    "vgm_gm_operators_lu": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, c_dp, C.c_size_t, C.c_void_p]),
    "vgm_spd_inverse": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_size_t, C.c_void_p]),
    "vgm_gp_fit_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vgm_gp_fit": (C.c_int, [c_dp, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, _P(C.c_double), C.c_int, c_dp,
                             c_dp, c_dp, C.c_size_t, C.c_void_p]),
    "vgm_lorenz96_rk4": (C.c_int, [c_dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp, C.c_int,
                                   c_dp, C.c_void_p]),
    "vgm_program_builtin": (C.c_int, [C.c_char_p, _P(C.c_void_p)]),
    "vgm_program_load": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.c_int, _P(C.c_void_p)]),
    "vgm_program_free": (C.c_int, [C.c_void_p]),
    "vgm_program_n_param": (C.c_int, [C.c_void_p]),
    "vgm_param_stats_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vgm_param_stats": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp,
                                  C.c_size_t, C.c_void_p]),
    "vgm_param_loss_ws_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vgm_param_loss": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip, c_dp, c_dp, c_dp,
                                 C.c_size_t, C.c_void_p]),
    "vgm_param_solve": (C.c_int, [c_dp, C.c_int, c_dp, C.c_void_p]),
    "vgm_param_cov_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vgm_param_cov": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_dp, c_dp, C.c_size_t,
                                C.c_void_p]),
    "vgm_state_sweep_ws_bytes": (C.c_size_t, [_P(SweepPlan)]),
    "vgm_state_sweep": (C.c_int, [C.c_void_p, _P(SweepPlan), _P(Operators), _P(SolverOpts), c_dp, c_dp, c_dp, c_dp,
                                  C.c_size_t, _P(SweepInfo), C.c_void_p]),
    "vgm_state_sweep_prepare": (C.c_int, [C.c_void_p, _P(SweepPlan), _P(Operators), c_dp, c_dp, c_dp, c_dp, C.c_size_t,
                                          _P(SweepInfo), C.c_void_p]),
    "vgm_state_sweep_level": (C.c_int, [_P(SweepPlan), _P(Operators), _P(SolverOpts), c_dp, C.c_int, c_dp, C.c_size_t,
                                        _P(SweepInfo), C.c_void_p]),
    "vgm_state_sweep_levels": (C.c_int, [_P(SweepPlan), _P(Operators), _P(SolverOpts), c_dp, c_dp, C.c_size_t,
                                         _P(SweepInfo), C.c_void_p]),
    "vgm_state_assemble": (C.c_int, [C.c_void_p, _P(SweepPlan), c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                     C.c_void_p]),
    "vgm_pcg_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vgm_pcg_solve": (C.c_int, [_P(Operators), C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_int, c_ip, C.c_int, C.c_int,
                                C.c_int, _P(SolverOpts), c_dp, C.c_size_t, _P(SweepInfo), C.c_void_p]),
    "vgm_state_cov_ws_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "vgm_state_cov": (C.c_int, [C.c_void_p, _P(SweepPlan), _P(Operators), c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp,
                                C.c_size_t, C.c_void_p]),
    "vgm_iteration_ws_bytes": (C.c_size_t, [_P(ModelTables), _P(SweepPlan)]),
    "vgm_iteration_device": (C.c_int, [C.c_void_p, _P(ModelTables), _P(SweepPlan), _P(Operators), _P(SolverOpts),
                                       c_dp, c_dp, c_dp, c_dp, c_dp, C.c_size_t, _P(SweepInfo), C.c_void_p]),
    "vgm_iteration_host": (C.c_int, [C.c_void_p, _P(ModelTables), _P(SweepPlan), _P(Operators), _P(SolverOpts),
                                     c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_size_t, _P(SweepInfo),
                                     C.c_void_p]),
    "vgm_comm_nccl_version": (C.c_int, [_P(C.c_int)]),
    "vgm_comm_unique_id": (C.c_int, [C.c_char_p]),
    "vgm_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int, _P(C.c_void_p)]),
    "vgm_comm_free": (C.c_int, [C.c_void_p]),
    "vgm_comm_rank": (C.c_int, [C.c_void_p, _P(C.c_int), _P(C.c_int)]),
    "vgm_comm_allreduce_stats": (C.c_int, [C.c_void_p, c_dp, C.c_int, C.c_void_p]),
    "vgm_comm_halo_ws_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "vgm_comm_halo_exchange": (C.c_int, [C.c_void_p, c_dp, C.c_int, c_ip, C.c_int, C.c_int, c_ip, c_ip, C.c_int, c_dp,
                                         C.c_size_t, C.c_void_p]),
}

_lib = None


def load():
    """Load libvgm_b200.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C variational_gradient_matching_for_dynamical_systems_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    if os.environ.get("VGM_DGEMM_VARIANT"):          # tuning knob, see vgm_dgemm_set_variant
        lib.vgm_dgemm_set_variant(int(os.environ["VGM_DGEMM_VARIANT"]))
    return lib


def check(code, allow_notconv=False):
    if code == VGM_OK:
        return code
    msg = load().vgm_last_error().decode(errors="replace")
    if code == VGM_ENOTCONV:
        if allow_notconv:
            return code
        raise NotConverged(code, msg)
    raise VGMError(code, msg)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("variational_gradient_matching_for_dynamical_systems_b200 needs a CUDA device "
                           "(B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())
