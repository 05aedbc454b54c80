"""Drop-in for the reference's ``VGM_modules/GP_regression.py`` (same function names,
positional order, keyword names and defaults), computed on the GPU through libvgm_b200.

* ``kernel_function``            <- GP_regression.py:160-319
* ``fitting_state_observations`` <- GP_regression.py:33-51

Differences that are deliberate and documented:
the two ``np.random.multivariate_normal`` draws that only feed a plot (:309-312) and all
matplotlib side effects are omitted; ``D = dC C^-1`` is obtained from a Cholesky
factorisation on the device instead of LAPACK ``gesv`` (agreement ~ cond(C) * eps).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

# kernel_type strings of GP_regression.py:176-231
_KTYPES = {"rbf": _lib.KERNEL_RBF, "RQ": _lib.KERNEL_RQ, "periodic": _lib.KERNEL_PERIODIC,
           "locally_periodic": _lib.KERNEL_LOCALLY_PERIODIC, "rbf+lin": _lib.KERNEL_RBF_LIN,
           "sigmoid": _lib.KERNEL_SIGMOID, "sigmoid2": _lib.KERNEL_SIGMOID2,
           "exp_sin_squared": _lib.KERNEL_EXP_SIN_SQUARED, "matern": _lib.KERNEL_MATERN}


def _round_up(x, m):
    return (x + m - 1) // m * m


def kernel_function_device(time_points, time_points2, kernel_type="RQ", kernel_param=[0.1, 2, 5], batch_params=None,
                           want_inverse=False):
    """Device-resident version: returns a dict of padded CUDA tensors
    ``D, A, C, dC, ddC, Cobs, Cso`` (+ ``Cinv``), each ``[rows, ld]`` with ``ld`` = columns
    rounded up to 8, plus ``T``/``T2``."""
    if kernel_type not in _KTYPES:
        # the reference falls through its if/elif chain and fails on the unassigned name at :232
        raise UnboundLocalError("cannot access local variable 'kernel' where it is not associated with a value "
                                "(unknown kernel_type %r)" % (kernel_type,))
    if kernel_type == "matern" and float(kernel_param[1]) not in (0.5, 1.5, 2.5):
        raise NotImplementedError("matern with nu=%r: the reference's general-nu branch (GP_regression.py:224-230) "
                                  "indexes a SymPy expression and cannot run" % (kernel_param[1],))
    lib = _lib.load()
    dev = _lib.require_cuda()
    t = torch.as_tensor(np.asarray(time_points, dtype=np.float64), device=dev)
    t2 = torch.as_tensor(np.asarray(time_points2, dtype=np.float64), device=dev)
    T, T2 = t.numel(), t2.numel()
    ld, ld2 = _round_up(max(T, 1), 8), _round_up(max(T2, 1), 8)
    z = lambda r, l: torch.zeros((max(r, 1), l), dtype=torch.float64, device=dev)
    out = dict(T=T, T2=T2, ld=ld, ld2=ld2, C=z(T, ld), dC=z(T, ld), ddC=z(T, ld), Cobs=z(T2, ld2), Cso=z(T, ld2),
               D=z(T, ld), A=z(T, ld))
    par = (C.c_double * len(kernel_param))(*[float(v) for v in kernel_param])
    s = _lib.stream_ptr()
    _lib.check(lib.vgm_kernel_build(_KTYPES[kernel_type], par, len(kernel_param), t.data_ptr(), T, t2.data_ptr(), T2,
                                    ld, ld2, out["C"].data_ptr(), out["dC"].data_ptr(), out["ddC"].data_ptr(),
                                    out["Cobs"].data_ptr(), out["Cso"].data_ptr(), s))
    n = lib.vgm_gm_operators_ws_bytes(T, ld)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    cinv = z(T, ld) if want_inverse else None
    rc = lib.vgm_gm_operators(out["C"].data_ptr(), out["dC"].data_ptr(), out["ddC"].data_ptr(), T, ld, 1,
                              out["D"].data_ptr(), out["A"].data_ptr(), _lib.ptr(cinv), ws.data_ptr(), n, s)
    if rc == _lib.VGM_ENOTSPD and not want_inverse:
        # indefinite C ('sigmoid2', the reference's Matern 1.5/2.5 on the squared distance): LU with partial
        # pivoting, which is what the reference's np.linalg.solve does for every kernel (GP_regression.py:306)
        rc = lib.vgm_gm_operators_lu(out["C"].data_ptr(), out["dC"].data_ptr(), out["ddC"].data_ptr(), T, ld,
                                     out["D"].data_ptr(), out["A"].data_ptr(), ws.data_ptr(), n, s)
        out["factorisation"] = "lu"
    _lib.check(rc)
    if want_inverse:
        out["Cinv"] = cinv
    return out


def kernel_function(time_points, time_points2, kernel_type='RQ', kernel_param=[0.1, 2, 5]):
    '''Populates the GP covariance matrix and it's derivatives (GP_regression.py:160).
    Returns ``cov_diff_times_inv_cov, eps_cov, cov, cov_obs, cov_state_obs`` as numpy arrays.'''
    # error handling, verbatim behaviour of GP_regression.py:168-171
    if type(time_points) is not np.ndarray:
        raise ValueError('time points is not a numpy array')
    elif len(time_points.shape) > 1:
        raise ValueError('time points must be a one dimensional numpy array')
    if kernel_type in ("periodic", "locally_periodic", "exp_sin_squared"):
        # sqrt((t0-t1)**2) makes the reference's derivative grids NaN on the diagonal; its plot-only
        # multivariate_normal draw (:310-311) then fails.  kernel_function_device still builds the grids.
        raise np.linalg.LinAlgError("SVD did not converge")
    o = kernel_function_device(time_points, np.asarray(time_points2, dtype=float), kernel_type, kernel_param)
    T, T2 = o["T"], o["T2"]
    c = lambda m, r, n: m[:r, :n].cpu().numpy()
    return c(o["D"], T, T), c(o["A"], T, T), c(o["C"], T, T), c(o["Cobs"], T2, T2), c(o["Cso"], T, T2)


class LazyPseudoInverse:
    """``state_pred_inv_cov`` of ``fitting_state_observations``: the Moore-Penrose inverse of the predictive
    covariance (GP_regression.py:46), computed on first use.  The analytical updates never read it
    (proxies_for_ode_parameters_and_states.py:265-266 is commented out), so the default path does not pay for a
    T x T SVD on the host; ``np.asarray(x)`` (or any indexing / arithmetic through NumPy) materialises it with
    ``np.linalg.pinv``, exactly what the reference calls.  ``pred_cov`` is the matrix it inverts."""

    def __init__(self, pred_cov):
        self.pred_cov = np.asarray(pred_cov, dtype=np.float64)
        self._pinv = None

    shape = property(lambda self: self.pred_cov.shape)
    dtype = property(lambda self: self.pred_cov.dtype)
    ndim = 2

    def __array__(self, dtype=None, copy=None):
        if self._pinv is None:
            self._pinv = np.linalg.pinv(self.pred_cov)           # GP_regression.py:46
        return self._pinv if dtype is None else self._pinv.astype(dtype)

    def __getitem__(self, idx):
        return self.__array__()[idx]

    def __len__(self):
        return self.pred_cov.shape[0]

    def dot(self, other):
        return self.__array__().dot(other)

    def __matmul__(self, other):
        return self.__array__() @ other

    def __rmatmul__(self, other):
        return other @ self.__array__()


def fitting_state_observations(observations, hidden_states, SNR, time_points, cov, cov_obs, cov_func_obs,
                               fig_shape=(10, 8)):
    """GP fit of the observed states = initial state proxy (GP_regression.py:33-51).

    Reproduces the reference's accumulating-noise behaviour (``cov_obs`` is reassigned inside
    the loop at :42, so observed state i sees sum_{j<=i} sigma_j^2), leaves unobserved states
    at zero (:38), adds a column for an observed state that is not in ``hidden_states`` (as the reference's
    ``state_pred_mean[state] = ...`` does) and returns ``(state_pred_mean DataFrame T x K, state_pred_inv_cov)``.
    Everything runs in ``vgm_gp_fit``: Cholesky SOLVES of cov_obs (no explicit inverse) and DMMA GEMMs;
    ``state_pred_inv_cov`` is a :class:`LazyPseudoInverse` (the SVD-based pinv of :46 only on request)."""
    import pandas as pd
    lib = _lib.load()
    dev = _lib.require_cuda()
    Y = np.array(observations.values, dtype=np.float64, copy=True)
    obs_variance = (np.mean(Y, axis=0) / SNR) ** 2 if Y.shape[1] else np.zeros(0)
    T, T2, n_obs = len(time_points), np.asarray(cov_obs).shape[0], Y.shape[1]
    names = list(map(str, hidden_states))
    ld, ld2 = _round_up(T, 8), _round_up(T2, 8)

    def padded(M, rows, cols, l):
        out = torch.zeros((max(rows, 1), l), dtype=torch.float64, device=dev)
        out[:rows, :cols] = torch.as_tensor(np.array(M, dtype=np.float64, copy=True), device=dev)
        return out
    Co, Cso, Cm = padded(cov_obs, T2, T2, ld2), padded(cov_func_obs, T, T2, ld2), padded(cov, T, T, ld)
    Yd = padded(Y.T, n_obs, T2, ld2)
    mean_d = torch.zeros((max(n_obs, 1), ld), dtype=torch.float64, device=dev)
    pc_d = torch.zeros((T, ld), dtype=torch.float64, device=dev)
    n = lib.vgm_gp_fit_ws_bytes(T, T2, ld2)
    ws = torch.empty(n + 256, dtype=torch.uint8, device=dev)
    var = (C.c_double * max(n_obs, 1))(*[float(v) for v in obs_variance])
    _lib.check(lib.vgm_gp_fit(Co.data_ptr(), Cso.data_ptr(), Cm.data_ptr(), T, T2, ld, ld2, Yd.data_ptr(), var, n_obs,
                              mean_d.data_ptr(), pc_d.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr()))
    fit = mean_d[:n_obs, :T].cpu().numpy()
    cols = list(names)
    mean = np.zeros((T, len(names)))
    for i, state in enumerate(list(observations.columns)):
        key = str(state)
        if key not in cols:                     # the reference's column assignment creates it
            cols.append(key)
            mean = np.concatenate([mean, np.zeros((T, 1))], axis=1)
        mean[:, cols.index(key)] = fit[i]
    state_pred_mean = pd.DataFrame(mean, columns=cols, index=time_points).rename_axis('time')
    return state_pred_mean, LazyPseudoInverse(pc_d[:, :T].cpu().numpy())
