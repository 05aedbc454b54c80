"""Drop-in surface for the reference's ``VGM_modules/kernels.py``.

The reference file is a verbatim copy of scikit-learn's ``gaussian_process/kernels.py`` that
nothing imports (and that cannot be imported as shipped: relative imports at ``kernels.py:36-39``).
The arithmetic is scikit-learn's, so this module re-exports the installed scikit-learn kernel
classes under the same names (identical constructors, ``__call__(X, Y=None, eval_gradient=False)``,
``diag``, ``is_stationary``, ``theta``/``bounds``, ``+ * **`` operators) and overrides
``RBF.__call__`` and ``RationalQuadratic.__call__`` for one-dimensional inputs so that they hit
the same fused CUDA kernel as ``GP_regression.kernel_function`` (``vgm_kernel_build``,
RBF: kernels.py:1030-1130, RationalQuadratic: kernels.py:1270-1367).  Anything the CUDA kernel
does not cover (gradients w.r.t. hyper-parameters, anisotropic or multi-dimensional inputs)
defers to scikit-learn's own implementation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from sklearn.gaussian_process import kernels as _sk
from sklearn.gaussian_process.kernels import (CompoundKernel, ConstantKernel, DotProduct, ExpSineSquared,  # noqa: F401
                                              Exponentiation, Hyperparameter, Kernel, KernelOperator, Matern,
                                              NormalizedKernelMixin, PairwiseKernel, Product,
                                              StationaryKernelMixin, Sum, WhiteKernel)

from . import _lib


def _gpu_gram(kernel_type, params, X, Y):
    import torch
    lib = _lib.load()
    dev = _lib.require_cuda()
    x = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float64).ravel(), device=dev)
    par = (C.c_double * len(params))(*[float(v) for v in params])
    n = x.numel()
    if Y is None:
        ld = (n + 7) // 8 * 8
        out = torch.zeros((n, ld), dtype=torch.float64, device=dev)
        _lib.check(lib.vgm_kernel_build(kernel_type, par, len(params), x.data_ptr(), n, None, 0, ld, 8,
                                        out.data_ptr(), None, None, None, None, _lib.stream_ptr()))
        return out[:, :n].cpu().numpy()
    y = torch.as_tensor(np.ascontiguousarray(Y, dtype=np.float64).ravel(), device=dev)
    m = y.numel()
    ld, ld2 = (n + 7) // 8 * 8, (m + 7) // 8 * 8
    out = torch.zeros((n, ld2), dtype=torch.float64, device=dev)
    _lib.check(lib.vgm_kernel_build(kernel_type, par, len(params), x.data_ptr(), n, y.data_ptr(), m, ld, ld2,
                                    None, None, None, None, out.data_ptr(), _lib.stream_ptr()))
    return out[:, :m].cpu().numpy()


def _one_dim(X, Y):
    X = np.atleast_2d(X)
    if X.ndim != 2 or X.shape[1] != 1:
        return False
    if Y is not None:
        Y = np.atleast_2d(Y)
        if Y.ndim != 2 or Y.shape[1] != 1:
            return False
    return True


class RBF(_sk.RBF):
    """k(x, y) = exp(-|x - y|^2 / (2 l^2)); 1-D isotropic inputs run on the GPU."""

    def __call__(self, X, Y=None, eval_gradient=False):
        if eval_gradient or self.anisotropic or not _one_dim(X, Y):
            return super().__call__(X, Y, eval_gradient)
        return _gpu_gram(_lib.KERNEL_RBF, [float(np.squeeze(self.length_scale)), 1.0], X, Y)


class RationalQuadratic(_sk.RationalQuadratic):
    """k(x, y) = (1 + |x - y|^2 / (2 alpha l^2))^-alpha; 1-D inputs run on the GPU."""

    def __call__(self, X, Y=None, eval_gradient=False):
        if eval_gradient or not _one_dim(X, Y):
            return super().__call__(X, Y, eval_gradient)
        return _gpu_gram(_lib.KERNEL_RQ, [float(self.alpha), float(self.length_scale), 1.0], X, Y)
