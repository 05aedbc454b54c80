"""TEST INFRASTRUCTURE ONLY -- live import of the *unmodified* reference modules.

This file is part of ``oracle/`` and must never be imported by the product package
(``variational_gradient_matching_for_dynamical_systems_b200``).  Only ``tests/``,
``oracle/make_golden.py``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline legs
may touch ``oracle/``.

What it does
------------
The reference (``/root/reference/VGM_modules``) cannot be imported as shipped in this
image:

* ``matplotlib`` is not installed (imported at ``proxies_for_ode_parameters_and_states.py:11,22-23``,
  ``GP_regression.py:12``, ``plotting.py:6``);
* ``import scipy.optimize.nnls as nnls`` (``proxies_for_ode_parameters_and_states.py:16``) is
  not a module in SciPy >= 1.12;
* pandas 3 returns a read-only ``DataFrame.values`` so the in-place writes at
  ``proxies_for_ode_parameters_and_states.py:185,262`` raise on the second sweep.

``load_reference()`` therefore (i) installs no-op stub modules for matplotlib / plotting /
``scipy.optimize.nnls`` in ``sys.modules``, (ii) puts ``<ref>/VGM_modules`` on ``sys.path``
and imports the reference modules *unchanged*, and (iii) offers ``WritableFrame``, a
``pandas.DataFrame`` subclass whose ``.values`` is a writable array that aliases nothing in
pandas (so the reference's live-update aliasing ``global_mean = state_proxy[:]`` behaves as
it did under the author's pandas).

Two variants of the proxies module are provided:

* ``strict``  -- the file exactly as shipped (hard-coded ``ode_param_proxy = [8]`` at
  ``proxies_for_ode_parameters_and_states.py:211``);
* ``patched`` -- the same source with exactly that one line replaced by ``pass`` so that
  multi-parameter models run with the ODE-parameter proxy that was passed in.

The reference lives only in the build container; on the GPU box this module raises
``ReferenceUnavailable`` and tests fall back to ``tests/golden/*.npz``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import pandas as pd

REFERENCE_DIR = os.environ.get("VGM_REFERENCE_DIR", "/root/reference")
_MODULES_DIR = os.path.join(REFERENCE_DIR, "VGM_modules")


class ReferenceUnavailable(RuntimeError):
    pass


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_MODULES_DIR, "proxies_for_ode_parameters_and_states.py"))


class WritableFrame(pd.DataFrame):
    """DataFrame whose ``.values`` is a persistent *writable* ndarray.

    The reference mutates ``state_proxy.values`` in place
    (``proxies_for_ode_parameters_and_states.py:185,262``); pandas >= 3 hands out read-only
    views.  We keep one private C-contiguous array and return it from ``.values`` so that
    the aliasing semantics (live Gauss-Seidel reads at ``:225``) are those of the author's
    environment.
    """

    _metadata = ["_vgm_values"]

    @property
    def _constructor(self):
        return pd.DataFrame

    @property
    def values(self):  # type: ignore[override]
        v = getattr(self, "_vgm_values", None)
        if v is None:
            v = np.array(pd.DataFrame.to_numpy(self), dtype=float, copy=True)
            object.__setattr__(self, "_vgm_values", v)
        return v


def writable_frame(values, index, columns, axis_name="time") -> WritableFrame:
    df = WritableFrame(np.array(values, dtype=float, copy=True), index=index, columns=list(columns))
    df = WritableFrame(df.rename_axis(axis_name))
    object.__setattr__(df, "_vgm_values", np.array(values, dtype=float, copy=True))
    return df


def _noop(*_a, **_k):
    return None


class _Anything(types.ModuleType):
    """Stub module: every attribute is a no-op callable (that returns another stub)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _noop


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image", "matplotlib.cm",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]
    # ``from plotting import *`` -- the reference calls these from inside every hot-path
    # function (proxies...py:137,373; GP_regression.py:49,312).  Pure side effects.
    plotting = types.ModuleType("plotting")
    for fn in ("plot_ode_parameters", "plot_states", "plot_states2", "plot_trajectories"):
        setattr(plotting, fn, _noop)
    plotting.__all__ = ["plot_ode_parameters", "plot_states", "plot_states2", "plot_trajectories"]
    sys.modules["plotting"] = plotting
    # ``import scipy.optimize.nnls as nnls`` needs a *module* of that name.
    import scipy.optimize as so
    if "scipy.optimize.nnls" not in sys.modules:
        m = types.ModuleType("scipy.optimize.nnls")
        m.nnls = getattr(so, "nnls", _noop)
        sys.modules["scipy.optimize.nnls"] = m


_CACHE: dict = {}


def load_reference():
    """Return a namespace with the live reference modules.

    Attributes: ``import_odes``, ``GP_regression``, ``rewrite`` (rewrite_odes_as_local_linear_combinations),
    ``proxies_strict``, ``proxies_patched``.
    """
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise ReferenceUnavailable(f"reference not found under {_MODULES_DIR}")
    _install_stubs()
    if _MODULES_DIR not in sys.path:
        sys.path.insert(0, _MODULES_DIR)
    ns = types.SimpleNamespace()
    ns.import_odes = importlib.import_module("import_odes")
    ns.GP_regression = importlib.import_module("GP_regression")
    ns.rewrite = importlib.import_module("rewrite_odes_as_local_linear_combinations")
    ns.proxies_strict = importlib.import_module("proxies_for_ode_parameters_and_states")
    # patched variant: one-line edit of the shipped source, exec'd into a fresh module
    src_path = os.path.join(_MODULES_DIR, "proxies_for_ode_parameters_and_states.py")
    with open(src_path) as f:
        src = f.read()
    needle = "                ode_param_proxy = [8]\n"
    if src.count(needle) != 1:
        raise ReferenceUnavailable("expected exactly one live 'ode_param_proxy = [8]' line")
    patched = types.ModuleType("proxies_for_ode_parameters_and_states_patched")
    patched.__file__ = src_path
    exec(compile(src.replace(needle, "                pass\n"), src_path, "exec"), patched.__dict__)
    ns.proxies_patched = patched
    # observation-term variant: the reference's OWN commented-out statement (:265-266) put in the place of the
    # active solve (:262); nothing else changes (the literal [8] stays, so it serves one-parameter models)
    active = "                global_mean[:,u] = np.squeeze(np.linalg.solve(local_scaling,local_mean)) \n"
    commented = ("#                global_mean[:,u] = np.squeeze(np.linalg.solve(local_scaling + state_pred_inv_cov,\\\n"
                 "#                           local_mean + np.dot(state_pred_inv_cov,state_pred_mean.iloc[:,u])))\n")
    if src.count(active) == 1 and src.count(commented) >= 1:
        stmt = ("                global_mean[:,u] = np.squeeze(np.linalg.solve(local_scaling + state_pred_inv_cov,"
                "local_mean + np.dot(state_pred_inv_cov,state_pred_mean.iloc[:,u])))\n")
        obs = types.ModuleType("proxies_for_ode_parameters_and_states_obs_term")
        obs.__file__ = src_path
        exec(compile(src.replace(active, stmt), src_path, "exec"), obs.__dict__)
        ns.proxies_obs_term = obs
    else:
        ns.proxies_obs_term = None
    _CACHE["ns"] = ns
    return ns


class Symbols:
    """Stand-in for the reference's ``symbols`` namespace class (``*_declarations.py``)."""

    def __init__(self, state, param):
        self.state = list(state)
        self.param = list(param)


class LocallyLinearODEs:
    """Stand-in for the ``locally_linear_odes`` namespace (``Lorenz96_declarations.py:44-50``)."""

    class _NS:
        pass

    def __init__(self):
        self.ode_param = self._NS()
        self.state = self._NS()
        self.ode_param.B = None
        self.ode_param.b = None
        self.state.R = None
        self.state.r = None
