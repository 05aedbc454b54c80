"""Drop-in for the reference's ``VGM_modules/rewrite_odes_as_local_linear_combinations.py``.

Same two functions, same positional arguments, same return shapes -- but the returned
"lists of callables" are thin views on an :class:`~.codegen.OdeModel` whose coefficient
expressions are compiled to CUDA device functions (NVRTC, sm_100a) when the proxies are
evaluated; ``proxies_for_ode_parameters_and_states`` never calls them element by element.

* ``B[k](*rows_of_[X|1]^T)`` -> array (1, p, T),  ``b[k](...)`` -> array (1, 1, T)
  (rewrite...py:55-56; symbol-free entries REPLACED by the ones vector, :49-52);
* ``R[u][k](*theta, *rows, ones)`` / ``r[u][k](...)`` -> arrays (1, 1, T), ``[]`` where ODE k does
  not contain state u (rewrite...py:101-119; symbol-free entries MULTIPLIED by ones, :112-115).

Calling one of these objects evaluates the SymPy expression with NumPy on the host; that is
an inspection/compatibility helper, not the compute path.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np
import sympy as sym

from .codegen import OdeModel

ONE = sym.Symbol("one_vector")


def _model_from(odes, state_symbols, ode_param_symbols, zero_fill=False):
    model = getattr(odes, "model", None)
    if isinstance(model, OdeModel) and model.zero_fill == bool(zero_fill) and model.K == len(state_symbols):
        return model
    if isinstance(model, OdeModel):
        return OdeModel(model.lines, model.state_names, model.param_names, zero_fill=zero_fill)
    # foreign callable (e.g. the reference's lambdified odes): evaluate it on the symbols,
    # exactly as rewrite...py:45 does
    exprs = odes(*state_symbols, *ode_param_symbols)
    return OdeModel([str(e) for e in exprs], [str(s) for s in state_symbols], [str(s) for s in ode_param_symbols],
                    zero_fill=zero_fill)


class _HostCallable:
    def __init__(self, args, exprs, shape):
        self._args, self._exprs, self._shape = args, exprs, shape
        self._f = None

    @property
    def expr(self):
        return self._exprs

    def __call__(self, *vals):
        if self._f is None:
            self._f = sym.lambdify(self._args, self._exprs, "numpy")
        out = self._f(*vals)
        T = np.size(vals[-1])                      # the ones vector is always the last argument
        items = [x for row in out for x in row]
        return np.stack([np.broadcast_to(np.asarray(v, dtype=float), (T,)) for v in items]).reshape(self._shape + (T,))


class ParamCoefficients(Sequence):
    """``B`` or ``b`` of the parameter rewrite: one entry per ODE."""

    def __init__(self, model: OdeModel, which, state_symbols):
        self.model, self.which = model, which
        self._state_symbols = list(state_symbols)

    def __len__(self):
        return self.model.K

    def __getitem__(self, k):
        m = self.model
        c = m.classes[int(m.ode_class[k])]
        subs = {ph: self._state_symbols[int(m.ode_idx[k, j])] for j, ph in enumerate(c.placeholders)}
        args = self._state_symbols + [ONE]
        vec = lambda e: e.xreplace(subs) if len(e.free_symbols) else e * ONE
        if self.which == "B":
            return _HostCallable(args, [[vec(e) for e in c.B]], (1, m.p))
        return _HostCallable(args, [[vec(c.b)]], (1, 1))


class _StateRow(Sequence):
    def __init__(self, parent, u):
        self.parent, self.u = parent, u

    def __len__(self):
        return self.parent.model.K

    def __getitem__(self, k):
        return self.parent._entry(self.u, k)


class StateCoefficients(Sequence):
    """``R`` or ``r`` of the state rewrite: ``R[u][k]`` (lazy K x K table)."""

    def __init__(self, model: OdeModel, which, state_symbols, param_symbols, couplings=None):
        self.model, self.which = model, which
        self._state_symbols, self._param_symbols = list(state_symbols), list(param_symbols)
        self.couplings = couplings

    def __len__(self):
        return self.model.K

    def __getitem__(self, u):
        return _StateRow(self, u)

    def _entry(self, u, k):
        m = self.model
        c = m.classes[int(m.ode_class[k])]
        row = m.ode_idx[k, :c.arity].tolist()
        if self.couplings is not None and k not in list(self.couplings[u]):
            return []
        if u not in row or not c.present[row.index(u)]:
            return []
        j = row.index(u)
        if c.nonlinear[j]:
            raise ValueError("ODE %d is not linear in state %s" % (k, m.state_names[u % m.K1]))
        e = c.R[j] if self.which == "R" else c.r[j]
        subs = {ph: self._state_symbols[int(m.ode_idx[k, jj])] for jj, ph in enumerate(c.placeholders)}
        e = e.xreplace(subs)
        if len(e.free_symbols) == 0:
            e = e * ONE                                            # rewrite...py:112-115
        return _HostCallable(self._param_symbols + self._state_symbols + [ONE], [[e]], (1, 1))


def rewrite_odes_as_linear_combination_in_parameters(odes, state_symbols, ode_param_symbols, zero_fill=False):
    '''Rewrites each ODE as a Linear Combination in all ODE Parameters (rewrite...py:28-58).

    ``zero_fill=True`` (extension) keeps symbol-free coefficients as constants instead of
    replacing them by the ones vector -- the mathematically intended rewrite of the MATLAB
    original; the default reproduces the Python reference as written.'''
    state_symbols = list(state_symbols)            # never mutate the caller's list (rewrite...py:36)
    model = _model_from(odes, state_symbols, ode_param_symbols, zero_fill=zero_fill)
    model.require_linear_in_parameters()           # sympy NonlinearError, as rewrite...py:45 raises it
    return ParamCoefficients(model, "B", state_symbols), ParamCoefficients(model, "b", state_symbols)


def rewrite_odes_as_linear_combination_in_states(odes, state_symbols, ode_param_symbols, observed_states,
                                                 state_couplings, clamp_states_to_observation_fit=1):
    '''Rewrite each ODE as a Linear Combination in an Individual State (rewrite...py:76-124).
    As in the reference the clamp flag does not restrict the rewrite (its hidden-state list at
    :92-96 is never used); entries exist for every u and every k in ``state_couplings[u]``.'''
    state_symbols = list(state_symbols)
    model = _model_from(odes, state_symbols, ode_param_symbols)
    own = getattr(state_couplings, "_ptr", None) is not None      # produced by our find_state_couplings_in_odes
    couplings = None if own else state_couplings
    if couplings is not None:
        # like the reference, fail at rewrite time if a listed pair is not linear in x_u
        for u, lst in enumerate(couplings):
            for k in lst:
                c = model.classes[int(model.ode_class[k])]
                row = model.ode_idx[k, :c.arity].tolist()
                if u in row and c.nonlinear[row.index(u)]:
                    raise ValueError("ODE %d is not linear in state %s" % (k, model.state_names[u % model.K1]))
    R = StateCoefficients(model, "R", state_symbols, ode_param_symbols, couplings)
    r = StateCoefficients(model, "r", state_symbols, ode_param_symbols, couplings)
    return R, r
