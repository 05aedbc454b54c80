"""Drop-in for the reference's ``VGM_modules/import_odes.py`` (same function names and
signatures) with an O(K) front-end.

The reference sympifies, factors and lambdifies every line (``import_odes.py:22-29``) and
finds couplings through K x K membership tests (``:73-87``); at K = 100 000 that is hours of
SymPy.  Here lines are hashed into structural classes first (``codegen.OdeModel``), so the
symbolic work is done once per class and the per-ODE data is an ``int32`` gather table.
The returned objects still honour the reference's calling conventions:

* ``odes(*state_symbols, *param_symbols)`` returns the list of SymPy right-hand sides
  (what ``rewrite_odes_as_local_linear_combinations`` calls, ``rewrite...py:45,108``);
  with numeric arguments it evaluates the right-hand sides;
* ``odes_sym`` is a lazy sequence of factored SymPy expressions.
"""
from __future__ import annotations

from collections.abc import Sequence

import numpy as np
import sympy as sym

from .codegen import OdeModel

ONE = sym.Symbol("one_vector")


class LazyExprList(Sequence):
    """Sequence of the factored right-hand sides; element k is built on first access."""

    def __init__(self, model: OdeModel, state_symbols):
        self.model = model
        self._state_symbols = list(state_symbols)
        self._cache = {}

    def __len__(self):
        return self.model.K

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        if k < 0:
            k += len(self)
        if k not in self._cache:
            m = self.model
            c = m.classes[int(m.ode_class[k])]
            subs = {ph: self._state_symbols[int(m.ode_idx[k, j])] for j, ph in enumerate(c.placeholders)}
            self._cache[k] = c.expr.xreplace(subs)
        return self._cache[k]


class OdeSystem:
    """Callable returned as ``odes``; carries the parsed model for the GPU path."""

    def __init__(self, model: OdeModel, state_symbols, param_symbols):
        self.model = model
        self.state_symbols = list(state_symbols)
        self.param_symbols = list(param_symbols)
        self.odes_sym = LazyExprList(model, state_symbols)
        self._numeric = None

    def __call__(self, *args):
        K, p = self.model.K, self.model.p
        if len(args) != K + p:
            raise TypeError("odes() takes %d positional arguments (%d states + %d parameters) but %d were given"
                            % (K + p, K, p, len(args)))
        if any(isinstance(a, sym.Basic) for a in args):
            subs = dict(zip(self.state_symbols + self.param_symbols, args))
            same = all(k is v or k == v for k, v in subs.items())
            return [e if same else e.xreplace(subs) for e in self.odes_sym]
        if self._numeric is None:
            fs = []
            for c in self.model.classes:
                fs.append(sym.lambdify(list(c.placeholders) + self.model.param_symbols, c.expr, "numpy"))
            self._numeric = fs
        m = self.model
        states, params = args[:K], args[K:]
        out = []
        for k in range(K):
            c = m.classes[int(m.ode_class[k])]
            out.append(self._numeric[int(m.ode_class[k])](*[states[int(m.ode_idx[k, j])] for j in range(c.arity)],
                                                          *params))
        return out


def import_odes(symbols, odes_path):
    '''Imports ODEs from file specified by odes_path (import_odes.py:18-31)'''
    with open(odes_path) as f:
        odes_string = f.read().splitlines()
    model = OdeModel(odes_string, [str(s) for s in symbols.state], [str(s) for s in symbols.param])
    odes = OdeSystem(model, symbols.state, symbols.param)
    return odes, odes.odes_sym


def gradient_of_odes(symbols, odes_sym):
    '''Computes gradient of ODEs (import_odes.py:33-67; only the 'minimizer' optimiser uses it).
    Every derivative is multiplied by the ones vector; exact zeros become eps * ones (:45,:59).'''
    symbols_all = list(symbols.param) + list(symbols.state) + [ONE]
    eps = np.finfo(float).eps

    def vec(e):
        e = e * ONE
        return eps * ONE if e == 0 else e
    d_param = [[vec(o.diff(s)) for s in symbols.param] for o in odes_sym]
    odes_diff_param = sym.lambdify(symbols_all, d_param)
    odes_diff_states = []
    for u in range(len(symbols.state)):
        odes_diff_states.append(sym.lambdify(symbols_all, [vec(o.diff(symbols.state[u])) for o in odes_sym]))
    return odes_diff_param, odes_diff_states


class CouplingLists(Sequence):
    """``state_couplings`` as a lazy list of lists (CSR underneath), O(K) memory."""

    def __init__(self, ptr, idx):
        self._ptr, self._idx = ptr, idx

    def __len__(self):
        return len(self._ptr) - 1

    def __getitem__(self, u):
        if isinstance(u, slice):
            return [self[i] for i in range(*u.indices(len(self)))]
        if u < 0:
            u += len(self)
        return self._idx[self._ptr[u]:self._ptr[u + 1]].tolist()

    def __eq__(self, other):
        try:
            return len(self) == len(other) and all(self[i] == list(other[i]) for i in range(len(self)))
        except TypeError:
            return NotImplemented


def find_state_couplings_in_odes(odes_sym, symbols):
    '''Find Couplings between States across ODEs (import_odes.py:73-87).
    Returns (state_couplings, odes_couplings_to_states): c(u) = ODEs whose RHS contains x_u, and
    for every ODE the list of state symbols it contains.'''
    if isinstance(odes_sym, LazyExprList):
        m = odes_sym.model
        uu, kk, _ = m.pairs()
        ptr = np.zeros(m.K + 1, dtype=np.int64)
        np.add.at(ptr, uu + 1, 1)
        state_couplings = CouplingLists(np.cumsum(ptr), kk)
        # per ODE: the states it contains, in the order of symbols.state
        order = np.lexsort((uu, kk))
        ptr2 = np.zeros(m.K + 1, dtype=np.int64)
        np.add.at(ptr2, kk + 1, 1)
        ptr2 = np.cumsum(ptr2)
        st = list(symbols.state)
        us = uu[order]

        class _Rows(Sequence):
            def __len__(self):
                return m.K

            def __getitem__(self, k):
                return [st[int(u)] for u in us[ptr2[k]:ptr2[k + 1]]]
        return state_couplings, _Rows()
    odes_couplings_to_states = []
    for u in range(len(symbols.state)):
        free = odes_sym[u].free_symbols
        odes_couplings_to_states.append([s for s in symbols.state if s in free])
    state_couplings = [[i for i, item in enumerate(odes_couplings_to_states) if s in item] for s in symbols.state]
    return state_couplings, odes_couplings_to_states


def ismember(a, b):
    '''import_odes.py:93-97'''
    bind = {}
    for elt in b:
        if elt not in bind:
            bind[elt] = 1
    return [bind.get(itm, None) for itm in a]
