"""Drop-in for the reference's ``VGM_modules/proxies_for_ode_parameters_and_states.py``:
the two coordinate-ascent updates with their verbatim signatures, computed on the GPU.

* ``proxy_for_ode_parameters``  <- proxies_for_ode_parameters_and_states.py:41-139
* ``proxy_for_ind_states``      <- proxies_for_ode_parameters_and_states.py:159-375

Scope: ``optimizer='analytical'`` -- the closed-form Gaussian updates -- for both functions.
``proxy_for_ode_parameters(optimizer='minimizer')`` keeps the reference's SciPy optimiser calls
and evaluates the squared gradient-matching loss and its gradient (:379-392) on the device, which also covers models
that are not linear in a parameter; its tolerance is the minimiser's.  The 'minimizer' STATE update (:278-366, SciPy
L-BFGS-B per state) raises ``NotImplementedError``.  Plotting side effects are omitted.

Reference behaviours that are reproduced on purpose:

* the state update evaluates R_uk, r_uk on a snapshot of the proxies taken once per sweep but
  reads ``D x_k`` from the live, already updated columns (:185,:191,:225);
* the state update ignores the estimated ODE parameters and plugs in the literal ``[8]``
  (:211).  ``settings.theta_mode = 'reference'`` (default) does the same for one-parameter
  models; for models with several parameters the shipped reference raises ``TypeError``, so
  the passed ``ode_param_proxy`` is used there (what the one-line-patched reference does) and
  a warning is issued once.  ``settings.theta_mode = 'proxy'`` always uses ``ode_param_proxy``;
* if state u does not occur in its own ODE the diagonal is doubled (:201,:245).
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from .engine import Operators, VGMEngine


class settings:
    """Module-level switches (the reference's signatures have no room for them)."""
    theta_mode = "reference"        # 'reference' | 'proxy'
    theta_literal = 8.0             # proxies...py:211
    tol = 1e-13                     # relative residual of the per-state solves
    solver = "auto"
    schedule = "reference"          # 'reference' | 'colour' (opt-in parallel multi-colour schedule)
    observation_term = False        # True: solve(local_scaling + state_pred_inv_cov, local_mean + state_pred_inv_cov
                                    # . state_pred_mean_u), the statement commented out at proxies...py:265-266
    _warned = False


_engines = {}          # content key -> VGMEngine (the engine holds the model, so nothing in the key can be recycled)


def _fingerprint(M):
    """Content hash of an operator: every byte counts (two kernels that agree in a few entries are different)."""
    import hashlib
    M = np.ascontiguousarray(M, dtype=np.float64)
    return (M.shape, hashlib.blake2b(M.tobytes(), digest_size=16).hexdigest())


def _model_key(model):
    """The model by CONTENT: its text, symbols and replication (id() is recycled after garbage collection)."""
    return (tuple(model.lines), tuple(model.state_names), tuple(model.param_names), model.zero_fill, model.n_copies)


def _couplings_key(couplings):
    return None if couplings is None else tuple(tuple(int(k) for k in lst) for lst in couplings)


def _model_of(locally_linear_odes):
    for holder in (getattr(locally_linear_odes, "ode_param", None), getattr(locally_linear_odes, "state", None)):
        for name in ("B", "b", "R", "r"):
            obj = getattr(holder, name, None) if holder is not None else None
            if obj is not None and hasattr(obj, "model"):
                return obj.model
    raise TypeError("locally_linear_odes must hold the objects returned by this package's "
                    "rewrite_odes_as_linear_combination_in_parameters / _in_states")


def _model_of_any(locally_linear_odes, odes):
    """The model for the 'minimizer' path: from the rewrites when they exist, else from ``odes`` (import_odes); a
    model that is not linear in its parameters has no parameter rewrite but still has a loss."""
    try:
        return _model_of(locally_linear_odes)
    except TypeError:
        model = getattr(odes, "model", None)
        if model is None:
            raise
        return model


def _engine(model, D, A, hidden, couplings, theta_mode, obs_inv_cov=None):
    key = (_model_key(model), _fingerprint(D), tuple(hidden), _couplings_key(couplings), theta_mode,
           float(settings.theta_literal), settings.tol, settings.solver, settings.schedule,
           None if obs_inv_cov is None else _fingerprint(obs_inv_cov))
    eng = _engines.get(key)
    if eng is None:
        while len(_engines) >= 8:
            _engines.pop(next(iter(_engines)))                 # oldest first; the others stay valid (content keys)
        ops = Operators(D, A, obs_inv_cov=obs_inv_cov)
        literal = tuple([float(settings.theta_literal)] * model.p)
        eng = VGMEngine(model, ops, None, hidden=hidden, couplings=couplings, theta_mode=theta_mode,
                        theta_literal=literal, tol=settings.tol, solver=settings.solver,
                        schedule=settings.schedule)
        _engines[key] = eng
    return eng


def _engine_for_parameters(model, D):
    """Any cached engine of this model and operator serves the theta update (it only needs X, D.X and the
    statistics kernels); otherwise one without hidden states is built."""
    mk, fp = _model_key(model), _fingerprint(D)
    for k, e in _engines.items():
        if k[0] == mk and k[1] == fp:
            return e
    return _engine(model, D, None, (), None, "proxy")


def proxy_for_ode_parameters(state_proxy, locally_linear_odes, dC_times_inv_C, eps_cov, ode_param_symbols,
                             ode_param_true, odes, odes_gradient, constraints=None, optimizer='minimizer',
                             fig_shape=(7, 4), init=None):
    '''Estimates proxy for ODE parameters'''
    # error handling, proxies...py:47-50
    if state_proxy.shape[0] != dC_times_inv_C.shape[0]:
        raise ValueError('Either state_proxy or dC_times_invC have the wrong shape')
    elif dC_times_inv_C.shape[0] != dC_times_inv_C.shape[1]:
        raise ValueError('dC_times_invC is not a square matrix')
    if optimizer not in ('analytical', 'minimizer'):
        raise NotImplementedError("optimizer=%r" % (optimizer,))
    model = _model_of(locally_linear_odes) if optimizer == 'analytical' else _model_of_any(locally_linear_odes, odes)
    eng = _engine_for_parameters(model, dC_times_inv_C)          # reuses the sweep engine when there is one
    eng.set_states(np.asarray(state_proxy.values, dtype=np.float64))
    eng.compute_DX()
    if optimizer == 'minimizer':
        # proxies...py:114-128: the reference hands squared_loss / squared_loss_gradient (:379-392) to SciPy.  Same
        # optimiser calls here (BFGS without bounds, L-BFGS-B with them, same start 0.1 or `init`), the loss and its
        # gradient evaluated on the device (vgm_param_loss).  Own tolerance: the minimiser's, not 1e-9.
        from scipy.optimize import minimize
        x0 = 0.1 * np.ones(len(ode_param_symbols)) if init is None else np.asarray(init, dtype=np.float64)
        if constraints not in (None, 'shrinkage', 'nonnegative'):
            raise UnboundLocalError("cannot access local variable 'global_mean' where it is not associated with a value "
                                    "(constraints=%r has no 'minimizer' branch in the reference)" % (constraints,))
        alpha = 0.3 if constraints == 'shrinkage' else 0.0
        cache = {}

        def both(th):
            key = th.tobytes()
            if cache.get("k") != key:
                cache["k"], cache["v"] = key, eng.param_loss(th, alpha)
            return cache["v"]
        kw = {"bounds": tuple([(0, None)] * len(x0))} if constraints == 'nonnegative' else {}
        global_mean = minimize(lambda th: both(th)[0], x0, jac=lambda th: both(th)[1], **kw).x
        return pd.DataFrame(global_mean, columns=['value'], index=map(str, ode_param_symbols)).rename_axis(
            'ODE parameter symbols')
    stats = eng.param_stats().cpu().numpy()
    p = model.p
    local_mean, local_scaling = stats[:p], stats[p:].reshape(p, p)
    if constraints is None:
        global_mean = eng.param_solve()[:p].cpu().numpy()            # proxies...py:93
    else:
        # the constrained variants are scikit-learn fits on the p x p normal equations (:97-109)
        from sklearn.linear_model import Lasso, Ridge
        if constraints == 'nonnegative':
            reg = Lasso(alpha=0.00001, positive=True)
        elif constraints == 'shrinkage':
            reg = Ridge(alpha=1)
        elif constraints == 'sparse':
            reg = Lasso(alpha=1)
        elif constraints == 'sparse+nonnegative':
            reg = Lasso(alpha=1, positive=True)
        else:
            raise ValueError("unknown constraints %r" % (constraints,))
        global_mean = reg.fit(local_scaling, np.squeeze(local_mean)).coef_
    return pd.DataFrame(global_mean, columns=['value'], index=map(str, ode_param_symbols)).rename_axis(
        'ODE parameter symbols')


def covariance_for_ode_parameters(state_proxy, locally_linear_odes, dC_times_inv_C, eps_cov):
    """Opt-in (not in the reference's surface): the p x p matrix ``sum_k B_k^T A B_k`` that
    ``proxy_for_ode_parameters`` computes as ``global_cov`` (proxies...py:79) and never returns."""
    model = _model_of(locally_linear_odes)
    ops = Operators(dC_times_inv_C, eps_cov)
    eng = VGMEngine(model, ops, None, hidden=(), couplings=None, theta_mode="proxy")
    eng.set_states(np.asarray(state_proxy.values, dtype=np.float64))
    return eng.param_cov()


def proxy_for_ind_states(state_proxy, ode_param_proxy, odes, odes_gradient_states, locally_linear_odes,
                         dC_times_inv_C, eps_cov, state_pred_mean, state_pred_inv_cov, observations, true_states,
                         state_couplings, burnin=5, clamp_states_to_observation_fit=True, constraints=None,
                         optimizer='analytical', fig_shape=(10, 8)):
    '''Estimates proxy for each individual state'''
    if optimizer != 'analytical':
        raise NotImplementedError("optimizer=%r: only the closed-form 'analytical' sweep runs on the B200 path "
                                  "(proxies...py:278-366 is SciPy L-BFGS-B, out of scope)" % (optimizer,))
    if constraints is not None:
        raise NotImplementedError("constraints=%r: the scikit-learn Lasso branch (proxies...py:270-272) is out of "
                                  "scope" % (constraints,))
    time_points = np.array(state_proxy.index)
    state_symbols = state_proxy.columns.values
    X = np.asarray(state_proxy.values, dtype=np.float64)
    model = _model_of(locally_linear_odes)
    if clamp_states_to_observation_fit == True:  # noqa: E712  (same test as proxies...py:175)
        obs_cols = set(observations.columns.values)
        hidden = [u for u in range(len(state_symbols)) if state_symbols[u] not in obs_cols]
    else:
        hidden = list(range(len(state_symbols)))
    theta = np.asarray(getattr(ode_param_proxy, "values", ode_param_proxy), dtype=np.float64).ravel()
    theta_mode = settings.theta_mode
    if theta_mode == "reference" and model.p != 1:
        if not settings._warned:
            warnings.warn("the reference plugs the literal [8] into the state update (proxies...py:211) and raises "
                          "TypeError for models with %d parameters; using ode_param_proxy instead" % model.p)
            settings._warned = True
        theta_mode = "proxy"
    own = getattr(state_couplings, "_ptr", None) is not None
    couplings = None if own else state_couplings
    Q = np.asarray(state_pred_inv_cov, dtype=np.float64) if settings.observation_term else None
    eng = _engine(model, dC_times_inv_C, None, tuple(hidden), couplings, theta_mode, obs_inv_cov=Q)
    eng.set_observation_mean(np.asarray(state_pred_mean.values, dtype=np.float64) if settings.observation_term else None)
    eng.set_states(X)
    eng.compute_DX()
    eng.state_sweep(theta_star=None if theta_mode == "reference" else theta)
    global_mean = eng.get_states()
    return pd.DataFrame(global_mean, columns=map(str, state_symbols), index=time_points).rename_axis('time')


def _minimizer_only(*_a, **_k):
    raise NotImplementedError("squared_loss* belong to the SciPy 'minimizer' path (proxies...py:379-434), which is out "
                              "of scope of the B200 hot path")


squared_loss = squared_loss_gradient = squared_loss_states = squared_loss_states_gradient = _minimizer_only
