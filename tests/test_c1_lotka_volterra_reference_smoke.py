"""Config C1 of BASELINE.json ("Lotka-Volterra (2 states, 4 params) via VGM_for_Lotka_Volterra.py on CPU, reference
defaults"): an end-to-end CPU smoke of the LIVE reference through the oracle harness.

The shipped .py driver does not run against the shipped modules; its up-to-date form is the notebook,
so this test replays the notebook's call sequence (VGM_for_Lotka_Volterra.ipynb cells 25-41) with the user inputs of
VGM_for_Lotka_Volterra.py:42-107 / the notebook cells 6-23 -- theta = [2, 1, 4, 1], x0 = [5, 3], both states observed
at t = [0, .1, .5, .6, .8, 1, 1.4, 1.6, 2], for_estimation = arange(0, 4, 0.08) (T = 50), kernel_function defaults
(RQ [0.1, 2, 5]), 'minimizer' optimizer with the 'nonnegative' parameter constraint, 10 iterations -- and then the
closed-form ('analytical') parameter update this repository's hot path replaces, on the same proxies.
Data generation is done here (seeded) instead of setup_simulation, which writes ./data/*.csv and draws unseeded noise.
Runs only where /root/reference exists (the build container)."""
import os
import tempfile

import numpy as np
import pandas as pd
import pytest
import sympy as sym

from oracle import reference_shim as rs
from oracle import vgm_oracle as vo

pytestmark = pytest.mark.skipif(not rs.reference_available(), reason="live reference not present")


def _simulate(theta, x0, t):
    from scipy.integrate import odeint
    f = lambda x, _t: [theta[0] * x[0] - theta[1] * x[0] * x[1], -theta[2] * x[1] + theta[3] * x[0] * x[1]]
    return odeint(f, x0, t)


def test_notebook_call_sequence_runs_end_to_end():
    ref = rs.load_reference()
    names, pnames = ["_x_1", "_x_2"], ["_theta_1", "_theta_2", "_theta_3", "_theta_4"]
    symbols = rs.Symbols(sym.symbols(names), sym.symbols(pnames))
    theta_true, x0, snr = [2.0, 1.0, 4.0, 1.0], [5.0, 3.0], 40.0
    t_est = np.arange(0, 4.0, 0.08)                                     # cell 21
    t_obs = np.array([0, 0.1, 0.5, 0.6, 0.8, 1.0, 1.4, 1.6, 2.0])       # cell 14
    # cell 25
    path = os.path.join("/root/reference", "Lotka_Volterra_ODEs.txt")
    odes, odes_sym = ref.import_odes.import_odes(symbols, path)
    state_couplings, _ = ref.import_odes.find_state_couplings_in_odes(odes_sym, symbols)
    grad_param, grad_states = ref.import_odes.gradient_of_odes(symbols, odes_sym)
    assert [list(c) for c in state_couplings] == [[0, 1], [0, 1]]
    # cell 27 (seeded stand-in for setup_simulation: simulate_state_dynamics.py:79-92,156-187)
    fine = np.arange(0, 4.0, 0.01)
    truth = pd.DataFrame(_simulate(theta_true, x0, fine), index=fine, columns=names).rename_axis("time")
    Xo = _simulate(theta_true, x0, np.concatenate([[0.0], t_obs[1:]]))
    rng = np.random.default_rng(0)
    Yo = Xo + np.sqrt((Xo.mean(axis=0) / snr) ** 2) * rng.standard_normal(Xo.shape)
    observations = pd.DataFrame(Yo, index=t_obs, columns=names).rename_axis("time")
    # cell 30: reference defaults RQ [0.1, 2, 5]
    D, A, cov, cov_obs, cov_so = ref.GP_regression.kernel_function(t_est, observations.index)
    assert D.shape == (50, 50) and cov_obs.shape == (9, 9) and cov_so.shape == (50, 9)
    # cells 32, 35
    lin = rs.LocallyLinearODEs()
    lin.ode_param.B, lin.ode_param.b = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(
        odes, symbols.state, symbols.param)
    lin.state.R, lin.state.r = ref.rewrite.rewrite_odes_as_linear_combination_in_states(
        odes, symbols.state, symbols.param, symbols.state, state_couplings, 1)
    # cell 39
    mean, inv_cov = ref.GP_regression.fitting_state_observations(observations, symbols.state, snr, t_est, cov, cov_obs,
                                                                 cov_so)
    assert mean.shape == (50, 2) and inv_cov.shape == (50, 50)
    om, _ = vo.fit_state_observations(Yo, [0, 1], 2, snr, cov, cov_obs, cov_so)
    assert np.max(np.abs(mean.values - om)) < 1e-9 * np.max(np.abs(om))
    # cell 41: 10 iterations of the reference's default ('minimizer') loop
    state = rs.writable_frame(mean.values, t_est, names)
    P = ref.proxies_patched
    for _ in range(10):
        param = P.proxy_for_ode_parameters(state, lin, D, A, symbols.param, theta_true, odes, grad_param, 'nonnegative',
                                           'minimizer')
        out = P.proxy_for_ind_states(state, param, odes, grad_states, lin, D, A, mean, inv_cov, observations, truth,
                                     state_couplings, False, optimizer='minimizer', constraints=None)
        state = rs.writable_frame(out.values, t_est, names)
    th = param.values.ravel()
    assert th.shape == (4,) and np.all(np.isfinite(th)) and np.all(th >= 0.0)
    assert np.all(np.isfinite(state.values)) and state.shape == (50, 2)
    assert list(param.index) == pnames and param.index.name == "ODE parameter symbols"
    # the closed-form update on the same proxies (the path this repository accelerates): live reference == oracle
    th_ref = P.proxy_for_ode_parameters(state, lin, D, A, symbols.param, theta_true, odes, grad_param, None,
                                        'analytical').values.ravel()
    lin_o = vo.Linearisation(vo.parse_ode_lines(open(path).read().splitlines()), symbols.state, symbols.param)
    th_o = vo.theta_step(state.values, lin_o, D)[0]
    assert np.max(np.abs(th_ref - th_o)) < 1e-9 * np.max(np.abs(th_o))
