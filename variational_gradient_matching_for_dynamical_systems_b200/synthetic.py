"""Synthetic inputs for tests and benchmarks (data generation is not part of the hot path;
the reference uses scipy odeint with unseeded noise, simulate_state_dynamics.py:79-187)."""
from __future__ import annotations

import numpy as np
import torch


def lorenz96_trajectory(K, T, dt_grid=0.05, alpha=8.0, seed=0, dt=0.01, spinup=5.0, device="cuda"):
    """RK4 Lorenz96 trajectory sampled at t_i = i*dt_grid.  Returns (t [T] numpy, X [K, T] torch, state-major).
    One CUDA launch per grid interval (``vgm_lorenz96_rk4``, csrc/simulate.cu); the initial state alpha + N(0, 1)
    is drawn on the host with a seeded generator, so a (K, seed) pair always gives the same trajectory."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    dev = torch.device(device)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x0 = (alpha + torch.randn(K, generator=gen, dtype=torch.float64)).to(dev)
    sub = max(1, int(round(dt_grid / dt)))
    h = dt_grid / sub
    if sub > 8:
        raise ValueError("dt_grid / dt = %d sub-steps per grid interval; the kernel holds at most 8" % sub)
    ld = (T + 7) // 8 * 8
    X = torch.zeros((K, ld), dtype=torch.float64, device=dev)
    scratch = torch.empty(2 * K, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.vgm_lorenz96_rk4(x0.data_ptr(), K, C.c_double(alpha), C.c_double(h), sub, int(round(spinup / h)), T,
                                        X.data_ptr(), ld, scratch.data_ptr(), _lib.stream_ptr()))
    return np.arange(T) * dt_grid, X[:, :T].contiguous()


def sample_observations(X_KT, observed_states, observed_time_idx, SNR, seed=0):
    """Noisy observations of a state-major trajectory (simulate_state_dynamics.py:156-187): the noise variance of
    observed state s is (mean over the observed time points of x_s / SNR)^2 (:176), the noise is N(0, 1) scaled by
    its square root (:179; the reference draws it unseeded, here a seeded host generator).  Returns
    (Y [n_obs_states, n_obs_times], obs_variance [n_obs_states]) on the device of ``X_KT``."""
    dev = X_KT.device
    rows = torch.as_tensor(np.asarray(observed_states, dtype=np.int64), device=dev)
    cols = torch.as_tensor(np.asarray(observed_time_idx, dtype=np.int64), device=dev)
    sub = X_KT.index_select(0, rows).index_select(1, cols)
    var = (sub.mean(dim=1) / SNR) ** 2
    gen = torch.Generator(device="cpu").manual_seed(seed)
    noise = torch.randn(sub.shape, generator=gen, dtype=torch.float64).to(dev)
    return sub + var.sqrt()[:, None] * noise, var


def lorenz96_problem(K, T, dt_grid=0.05, seed=0, hidden_noise=0.5, obs_snr=None, device="cuda"):
    """Config C3/C5 of BASELINE.json: states with even 0-based index observed (= _x_1,_x_3,...,
    VGM_for_Lorenz96.py:53) and clamped; hidden states start at truth + N(0, hidden_noise^2).
    Returns (t, X0 [K, T] torch, hidden indices)."""
    t, X = lorenz96_trajectory(K, T, dt_grid, seed=seed, device=device)
    gen = torch.Generator(device="cpu").manual_seed(seed + 1)
    hidden = np.arange(1, K, 2)
    noise = torch.randn((len(hidden), T), generator=gen, dtype=torch.float64).to(device)
    X0 = X.clone()
    X0[1::2] += hidden_noise * noise
    if obs_snr:
        gen2 = torch.Generator(device="cpu").manual_seed(seed + 2)
        sd = (X[0::2].mean(dim=1, keepdim=True) / obs_snr).abs()
        X0[0::2] += sd * torch.randn(X[0::2].shape, generator=gen2, dtype=torch.float64).to(device)
    return t, X0, hidden
