"""Synthetic inputs for tests and benchmarks (data generation is not part of the hot path;
the reference uses scipy odeint with unseeded noise, simulate_state_dynamics.py:79-187)."""
from __future__ import annotations

import numpy as np
import torch


def lorenz96_rhs(x, alpha=8.0):
    return (torch.roll(x, -1, -1) - torch.roll(x, 2, -1)) * torch.roll(x, 1, -1) - x + alpha


def lorenz96_trajectory(K, T, dt_grid=0.05, alpha=8.0, seed=0, dt=0.01, spinup=5.0, device="cuda"):
    """RK4 Lorenz96 trajectory sampled at t_i = i*dt_grid.  Returns (t [T] numpy, X [K, T] torch, state-major)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = (alpha + torch.randn(K, generator=gen, dtype=torch.float64)).to(device)
    sub = max(1, int(round(dt_grid / dt)))
    h = dt_grid / sub

    def step(x):
        k1 = lorenz96_rhs(x, alpha)
        k2 = lorenz96_rhs(x + 0.5 * h * k1, alpha)
        k3 = lorenz96_rhs(x + 0.5 * h * k2, alpha)
        k4 = lorenz96_rhs(x + h * k3, alpha)
        return x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    for _ in range(int(round(spinup / h))):
        x = step(x)
    out = torch.empty((T, K), dtype=torch.float64, device=device)
    for i in range(T):
        out[i] = x
        for _ in range(sub):
            x = step(x)
    return np.arange(T) * dt_grid, out.t().contiguous()


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
