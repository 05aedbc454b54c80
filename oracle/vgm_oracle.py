"""TEST INFRASTRUCTURE ONLY -- NumPy restatement (CPU oracle) of the reference hot path.

Nothing in the product package may import this file.  It restates, in plain NumPy/SymPy,
the arithmetic of the mean-field variational gradient-matching loop of
``ngorbach/Variational_Gradient_Matching_for_Dynamical_Systems`` (``VGM_modules/``):

=====================  ==========================================================
function here          reference site it follows
=====================  ==========================================================
kernel_matrices        GP_regression.py:160-260  (kernel + derivative grids)
gm_operators           GP_regression.py:306,317  (D = dC C^-1 by LU solve, A)
kernel_function        GP_regression.py:160-319  (whole function, no RNG/plot)
Linearisation          rewrite_odes_as_local_linear_combinations.py:28-58,76-124
theta_step             proxies_for_ode_parameters_and_states.py:55-93
state_sweep            proxies_for_ode_parameters_and_states.py:185-262
state_cov              proxies_for_ode_parameters_and_states.py:190,239
lorenz96_*             same sites, specialised to the Lorenz96 structure
                       (Lorenz96_ODEs.txt:1-10) for K beyond the reach of SymPy
fit_state_observations GP_regression.py:33-51
=====================  ==========================================================

Parity status: PINNED.  ``oracle/make_golden.py`` executes the *live, unmodified*
reference (through ``oracle/reference_shim.py``) in the build container and stores its
outputs in ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` checks every function
here against those vectors (and, when ``/root/reference`` is present, against the live
reference directly).  The reference itself ships no tests or golden vectors (SURVEY.md
section 4), so live execution is the only possible pin.
"""
from __future__ import annotations

import numpy as np
import sympy as sym

# --------------------------------------------------------------------------------------
# G1: kernel grids  (GP_regression.py:160-260)
# --------------------------------------------------------------------------------------

def _kernel_closed_form(kernel_type, kernel_param):
    """Return f(d) -> (k, dk/dt0, d2k/dt0dt1) for d = t0 - t1 (stationary kernels)."""
    sig2 = float(kernel_param[-1]) ** 2          # GP_regression.py:232
    if kernel_type == "rbf":                      # GP_regression.py:176-179
        ell = float(kernel_param[0])

        def f(d):
            k = sig2 * np.exp(-0.5 * (d / ell) ** 2)
            return k, -(d / ell ** 2) * k, (1.0 / ell ** 2 - d ** 2 / ell ** 4) * k
        return f
    if kernel_type == "RQ":                       # GP_regression.py:205-209
        alpha, ell = float(kernel_param[0]), float(kernel_param[1])

        def f(d):
            base = 1.0 + d ** 2 / (2.0 * alpha * ell ** 2)
            k = sig2 * base ** (-alpha)
            dk = -sig2 * d / ell ** 2 * base ** (-alpha - 1.0)
            ddk = sig2 / ell ** 2 * base ** (-alpha - 1.0) \
                - sig2 * d ** 2 / ell ** 4 * (alpha + 1.0) / alpha * base ** (-alpha - 2.0)
            return k, dk, ddk
        return f
    return None


def _kernel_sympy(kernel_type, kernel_param):
    """Generic route for the remaining kernel types: the same symbolic expressions as
    GP_regression.py:181-231, differentiated (:236-237) and evaluated on whole grids."""
    t0, t1 = sym.symbols("t0 t1")
    p = kernel_param
    if kernel_type == "periodic":
        dist = (t0 - t1) / p[1]
        k = sym.exp(-2 * sym.sin(np.pi * sym.sqrt(dist ** 2) / p[1]) ** 2 / p[0] ** 2)
    elif kernel_type == "locally_periodic":
        dist = t0 - t1
        k = sym.exp(-dist ** 2 / 2 * p[1] ** 2) * sym.exp(
            -2 * sym.sin(np.pi * sym.sqrt((t0 - t1) ** 2) / p[2]) ** 2 / p[1] ** 2)
    elif kernel_type == "rbf+lin":
        k = sym.exp(-(t0 - t1) ** 2 / p[1] ** 2) + p[2] + p[3] * (t0 - p[4]) * (t1 - p[4])
    elif kernel_type == "sigmoid":
        k = sym.asin((p[1] + p[2] * t0 * t1) / sym.sqrt((p[1] + p[2] * t0 ** 2 + 1) * (p[1] + p[2] * t1 ** 2 + 1)))
    elif kernel_type == "exp_sin_squared":
        dist = sym.sqrt((t0 - t1) ** 2)
        k = sym.exp(-2 * (sym.sin(np.pi * dist / p[1]) / p[0]) ** 2)
    elif kernel_type == "sigmoid2":
        k = sym.tanh(p[0] * t0 * t1 + p[1])
    elif kernel_type == "matern" and p[1] in (0.5, 1.5, 2.5):
        dist = ((t0 - t1) / p[0]) ** 2
        if p[1] == 0.5:
            k = sym.exp(-dist)
        elif p[1] == 1.5:
            kk = dist * np.sqrt(3)
            k = (1.0 + kk) * sym.exp(-kk)
        else:
            kk = dist * np.sqrt(5)
            k = (1.0 + kk + kk ** 2 / 3.0) * sym.exp(-kk)
    else:
        raise NotImplementedError(kernel_type)
    k = k * p[-1] ** 2
    dk = k.diff(t0)
    ddk = dk.diff(t1)
    return tuple(sym.lambdify((t0, t1), e, "numpy") for e in (k, dk, ddk))


def kernel_matrices(time_points, time_points2, kernel_type="RQ", kernel_param=(0.1, 2, 5)):
    """C, dC, ddC on ``time_points``; C_obs on ``time_points2``; C_state_obs (cross)."""
    t = np.asarray(time_points, dtype=float)
    t2 = np.asarray(time_points2, dtype=float)
    f = _kernel_closed_form(kernel_type, list(kernel_param))
    if f is not None:
        C, dC, ddC = f(t[:, None] - t[None, :])
        Cobs = f(t2[:, None] - t2[None, :])[0]
        Cso = f(t[:, None] - t2[None, :])[0]
    else:
        fk, fdk, fddk = _kernel_sympy(kernel_type, list(kernel_param))
        g = lambda fn, a, b: np.broadcast_to(fn(a[:, None], b[None, :]), (len(a), len(b))).astype(float)
        C, dC, ddC = g(fk, t, t), g(fdk, t, t), g(fddk, t, t)
        Cobs, Cso = g(fk, t2, t2), g(fk, t, t2)
    return C, dC, ddC, Cobs, Cso


def gm_operators(C, dC, ddC):
    """D = dC C^-1 via LU solve on the transposes (GP_regression.py:306);
    A = ddC - D dC^T (GP_regression.py:317)."""
    D = np.linalg.solve(C.T, dC.T).T
    A = ddC - D.dot(dC.T)
    return D, A


def kernel_function(time_points, time_points2, kernel_type="RQ", kernel_param=(0.1, 2, 5)):
    """Same return tuple as GP_regression.kernel_function (:319).  The two
    ``np.random.multivariate_normal`` draws (:310-311) only feed a plot and are omitted."""
    if type(time_points) is not np.ndarray:
        raise ValueError("time points is not a numpy array")
    if len(time_points.shape) > 1:
        raise ValueError("time points must be a one dimensional numpy array")
    C, dC, ddC, Cobs, Cso = kernel_matrices(time_points, time_points2, kernel_type, kernel_param)
    D, A = gm_operators(C, dC, ddC)
    return D, A, C, Cobs, Cso


def fit_state_observations(obs_values, obs_cols, n_states, SNR, cov, cov_obs, cov_state_obs):
    """GP_regression.py:33-51.  ``obs_values`` is T2 x n_obs, ``obs_cols`` the state index
    of each observed column.  Reproduces the accumulating-noise quirk (:42: ``cov_obs`` is
    reassigned inside the loop, so observed state i sees sum_{j<=i} sigma_j^2) and the
    final-matrix predictive covariance (:45)."""
    obs_values = np.asarray(obs_values, dtype=float)
    obs_var = (np.mean(obs_values, axis=0) / SNR) ** 2
    mean = np.zeros((cov.shape[0], n_states))
    Co = np.array(cov_obs, dtype=float)
    for i, col in enumerate(obs_cols):
        Co = Co + np.diag([obs_var[i]] * Co.shape[1])
        mean[:, col] = cov_state_obs.dot(np.linalg.solve(Co, obs_values[:, i]))
    pred_cov = cov - cov_state_obs.dot(np.linalg.solve(Co, cov_state_obs.T))
    return mean, np.linalg.pinv(pred_cov)


def predictive_cov(obs_values, SNR, cov, cov_obs, cov_state_obs):
    """``state_pred_cov`` of GP_regression.py:45 (the matrix whose pinv the function returns)."""
    obs_values = np.asarray(obs_values, dtype=float)
    obs_var = (np.mean(obs_values, axis=0) / SNR) ** 2
    Co = np.array(cov_obs, dtype=float) + np.sum(obs_var) * np.eye(np.asarray(cov_obs).shape[0])
    return cov - cov_state_obs.dot(np.linalg.solve(Co, cov_state_obs.T))


# --------------------------------------------------------------------------------------
# G3 / G4: locally linear rewrites  (rewrite_odes_as_local_linear_combinations.py)
# --------------------------------------------------------------------------------------

ONE = sym.Symbol("one_vector")


def parse_ode_lines(lines):
    """import_odes.py:22-25 -- sympify then factor each right-hand side."""
    return [sym.sympify(s).factor() for s in lines if s.strip()]


def state_couplings(odes_sym, state_symbols):
    """import_odes.py:73-87: c(u) = ODE indices whose RHS contains x_u (ascending)."""
    free = [e.free_symbols for e in odes_sym]
    return [[k for k in range(len(state_symbols)) if s in free[k]] for s in state_symbols]


class Linearisation:
    """Symbolic B,b (per ODE, linear in theta) and R,r (per state x ODE, linear in x_u)
    with the reference's constant-handling quirks:

    * parameter rewrite: an entry with no free symbols is *replaced by* ``one_vector``
      (rewrite...py:49-52) -- zero becomes one;
    * state rewrite: an entry with no free symbols is *multiplied by* ``one_vector``
      (rewrite...py:112-115) -- zero stays zero.

    ``zero_fill=True`` switches the parameter rewrite to the mathematically intended
    behaviour of the MATLAB original (constants are multiplied by the ones vector).
    """

    def __init__(self, odes_sym, state_symbols, param_symbols, couplings=None, zero_fill=False):
        self.state_symbols = list(state_symbols)
        self.param_symbols = list(param_symbols)
        self.K = len(self.state_symbols)
        self.p = len(self.param_symbols)
        self.odes = [sym.sympify(e).expand() for e in odes_sym]      # rewrite...py:45,108 (.expand())
        self.couplings = couplings if couplings is not None else state_couplings(odes_sym, state_symbols)
        self.B, self.b = [], []
        for k in range(self.K):
            eB, eb = sym.linear_eq_to_matrix([self.odes[k]], self.param_symbols)
            eb = -eb
            eB, eb = list(eB), list(eb)
            for i in range(len(eB)):
                if len(eB[i].free_symbols) == 0:
                    eB[i] = eB[i] * ONE if zero_fill else ONE
            for i in range(len(eb)):
                if len(eb[i].free_symbols) == 0:
                    eb[i] = eb[i] * ONE if zero_fill else ONE
            self.B.append(eB)
            self.b.append(eb[0])
        self.R = [dict() for _ in range(self.K)]
        self.r = [dict() for _ in range(self.K)]
        for u in range(self.K):
            for k in self.couplings[u]:
                eR, er = sym.linear_eq_to_matrix([self.odes[k]], [self.state_symbols[u]])
                eR, er = eR[0], -er[0]
                if len(eR.free_symbols) == 0:
                    eR = eR * ONE
                if len(er.free_symbols) == 0:
                    er = er * ONE
                self.R[u][k] = eR
                self.r[u][k] = er
        self._fB = self._fb = self._fR = self._fr = None

    # evaluation helpers ----------------------------------------------------------------
    def _compile(self):
        xs = self.state_symbols + [ONE]
        self._fB = [sym.lambdify(xs, e, "numpy") for e in self.B]
        self._fb = [sym.lambdify(xs, e, "numpy") for e in self.b]
        allsym = self.param_symbols + self.state_symbols + [ONE]
        self._fR = [{k: sym.lambdify(allsym, e, "numpy") for k, e in d.items()} for d in self.R]
        self._fr = [{k: sym.lambdify(allsym, e, "numpy") for k, e in d.items()} for d in self.r]

    def eval_B_b(self, k, S):
        """S: T x K snapshot.  Returns B (T x p), b (T)."""
        if self._fB is None:
            self._compile()
        T = S.shape[0]
        args = [S[:, j] for j in range(self.K)] + [np.ones(T)]
        B = np.stack([np.broadcast_to(np.asarray(v, dtype=float), (T,)) for v in self._fB[k](*args)], axis=1)
        b = np.broadcast_to(np.asarray(self._fb[k](*args), dtype=float), (T,))
        return B, b

    def eval_R_r(self, u, k, theta, S):
        if self._fR is None:
            self._compile()
        T = S.shape[0]
        args = list(theta) + [S[:, j] for j in range(self.K)] + [np.ones(T)]
        R = np.broadcast_to(np.asarray(self._fR[u][k](*args), dtype=float), (T,))
        r = np.broadcast_to(np.asarray(self._fr[u][k](*args), dtype=float), (T,))
        return R, r


# --------------------------------------------------------------------------------------
# P2 / P3: theta step  (proxies...py:55-93)
# --------------------------------------------------------------------------------------

def theta_step(X, lin: Linearisation, D, A=None):
    """Returns (theta, local_mean, local_scaling, global_cov).  ``constraints=None`` path:
    theta = solve(local_scaling, local_mean) (proxies...py:93).  ``global_cov`` is computed
    by the reference (:79) but never returned; we return it for the covariance parity."""
    X = np.asarray(X, dtype=float)
    p = lin.p
    m = np.zeros(p)
    Sc = np.zeros((p, p))
    Cth = np.zeros((p, p))
    for k in range(lin.K):
        B, b = lin.eval_B_b(k, X)
        m += B.T.dot(D.dot(X[:, k]) - b)
        Sc += B.T.dot(B)
        if A is not None:
            Cth += B.T.dot(A.dot(B))
    theta = np.linalg.solve(Sc, m)
    return theta, m, Sc, Cth


def squared_loss(theta, X, D, odes_sym, state_symbols, param_symbols, alpha=0.0):
    """Loss and gradient of the 'minimizer' parameter update (proxies...py:379-392):
    ``sum (f(X, theta) - D X)^2 + alpha |theta|^2`` and ``2 sum (f - D X) df/dtheta + 2 alpha theta``
    (the reference's gradient_of_odes replaces exact-zero derivatives by machine epsilon, import_odes.py:45; that
    changes the gradient by ~1e-16 relative and is not reproduced).  Returns (loss, grad)."""
    X = np.asarray(X, dtype=float)
    theta = np.asarray(theta, dtype=float)
    DX = np.asarray(D).dot(X)
    args = list(state_symbols) + list(param_symbols)
    vals = [X[:, i] for i in range(X.shape[1])] + [float(v) for v in theta]
    loss, grad = 0.0, np.zeros(len(theta))
    for k, e in enumerate(odes_sym):
        f = np.broadcast_to(sym.lambdify(args, e, "numpy")(*vals), (X.shape[0],))
        res = f - DX[:, k]
        loss += float(np.sum(res ** 2))
        for i, ps in enumerate(param_symbols):
            df = np.broadcast_to(sym.lambdify(args, sym.diff(e, ps), "numpy")(*vals), (X.shape[0],))
            grad[i] += 2.0 * float(np.sum(res * df))
    return loss + alpha * float(np.sum(theta ** 2)), grad + 2.0 * alpha * theta


# --------------------------------------------------------------------------------------
# P4: state sweep  (proxies...py:185-262)
# --------------------------------------------------------------------------------------

def state_sweep(X, theta_star, lin: Linearisation, D, hidden, A=None, cov_states=(), obs_inv_cov=None, obs_mean=None):
    """One sweep over ``hidden`` (ascending index order, proxies...py:195).

    ``obs_inv_cov`` / ``obs_mean`` (opt-in): the observation term the reference keeps commented out at
    proxies...py:265-266, ``solve(local_scaling + state_pred_inv_cov, local_mean + state_pred_inv_cov .
    state_pred_mean[:, u])`` (active in the MATLAB original, proxy_for_ind_states.m:92-93).

    * coefficients R, r are evaluated on a snapshot of X taken once (:191);
    * ``D @ X[:, k]`` reads the live array (alias at :185, read at :225);
    * if u is not in c(u) the reference's aliasing (:201,:245) doubles the diagonal.
    Returns (X_new, {u: Cov_u for u in cov_states}).
    """
    X = np.array(X, dtype=float, copy=True)
    S = X.copy()
    T = X.shape[0]
    covs = {}
    for u in hidden:
        lam = np.zeros(T)
        rhs = np.zeros(T)
        Mu = None
        cov = np.zeros((T, T)) if (A is not None and u in cov_states) else None
        for k in lin.couplings[u]:
            Rv, rv = lin.eval_R_r(u, k, theta_star, S)
            if k != u:
                bb = rv - D.dot(X[:, k])
                rhs += -Rv * bb
                lam += Rv * Rv
                Buk = np.diag(Rv)
            else:
                Buk = np.diag(Rv) - D
                rhs += -Buk.T.dot(rv)
                Mu = Buk.T.dot(Buk)
            if cov is not None:
                cov += Buk.T.dot(A.dot(Buk))
        if Mu is None:
            Aeff = 2.0 * np.diag(lam)
        else:
            Aeff = Mu + np.diag(lam)
        if obs_inv_cov is not None:
            Aeff = Aeff + obs_inv_cov
            rhs = rhs + np.asarray(obs_inv_cov).dot(np.asarray(obs_mean)[:, u])
        X[:, u] = np.linalg.solve(Aeff, rhs)
        if cov is not None:
            covs[u] = cov
    return X, covs


# --------------------------------------------------------------------------------------
# Lorenz96 specialisation (vectorised; any K >= 4).  Lorenz96_ODEs.txt:1-10:
#   f_k = (x_{k+1} - x_{k-2}) x_{k-1} - x_k + alpha         (indices mod K)
# --------------------------------------------------------------------------------------

def lorenz96_theta_step(X, D, A=None):
    """theta step for Lorenz96: B_k == 1, b_k = f_k - alpha."""
    X = np.asarray(X, dtype=float)
    T, K = X.shape
    DX = D.dot(X)
    xp1, xm1, xm2 = np.roll(X, -1, 1), np.roll(X, 1, 1), np.roll(X, 2, 1)
    b = xp1 * xm1 - xm2 * xm1 - X          # expanded form, as sympy's .expand() orders terms
    m = float(np.sum(DX - b))
    Sc = float(K * T)
    Cth = float(K * np.sum(A)) if A is not None else 0.0
    return m / Sc, m, Sc, Cth


def lorenz96_state_sweep(X, alpha_star, D, hidden, return_info=False):
    """Lorenz96 state sweep in the reference's sequential order with live D.x_k reads.

    For state u the four coupled ODEs are k = u-1, u, u+1, u+2 (c(u) sorted ascending in the
    reference; the order only changes the summation order of three terms)."""
    X = np.array(X, dtype=float, copy=True)
    S = X.copy()
    T, K = X.shape
    W = np.eye(T) + D
    M = W.T.dot(W)
    a = float(alpha_star)
    hidden = list(hidden)
    # D @ X for every column, refreshed column-wise as states are updated (live semantics)
    DX = D.dot(X)
    for u in hidden:
        s = lambda j: S[:, (u + j) % K]
        terms = {}
        # k = u+1 : f = (x_{u+2} - x_{u-1}) x_u - x_{u+1} + a
        terms[(u + 1) % K] = (s(2) - s(-1), a - s(1))
        # k = u-1 : f = (x_u - x_{u-3}) x_{u-2} - x_{u-1} + a
        terms[(u - 1) % K] = (s(-2), a - s(-1) - s(-3) * s(-2))
        # k = u+2 : f = (x_{u+3} - x_u) x_{u+1} - x_{u+2} + a
        terms[(u + 2) % K] = (-s(1), a + s(3) * s(1) - s(2))
        r_uu = a + s(1) * s(-1) - s(-2) * s(-1)
        lam = np.zeros(T)
        rhs = np.zeros(T)
        for k in sorted(set(list(terms.keys()) + [u])):
            if k == u:
                rhs += W.T.dot(r_uu)             # -B^T r with B = -I - D
            else:
                Rv, rv = terms[k]
                rhs += -Rv * (rv - DX[:, k])
                lam += Rv * Rv
        X[:, u] = np.linalg.solve(M + np.diag(lam), rhs)
        DX[:, u] = D.dot(X[:, u])
    return X


def lorenz96_column_needs(cols, K, hidden):
    """States whose rows ``lorenz96_sweep_columns`` reads for the hidden columns ``cols``:
    (snapshot states, live states).  Snapshot: u-3 .. u+3 (the coefficient stencil);
    live: the hidden k in {u-1, u+1, u+2} that precede u in the ascending update order (proxies...py:195),
    whose D.x_k the reference reads from the already-updated column (:185 alias, :225)."""
    hid = set(int(h) for h in hidden)
    snap, live = set(), set()
    for u in cols:
        u = int(u)
        for j in range(-3, 4):
            snap.add((u + j) % K)
        for j in (-1, 1, 2):
            k = (u + j) % K
            if k in hid and k < u:
                live.add(k)
    return sorted(snap), sorted(live)


def lorenz96_sweep_columns(snap, live, K, alpha_star, D, cols, hidden):
    """``lorenz96_state_sweep`` restricted to the hidden columns ``cols`` (the C5 check of SURVEY.md 8d:
    "restatement on 64 random hidden columns"); same statements as proxies...py:195-262.

    snap : {state index: T-vector} sweep-start snapshot rows (at least ``lorenz96_column_needs(...)[0]``)
    live : {state index: T-vector} UPDATED rows of the hidden states that precede a requested column in the
           update order and are read live by it (``lorenz96_column_needs(...)[1]``)
    Returns {u: x_u}.  Each column costs one dense T x T solve, like the reference."""
    D = np.asarray(D, dtype=float)
    T = D.shape[0]
    W = np.eye(T) + D
    M = W.T.dot(W)
    a = float(alpha_star)
    hid = set(int(h) for h in hidden)
    out = {}
    for u in cols:
        u = int(u)
        s = lambda j: np.asarray(snap[(u + j) % K], dtype=float)

        def dx(j):
            k = (u + j) % K
            xk = live[k] if (k in hid and k < u) else snap[k]
            return D.dot(np.asarray(xk, dtype=float))
        terms = {(u + 1) % K: (s(2) - s(-1), a - s(1), dx(1)),
                 (u - 1) % K: (s(-2), a - s(-1) - s(-3) * s(-2), dx(-1)),
                 (u + 2) % K: (-s(1), a + s(3) * s(1) - s(2), dx(2))}
        r_uu = a + s(1) * s(-1) - s(-2) * s(-1)
        lam = np.zeros(T)
        rhs = np.zeros(T)
        for k in sorted(set(list(terms.keys()) + [u])):       # ascending ODE index, as c(u) is sorted
            if k == u:
                rhs += W.T.dot(r_uu)
            else:
                Rv, rv, dxk = terms[k]
                rhs += -Rv * (rv - dxk)
                lam += Rv * Rv
        out[u] = np.linalg.solve(M + np.diag(lam), rhs)
    return out


def lorenz96_state_cov(S, u, D, A):
    """Cov_u = (I+D)^T A (I+D) + A o sum_k R_k R_k^T  (closed form of proxies...py:239)."""
    T, K = S.shape
    W = np.eye(T) + D
    s = lambda j: S[:, (u + j) % K]
    Rs = [s(2) - s(-1), s(-2), -s(1)]
    G = W.T.dot(A.dot(W))
    return G + A * sum(np.outer(R, R) for R in Rs)


# --------------------------------------------------------------------------------------
# synthetic-input helpers shared by tests and make_golden (deterministic, seeded)
# --------------------------------------------------------------------------------------

def lorenz96_rhs(x, alpha=8.0):
    return (np.roll(x, -1, -1) - np.roll(x, 2, -1)) * np.roll(x, 1, -1) - x + alpha


def lorenz96_trajectory(K, T, dt_grid, alpha=8.0, seed=0, dt=0.01, spinup=5.0):
    """RK4 trajectory sampled on t_i = i*dt_grid, i < T.  Returns (t, X[T,K])."""
    rng = np.random.default_rng(seed)
    x = alpha + rng.standard_normal(K)
    sub = max(1, int(round(dt_grid / dt)))
    h = dt_grid / sub

    def step(x):
        k1 = lorenz96_rhs(x, alpha)
        k2 = lorenz96_rhs(x + 0.5 * h * k1, alpha)
        k3 = lorenz96_rhs(x + 0.5 * h * k2, alpha)
        k4 = lorenz96_rhs(x + h * k3, alpha)
        return x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)

    for _ in range(int(round(spinup / h))):
        x = step(x)
    out = np.empty((T, K))
    for i in range(T):
        out[i] = x
        for _ in range(sub):
            x = step(x)
    return np.arange(T) * dt_grid, out


def lorenz96_ode_lines(K):
    """Text of a Lorenz96 ``*_ODEs.txt`` file with K states (format of Lorenz96_ODEs.txt:1-10)."""
    n = lambda j: "_x_%d" % ((j % K) + 1)
    return ["(%s - %s) * %s - %s + _alpha" % (n(k + 1), n(k - 2), n(k - 1), n(k)) for k in range(K)]
