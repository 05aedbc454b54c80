"""Device-resident VGM engine: keeps the state proxies, the gradient-matching operators and
all index tables in HBM across coordinate-ascent iterations and drives the C ABI
(``include/vgm_b200.h``).  PyTorch is used only for device memory and streams.

Layout: the state proxy means are stored state-major, ``X[k, t]`` (``[K, ld]``, ``ld`` =
T rounded up to a multiple of 8), so one state's trajectory is contiguous -- the transpose of
the reference's ``T x K`` DataFrame (proxies_for_ode_parameters_and_states.py:53,168).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .codegen import OdeModel, Program, build_host_plan


def _round_up(x, m):
    return (x + m - 1) // m * m


def _i32(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=dev)


def _host_i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    arr = (C.c_int * max(1, a.size))(*a.tolist())
    return arr


class Operators:
    """D = dC C^-1 (T x T) and friends on the device, padded to leading dimension ld."""

    def __init__(self, D, A=None, ld=None, device=None, obs_inv_cov=None):
        """``obs_inv_cov`` (T x T, symmetric positive semi-definite; opt-in) adds the observation term the reference
        keeps commented out at proxies...py:265-266: every system matrix becomes ``local_scaling + obs_inv_cov``."""
        self.dev = device or _lib.require_cuda()
        D = np.asarray(D, dtype=np.float64) if not torch.is_tensor(D) else D
        T = D.shape[0]
        if D.shape[0] != D.shape[1]:
            raise ValueError("dC_times_invC is not a square matrix")
        self.T = T
        # rows start on 256-byte boundaries in every precision the solver uses (FP64, FP32, bf16): whole DRAM
        # sectors / TMA lines per row; at T=1000 that is 2.4 % of padding for ~4 % faster streaming kernels
        self.ld = ld or (_round_up(T, 128) if T >= 512 else _round_up(T, 8))
        self.D = self._pad(D)
        self.DT = self._pad(self.D[:, :T].t().contiguous())
        self.A = self._pad(A) if A is not None else None
        self.M = None
        self.M_c = None
        self.G = None
        self.G_bf16 = None
        self.M_bf16x2 = None
        self.G_shift = None
        self.band_M = self.band_M_lp = self.band_G = 0
        self.band_D = self.band_of(self.D)
        self.DtD = None
        self.Q = self._pad(obs_inv_cov) if obs_inv_cov is not None else None

    # |B[i][j]| <= 2^-55 max|B| beyond the band.  The operators decay geometrically away from the diagonal (a factor
    # 2^-0.45 per entry for the RBF kernel at the bench spacing), so what a row drops sums to < 2 units of round-off of
    # max|B| max|x| -- less than the rounding error the kept ~300-term FP64 accumulation commits itself.  Measured at
    # K = 100 000: band(M) 146 -> 126 against 2^-64, the residual GEMMs 10 % shorter, parity against the dense oracle
    # unchanged (5.4e-13).  VGM_BAND_LOG2=-64 restores the threshold at which the dropped part is invisible in FP64.
    EXACT_BAND = 2.0 ** float(os.environ.get("VGM_BAND_LOG2", "-55"))

    def band_of(self, B, rel=None):
        """Half band width of a T x T operator on the device (vgm_band_halfwidth); 0 = treat as dense."""
        out = C.c_int(0)
        _lib.check(_lib.load().vgm_band_halfwidth(B.data_ptr(), self.T, self.T, self.ld,
                                                   self.EXACT_BAND if rel is None else rel, C.byref(out),
                                                   _lib.stream_ptr()))
        band = int(out.value)
        return band if 0 < band < self.T - 64 else 0

    def _pad(self, M):
        out = torch.zeros((self.T, self.ld), dtype=torch.float64, device=self.dev)
        out[:, :self.T] = torch.as_tensor(M, dtype=torch.float64, device=self.dev)[:, :self.T]
        return out

    DENSE_MAX_T = 160       # csrc/dense_solve.cuh

    def ensure_DtD(self):
        """D^T D for the direct per-CTA solver of systems with time-varying R_uu (T <= DENSE_MAX_T)."""
        if self.DtD is None:
            lib = _lib.load()
            self.DtD = torch.zeros_like(self.D)
            g = _lib.GemmArgs()
            g.A, g.B, g.C = self.DT.data_ptr(), self.DT.data_ptr(), self.DtD.data_ptr()
            g.lda = g.ldb = g.ldc = self.ld
            g.N, g.M, g.K = self.T, self.T, self.T
            g.alpha = 1.0
            _lib.check(lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr()))
            if self.Q is not None:
                self.DtD += self.Q
        return self.DtD

    def ensure_shared_M(self, c):
        """M = (c I - D)^T (c I - D): the shared system matrix when R_uu == c for every state
        (proxies...py:229,232 with R = c I)."""
        if self.M is not None and self.M_c == c:
            return self.M
        if self.M is not None:
            # engines keep raw device pointers to M, G and their bf16 copies: replacing them here would leave an
            # earlier engine reading freed memory
            raise ValueError("this Operators object already holds M = (cI - D)^T (cI - D) for c = %r; a model whose "
                             "constant R_uu is %r needs its own Operators(D, A)" % (self.M_c, c))
        lib = _lib.load()
        T, ld = self.T, self.ld
        BT = -self.DT.clone()                      # (c I - D)^T = c I - D^T, rows K-contiguous
        BT[:, :T] += c * torch.eye(T, dtype=torch.float64, device=self.dev)
        M = torch.zeros_like(BT)
        g = _lib.GemmArgs()
        g.A, g.B, g.C = BT.data_ptr(), BT.data_ptr(), M.data_ptr()
        g.lda = g.ldb = g.ldc = ld
        g.N = g.M = g.K = T
        g.alpha = 1.0
        _lib.check(lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr()))
        if self.Q is not None:
            M += self.Q
        self.M, self.M_c = M, c
        self.band_M = self.band_of(M)
        self.band_M_lp = self.band_of(M, rel=2.0 ** -20)
        self.G = None
        return M

    def ensure_preconditioner(self, shift):
        """G = (M + shift I)^-1 by Cholesky on the device (vgm_spd_inverse)."""
        if self.G is not None and self.G_shift is not None and 0.5 * self.G_shift <= shift <= 2.0 * self.G_shift:
            return self.G
        lib = _lib.load()
        T, ld = self.T, self.ld
        S = self.M.clone()
        S[:, :T] += shift * torch.eye(T, dtype=torch.float64, device=self.dev)
        G = torch.zeros_like(S)
        nbytes = lib.vgm_gm_operators_ws_bytes(T, ld)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        _lib.check(lib.vgm_spd_inverse(S.data_ptr(), T, ld, G.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()))
        self.G, self.G_shift = G, float(shift)
        # low-precision operands of the mixed-precision solver: bf16 G (a preconditioner only has to be a
        # fixed SPD approximation, so entries below the bf16 resolution are dropped as well) and the
        # bf16 (hi, lo) split of M
        self.G_bf16 = torch.zeros((T, ld), dtype=torch.bfloat16, device=self.dev)
        _lib.check(lib.vgm_convert_bf16(G.data_ptr(), self.G_bf16.data_ptr(), T, ld, _lib.stream_ptr()))
        self.band_G = self.band_of(G, rel=2.0 ** -6)
        self.M_bf16x2 = torch.zeros((2, T, ld), dtype=torch.bfloat16, device=self.dev)
        _lib.check(lib.vgm_split_bf16(self.M.data_ptr(), self.M_bf16x2.data_ptr(), T, ld, _lib.stream_ptr()))
        return G


class DevicePlan:
    """``vgm_sweep_plan`` with its device/host arrays kept alive."""

    def __init__(self, hp, T, ld, tables, device=None):
        dev = device or _lib.require_cuda()
        self.host = hp
        keep = self._keep = {}
        for name in ("slot_state", "pair_ptr", "pair_ode", "pair_code", "pair_live", "pair_self", "slot_has_self",
                     "level_slots", "live_state", "live_pair_slot", "live_pair_src", "early_slots", "bulk_slots"):
            arr = getattr(hp, name, np.zeros(0, dtype=np.int32))
            keep[name] = _i32(arr if len(arr) else np.zeros(1, dtype=np.int32), dev)
        for name in ("level_ptr", "live_level_ptr", "live_pair_level_ptr"):
            keep[name + "_host"] = _host_i32(getattr(hp, name))
        p = _lib.SweepPlan()
        p.T, p.ld, p.n_hidden = T, ld, hp.n_hidden
        p.slot_state = keep["slot_state"].data_ptr()
        p.pair_ptr = keep["pair_ptr"].data_ptr()
        p.pair_ode = keep["pair_ode"].data_ptr()
        p.pair_code = keep["pair_code"].data_ptr()
        p.pair_live = keep["pair_live"].data_ptr()
        p.pair_self = keep["pair_self"].data_ptr()
        p.ode_idx = tables["ode_idx"].data_ptr()
        p.ode_state = tables["ode_state"].data_ptr()
        p.slot_has_self = keep["slot_has_self"].data_ptr()
        p.n_levels = hp.n_levels
        p.level_ptr_host = C.cast(keep["level_ptr_host"], C.POINTER(C.c_int))
        p.level_slots = keep["level_slots"].data_ptr()
        p.n_live = hp.n_live
        p.live_state = keep["live_state"].data_ptr()
        p.live_level_ptr_host = C.cast(keep["live_level_ptr_host"], C.POINTER(C.c_int))
        p.n_live_pairs = hp.n_live_pairs
        p.live_pair_slot = keep["live_pair_slot"].data_ptr()
        p.live_pair_src = keep["live_pair_src"].data_ptr()
        p.live_pair_level_ptr_host = C.cast(keep["live_pair_level_ptr_host"], C.POINTER(C.c_int))
        p.ruu_const, p.ruu_value = hp.ruu_const, hp.ruu_value
        p.n_early = int(getattr(hp, "n_early", 0))
        p.early_slots = keep["early_slots"].data_ptr() if p.n_early else None
        p.bulk_slots = keep["bulk_slots"].data_ptr() if p.n_early else None
        self.c = p
        self.n_hidden = hp.n_hidden
        self.hidden = np.asarray(hp.slot_state, dtype=np.int64)


class LocalProblem:
    """What one device works on: rows of X, ODE rows (gather tables) and the sweep plan.
    For a single GPU this is the whole model; ``distributed.localize`` builds the per-rank
    restriction (owned states first, then halo rows)."""

    def __init__(self, n_rows, n_stats_odes, ode_class, ode_idx, ode_state, host_plan, p):
        self.n_rows, self.n_stats_odes, self.p = int(n_rows), int(n_stats_odes), int(p)
        self.ode_class = np.ascontiguousarray(ode_class, dtype=np.int32)
        self.ode_idx = np.ascontiguousarray(ode_idx, dtype=np.int32)
        self.ode_state = np.ascontiguousarray(ode_state, dtype=np.int32)
        self.n_odes = len(self.ode_class)
        self.host_plan = host_plan

    @classmethod
    def from_model(cls, model: OdeModel, hidden, couplings=None, schedule="reference"):
        hp = build_host_plan(model, hidden, couplings=couplings, schedule=schedule)
        return cls(model.K, model.K, model.ode_class, model.ode_idx, model.ode_state, hp, model.p)


def pipeline_ranges(problem, n_chunks, theta_mode="reference"):
    """State ranges for the pipelined ``vgm_iteration_host`` (vgm_model_tables.pipe_*), or None when the model /
    plan does not allow it.  Pure host logic (NumPy): ``problem`` is a ``LocalProblem``.  ``n_chunks`` is either
    the number of equal ranges or the increasing upper ends of the ranges as fractions of the states, ending in
    1.0 (a short first range lets the compute start early, a short last one shortens the final copy-out).

    Returns a dict with
      kp  [n+1]  state boundaries of the ranges (multiples of 8): range c is copied in and STAGED as a whole
      sp  [n+1]  hidden slots of each staged range
      op  [n+1]  hidden slots of each SOLVE range: the systems whose state lies in [kp[c] - hx, kp[c+1] - hx); op[0]
                 is the seam (slots of the states < hx, which the last range reads cyclically and which are solved last)
      rp  [n+1]  the same boundaries as positions within dependency level 0
      hx, hd     how far (in states, cyclically) the coefficients of a state reach into X and into D.X
    Invariants (tests/test_host_logic.py): every state a range's staging reads is either not a hidden state solved
    before that staging, i.e. all coefficients come from the sweep-start snapshot (proxies...py:191)."""
    K, hp = problem.n_rows, problem.host_plan
    hidden = np.asarray(hp.slot_state, dtype=np.int64)
    fractions = None
    if not np.isscalar(n_chunks):
        fractions = np.asarray(list(n_chunks), dtype=np.float64)
        if len(fractions) < 2 or np.any(np.diff(fractions) <= 0) or fractions[0] <= 0 or abs(fractions[-1] - 1.0) > 1e-12:
            raise ValueError("range ends must increase and finish at 1.0")
        n_chunks = len(fractions)
    if (n_chunks < 2 or theta_mode != "reference" or problem.n_stats_odes != K or problem.n_odes != K
            or hp.n_hidden < 2 or np.any(np.diff(hidden) <= 0)
            or not np.array_equal(problem.ode_state, np.arange(K))):
        return None
    arity = int(np.asarray(problem.ode_idx).size) // max(problem.n_odes, 1)

    def cyc(d):
        d = np.abs(d) % K
        return np.minimum(d, K - d)
    h = int(cyc(np.asarray(problem.ode_idx).reshape(K, arity).astype(np.int64) - np.arange(K)[:, None]).max()) if arity else 0
    pair_slot = np.repeat(np.arange(hp.n_hidden), np.diff(hp.pair_ptr))
    hc = int(cyc(hidden[pair_slot] - problem.ode_state[hp.pair_ode].astype(np.int64)).max()) if len(pair_slot) else 0
    hx, hd = h + hc, hc
    n_chunks = min(int(n_chunks), 32)
    ends = np.arange(n_chunks + 1) / n_chunks if fractions is None else np.concatenate([[0.0], fractions])
    kp = (np.round(ends * K / 8).astype(np.int64) * 8)
    kp[0], kp[-1] = 0, K
    if np.any(np.diff(kp) < 2 * hx + 8) or hx <= 0:
        return None
    lvl0 = np.asarray(hp.level_slots[:hp.level_ptr[1]], dtype=np.int64)
    later = np.asarray(hp.level_slots[hp.level_ptr[1]:], dtype=np.int64)
    sp = np.searchsorted(hidden, kp)
    if np.any(np.diff(lvl0) <= 0) or (len(later) and later.min() < sp[-2]):
        return None
    # solve ranges: shifted left by the halo; they start at the seam (states < hx) and end with the level
    ko = kp - hx
    ko[0], ko[-1] = hx, K
    op = np.searchsorted(hidden, ko)
    rp = np.searchsorted(lvl0, op)
    rp[-1] = len(lvl0)
    last_rows = np.concatenate([lvl0[rp[-2]:], lvl0[:rp[0]]]).astype(np.int32)
    return dict(n=n_chunks, hx=hx, hd=hd, kp=kp, sp=sp, rp=rp, op=op, last_rows=last_rows)


class VGMEngine:
    """Mean-field variational gradient matching on one GPU.

    Parameters
    ----------
    model : OdeModel
    D, A : T x T operators ``dC_times_inv_C`` and ``eps_cov`` (GP_regression.py:306,317)
    hidden : state indices to update, in the reference's order (proxies...py:175-179,195)
    theta_mode : 'reference' plugs the literal ``theta_literal`` (default [8.0], the value
        hard-coded at proxies...py:211) into R_uk, r_uk; 'proxy' plugs the current estimate.
    solver : 'mixed' = FP64 residuals on the DMMA pipe + FP32 PCG corrections whose operator products run
        on the tcgen05 tensor cores (bf16-split operands); 'pcg' = FP64 preconditioned CG only;
        'auto' = the direct per-CTA solver ('dense', csrc/dense_solve.cu) for T <= 160, else 'mixed' whenever the
        model allows it (R_uu constant, T >= MIXED_MIN_T), else 'pcg'.  ``self.solver`` holds the resolved choice.
    """
    MIXED_MIN_T = 128
    # preconditioner S G S of the mixed-precision solver: G = (M + SHIFT_FRAC mean(Lambda) I)^-1,
    # S = diag sqrt((shift + c) / (Lambda + c)), c = SCALE_FRAC mean(Lambda); spec(S G S A) then has a
    # condition number of ~6-7 instead of ~20 for the plain shifted inverse (Lorenz96, C3/C5 of BASELINE.json)
    SHIFT_FRAC, SCALE_FRAC = 0.55, 1.5

    def __init__(self, model: OdeModel, D, A=None, hidden=None, couplings=None, schedule="reference",
                 theta_mode="reference", theta_literal=(8.0,), tol=1e-13, max_iter=0, check_every=8,
                 precondition=True, solver="auto", inner_iters=10, optimistic_passes=4, small_level_solver=False,
                 prefer_builtin=True,
                 device=None, problem=None):
        self.lib = _lib.load()
        self.dev = device or _lib.require_cuda()
        self.model = model
        self.ops = D if isinstance(D, Operators) else Operators(D, A, device=self.dev)
        self.T, self.ld = self.ops.T, self.ops.ld
        if problem is None:
            if hidden is None:
                hidden = np.arange(model.K)
            problem = LocalProblem.from_model(model, hidden, couplings=couplings, schedule=schedule)
        self.problem = problem
        self.K, self.p = problem.n_rows, model.p
        self.program = Program(model, prefer_builtin=prefer_builtin)
        self.tables = {"ode_class": _i32(problem.ode_class, self.dev), "ode_idx": _i32(problem.ode_idx, self.dev),
                       "ode_state": _i32(problem.ode_state, self.dev)}
        self.plan = DevicePlan(problem.host_plan, self.T, self.ld, self.tables, device=self.dev)
        if theta_mode not in ("reference", "proxy"):
            raise ValueError("theta_mode must be 'reference' or 'proxy'")
        self.theta_mode = theta_mode
        self.theta_literal = tuple(float(v) for v in theta_literal)
        if theta_mode == "reference" and len(self.theta_literal) != self.p:
            raise TypeError("the reference plugs the literal [8] into R_uk, r_uk (proxies...py:211); a model with %d "
                            "parameters needs theta_mode='proxy' or a theta_literal of that length" % self.p)
        self.precondition = bool(precondition)
        if solver not in ("auto", "mixed", "pcg"):
            raise ValueError("solver must be 'auto', 'mixed' or 'pcg'")
        # 'mixed': FP64 residuals + FP32/bf16-split tensor-core corrections (needs a shared system matrix, i.e.
        # R_uu constant, and the preconditioner); 'pcg': FP64 preconditioned CG on the DMMA pipe only
        can_mix = bool(self.plan.c.ruu_const) and self.precondition and self.T >= self.MIXED_MIN_T
        if solver == "mixed" and not can_mix:
            raise ValueError("solver='mixed' needs constant R_uu, precondition=True and T >= %d" % self.MIXED_MIN_T)
        # 'dense' (only through solver='auto'): T <= 160, every system is formed and factorised by its own CTA
        can_dense = solver == "auto" and self.T <= Operators.DENSE_MAX_T
        self.solver = "dense" if can_dense else ("mixed" if (solver in ("auto", "mixed") and can_mix) else "pcg")
        # solver='auto' lets short time grids (T <= 160) take the direct per-CTA solver; an explicit choice is honoured
        self.opts = _lib.SolverOpts(tol, max_iter, check_every, int(inner_iters), int(optimistic_passes),
                                    1 if small_level_solver else 0, 0 if solver == "auto" else -1)
        self.X = torch.zeros((self.K, self.ld), dtype=torch.float64, device=self.dev)
        self.DX = torch.zeros_like(self.X)
        self.theta = torch.zeros(max(self.p, 1), dtype=torch.float64, device=self.dev)
        self.theta_star = torch.zeros(max(self.p, 1), dtype=torch.float64, device=self.dev)
        self.stats = torch.zeros(self.p + self.p * self.p, dtype=torch.float64, device=self.dev)
        self._cops = _lib.Operators()
        self._obs_rhs = None
        self._cops.D, self._cops.DT = self.ops.D.data_ptr(), self.ops.DT.data_ptr()
        self._cops.band_D = self.ops.band_D
        if not self.plan.c.ruu_const and self.T <= Operators.DENSE_MAX_T:
            self._cops.DtD = self.ops.ensure_DtD().data_ptr()
        if self.plan.c.ruu_const:
            self._cops.M = self.ops.ensure_shared_M(self.plan.c.ruu_value).data_ptr()
            self._cops.band_M = self.ops.band_M
        self._mt = _lib.ModelTables()
        self._mt.n_states, self._mt.n_odes, self._mt.n_param = self.K, problem.n_stats_odes, self.p
        self._mt.ode_class = self.tables["ode_class"].data_ptr()
        self._mt.ode_idx = self.tables["ode_idx"].data_ptr()
        self._mt.ode_state = self.tables["ode_state"].data_ptr()
        self._mt.theta_mode = 0 if theta_mode == "reference" else 1
        for i, v in enumerate(self.theta_literal[:8]):
            self._mt.theta_literal[i] = v
        n_ws = max(self.lib.vgm_iteration_ws_bytes(C.byref(self._mt), C.byref(self.plan.c)), 1024)
        self.ws = torch.empty(n_ws, dtype=torch.uint8, device=self.dev)
        self.info = _lib.SweepInfo()
        self.last_info = None
        self._pipe_cache = {}
        self.last_pipeline = 0

    # ---- opt-in observation term -------------------------------------------------------------
    def set_observation_mean(self, mu_TK):
        """Observation term of the state update (opt-in): with ``Operators(obs_inv_cov=Q)`` and ``mu`` (T x K, the GP
        fit ``state_pred_mean``) state u solves ``(local_scaling + Q) x = local_mean + Q mu_u`` -- the statement the
        reference keeps commented out (proxies...py:265-266; active in the MATLAB original,
        archive/VGM_functions/proxy_for_ind_states.m:92-93).  ``None`` switches it off again."""
        if mu_TK is None:
            self._obs_rhs = None
            self._cops.obs_rhs = None
            return self
        if self.ops.Q is None:
            raise ValueError("the observation term needs Operators(..., obs_inv_cov=Q)")
        if not self.plan.c.ruu_const and self.T > Operators.DENSE_MAX_T:
            raise NotImplementedError("observation term with a time-varying R_uu needs T <= %d (direct per-CTA solver)"
                                      % Operators.DENSE_MAX_T)
        mu = torch.as_tensor(np.asarray(mu_TK, dtype=np.float64), device=self.dev)
        if mu.shape != (self.T, self.K):
            raise ValueError("state_pred_mean must be T x K")
        nh = max(self.plan.n_hidden, 1)
        rows = torch.zeros((nh, self.ld), dtype=torch.float64, device=self.dev)
        rows[:self.plan.n_hidden, :self.T] = mu.t()[torch.as_tensor(self.plan.hidden, device=self.dev)]
        out = torch.zeros_like(rows)
        g = _lib.GemmArgs()
        g.A, g.B, g.C = rows.data_ptr(), self.ops.Q.data_ptr(), out.data_ptr()      # Q is symmetric: rows . Q^T
        g.lda = g.ldb = g.ldc = self.ld
        g.N, g.M, g.K = self.plan.n_hidden, self.T, self.T
        g.alpha = 1.0
        if self.plan.n_hidden:
            _lib.check(self.lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr()))
        self._obs_rhs = out
        self._cops.obs_rhs = out.data_ptr()
        return self

    # ---- data movement ---------------------------------------------------------------------
    def set_states(self, X_TK):
        """Upload a T x K array (reference orientation)."""
        X = torch.from_numpy(np.array(X_TK, dtype=np.float64, copy=True)) if not torch.is_tensor(X_TK) else X_TK
        if X.shape != (self.T, self.K):
            raise ValueError("Either state_proxy or dC_times_invC have the wrong shape")
        self.X.zero_()
        self.X[:, :self.T] = X.to(self.dev, dtype=torch.float64).t()
        return self

    def set_states_state_major(self, X_KT):
        self.X.zero_()
        self.X[:, :self.T] = torch.as_tensor(X_KT, dtype=torch.float64, device=self.dev)[:, :self.T]
        return self

    def get_states(self):
        """T x K numpy array (reference orientation)."""
        return self.X[:, :self.T].t().contiguous().cpu().numpy()

    def get_theta(self):
        return self.theta[:self.p].cpu().numpy()

    # ---- preconditioner -----------------------------------------------------------------------
    def _prepare_preconditioner(self):
        if not (self.precondition and self.plan.c.ruu_const) or self.plan.n_hidden == 0 or self.solver == "dense":
            self._cops.G = None
            self._cops.G_bf16 = None
            self._cops.M_bf16x2 = None
            return
        if self.ops.G is None:
            # shift = mean of Lambda over all hidden states/time points of the current proxy
            nh, ld = self.plan.n_hidden, self.ld
            lam = torch.empty((nh, ld), dtype=torch.float64, device=self.dev)
            tmp = [torch.empty_like(lam) for _ in range(2)]
            nlp = max(self.plan.c.n_live_pairs, 1)
            rlive = torch.empty((nlp, ld), dtype=torch.float64, device=self.dev)
            self._set_theta_star()
            _lib.check(self.lib.vgm_state_assemble(self.program.handle, C.byref(self.plan.c), self.X.data_ptr(),
                                                   self.DX.data_ptr(), self.theta_star.data_ptr(), lam.data_ptr(),
                                                   tmp[0].data_ptr(), tmp[1].data_ptr(), None, rlive.data_ptr(),
                                                   _lib.stream_ptr()))
            lam_mean = float(lam[:, :self.T].mean().item())
            if not np.isfinite(lam_mean) or lam_mean <= 0:
                lam_mean = 1.0
            self.lam_mean = lam_mean
            self.ops.ensure_preconditioner((self.SHIFT_FRAC if self.solver == "mixed" else 1.0) * lam_mean)
        mixed = self.solver == "mixed"
        self._cops.G = self.ops.G.data_ptr()
        self._cops.G_bf16 = self.ops.G_bf16.data_ptr() if mixed else None
        self._cops.M_bf16x2 = self.ops.M_bf16x2.data_ptr() if mixed else None
        self._cops.G_shift = self.ops.G_shift
        self._cops.precond_c = self.SCALE_FRAC * self.ops.G_shift / self.SHIFT_FRAC if mixed else 0.0
        self._cops.band_G = self.ops.band_G
        self._cops.band_M_lp = self.ops.band_M_lp

    def _set_theta_star(self, theta_star=None):
        if theta_star is not None:
            self.theta_star[:self.p] = torch.as_tensor(np.asarray(theta_star, dtype=np.float64).ravel(), device=self.dev)
        elif self.theta_mode == "reference":
            self.theta_star[:self.p] = torch.tensor(self.theta_literal, dtype=torch.float64, device=self.dev)
        else:
            self.theta_star[:self.p] = self.theta[:self.p]

    # ---- the three building blocks --------------------------------------------------------------
    def compute_DX(self):
        """DX[k, :] = D @ x_k for every state (one DMMA GEMM)."""
        g = _lib.GemmArgs()
        g.A, g.B, g.C = self.X.data_ptr(), self.ops.D.data_ptr(), self.DX.data_ptr()
        g.lda = g.ldb = g.ldc = self.ld
        g.N, g.M, g.K = self.K, self.T, self.T
        g.alpha = 1.0
        g.band = self.ops.band_D
        _lib.check(self.lib.vgm_dgemm_nt(C.byref(g), _lib.stream_ptr()))

    def param_stats(self):
        """Sufficient statistics [m | S] of the theta update on the device."""
        self.model.require_linear_in_parameters()
        n_odes = self.problem.n_stats_odes
        n = self.lib.vgm_param_stats_ws_bytes(n_odes, self.T, self.p)
        _lib.check(self.lib.vgm_param_stats(self.program.handle, self.X.data_ptr(), self.DX.data_ptr(), self.ld, self.T,
                                            n_odes, self._mt.ode_class, self._mt.ode_idx, self._mt.ode_state,
                                            self.stats.data_ptr(), self.ws.data_ptr(), n, _lib.stream_ptr()))
        return self.stats

    def param_loss(self, theta, alpha=0.0):
        """Squared gradient-matching loss of the 'minimizer' parameter update and its gradient
        (proxies...py:379-392): ``sum (f(X, theta) - D X)^2 + alpha |theta|^2`` on the current states; D.X must be
        current (``compute_DX``).  Works for models that are not linear in their parameters.  Returns (loss, grad)."""
        th = torch.as_tensor(np.array(theta, dtype=np.float64, copy=True).ravel(), device=self.dev)
        if th.numel() != self.p:
            raise ValueError("theta must have %d entries" % self.p)
        n_odes = self.problem.n_stats_odes
        n = self.lib.vgm_param_loss_ws_bytes(n_odes, self.p)
        if getattr(self, "_loss_ws", None) is None or self._loss_ws.numel() < n:
            self._loss_ws = torch.empty(max(n, 256), dtype=torch.uint8, device=self.dev)
            self._loss_out = torch.zeros(self.p + 1, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.vgm_param_loss(self.program.handle, self.X.data_ptr(), self.DX.data_ptr(), self.ld, self.T,
                                           n_odes, self._mt.ode_class, self._mt.ode_idx, self._mt.ode_state, th.data_ptr(),
                                           self._loss_out.data_ptr(), self._loss_ws.data_ptr(), n, _lib.stream_ptr()))
        out = self._loss_out.cpu().numpy()
        thn = th.cpu().numpy()
        return float(out[0] + alpha * np.sum(thn ** 2)), out[1:] + 2.0 * alpha * thn

    def param_solve(self):
        _lib.check(self.lib.vgm_param_solve(self.stats.data_ptr(), self.p, self.theta.data_ptr(), _lib.stream_ptr()))
        return self.theta

    def theta_step(self):
        """proxy_for_ode_parameters(optimizer='analytical', constraints=None)."""
        self.compute_DX()
        self.param_stats()
        self.param_solve()
        return self.theta[:self.p]

    def state_sweep(self, theta_star=None, dx_is_current=True, allow_unconverged=False):
        """proxy_for_ind_states(optimizer='analytical', constraints=None): one sweep."""
        if not dx_is_current:
            self.compute_DX()
        self._set_theta_star(theta_star)
        self._prepare_preconditioner()
        n = self.lib.vgm_state_sweep_ws_bytes(C.byref(self.plan.c))
        rc = self.lib.vgm_state_sweep(self.program.handle, C.byref(self.plan.c), C.byref(self._cops),
                                      C.byref(self.opts), self.X.data_ptr(), self.DX.data_ptr(),
                                      self.theta_star.data_ptr(), self.ws.data_ptr(), n, C.byref(self.info),
                                      _lib.stream_ptr())
        self.last_info = {k: getattr(self.info, k) for k, _ in _lib.SweepInfo._fields_}
        _lib.check(rc, allow_notconv=allow_unconverged)
        return self.last_info

    def sweep_prepare(self, theta_star=None):
        """Phase 1 of a sweep (vgm_state_sweep_prepare)."""
        self._set_theta_star(theta_star)
        self._prepare_preconditioner()
        n = self.lib.vgm_state_sweep_ws_bytes(C.byref(self.plan.c))
        _lib.check(self.lib.vgm_state_sweep_prepare(self.program.handle, C.byref(self.plan.c), C.byref(self._cops),
                                                    self.X.data_ptr(), self.DX.data_ptr(), self.theta_star.data_ptr(),
                                                    self.ws.data_ptr(), n, C.byref(self.info), _lib.stream_ptr()))

    def sweep_level(self, level, allow_unconverged=False):
        """Phase 2 of a sweep for one dependency level (vgm_state_sweep_level)."""
        n = self.lib.vgm_state_sweep_ws_bytes(C.byref(self.plan.c))
        rc = self.lib.vgm_state_sweep_level(C.byref(self.plan.c), C.byref(self._cops), C.byref(self.opts),
                                            self.X.data_ptr(), int(level), self.ws.data_ptr(), n, C.byref(self.info),
                                            _lib.stream_ptr())
        self.last_info = {k: getattr(self.info, k) for k, _ in _lib.SweepInfo._fields_}
        _lib.check(rc, allow_notconv=allow_unconverged)

    def sweep_levels(self, allow_unconverged=False):
        """Phase 2 of a sweep for every level in order (vgm_state_sweep_levels), tail overlap included."""
        n = self.lib.vgm_state_sweep_ws_bytes(C.byref(self.plan.c))
        rc = self.lib.vgm_state_sweep_levels(C.byref(self.plan.c), C.byref(self._cops), C.byref(self.opts),
                                             self.X.data_ptr(), self.ws.data_ptr(), n, C.byref(self.info),
                                             _lib.stream_ptr())
        self.last_info = {k: getattr(self.info, k) for k, _ in _lib.SweepInfo._fields_}
        _lib.check(rc, allow_notconv=allow_unconverged)

    def iteration(self, allow_unconverged=False):
        """theta update followed by the state sweep, all on the device (vgm_iteration_device)."""
        self.model.require_linear_in_parameters()
        if self._cops.G is None and self.precondition and self.plan.c.ruu_const and self.plan.n_hidden:
            self.compute_DX()
            self._prepare_preconditioner()
        rc = self.lib.vgm_iteration_device(self.program.handle, C.byref(self._mt), C.byref(self.plan.c),
                                           C.byref(self._cops), C.byref(self.opts), self.X.data_ptr(),
                                           self.DX.data_ptr(), self.theta.data_ptr(), self.stats.data_ptr(),
                                           self.ws.data_ptr(), self.ws.numel(), C.byref(self.info), _lib.stream_ptr())
        self.last_info = {k: getattr(self.info, k) for k, _ in _lib.SweepInfo._fields_}
        _lib.check(rc, allow_notconv=allow_unconverged)
        return self.last_info

    def _pipeline_tables(self, n_chunks):
        """ctypes view of ``pipeline_ranges`` for this engine (or None)."""
        pr = pipeline_ranges(self.problem, n_chunks, self.theta_mode)
        if pr is None:
            return None
        return dict(n=pr["n"], hx=pr["hx"], hd=pr["hd"], kp=_host_i32(pr["kp"]), sp=_host_i32(pr["sp"]),
                    rp=_host_i32(pr["rp"]), op=_host_i32(pr["op"]), last_rows=_i32(pr["last_rows"], self.dev))

    def iteration_host(self, X_host_KT: torch.Tensor, theta_host: torch.Tensor, X_host_out: torch.Tensor = None,
                       allow_unconverged=False, updated_rows_only=False, pipeline_chunks=0):
        """End-to-end iteration with HOST buffers (vgm_iteration_host): ``X_host_KT`` is a
        (pinned) CPU tensor [K, ld]; the updated proxies go to ``X_host_out`` (default: in
        place); ``theta_host`` a CPU tensor [p].  ``updated_rows_only``: copy back only the rows of the
        hidden states (the others are not changed by the sweep and are left as they are in ``X_host_out``);
        needs the hidden states to be equally spaced rows, otherwise everything is copied.
        ``pipeline_chunks`` >= 2: cut the states into that many ranges and overlap the host->device copy of range
        c+1 and the device->host copy of range c-1 with the solve of range c (falls back to the plain path when the
        model does not allow it; ``self.last_pipeline`` tells which ran)."""
        self.model.require_linear_in_parameters()
        if X_host_out is None:
            X_host_out = X_host_KT
        pkey = pipeline_chunks if np.isscalar(pipeline_chunks) else tuple(pipeline_chunks)
        if pkey and pkey not in self._pipe_cache:
            self._pipe_cache[pkey] = self._pipeline_tables(pipeline_chunks)
        pt = self._pipe_cache.get(pkey) if pkey else None
        self.last_pipeline = pt["n"] if pt else 0
        m = self._mt
        m.pipe_chunks = 0
        if pt:
            m.pipe_chunks, m.pipe_halo_x, m.pipe_halo_dx = pt["n"], pt["hx"], pt["hd"]
            m.pipe_state_ptr_host = C.cast(pt["kp"], C.POINTER(C.c_int))
            m.pipe_slot_ptr_host = C.cast(pt["sp"], C.POINTER(C.c_int))
            m.pipe_row_ptr_host = C.cast(pt["rp"], C.POINTER(C.c_int))
            m.pipe_out_slot_ptr_host = C.cast(pt["op"], C.POINTER(C.c_int))
            m.pipe_last_rows = pt["last_rows"].data_ptr()
            self._pipe_keep = pt
        self._mt.out_n_rows = 0
        h = self.plan.hidden
        if updated_rows_only and len(h) > 1:
            step = int(h[1] - h[0])
            if step > 0 and np.all(np.diff(np.sort(h)) == step):
                self._mt.out_row0, self._mt.out_row_stride, self._mt.out_n_rows = int(h.min()), step, len(h)
        if self._cops.G is None and self.precondition and self.plan.c.ruu_const and self.plan.n_hidden:
            self.X.copy_(X_host_KT, non_blocking=True)
            self.compute_DX()
            self._prepare_preconditioner()
        rc = self.lib.vgm_iteration_host(self.program.handle, C.byref(self._mt), C.byref(self.plan.c),
                                         C.byref(self._cops), C.byref(self.opts), X_host_KT.data_ptr(),
                                         X_host_out.data_ptr(), theta_host.data_ptr(), self.X.data_ptr(), self.DX.data_ptr(),
                                         self.theta.data_ptr(), self.stats.data_ptr(), self.ws.data_ptr(),
                                         self.ws.numel(), C.byref(self.info), _lib.stream_ptr())
        self.last_info = {k: getattr(self.info, k) for k, _ in _lib.SweepInfo._fields_}
        m.pipe_chunks = 0
        _lib.check(rc, allow_notconv=allow_unconverged)
        return self.last_info

    # ---- opt-in covariance (P5) -------------------------------------------------------------------
    def param_cov(self):
        """p x p matrix sum_k B_k^T A B_k of the parameter update (proxies...py:79; the reference computes it and
        drops it).  Evaluated on the current states."""
        if self.ops.A is None:
            raise ValueError("eps_cov (A) was not given")
        self.model.require_linear_in_parameters()
        n_odes = self.problem.n_stats_odes
        n = self.lib.vgm_param_cov_ws_bytes(n_odes, self.ld, self.p)
        ws = torch.empty(n + 256, dtype=torch.uint8, device=self.dev)
        cov = torch.zeros(self.p * self.p, dtype=torch.float64, device=self.dev)
        _lib.check(self.lib.vgm_param_cov(self.program.handle, self.X.data_ptr(), self.ops.A.data_ptr(), self.ld, self.T,
                                          n_odes, self._mt.ode_class, self._mt.ode_idx, cov.data_ptr(), ws.data_ptr(), n,
                                          _lib.stream_ptr()))
        return cov.cpu().numpy().reshape(self.p, self.p)

    def state_cov(self, state_index, theta_star=None):
        """T x T matrix sum_k B_uk^T A B_uk of one hidden state (proxies...py:239)."""
        if self.ops.A is None:
            raise ValueError("eps_cov (A) was not given")
        slots = np.nonzero(self.plan.hidden == int(state_index))[0]
        if len(slots) == 0:
            raise ValueError("state %r is not a hidden state of this plan" % (state_index,))
        self._set_theta_star(theta_star)
        cov = torch.zeros((self.T, self.ld), dtype=torch.float64, device=self.dev)
        n = self.lib.vgm_state_cov_ws_bytes(self.T, self.ld)
        ws = torch.empty(n + 256, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.vgm_state_cov(self.program.handle, C.byref(self.plan.c), C.byref(self._cops),
                                          self.ops.A.data_ptr(), self.X.data_ptr(), self.theta_star.data_ptr(),
                                          int(slots[0]), cov.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr()))
        return cov[:, :self.T].cpu().numpy()
