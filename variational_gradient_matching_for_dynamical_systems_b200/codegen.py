"""Model front-end and CUDA code generator.

Replaces ``sympy.lambdify`` in the reference's rewrites
(``rewrite_odes_as_local_linear_combinations.py:28-58`` for B_k, b_k and ``:76-124`` for
R_uk, r_uk) by CUDA device functions compiled with NVRTC for sm_100a, and replaces the
reference's O(K^2) SymPy set-up (``import_odes.py:18-29,73-87``; ``rewrite...py:106-119``
re-evaluates every ODE for every (state, ODE) pair) by *structural classes*: ODE lines that
are identical up to renaming of their state symbols share one symbolic rewrite and one device
function, and differ only in a row of an ``int32`` gather table.  Lorenz96 with any K is a
single class of arity 4.

The symbolic steps are the reference's own -- ``sympify(...).factor()`` (import_odes.py:25),
``.expand()`` + ``linear_eq_to_matrix`` (rewrite...py:45,108) -- including its handling of
symbol-free entries:

* parameter rewrite: a symbol-free entry of B_k or b_k is REPLACED by the ones vector
  (rewrite...py:49-52), i.e. 0 -> 1 and c -> 1.  ``zero_fill=True`` keeps the constant instead
  (the behaviour of the MATLAB original);
* state rewrite: a symbol-free R_uk / r_uk is MULTIPLIED by the ones vector (rewrite...py:112-115).
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import sympy as sym

from . import _lib

# identifiers, but not the exponent of a numeric literal ("1e2" must not be read as the state name "e2")
_IDENT = re.compile(r"(?<![0-9.])[A-Za-z_][A-Za-z_0-9]*")
_HERE = os.path.dirname(os.path.abspath(__file__))
_TEMPLATE = os.path.join(_HERE, "csrc", "vgm_program_template.cuh")


# --------------------------------------------------------------------------------------
# CUDA expression printer
# --------------------------------------------------------------------------------------
from sympy.printing.c import C99CodePrinter  # noqa: E402


class _CudaPrinter(C99CodePrinter):
    def _print_Pow(self, expr):
        b, e = expr.as_base_exp()
        if e.is_Integer and 2 <= int(e) <= 4:
            s = self._print(b)
            if not b.is_Atom:
                s = "(" + s + ")"
            return "(" + "*".join([s] * int(e)) + ")"
        if e.is_Integer and -4 <= int(e) <= -1:
            s = self._print(b)
            if not b.is_Atom:
                s = "(" + s + ")"
            return "(1.0/(" + "*".join([s] * (-int(e))) + "))"
        return super()._print_Pow(expr)

    def _print_Integer(self, expr):
        return "%d.0" % int(expr)

    def _print_Rational(self, expr):
        return "(%d.0/%d.0)" % (int(expr.p), int(expr.q))


_printer = _CudaPrinter({"precision": 17})


def _ccode(expr, subs):
    return _printer.doprint(sym.sympify(expr).xreplace(subs))


# --------------------------------------------------------------------------------------
# Structural classes
# --------------------------------------------------------------------------------------
class OdeClass:
    """One structural class of ODE right-hand sides (placeholders s0..s_{a-1} for states)."""

    def __init__(self, template, placeholders, param_symbols, zero_fill):
        self.template = template
        self.arity = len(placeholders)
        self.placeholders = placeholders
        expr = sym.sympify(template).factor()                      # import_odes.py:25
        self.expr = expr
        self.present = [s in expr.free_symbols for s in placeholders]
        ex = expr.expand()                                         # rewrite...py:45,108
        # ---- linear in the parameters --------------------------------------------------
        # A right-hand side that is NOT linear in a parameter (Protein Transduction: 1/(_K_m + _R_pp),
        # FitzHugh-Nagumo: /_tau) still imports, has couplings and a state rewrite in the reference; only its
        # parameter rewrite raises (rewrite...py:45).  Same here: the error is kept and raised by the parameter
        # rewrite / the theta statistics, everything else works (state sweep with theta_mode='proxy').
        one = sym.Integer(1)
        self.param_nonlinear, self.param_error = False, None
        self.B, self.b = [sym.Integer(0)] * len(param_symbols), sym.Integer(0)
        try:
            eB, eb = sym.linear_eq_to_matrix([ex], list(param_symbols))
        except Exception as err:                                   # sympy NonlinearError
            self.param_nonlinear, self.param_error = True, err
            eB = None
        if eB is not None:
            eb = -eb
            self.B = []
            for e in list(eB):
                if len(e.free_symbols) == 0:
                    e = e if zero_fill else one                    # rewrite...py:49-50
                self.B.append(e)
            e = eb[0]
            if len(e.free_symbols) == 0:
                e = e if zero_fill else one                        # rewrite...py:51-52
            self.b = e
        # ---- linear in each individual state ---------------------------------------------
        self.R, self.r, self.nonlinear = [], [], []
        for j, s in enumerate(placeholders):
            if not self.present[j]:
                self.R.append(sym.Integer(0)); self.r.append(sym.Integer(0)); self.nonlinear.append(False)
                continue
            try:
                eR, er = sym.linear_eq_to_matrix([ex], [s])        # rewrite...py:108
                self.R.append(eR[0]); self.r.append(-er[0]); self.nonlinear.append(False)
            except Exception:                                      # NonlinearError: not linear in x_u
                self.R.append(sym.Integer(0)); self.r.append(sym.Integer(0)); self.nonlinear.append(True)


class OdeModel:
    """A ``*_ODEs.txt`` model: one right-hand side per line, line i <-> ``state_names[i]``
    (import_odes.py:22-29).  ``n_copies`` > 1 replicates the model for a batch of independent
    trajectories: copy n owns states n*K .. n*K+K-1 (the ODE parameters are shared)."""

    def __init__(self, lines, state_names, param_names, zero_fill=False, n_copies=1):
        lines = [ln for ln in lines if ln.strip()]
        self.state_names = [str(s) for s in state_names]
        self.param_names = [str(s) for s in param_names]
        if len(lines) != len(self.state_names):
            raise ValueError("number of ODE lines (%d) != number of states (%d)" % (len(lines), len(self.state_names)))
        self.lines = lines
        self.K1 = len(lines)                       # states of ONE trajectory
        self.n_copies = int(n_copies)
        self.K = self.K1 * self.n_copies
        self.p = len(self.param_names)
        self.zero_fill = bool(zero_fill)
        self.param_symbols = [sym.Symbol(n) for n in self.param_names]
        index = {n: i for i, n in enumerate(self.state_names)}
        tmpl_id, templates, rows = {}, [], []
        ode_class = np.empty(self.K1, dtype=np.int32)
        for k, ln in enumerate(lines):
            order = {}

            def rep(m):
                name = m.group(0)
                if name in index:
                    if name not in order:
                        order[name] = len(order)
                    return "vgmS%d" % order[name]
                return name
            tmpl = _IDENT.sub(rep, ln.strip())
            tmpl = re.sub(r"\s+", "", tmpl)
            if tmpl not in tmpl_id:
                tmpl_id[tmpl] = len(templates)
                templates.append((tmpl, len(order)))
            ode_class[k] = tmpl_id[tmpl]
            rows.append([index[n] for n in order])
        self.arity = max(1, max(a for _, a in templates))
        self.classes = []
        for tmpl, a in templates:
            ph = [sym.Symbol("vgmS%d" % j) for j in range(a)]
            self.classes.append(OdeClass(tmpl, ph, self.param_symbols, self.zero_fill))
        idx = np.zeros((self.K1, self.arity), dtype=np.int32)
        for k, r in enumerate(rows):
            idx[k, :len(r)] = r
        self.code_base = np.cumsum([0] + [c.arity for c in self.classes]).astype(np.int32)
        if self.n_copies > 1:
            off = (np.arange(self.n_copies, dtype=np.int32) * self.K1)[:, None, None]
            self.ode_idx = (idx[None, :, :] + off).reshape(self.K, self.arity).astype(np.int32)
            self.ode_class = np.tile(ode_class, self.n_copies)
        else:
            self.ode_idx, self.ode_class = idx, ode_class
        self.ode_state = np.arange(self.K, dtype=np.int32)
        self._arity_of = np.array([c.arity for c in self.classes], dtype=np.int32)
        self._pairs = None

    @property
    def param_nonlinear(self):
        """True if some right-hand side is not linear in the ODE parameters (no closed-form theta update)."""
        return any(c.param_nonlinear for c in self.classes)

    def require_linear_in_parameters(self):
        """Raise what the reference's parameter rewrite raises (rewrite...py:45) for such a model."""
        for c in self.classes:
            if c.param_nonlinear:
                raise c.param_error

    # ---- construction helpers -----------------------------------------------------------
    @classmethod
    def from_file(cls, path, state_names, param_names, **kw):
        with open(path) as f:
            return cls(f.read().splitlines(), state_names, param_names, **kw)

    @classmethod
    def lorenz96(cls, K, **kw):
        """Lorenz96 with K states in the format of ``Lorenz96_ODEs.txt:1-10``."""
        return cls(lorenz96_ode_lines(K), ["_x_%d" % (i + 1) for i in range(K)], ["_alpha"], **kw)

    # ---- couplings ------------------------------------------------------------------------
    def pairs(self):
        """All (state u, ODE k, code) triples with x_u present in f_k, sorted by (u, k):
        the reference's ``state_couplings`` (import_odes.py:85) in CSR form."""
        if self._pairs is None:
            present = np.zeros((len(self.classes), self.arity), dtype=bool)
            for ci, c in enumerate(self.classes):
                present[ci, :c.arity] = c.present
            cls_k = self.ode_class
            mask = present[cls_k]                                      # [K, arity]
            kk, jj = np.nonzero(mask)
            uu = self.ode_idx[kk, jj]
            code = self.code_base[cls_k[kk]] + jj
            order = np.lexsort((kk, uu))
            self._pairs = (uu[order].astype(np.int64), kk[order].astype(np.int64), code[order].astype(np.int32))
        return self._pairs

    def state_couplings(self):
        """List-of-lists form of import_odes.find_state_couplings_in_odes (small models only)."""
        uu, kk, _ = self.pairs()
        out = [[] for _ in range(self.K)]
        for u, k in zip(uu.tolist(), kk.tolist()):
            out[u].append(k)
        return out

    def is_builtin_lorenz96(self):
        if self.p != 1 or len(self.classes) != 1 or self.classes[0].arity != 4 or self.zero_fill:
            return False
        s = self.classes[0].placeholders
        target = (s[0] - s[1]) * s[2] - s[3] + self.param_symbols[0]
        return sym.expand(self.classes[0].expr - target) == 0

    # ---- code generation ---------------------------------------------------------------------
    def cuda_source(self):
        """CUDA source of the model's coefficient device functions + the kernel template."""
        subs_all = {}
        for i, ps in enumerate(self.param_symbols):
            subs_all[ps] = sym.Symbol("th[%d]" % i)
        out = ["// generated by variational_gradient_matching_for_dynamical_systems_b200.codegen",
               "#define VGM_PROG(name) vgm_prog_##name",
               "#define VGM_NPARAM %d" % self.p,
               "#define VGM_ARITY %d" % self.arity,
               "__device__ __forceinline__ void vgm_theta_coeffs(int cls, const double (&s)[VGM_ARITY], "
               "double (&B)[VGM_NPARAM], double& b) {",
               "  switch (cls) {"]
        for ci, c in enumerate(self.classes):
            subs = {ph: sym.Symbol("s[%d]" % j) for j, ph in enumerate(c.placeholders)}
            out.append("    case %d: {" % ci)
            if c.param_nonlinear:                  # never evaluated: the host refuses the theta statistics
                out.append("      for (int i = 0; i < VGM_NPARAM; ++i) B[i] = 0.0; b = 0.0;")
                out.append("    } break;")
                continue
            for i, e in enumerate(c.B):
                out.append("      B[%d] = %s;" % (i, _ccode(e, subs)))
            out.append("      b = %s;" % _ccode(c.b, subs))
            out.append("    } break;")
        out.append("    default: {")
        for i in range(self.p):
            out.append("      B[%d] = 0.0;" % i)
        out += ["      b = 0.0;", "    } break;", "  }", "}",
                "__device__ __forceinline__ void vgm_state_coeffs(int code, const double* th, "
                "const double (&s)[VGM_ARITY], double& R, double& r) {",
                "  switch (code) {"]
        for ci, c in enumerate(self.classes):
            subs = {ph: sym.Symbol("s[%d]" % j) for j, ph in enumerate(c.placeholders)}
            subs.update(subs_all)
            for j in range(c.arity):
                if not c.present[j] or c.nonlinear[j]:
                    continue
                out.append("    case %d: { R = %s; r = %s; } break;" % (
                    int(self.code_base[ci]) + j, _ccode(c.R[j], subs), _ccode(c.r[j], subs)))
        out += ["    default: { R = 0.0; r = 0.0; } break;", "  }", "}",
                "__device__ __forceinline__ void vgm_ode_eval(int cls, const double* th, const double (&s)[VGM_ARITY], "
                "double& f, double (&df)[VGM_NPARAM]) {",
                "  switch (cls) {"]
        for ci, c in enumerate(self.classes):
            subs = {ph: sym.Symbol("s[%d]" % j) for j, ph in enumerate(c.placeholders)}
            subs.update(subs_all)
            out.append("    case %d: {" % ci)
            out.append("      f = %s;" % _ccode(c.expr, subs))
            for i, ps in enumerate(self.param_symbols):
                out.append("      df[%d] = %s;" % (i, _ccode(sym.diff(c.expr, ps), subs)))     # import_odes.py:40-47
            out.append("    } break;")
        out.append("    default: {")
        for i in range(self.p):
            out.append("      df[%d] = 0.0;" % i)
        out += ["      f = 0.0;", "    } break;", "  }", "}", ""]
        with open(_TEMPLATE) as f:
            out.append(f.read())
        return "\n".join(out)

    def compile_cubin(self):
        return nvrtc_compile(self.cuda_source())


def lorenz96_ode_lines(K):
    n = lambda j: "_x_%d" % ((j % K) + 1)
    return ["(%s - %s) * %s - %s + _alpha" % (n(k + 1), n(k - 2), n(k - 1), n(k)) for k in range(K)]


def write_lorenz96_odes(path, K):
    with open(path, "w") as f:
        f.write("\n".join(lorenz96_ode_lines(K)) + "\n")


def nvrtc_compile(source: str, arch: str = "sm_100a") -> bytes:
    """Compile CUDA source to a cubin for sm_100a with NVRTC (works without a GPU)."""
    from cuda.bindings import nvrtc
    err, prog = nvrtc.nvrtcCreateProgram(source.encode(), b"vgm_program.cu", 0, [], [])
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        raise RuntimeError("nvrtcCreateProgram failed: %s" % err)
    opts = [b"--gpu-architecture=" + arch.encode(), b"-lineinfo", b"--std=c++17"]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC compilation failed:\n" + log.decode(errors="replace"))
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * n
    nvrtc.nvrtcGetCUBIN(prog, cubin)
    nvrtc.nvrtcDestroyProgram(prog)
    return cubin


class Program:
    """Handle of a loaded coefficient program (``vgm_program*``)."""

    def __init__(self, model: OdeModel, prefer_builtin=True):
        lib = _lib.load()
        self.model = model
        h = C.c_void_p()
        if prefer_builtin and model.is_builtin_lorenz96():
            _lib.check(lib.vgm_program_builtin(b"lorenz96", C.byref(h)))
            self.kind = "builtin:lorenz96"
        else:
            cubin = model.compile_cubin()
            self._cubin = cubin
            _lib.check(lib.vgm_program_load(cubin, len(cubin), model.p, model.arity, C.byref(h)))
            self.kind = "nvrtc"
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().vgm_program_free(self.handle)
                self.handle = None
        except Exception:
            pass


# --------------------------------------------------------------------------------------
# Sweep plan (host side): pairs, dependency levels, live reads
# --------------------------------------------------------------------------------------
class HostPlan:
    """Index tables of one state sweep (see ``vgm_sweep_plan`` in include/vgm_b200.h)."""


def build_host_plan(model: OdeModel, hidden, couplings=None, schedule="reference"):
    """Build the sweep plan for ``hidden`` (state indices in the reference's update order,
    proxies...py:195).  ``couplings`` optionally overrides c(u) (list of lists, as the
    reference's ``state_couplings`` argument).  ``schedule``: 'reference' reproduces the
    sequential Gauss-Seidel order through dependency levels; 'colour' is the opt-in parallel
    multi-colour schedule (distance-1/2 conflicts never share a colour)."""
    hidden = np.asarray(list(hidden), dtype=np.int64)
    nh = len(hidden)
    K = model.K
    pos = np.full(K, -1, dtype=np.int64)
    pos[hidden] = np.arange(nh)
    uu, kk, code = model.pairs()
    if couplings is not None:
        K1 = model.K1
        if len(couplings) not in (K, K1):
            raise ValueError("state_couplings must have one list per state")
        if len(couplings) == K1 and K1 <= 4096:
            ok = np.zeros((K1, K1), dtype=bool)
            for u, lst in enumerate(couplings):
                for k in lst:
                    ok[u, k] = True
            keep = ok[uu % K1, kk % K1] & ((uu // K1) == (kk // K1))
            structural = np.zeros((K1, K1), dtype=bool)
            structural[uu % K1, kk % K1] = True
            missing = [(int(u), int(k)) for u, k in zip(*np.nonzero(ok & ~structural)) if pos[u] >= 0]
        else:
            allowed = set()
            for u, lst in enumerate(couplings):
                for k in lst:
                    allowed.add((u, k))
            keep = np.array([(int(u), int(k)) in allowed for u, k in zip(uu, kk)], dtype=bool)
            have = set(zip(uu[keep].tolist(), kk[keep].tolist()))
            missing = [pr for pr in allowed if pr not in have and pos[pr[0]] >= 0]
        if missing:
            raise ValueError("state_couplings lists (state, ODE) pairs where the state does not occur: %r" % missing[:4])
        uu, kk, code = uu[keep], kk[keep], code[keep]
    sel = pos[uu] >= 0
    uu, kk, code = uu[sel], kk[sel], code[sel]
    slot = pos[uu]
    order = np.lexsort((kk, slot))
    uu, kk, code, slot = uu[order], kk[order], code[order], slot[order]
    n_pairs = len(uu)
    # nonlinear check
    nonlin_codes = set()
    for ci, c in enumerate(model.classes):
        for j in range(c.arity):
            if c.nonlinear[j]:
                nonlin_codes.add(int(model.code_base[ci]) + j)
    if nonlin_codes:
        bad = np.isin(code, list(nonlin_codes))
        if bad.any():
            e = int(np.nonzero(bad)[0][0])
            raise ValueError("ODE %d is not linear in hidden state %s (no closed-form update); clamp it or remove the "
                             "pair from state_couplings" % (int(kk[e]), model.state_names[int(uu[e]) % model.K1]))
    pair_ptr = np.zeros(nh + 1, dtype=np.int64)
    np.add.at(pair_ptr, slot + 1, 1)
    pair_ptr = np.cumsum(pair_ptr)
    k_state = model.ode_state[kk].astype(np.int64)
    pair_self = (k_state == uu).astype(np.int32)
    has_self = np.zeros(nh, dtype=np.int32)
    has_self[slot[pair_self == 1]] = 1
    prod = pos[k_state]                       # producer slot of D.x_k (or -1 if that state is not updated)
    cand = (pair_self == 0) & (prod >= 0)
    # ---- levels ---------------------------------------------------------------------------
    level = np.zeros(nh, dtype=np.int64)
    if schedule == "reference":
        dep = cand & (prod < slot)
        if dep.any():
            e_idx = np.nonzero(dep)[0]
            # sequential pass in slot order (dependencies always point to smaller slots)
            by_slot_start = np.searchsorted(slot[e_idx], np.arange(nh), side="left")
            by_slot_end = np.searchsorted(slot[e_idx], np.arange(nh), side="right")
            pe = prod[e_idx]
            for i in np.unique(slot[e_idx]).tolist():
                a, b = by_slot_start[i], by_slot_end[i]
                level[i] = 1 + level[pe[a:b]].max()
        live_mask = dep
    elif schedule == "colour":
        level = colour_states(model, hidden, uu, kk, slot, prod, cand)
        live_mask = cand & (level[np.clip(prod, 0, None)] < level[slot])
    else:
        raise ValueError("unknown schedule %r" % schedule)
    n_levels = int(level.max()) + 1 if nh else 0
    lev_order = np.lexsort((np.arange(nh), level))
    level_ptr = np.zeros(n_levels + 1, dtype=np.int32)
    np.add.at(level_ptr, level + 1, 1)
    level_ptr = np.cumsum(level_ptr).astype(np.int32)
    # ---- live reads -------------------------------------------------------------------------
    e_live = np.nonzero(live_mask)[0]
    # order live pairs by (level of consumer, slot, k)
    if len(e_live):
        o = np.lexsort((kk[e_live], slot[e_live], level[slot[e_live]]))
        e_live = e_live[o]
    pair_live = np.full(n_pairs, -1, dtype=np.int32)
    pair_live[e_live] = np.arange(len(e_live), dtype=np.int32)
    prod_slots = np.unique(prod[e_live]) if len(e_live) else np.zeros(0, dtype=np.int64)
    if len(prod_slots):
        o = np.lexsort((prod_slots, level[prod_slots]))
        prod_slots = prod_slots[o]
    live_index = {int(s): i for i, s in enumerate(prod_slots.tolist())}
    live_state = hidden[prod_slots].astype(np.int32) if len(prod_slots) else np.zeros(0, dtype=np.int32)
    live_level_ptr = np.zeros(n_levels + 1, dtype=np.int32)
    if len(prod_slots):
        np.add.at(live_level_ptr, level[prod_slots] + 1, 1)
    live_level_ptr = np.cumsum(live_level_ptr).astype(np.int32)
    live_pair_slot = slot[e_live].astype(np.int32)
    live_pair_src = np.array([live_index[int(s)] for s in prod[e_live].tolist()], dtype=np.int32)
    live_pair_level_ptr = np.zeros(n_levels + 1, dtype=np.int32)
    if len(e_live):
        np.add.at(live_pair_level_ptr, level[slot[e_live]] + 1, 1)
    live_pair_level_ptr = np.cumsum(live_pair_level_ptr).astype(np.int32)
    # ---- constant R_uu? ------------------------------------------------------------------------
    ruu_const, ruu_value = 0, 0.0
    self_codes = np.unique(code[pair_self == 1])
    if len(self_codes):
        vals = []
        for cd in self_codes.tolist():
            ci = int(np.searchsorted(model.code_base, cd, side="right") - 1)
            j = cd - int(model.code_base[ci])
            e = model.classes[ci].R[j]
            vals.append(float(e) if len(e.free_symbols) == 0 else None)
        if all(v is not None for v in vals) and max(vals) == min(vals):
            ruu_const, ruu_value = 1, vals[0]
    hp = HostPlan()
    hp.n_hidden, hp.n_pairs = nh, n_pairs
    hp.slot_state = hidden.astype(np.int32)
    hp.pair_ptr = pair_ptr.astype(np.int32)
    hp.pair_ode = kk.astype(np.int32)
    hp.pair_code = code.astype(np.int32)
    hp.pair_live = pair_live
    hp.pair_self = pair_self
    hp.slot_has_self = has_self
    hp.n_levels = n_levels
    hp.level = level
    hp.level_ptr = level_ptr
    hp.level_slots = lev_order.astype(np.int32)
    hp.n_live = len(live_state)
    hp.live_state = live_state
    hp.live_level_ptr = live_level_ptr
    hp.n_live_pairs = len(e_live)
    hp.live_pair_slot = live_pair_slot
    hp.live_pair_src = live_pair_src
    hp.live_pair_level_ptr = live_pair_level_ptr
    hp.ruu_const, hp.ruu_value = ruu_const, ruu_value
    hp.n_early, hp.early_slots, hp.bulk_slots = tail_overlap_tables(nh, n_levels, level, level_ptr, prod_slots)
    return hp


TAIL_MAX_SYSTEMS = 8     # csrc/small_solve.cuh: SMALL_SOLVE_MAX_SYSTEMS


def tail_overlap_tables(nh, n_levels, level, level_ptr, prod_slots):
    """When everything after dependency level 0 is a handful of systems (Lorenz96 with every second state observed:
    the wrap-around state alone, which reads the updated x of the first hidden state, proxies...py:225), the sweep
    can solve that chain on a side stream while the bulk of level 0 runs: the level-0 systems the chain reads live
    (``early_slots``) are taken out of the bulk (``bulk_slots`` = the other level-0 slots) and solved first.
    Returns (n_early, early_slots, bulk_slots); n_early = 0 when the plan does not have that shape."""
    none = (0, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
    if n_levels < 2 or n_levels > TAIL_MAX_SYSTEMS:
        return none
    n_l0 = int(level_ptr[1])
    if nh - n_l0 > TAIL_MAX_SYSTEMS or n_l0 <= 4 * TAIL_MAX_SYSTEMS:
        return none
    prod_slots = np.asarray(prod_slots, dtype=np.int64)
    early = np.sort(prod_slots[level[prod_slots] == 0]) if len(prod_slots) else prod_slots
    if len(early) == 0 or len(early) > TAIL_MAX_SYSTEMS:
        return none
    lvl0 = np.nonzero(level == 0)[0]
    bulk = np.setdiff1d(lvl0, early)
    return len(early), early.astype(np.int32), bulk.astype(np.int32)


def colour_states(model, hidden, uu, kk, slot, prod, cand):
    """Greedy colouring of the hidden-state conflict graph (u ~ state of ODE k for every
    non-self pair, both hidden).  For Lorenz96 conflicts are index distances 1..2, giving 3
    colours when 3 | K (4 otherwise)."""
    nh = len(hidden)
    nbr = [[] for _ in range(nh)]
    for e in np.nonzero(cand)[0].tolist():
        a, b = int(slot[e]), int(prod[e])
        if a != b:
            nbr[a].append(b)
            nbr[b].append(a)
    colour = np.full(nh, -1, dtype=np.int64)
    for i in range(nh):
        used = {int(colour[j]) for j in nbr[i] if colour[j] >= 0}
        c = 0
        while c in used:
            c += 1
        colour[i] = c
    return colour
