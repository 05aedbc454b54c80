"""Host-side drop-in surface: same names / signatures as the reference's VGM_modules, model
parsing, structural classes, couplings and generated CUDA source (no GPU needed)."""
import inspect
import os

import numpy as np
import pytest
import sympy as sym

import variational_gradient_matching_for_dynamical_systems_b200 as pkg
from variational_gradient_matching_for_dynamical_systems_b200 import import_odes as io_
from variational_gradient_matching_for_dynamical_systems_b200 import rewrite_odes_as_local_linear_combinations as rw
from oracle import reference_shim as rs
from oracle import vgm_oracle as vo


class Sy:
    def __init__(self, state, param):
        self.state, self.param = sym.symbols(state), sym.symbols(param)


def _write(tmp_path, lines):
    p = os.path.join(str(tmp_path), "m_ODEs.txt")
    open(p, "w").write("\n".join(lines) + "\n")
    return p


@pytest.mark.skipif(not rs.reference_available(), reason="live reference not present")
def test_signatures_match_reference():
    ref = rs.load_reference()
    import importlib
    GP = importlib.import_module(pkg.__name__ + ".GP_regression")
    PX = importlib.import_module(pkg.__name__ + ".proxies_for_ode_parameters_and_states")
    pairs = [(GP.kernel_function, ref.GP_regression.kernel_function),
             (GP.fitting_state_observations, ref.GP_regression.fitting_state_observations),
             (PX.proxy_for_ode_parameters, ref.proxies_strict.proxy_for_ode_parameters),
             (PX.proxy_for_ind_states, ref.proxies_strict.proxy_for_ind_states),
             (io_.import_odes, ref.import_odes.import_odes),
             (io_.gradient_of_odes, ref.import_odes.gradient_of_odes),
             (io_.find_state_couplings_in_odes, ref.import_odes.find_state_couplings_in_odes),
             (rw.rewrite_odes_as_linear_combination_in_states, ref.rewrite.rewrite_odes_as_linear_combination_in_states)]
    for mine, theirs in pairs:
        a, b = inspect.signature(mine), inspect.signature(theirs)
        assert list(a.parameters) == list(b.parameters), mine.__name__
        for n in a.parameters:
            da, db = a.parameters[n].default, b.parameters[n].default
            assert (da is inspect._empty) == (db is inspect._empty) and (da is inspect._empty or da == db), (mine.__name__, n)
    # the parameter rewrite only adds a trailing keyword (zero_fill)
    a = list(inspect.signature(rw.rewrite_odes_as_linear_combination_in_parameters).parameters)
    b = list(inspect.signature(ref.rewrite.rewrite_odes_as_linear_combination_in_parameters).parameters)
    assert a[:len(b)] == b


def test_import_odes_and_couplings_lorenz96(tmp_path):
    K = 10
    s = Sy(["_x_%d" % (i + 1) for i in range(K)], ["_alpha"])
    odes, odes_sym = io_.import_odes(s, _write(tmp_path, vo.lorenz96_ode_lines(K)))
    assert len(odes_sym) == K and len(odes.model.classes) == 1
    sc, oc = io_.find_state_couplings_in_odes(odes_sym, s)
    assert sc == vo.state_couplings(vo.parse_ode_lines(vo.lorenz96_ode_lines(K)), s.state)
    assert [str(v) for v in oc[0]] == ["_x_1", "_x_2", "_x_9", "_x_10"]
    # symbolic call convention used by the rewrites, numeric call convention used by simulators
    ex = odes(*s.state, *s.param)
    assert sym.expand(ex[1] - sym.sympify(vo.lorenz96_ode_lines(K)[1])) == 0
    x = np.arange(1.0, K + 1)
    assert np.allclose(odes(*x, 8.0), vo.lorenz96_rhs(x, 8.0))


@pytest.mark.skipif(not rs.reference_available(), reason="live reference not present")
def test_rewrites_match_reference_callables(tmp_path):
    """B, b, R, r evaluated through the reference's calling convention agree with the reference's
    own lambdified rewrites (Lotka-Volterra: exercises the zero -> ones quirk)."""
    ref = rs.load_reference()
    lines = ["_theta_1 * _x_1 - _theta_2 * _x_1 * _x_2", "-_theta_3 * _x_2 + _theta_4 * _x_1 * _x_2"]
    s = Sy(["_x_1", "_x_2"], ["_theta_1", "_theta_2", "_theta_3", "_theta_4"])
    path = _write(tmp_path, lines)
    odes, odes_sym = io_.import_odes(s, path)
    r_odes, r_sym = ref.import_odes.import_odes(s, path)
    sc, _ = io_.find_state_couplings_in_odes(odes_sym, s)
    r_sc, _ = ref.import_odes.find_state_couplings_in_odes(r_sym, s)
    assert sc == r_sc
    B, b = rw.rewrite_odes_as_linear_combination_in_parameters(odes, s.state, s.param)
    rB, rb = ref.rewrite.rewrite_odes_as_linear_combination_in_parameters(r_odes, s.state, s.param)
    R, r = rw.rewrite_odes_as_linear_combination_in_states(odes, s.state, s.param, [], sc)
    rR, rr = ref.rewrite.rewrite_odes_as_linear_combination_in_states(r_odes, s.state, s.param, [], r_sc)
    rng = np.random.default_rng(0)
    T = 7
    X = rng.standard_normal((T, 2))
    state = np.append(X, np.ones((T, 1)), axis=1).T
    th = [2.0, 1.0, 4.0, 1.0]
    for k in range(2):
        assert np.allclose(np.squeeze(B[k](*state)), np.squeeze(rB[k](*state)))
        assert np.allclose(np.squeeze(b[k](*state)), np.squeeze(rb[k](*state)))
        for u in range(2):
            if k in r_sc[u]:
                assert np.allclose(np.squeeze(R[u][k](*th, *state)), np.squeeze(rR[u][k](*th, *state)) * np.ones(T))
                assert np.allclose(np.squeeze(r[u][k](*th, *state)), np.squeeze(rr[u][k](*th, *state)) * np.ones(T))
            else:
                assert R[u][k] == [] and rR[u][k] == []


def test_structural_classes_scale_and_codegen():
    m = pkg.OdeModel.lorenz96(5000)
    assert len(m.classes) == 1 and m.arity == 4 and m.is_builtin_lorenz96()
    src = m.cuda_source()
    assert "vgm_theta_coeffs" in src and "vgm_prog_##name" in src
    pt = ["- _k_1 * _S - _k_2 * _S * _R + _k_3 * _RS", "_k_1 * _S",
          "- _k_2 * _S * _R + _k_3 * _RS + _V * _R_pp /(0.3 + _R_pp)",
          "_k_2 * _S * _R -(_k_3 + _k_4) * _RS", "_k_4 * _RS - _V * _R_pp / (0.3 + _R_pp)"]
    m = pkg.OdeModel(pt, ["_S", "_dS", "_R", "_RS", "_R_pp"], ["_k_1", "_k_2", "_k_3", "_k_4", "_V"], n_copies=3)
    assert m.K == 15 and m.ode_idx.max() == 14
    assert len(pkg.nvrtc_compile(m.cuda_source())) > 1000          # NVRTC targets sm_100a without a GPU
    # hidden R_pp is nonlinear in its own ODE: no closed-form update
    with pytest.raises(ValueError, match="not linear"):
        pkg.build_host_plan(m, [4])
    hp = pkg.build_host_plan(m, [0, 3, 5, 8, 10, 13])
    assert hp.n_levels >= 1 and hp.ruu_const == 0


def test_parameter_nonlinearity_fails_only_where_the_reference_fails():
    """A model that is not linear in a parameter imports and has couplings and a state rewrite (as in the reference);
    only the parameter rewrite raises (rewrite...py:45, sympy NonlinearError)."""
    m = pkg.OdeModel(["_V * _x / (_K_m + _y)", "-_y"], ["_x", "_y"], ["_V", "_K_m"])
    assert m.param_nonlinear and m.state_couplings() == [[0], [0, 1]]
    assert len(pkg.nvrtc_compile(m.cuda_source())) > 1000
    xs, ps = [sym.Symbol("_x"), sym.Symbol("_y")], [sym.Symbol("_V"), sym.Symbol("_K_m")]

    class _O:
        model = m
    with pytest.raises(Exception, match="(?i)nonlinear"):
        rw.rewrite_odes_as_linear_combination_in_parameters(_O(), xs, ps)
    R, r = rw.rewrite_odes_as_linear_combination_in_states(_O(), xs, ps, [], [[0], [1]])    # y enters ODE 0 nonlinearly
    assert R[0][0].expr[0][0] == sym.sympify("_V/(_K_m + _y)") and r[0][0].expr is not None
    assert not pkg.OdeModel(["_a * _x"], ["_x"], ["_a"]).param_nonlinear


REF_ROOT = "/root/reference"
REF_MODELS = {
    "Lotka_Volterra_ODEs.txt": (["_x_1", "_x_2"], ["_theta_1", "_theta_2", "_theta_3", "_theta_4"]),
    "Lorenz_attractor_ODEs.txt": (["_x", "_y", "_z"], ["_sigma", "_rho", "_alpha"]),
    "Lorenz96_ODEs.txt": (["_x_%d" % (i + 1) for i in range(10)], ["_alpha"]),
    "Protein_Transduction_ODEs.txt": (["_S", "_dS", "_R", "_RS", "_R_pp"], ["_k_1", "_k_2", "_k_3", "_k_4", "_V", "_K_m"]),
    "FitzHugh_Nagumo_ODEs.txt": (["_v", "_w"], ["_I_ext", "_a", "_b", "_tau"]),
}


@pytest.mark.skipif(not os.path.isdir(REF_ROOT), reason="reference tree not present")
@pytest.mark.parametrize("fname", sorted(REF_MODELS))
def test_every_shipped_model_file_imports(fname):
    """import_odes + find_state_couplings_in_odes accept every model file the reference ships whose symbols its
    declarations name, including the two that are not linear in a parameter (drop-in behaviour: the reference
    imports them and fails only in the parameter rewrite)."""
    io = io_
    path = os.path.join(REF_ROOT, fname)
    txt = open(path).read()
    states, params = REF_MODELS[fname]
    n_lines = len([ln for ln in txt.splitlines() if ln.strip()])
    states = states[:n_lines]

    class S:
        state = [sym.Symbol(n) for n in states]
        param = [sym.Symbol(n) for n in params]
    odes, odes_sym = io.import_odes(S, path)
    sc, _ = io.find_state_couplings_in_odes(odes_sym, S)
    assert len(sc) == len(states) and len(odes_sym) == len(states)
    free = set().union(*[e.free_symbols for e in odes_sym])
    assert free <= set(S.state) | set(S.param), free - set(S.state) - set(S.param)
    nonlin = fname in ("Protein_Transduction_ODEs.txt", "FitzHugh_Nagumo_ODEs.txt")
    assert odes.model.param_nonlinear == nonlin


def test_identifier_scan_leaves_numeric_literals_alone():
    m = pkg.OdeModel(["1e2 * _a * e2 + 2.5e-1", "-e2"], ["x", "e2"], ["_a"])
    assert m.state_couplings() == [[], [0, 1]]


def test_kernel_function_errors_match_reference():
    from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function
    with pytest.raises(ValueError, match="not a numpy array"):
        kernel_function([0.0, 1.0], [0.0], 'rbf', [1, 1])
    with pytest.raises(ValueError, match="one dimensional"):
        kernel_function(np.zeros((2, 2)), [0.0], 'rbf', [1, 1])
