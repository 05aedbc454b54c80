#!/usr/bin/env python
"""Benchmark of the VGM coordinate-ascent loop (BASELINE.json metric:
"VGM coordinate-ascent state-time updates/sec (Lorenz96 K=100k, T=1000)").

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's algorithm

A *step* is one full coordinate-ascent iteration -- the ODE-parameter update
(proxy_for_ode_parameters, optimizer='analytical') followed by the sweep over all hidden states
(proxy_for_ind_states, optimizer='analytical') -- on synthetic Lorenz96 data of the named shape,
every step starting from the same initial state proxies (so every step does the same,
cold-start, amount of work).  value = K_hidden * T * steps / seconds, whole job.
N > 1 shards the K states into contiguous blocks (strong scaling, NCCL halo exchange + all-reduce).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "VGM coordinate-ascent state-time updates/sec (Lorenz96 K=100k, T=1000)"
UNIT = "state-time updates/s"
KERNEL = {"type": "rbf", "param": [0.0625, 1.0]}      # l = 1.25*dt: cond(C) ~ 1e3
DT = 0.05


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--K", type=int, default=100000)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--cpu-sample-K", type=int, default=384, help="states in the bounded CPU-baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precondition", action="store_true")
    ap.add_argument("--solver", default="auto", choices=["auto", "mixed", "pcg"])
    ap.add_argument("--inner-iters", type=int, default=10)
    ap.add_argument("--tol", type=float, default=1e-13)
    ap.add_argument("--optimistic-passes", type=int, default=4)
    ap.add_argument("--pipeline-chunks", default="0.12,0.5,0.88,1.0",
                    help="e2e: state ranges whose copies overlap the compute: a count (equal ranges) or the upper ends "
                         "of the ranges as fractions; a short first range starts the compute early, a short last "
                         "one shortens the final copy-out; 0 = no pipelining")
    ap.add_argument("--no-profile", action="store_true", help="skip the second, per-kernel-instrumented timed region")
    ap.add_argument("--no-colour", action="store_true", help="skip the opt-in multi-colour schedule block (N=1 only)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 trajectory-batch block")
    ap.add_argument("--c4-trajectories", type=int, default=65536)
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of sampled hidden columns + theta")
    ap.add_argument("--parity-columns", type=int, default=64)
    args = ap.parse_args()
    pc = str(args.pipeline_chunks)
    args.pipeline_chunks = tuple(float(v) for v in pc.split(",")) if "," in pc else int(pc)
    return args


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's algorithm (the Python reference itself cannot
# travel to the GPU box; oracle/vgm_oracle.py is pinned against it by tests/golden)
# ----------------------------------------------------------------------------------------------
def cpu_port_run(K_sample, T, steps, warmup):
    from oracle import vgm_oracle as vo
    try:
        from threadpoolctl import threadpool_info
        blas = [{k: d.get(k) for k in ("internal_api", "num_threads", "version")} for d in threadpool_info()]
    except Exception:  # noqa: BLE001
        blas = None
    t, X = vo.lorenz96_trajectory(K_sample, T, DT, seed=0)
    rng = np.random.default_rng(1)
    X0 = X.copy()
    X0[:, 1::2] += 0.5 * rng.standard_normal((T, K_sample // 2))
    D, A, *_ = vo.kernel_function(t, t, KERNEL["type"], KERNEL["param"])
    hidden = list(range(1, K_sample, 2))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        vo.lorenz96_theta_step(X0, D)
        vo.lorenz96_state_sweep(X0, 8.0, D, hidden)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = len(hidden) * T * len(times) / total
    return value, total, blas


def live_reference_record():
    """The LIVE reference timed in the build container (scripts/time_live_reference.py -> profiles/
    cpu_live_reference.json): it cannot travel to the GPU box, so its numbers are quoted, labelled, next to the port
    that is timed here."""
    try:
        with open(os.path.join(ROOT, "profiles", "cpu_live_reference.json")) as f:
            d = json.load(f)
        return {"source": "profiles/cpu_live_reference.json (scripts/time_live_reference.py, build container)",
                "cores": d["cores"], "rows": [{k: r[k] for k in ("K", "hidden", "ms_per_hidden_state",
                                                                   "state_time_updates_per_s")} for r in d["rows"]],
                "mean_ms_per_hidden_state": d["mean_ms_per_hidden_state"], "extrapolated": d["extrapolated"]}
    except Exception:  # noqa: BLE001
        return None


def workload_config(K, T, n_gpus):
    """`config` of the JSON line: the SAME dict in the native arm and in the reference (CPU) arm -- it names the
    workload; what is specific to one arm's run goes under `details` (native) or `cpu_baseline.sample` (CPU)."""
    K_h = K // 2
    ld = (T + 127) // 128 * 128 if T >= 512 else (T + 7) // 8 * 8
    return {"workload": "Lorenz96 K=%d states (%d hidden, every second state observed+clamped), T=%d; one step = "
                        "analytical theta update + full state sweep from the same initial proxies" % (K, K_h, T),
            "kernel": KERNEL, "dt": DT, "parallelism": "state-block x%d" % n_gpus,
            "l2": "inputs larger than L2: each state-shaped FP64 array of the solve (X, P, Q, R, Z) is %.0f MB per GPU "
                  "(126 MB of L2)" % (K_h / n_gpus * ld * 8 / 1e6)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count()
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    Ks = args.cpu_sample_K
    value, total, blas = cpu_port_run(Ks, args.T, steps, warmup)
    sample = ("Lorenz96 K=%d (%d hidden), T=%d: %d iteration(s) of the oracle port of the reference's analytical "
              "theta update + state sweep (dense LAPACK solve per hidden state); per-state cost is flat in K "
              "(BASELINE.md section 3)" % (Ks, Ks // 2, args.T, steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.K, args.T, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "blas": blas, "live_reference": live_reference_record(),
                             "note": "this arm times the ORACLE PORT of the reference's algorithm (one dense LAPACK solve "
                                     "per hidden state, without the reference's 12 dense T^3 GEMMs per state), not the "
                                     "reference itself, on a bounded sample of the workload"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# Parity of the measured configuration against the CPU oracle ("restatement on 64 random
# hidden columns + theta").  The oracle is the checker only: it never produces anything that is timed or returned.
# ----------------------------------------------------------------------------------------------
PARITY_TOL = 1e-9          # BASELINE.json north_star: 1e-9 relative in the ODE-parameter and state posterior means


def parity_block(args, eng, core, X0_local, world, rank, K, T, hidden, n_columns):
    """One step from the initial proxies, then: theta against oracle.lorenz96_theta_step on the full X0 and
    `n_columns` seeded random hidden columns (always the first hidden state and the wrap-around state K-1, which reads
    the first one live) against oracle.lorenz96_sweep_columns -- the statements of
    proxies_for_ode_parameters_and_states.py:58-93,195-262 with the oracle's OWN kernel_function operators.
    Every rank calls this (the rows are gathered from their owners); rank 0 returns the dict."""
    import torch
    import torch.distributed as dist
    from oracle import vgm_oracle as vo
    t_par = time.perf_counter()
    hidden = np.asarray(hidden, dtype=np.int64)
    rng = np.random.default_rng(20261020)
    forced = [int(hidden[0]), int(hidden[-1])]
    pool = np.setdiff1d(hidden, forced)
    extra = rng.choice(pool, size=max(0, min(n_columns, len(hidden)) - len(forced)), replace=False)
    cols = sorted(set(forced + extra.tolist()))
    need_s, need_l = vo.lorenz96_column_needs(cols, K, hidden)
    core.X.copy_(X0_local)
    eng.iteration()
    theta_gpu = float(core.get_theta()[0])
    res_states = cols + [k for k in need_l if k not in cols]
    if world > 1:
        snap_rows = eng.gather_rows(need_s, source=X0_local)[:, :T]
        res_rows = eng.gather_rows(res_states)[:, :T]
        keep = core.X.clone()
        core.X.copy_(X0_local)
        X0_full = eng.gather_states()                      # [K, T], for the theta check on rank 0
        core.X.copy_(keep)
    else:
        dev = core.X.device
        snap_rows = X0_local.index_select(0, torch.as_tensor(need_s, device=dev))[:, :T]
        res_rows = core.X.index_select(0, torch.as_tensor(res_states, device=dev))[:, :T]
        X0_full = X0_local[:, :T]
    out = None
    if rank == 0:
        snap = {k: snap_rows[i].cpu().numpy() for i, k in enumerate(need_s)}
        res = {k: res_rows[i].cpu().numpy() for i, k in enumerate(res_states)}
        tgrid = np.arange(T) * DT
        D, *_ = vo.kernel_function(tgrid, tgrid[:2], KERNEL["type"], KERNEL["param"])
        ref = vo.lorenz96_sweep_columns(snap, {k: res[k] for k in need_l}, K, 8.0, D, cols, hidden)
        rel, absmax, elem = 0.0, 0.0, 0.0
        worst = None
        for u in cols:
            d = np.abs(res[u] - ref[u])
            r = float(d.max() / max(np.abs(ref[u]).max(), 1e-300))
            if r > rel:
                rel, worst = r, u
            absmax = max(absmax, float(d.max()))
            # element-wise relative error with an absolute floor of 1e-3 (the states are O(1..10))
            elem = max(elem, float((d / (np.abs(ref[u]) + 1e-3)).max()))
        Xh = X0_full.t().contiguous().cpu().numpy()        # [T, K] reference orientation
        theta_ref = float(vo.lorenz96_theta_step(Xh, D)[0])
        del Xh
        rel_theta = abs(theta_gpu - theta_ref) / abs(theta_ref)
        ok = bool(rel < PARITY_TOL and rel_theta < PARITY_TOL and np.isfinite(rel) and np.isfinite(rel_theta))
        out = {"config": "C5" if (K == 100000 and T == 1000) else "Lorenz96 K=%d T=%d" % (K, T),
               "columns": len(cols), "includes_wrap_around_state": True, "live_columns": [int(k) for k in need_l],
               "max_rel_states": rel, "worst_column": worst, "max_abs_states": absmax,
               "max_rel_states_elementwise_floor1e-3": elem, "rel_theta": rel_theta, "theta_oracle": theta_ref,
               "tol": PARITY_TOL, "ok": ok, "n_gpus": world,
               "oracle": "oracle/vgm_oracle.py: kernel_function + lorenz96_theta_step (full X) + lorenz96_sweep_columns "
                         "(dense LAPACK solve per column, proxies...py:195-262)",
               "seconds": round(time.perf_counter() - t_par, 2)}
    if world > 1:
        flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device=core.X.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return out


# ----------------------------------------------------------------------------------------------
# Opt-in multi-colour schedule (schedule='colour'): NOT the reference's update order, so no 1e-9 parity is claimed for
# it.  Its stated tolerance is on the converged hidden-state RMSE against the ground-truth trajectory: within 5 % of
# the RMSE the reference order reaches after the same number of iterations.
# ----------------------------------------------------------------------------------------------
COLOUR_RMSE_TOL = 0.05


def colour_block(args, model, ops, hidden, ref_engine, X0_local, K, T, dev, n_conv=8):
    import torch
    from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine
    from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_trajectory
    t0 = time.perf_counter()
    col = VGMEngine(model, ops, None, hidden=hidden, schedule="colour", solver=args.solver, inner_iters=args.inner_iters,
                    tol=args.tol, optimistic_passes=args.optimistic_passes)
    _, Xtrue = lorenz96_trajectory(K, T, DT, seed=0, device=dev)          # the trajectory lorenz96_problem perturbs
    hid = torch.as_tensor(np.asarray(hidden, dtype=np.int64), device=dev)
    truth = Xtrue.index_select(0, hid)[:, :T]
    del Xtrue
    rmse = lambda X: float(((X.index_select(0, hid)[:, :T] - truth) ** 2).mean().sqrt().item())

    def step():
        col.X.copy_(X0_local)
        return col.iteration()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_it = max(3, args.steps)
    e0.record()
    for _ in range(n_it):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_it
    # converged RMSE of both schedules from the same initial proxies
    r0 = rmse(X0_local)
    col.X.copy_(X0_local)
    ref_engine.X.copy_(X0_local)
    for _ in range(n_conv):
        col.iteration()
        ref_engine.iteration()
    r_col, r_ref = rmse(col.X), rmse(ref_engine.X)
    ref_engine.X.copy_(X0_local)
    out = {"schedule": "colour (opt-in; hidden states that read each other's D.x never share a colour)",
           "levels": int(col.plan.host.n_levels), "level_sizes": np.diff(col.plan.host.level_ptr).tolist(),
           "ms_per_step": ms, "state_time_updates_per_s": len(hidden) * T / (ms * 1e-3),
           "iterations_to_converge": n_conv, "rmse_initial": r0, "rmse_colour": r_col, "rmse_reference_order": r_ref,
           "rmse_vs_reference_order": abs(r_col - r_ref) / r_ref, "tol": COLOUR_RMSE_TOL,
           "ok": bool(r_col < r0 and abs(r_col - r_ref) <= COLOUR_RMSE_TOL * r_ref + 1e-3),
           "seconds": round(time.perf_counter() - t0, 2)}
    del col
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
# Config C4 (BASELINE.json configs[3]): 65 536 independent Protein-Transduction trajectories sharing one theta,
# trajectories split across the ranks, ONE all-reduce of the pooled parameter statistics per iteration.
# Model as in SURVEY.md 8d: K_m -> 0.3 so that the system is linear in the remaining 5 parameters; states S and RS are
# hidden, dS, R, R_pp observed and clamped; T = 99.  Inputs: the golden single trajectory (generated from the live
# reference set-up, tests/golden/protein_transduction_T99.npz) scaled per trajectory and state by 1 + 0.05 N(0,1).
# ----------------------------------------------------------------------------------------------
def c4_block(args, world, rank, dev):
    import sympy as sym
    import torch
    import torch.distributed as dist
    from oracle import vgm_oracle as vo
    from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel
    from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine
    t_all = time.perf_counter()
    g = np.load(os.path.join(ROOT, "tests", "golden", "protein_transduction_T99.npz"))
    lines, params, names = [str(x) for x in g["lines"]], [str(x) for x in g["params"]], [str(x) for x in g["names"]]
    sc = [[int(v) for v in str(c).split(",") if v != ""] for c in g["couplings"]]
    observed = [str(x) for x in g["observed"]]
    K1, T, N = len(names), g["X0"].shape[0], args.c4_trajectories
    hidden1 = [u for u in range(K1) if names[u] not in observed]
    D, A = g["D"], g["A"]
    rng = np.random.default_rng(7)
    scale = 1.0 + 0.05 * rng.standard_normal((N, 1, K1))
    Xn = g["X0"][None, :, :] * scale                                   # [N, T, K1]
    X_KT = np.ascontiguousarray(Xn.transpose(0, 2, 1).reshape(N * K1, T))   # state-major, trajectory n = rows n*K1..
    model = OdeModel(lines, names, params, n_copies=N)
    hidden = np.array([n * K1 + u for n in range(N) for u in hidden1])
    t0 = time.perf_counter()
    if world > 1:
        from variational_gradient_matching_for_dynamical_systems_b200.distributed import DistributedVGMEngine
        eng = DistributedVGMEngine(model, D, A, hidden=hidden, align=K1, theta_mode="proxy")
        core = eng.engine
    else:
        eng = core = VGMEngine(model, D, A, hidden=hidden, couplings=sc, theta_mode="proxy")
    eng.set_states_state_major(torch.as_tensor(X_KT))
    X0_local = core.X.clone()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0

    def step():
        core.X.copy_(X0_local)
        return eng.iteration()

    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n_it = max(3, args.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_it):
        info = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_it
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    # ---- parity: pooled statistics / theta vs a vectorised NumPy restatement of proxies...py:58-93 (trajectories are
    # just more rows of the sums), and a few trajectories' sweeps vs oracle.state_sweep (proxies...py:195-262)
    theta_gpu = core.get_theta()
    stats_gpu = core.stats.cpu().numpy()
    pick = sorted(rng.choice(N, size=min(8, N), replace=False).tolist())
    rows = np.array([n * K1 + u for n in pick for u in range(K1)])
    if world > 1:
        got = eng.gather_rows(rows)[:, :T].cpu().numpy()
    else:
        got = core.X.index_select(0, torch.as_tensor(rows, device=dev))[:, :T].cpu().numpy()
    out = None
    if rank == 0:
        lin = vo.Linearisation(vo.parse_ode_lines(lines), sym.symbols(names), sym.symbols(params), couplings=sc)
        S_big = Xn.reshape(N * T, K1)
        DX_big = np.einsum("st,ntk->nsk", D, Xn).reshape(N * T, K1)
        m, Sc = np.zeros(lin.p), np.zeros((lin.p, lin.p))
        for k in range(K1):
            B, b = lin.eval_B_b(k, S_big)
            m += B.T.dot(DX_big[:, k] - b)
            Sc += B.T.dot(B)
        theta_ref = np.linalg.solve(Sc, m)
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300))
        worst = 0.0
        for i, n in enumerate(pick):
            Xr, _ = vo.state_sweep(Xn[n], theta_ref, lin, D, hidden1)          # [T, K1]
            worst = max(worst, rel(got[i * K1:(i + 1) * K1].T, Xr))
        par = {"rel_local_mean": rel(stats_gpu[:lin.p], m), "rel_local_scaling": rel(stats_gpu[lin.p:].reshape(lin.p, lin.p), Sc),
               "rel_theta": rel(theta_gpu, theta_ref), "max_rel_states_sampled": worst, "sampled_trajectories": len(pick),
               "tol": PARITY_TOL}
        par["ok"] = bool(max(par["rel_local_mean"], par["rel_local_scaling"], par["rel_theta"], worst) < PARITY_TOL)
        out = {"workload": "C4: %d Protein-Transduction trajectories (K_m -> 0.3: 5 states, 5 shared parameters), T=%d, "
                           "hidden S and RS; trajectories split over %d GPU(s), one all-reduce of p + p^2 = %d "
                           "statistics per iteration" % (N, T, world, lin.p + lin.p ** 2),
               "n_trajectories": N, "n_gpus": world, "hidden_systems": int(len(hidden)), "iterations_timed": n_it,
               "ms_per_iteration": ms, "trajectories_per_s": N / (ms * 1e-3),
               "state_time_updates_per_s": len(hidden) * T / (ms * 1e-3), "solver": core.solver,
               "solver_info": dict(info) if isinstance(info, dict) else None, "setup_s": round(setup_s, 2),
               "parity": par, "seconds": round(time.perf_counter() - t_all, 2)}
    del eng, core, X0_local
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for ln in out.strip().splitlines():
                    f = [x.strip() for x in ln.split(",")]
                    if len(f) >= 9:
                        self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)
        sm = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for n, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned host buffer
    is allocated: first-touch then places the e2e staging buffers on the GPU's own NUMA node.  torchrun does not do
    this; without it half of the ranks copy across the socket interconnect.  Returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (%d)" % (min(cpus), max(cpus), len(cpus))
    except Exception:  # noqa: BLE001
        pass
    return None


def native_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel, _lib
    from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device
    from variational_gradient_matching_for_dynamical_systems_b200.engine import Operators, VGMEngine
    from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    full_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    K, T = args.K, args.T

    # ---- one-time set-up (excluded from the metric, reported separately) ------------------------
    t_setup0 = time.perf_counter()
    tgrid = np.arange(T) * DT
    kf = kernel_function_device(tgrid, tgrid, KERNEL["type"], KERNEL["param"])
    ops = Operators(kf["D"][:, :T], None, device=dev)
    _, X0, hidden = lorenz96_problem(K, T, DT, seed=0, device=dev)        # [K, T] on the device
    model = OdeModel.lorenz96(K)
    kw = dict(precondition=not args.no_precondition, solver=args.solver, inner_iters=args.inner_iters, tol=args.tol,
              optimistic_passes=args.optimistic_passes)
    if world > 1:
        from variational_gradient_matching_for_dynamical_systems_b200.distributed import DistributedVGMEngine
        # VGM_COMM=native: both exchanges through the library's vgm_comm_* entry points instead of torch.distributed
        eng = DistributedVGMEngine(model, ops, None, hidden=hidden, comm=os.environ.get("VGM_COMM", "torch"), **kw)
        core = eng.engine
        eng.set_states_state_major(X0)
        X0_local = core.X.clone()
        n_hidden_local = core.plan.n_hidden
    else:
        eng = core = VGMEngine(model, ops, None, hidden=hidden, **kw)
        core.set_states_state_major(X0)
        X0_local = core.X.clone()
        n_hidden_local = core.plan.n_hidden
    del X0
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup0
    n_hidden_total = len(hidden)

    def barrier():
        if world > 1:
            dist.barrier()

    def step():
        core.X.copy_(X0_local)
        return eng.iteration()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device-resident ------------------------------------------------------------
    def timed_region(profile):
        """K steps between barriers, CUDA events on the launching stream; max over ranks."""
        lib.vgm_profile_enable(1 if profile else 0)
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_launch, its = 0, []
        e0.record()
        for _ in range(args.steps):
            info = step()
            n_launch += info["kernel_launches"] + 1
            its.append(info["iterations"])
        e1.record()
        torch.cuda.synchronize()
        barrier()
        t = e0.elapsed_time(e1)
        if world > 1:
            tmax = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            t = float(tmax.item())
        return t, n_launch, its

    sampler = ClockSampler(local_rank)
    sampler.start()
    # region 1: the metric (no per-kernel instrumentation)
    ms, launches, pcg_iters = timed_region(False)
    clocks = sampler.stop()
    # region 2: the same K steps with every launch of the kernel families bracketed by CUDA events on its
    # stream (two event records per launch cost a few % of the step, more on small per-GPU problems, which is
    # why the metric above is taken without them)
    fam = {}
    ms_prof = None
    if not args.no_profile:
        ms_prof, _, _ = timed_region(True)
        for fid, name in enumerate(("dgemm", "tcgemm", "vec_other", "assemble", "ir_update", "ir_dir")):
            fl, by, fms, fn = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
            lib.vgm_profile_collect(fid, C.byref(fl), C.byref(by), C.byref(fms), C.byref(fn))
            fam[name] = {"flops": fl.value, "bytes": by.value, "ms": fms.value, "launches": fn.value}
        lib.vgm_profile_enable(0)
    value = n_hidden_total * T * args.steps / (ms * 1e-3)
    theta = core.get_theta().tolist()
    # context, not the metric: the loop as a user runs it (each iteration starts from the previous proxies, so the
    # solves are warm-started and later iterations need fewer corrections)
    evolving = []
    core.X.copy_(X0_local)
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = eng.iteration()
        torch.cuda.synchronize()
        evolving.append({"ms": round((time.perf_counter() - t0) * 1e3, 2), "solver_iterations": info["iterations"]})

    # ---- end-to-end: host buffers in, host buffers out ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        ld = core.ld
        Xh_in = torch.empty((core.K, ld), dtype=torch.float64, pin_memory=True)
        Xh_in.copy_(X0_local)
        Xh_out = torch.empty((core.K, ld), dtype=torch.float64, pin_memory=True)
        th_h = torch.zeros(max(core.p, 1), dtype=torch.float64, pin_memory=True)
        Xh_out.copy_(X0_local)          # the clamped (observed) rows never change; only updated rows travel back
        h2d = Xh_in.numel() * 8
        d2h = core.plan.n_hidden * ld * 8 + core.p * 8
        if world > 1:
            # only the rows this rank updated travel back (as at N=1): gathered on the device, one contiguous copy
            hid_idx = torch.as_tensor(np.asarray(core.plan.hidden, dtype=np.int64), device=dev)
            stage = torch.empty((len(hid_idx), ld), dtype=torch.float64, device=dev)
            Xh_hid = torch.empty((len(hid_idx), ld), dtype=torch.float64, pin_memory=True)

        def e2e_step():
            if world > 1:
                core.X.copy_(Xh_in, non_blocking=True)
                eng.iteration()
                torch.index_select(core.X, 0, hid_idx, out=stage)
                Xh_hid.copy_(stage, non_blocking=True)
                th_h.copy_(core.theta, non_blocking=True)
                torch.cuda.synchronize()
            else:
                core.iteration_host(Xh_in, th_h, Xh_out, updated_rows_only=True, pipeline_chunks=args.pipeline_chunks)

        e2e_step()
        # what the host link gives this rank while every rank copies at once (context for the e2e figure)
        barrier()
        torch.cuda.synchronize()
        tc0 = time.perf_counter()
        for _ in range(3):
            core.X.copy_(Xh_in, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 3 * h2d / (time.perf_counter() - tc0) / 1e9
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tmax = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        e2e = {"value": n_hidden_total * T * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt / args.steps,
               "h2d_gbs_rank0_all_ranks_copying": h2d_gbs, "cpu_binding": numa,
               "api": ("vgm_iteration_host (C ABI, pinned host buffers; %d state ranges pipelined: copy-in, "
                       "solve and copy-out overlap)" % core.last_pipeline) if world == 1 else
                      "DistributedVGMEngine.iteration with pinned host copies per rank (all local rows in, updated rows out)"}

    os.sched_setaffinity(0, full_affinity)       # the CPU checks / baseline below use every host core again

    # ---- parity of this very configuration against the oracle (every rank takes part in the gathers) ------
    parity = None
    if not args.no_parity:
        parity = parity_block(args, eng, core, X0_local, world, rank, K, T, hidden, args.parity_columns)

    # ---- opt-in multi-colour schedule, reported separately with its own tolerance (north_star) ---------------
    colour = None
    if world == 1 and not args.no_colour:
        colour = colour_block(args, model, ops, hidden, core, X0_local, K, T, dev)

    # ---- config C4 under the same launch (trajectory batch + all-reduce of the pooled statistics) ------------
    c4 = None
    if not args.no_c4:
        c4 = c4_block(args, world, rank, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: every kernel family, the dominant one (largest share of the step) on top -------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peak.json")) as f:
            fp64_peak = json.load(f)["cublas_dgemm_8192_tflops"]
    except Exception:  # noqa: BLE001
        fp64_peak = 35.4
    hbm_peak = peaks.get("hbm_gbs", 6650.0)                   # fallback: B200_PROFILING.md
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)    # kernels timed inside a long step
    peak_src = ("MEASURED_PEAKS.json" if peaks else "fallback of B200_PROFILING.md") + \
               "; FP64: cuBLAS DGEMM 8192^3 measured on this pool (profiles/fp64_peak.json)"
    names = {"dgemm": "vgm::dgemm_nt_kernel (FP64 DMMA, banded)",
             "tcgemm": "vgm::tc_gemm_kernel<split|single> (tcgen05 bf16, banded, operator-resident)",
             "ir_update": "vgm::ir_update_kernel", "ir_dir": "vgm::ir_dir_stream_kernel (beta from the GEMM epilogue's partial dots)",
             "vec_other": "vgm::ir_inner_init/ir_finish/ir_prep/cg_gather_x", "assemble": "state_assemble + param_stats"}
    # DRAM bytes per launch: NOT measured in this run -- read from the committed `ncu --set full` capture of the same
    # kernels at the same shape (profiles/r02_ncu_traffic.json, made by scripts/ncu_summarise.py from
    # scripts/ncu_targets.py at K=100000, T=1000; table in profiles/r02_ncu_summary.md)
    fam_kernel = {"dgemm": "void dgemm_nt_kernel", "tcgemm": "void tc_gemm_kernel", "ir_update": "void ir_update_kernel",
                  "ir_dir": "void ir_dir_stream_kernel", "assemble": "vgm_lorenz96_state_assemble", "vec_other": "ir_inner_init_kernel"}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            cap = json.load(f)
        ncu_traffic = {fam_: {"dram_bytes_per_launch": cap[k_]["dram_bytes_per_launch"],
                              "note": "from committed capture profiles/%s (kernel %s, %.1f us under ncu), see "
                                      "profiles/r02_ncu_summary.md; not measured in this run" %
                                      (cap[k_]["capture"], cap[k_]["kernel"], cap[k_]["duration_us"])}
                       for fam_, k_ in fam_kernel.items() if k_ in cap}
    except Exception:  # noqa: BLE001
        ncu_traffic = {}
    kernels = {}
    for name, f in fam.items():
        if f["ms"] <= 0:
            continue
        sec = f["ms"] * 1e-3
        tf, gbs = f["flops"] / sec / 1e12, f["bytes"] / sec / 1e9
        fpeak = fp64_peak if name == "dgemm" else bf16_peak
        frac_t, frac_h = (tf / fpeak if f["flops"] else 0.0), gbs / hbm_peak
        tensor_bound = frac_t >= frac_h
        kernels[name] = {"kernel": names[name], "bound": "tensor" if tensor_bound else "hbm",
                         "achieved": tf if tensor_bound else gbs, "peak": fpeak if tensor_bound else hbm_peak,
                         "unit": "TFLOP/s" if tensor_bound else "GB/s", "frac": frac_t if tensor_bound else frac_h,
                         "tflops": tf, "gbs": gbs, "launches": f["launches"], "ms_per_step": f["ms"] / args.steps,
                         "share_of_step": f["ms"] / ms_prof}
    roofline = None
    if kernels:
        top = max(kernels, key=lambda k: kernels[k]["share_of_step"])
        roofline = dict(kernels[top])
        tr = ncu_traffic.get(top) if (world == 1 and K == 100000 and T == 1000) else None
        roofline["traffic"] = tr["dram_bytes_per_launch"] if tr else None
        roofline["traffic_note"] = tr.get("note") if tr else None
        roofline["peak_source"] = peak_src
        roofline["all_kernels"] = kernels
    K_h = n_hidden_total
    F_canon = 2.0 * T * T * K + 2.0 * T * T * K_h + K_h * (T ** 3 / 3.0 + 2.0 * T * T) + 40.0 * T * K
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, total, blas = cpu_port_run(args.cpu_sample_K, T, 1, 0)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "Lorenz96 K=%d (%d hidden), T=%d, 1 iteration of the oracle port (%.1f s)" %
                         (args.cpu_sample_K, args.cpu_sample_K // 2, T, total), "blas": blas,
               "live_reference": live_reference_record()}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(K, T, world),
            "details": {"solver": ("mixed precision (tol %g): FP64 residuals on the DMMA pipe, FP32 PCG corrections "
                                  "(%d steps each) with bf16-split tcgen05 operator products" % (args.tol, args.inner_iters))
                        if core.solver == "mixed" else
                        ("FP64 PCG (tol %g) on the DMMA pipe, preconditioner %s" %
                         (args.tol, "none" if args.no_precondition else "(M+shift I)^-1")),
                        "bands": {"D": core.ops.band_D, "M": core.ops.band_M, "M_bf16": core.ops.band_M_lp,
                                  "G_bf16": core.ops.band_G},
                        "comm": ("single GPU" if world == 1 else
                                 "vgm_comm_* (own NCCL communicator)" if os.environ.get("VGM_COMM") == "native"
                                 else "torch.distributed (NCCL)"),
                        "state_array_mb_per_gpu": n_hidden_local * core.ld * 8 / 1e6,
                        "profiled_ms_per_step": (ms_prof / args.steps) if ms_prof else None},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "c4": c4, "colour_schedule": colour, "gpu_launches": launches,
            "clocks": clocks,
            "solver_iterations": pcg_iters, "theta": theta, "setup_s": setup_s,
            "evolving_loop": evolving,
            "canonical_flops": {"per_step": F_canon, "tflops": F_canon * args.steps / (ms * 1e-3) / 1e12,
                                "note": "F of SURVEY.md 8(d) (dense Cholesky per state); the executed algorithm is "
                                        "iterative on banded operators, see roofline for executed work"}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY FAILURE against the oracle: %r\n" % (parity,))
        return 3
    if c4 is not None and not c4["parity"]["ok"]:
        sys.stderr.write("bench.py: C4 PARITY FAILURE against the oracle: %r\n" % (c4["parity"],))
        return 3
    return 0


def main():
    args = parse()
    # Libraries (NCCL's version banner, for one) print to stdout; the contract is ONE JSON line there.
    # Keep the real stdout for that line and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return reference_arm(args)
    return native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
