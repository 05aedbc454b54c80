"""Per-phase wall/GPU timing of DistributedVGMEngine.iteration (launch with torchrun).  Rank 0 prints JSON."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.distributed import DistributedVGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.engine import Operators  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, T = int(os.environ.get("K", 100000)), 1000
    tg = np.arange(T) * 0.05
    kf = kernel_function_device(tg, tg, "rbf", [0.0625, 1.0])
    ops = Operators(kf["D"][:, :T], None, device=dev)
    _, X0, hidden = lorenz96_problem(K, T, 0.05, seed=0, device=dev)
    eng = DistributedVGMEngine(OdeModel.lorenz96(K), ops, None, hidden=hidden)
    eng.set_states_state_major(X0)
    e = eng.engine
    X0l = e.X.clone()
    phases = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        phases[name] = phases.get(name, 0.0) + (time.perf_counter() - t0) * 1e3

    for it in range(5):
        e.X.copy_(X0l)
        if it == 2:
            phases.clear()
        dist.barrier()
        timed("halo0", lambda: eng.halo(e.X))
        timed("DX", e.compute_DX)
        timed("stats", e.param_stats)
        timed("allreduce", lambda: dist.all_reduce(e.stats, op=dist.ReduceOp.SUM))
        timed("solve", e.param_solve)
        timed("prepare", e.sweep_prepare)
        if not eng.cross_rank_live:
            timed("levels", e.sweep_levels)
        else:
            for lev in range(eng.n_levels):
                if lev > 0:
                    timed("halo%d" % lev, lambda: eng.halo(e.X))
                timed("level%d" % lev, lambda: e.sweep_level(lev))
        timed("barrier", dist.barrier)
    out = {k: round(v / 3, 3) for k, v in phases.items()}
    out["info"] = dict(e.last_info)
    out["n_hidden_local"] = e.plan.n_hidden
    out["tail_overlap"] = int(e.plan.c.n_early)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps({"world": world, "K": K, "per_rank_ms": gathered}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
