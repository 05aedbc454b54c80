"""Multi-rank parity check (launch with torchrun, one rank per GPU): the state-block sharded
engine must reproduce the single-GPU engine.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.distributed import DistributedVGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.engine import VGMEngine  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.GP_regression import kernel_function_device  # noqa: E402
from variational_gradient_matching_for_dynamical_systems_b200.synthetic import lorenz96_problem  # noqa: E402


def _sum_int(v, dev):
    t = torch.tensor([int(v)], device=dev)
    dist.all_reduce(t)
    return t.item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    for tag, K, T, allhidden in (("alt_K64_T100", 64, 100, False), ("allhidden_K16_T40", 16, 40, True),
                                 ("alt_K2000_T1000", 2000, 1000, False)):
        tg = np.arange(T) * 0.05
        kf = kernel_function_device(tg, tg, "rbf", [0.0625, 1.0])
        D = kf["D"][:, :T].cpu().numpy()
        _, X0, hidden = lorenz96_problem(K, T, 0.05, seed=3, device=dev)
        if allhidden:
            hidden = np.arange(K)
        model = OdeModel.lorenz96(K)
        single = VGMEngine(model, D, None, hidden=hidden)
        single.set_states_state_major(X0)
        multi = DistributedVGMEngine(model, D, None, hidden=hidden)
        multi.set_states_state_major(X0)
        # the same engine with both exchanges through the library's own C-ABI wrappers (vgm_comm_*)
        native = DistributedVGMEngine(model, D, None, hidden=hidden, comm="native")
        native.set_states_state_major(X0)
        errs, errs_native = [], []
        for it in range(2):
            single.iteration()
            multi.iteration()
            native.iteration()
            Xs = single.X[:, :T]
            th = single.get_theta()[0]
            for eng, out in ((multi, errs), (native, errs_native)):
                Xm = eng.gather_states()
                out.append((((Xs - Xm).abs().max() / Xs.abs().max()).item(), abs(th - eng.get_theta()[0]) / abs(th)))
        # pooled 'minimizer' loss / gradient: sum over the ranks' ODE rows == single-GPU value
        single.compute_DX()
        multi.halo(multi.engine.X)
        multi.engine.compute_DX()
        ls, gs = single.param_loss([7.5], alpha=0.3)
        lm, gm = multi.param_loss([7.5], alpha=0.3)
        errs.append((0.0, max(abs(ls - lm) / abs(ls), float(np.max(np.abs(gs - gm)) / np.max(np.abs(gs))))))
        res[tag] = {"state_rel_err": max(e[0] for e in errs), "theta_rel_err": max(e[1] for e in errs),
                    "native_comm_state_rel_err": max(e[0] for e in errs_native),
                    "native_comm_theta_rel_err": max(e[1] for e in errs_native),
                    "levels": multi.n_levels, "shift": multi.shift, "cross_rank_live": multi.cross_rank_live,
                    "tail_overlap_ranks": int(_sum_int(multi.engine.plan.c.n_early > 0, dev))}
    ok = all(v["state_rel_err"] < 1e-10 and v["theta_rel_err"] < 1e-11 and v["native_comm_state_rel_err"] < 1e-10
             and v["native_comm_theta_rel_err"] < 1e-11 for v in res.values())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "ok": bool(flag.item()), "cases": res}))
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
