"""Config C4 sharded over GPUs (launch with torchrun, one rank per GPU): trajectories are split across ranks, the
only collective is the all-reduce of the pooled parameter statistics.  Checks the sharded run against
a single-GPU engine on the same batch and prints one JSON line on rank 0 (device time, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/c4_distributed.py --trajectories 65536
"""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectories", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = np.load(os.path.join(ROOT, "tests", "golden", "protein_transduction_T99.npz"))
    lines, params, names = [str(s) for s in g["lines"]], [str(s) for s in g["params"]], [str(s) for s in g["names"]]
    observed = [str(s) for s in g["observed"]]
    K1, T, N = len(names), g["X0"].shape[0], a.trajectories
    hidden1 = [u for u in range(K1) if names[u] not in observed]
    rng = np.random.default_rng(7)
    scale = 1.0 + 0.05 * rng.standard_normal((N, 1, K1))
    X_KT = np.ascontiguousarray((g["X0"][None, :, :] * scale).transpose(0, 2, 1).reshape(N * K1, T))   # state-major
    model = OdeModel(lines, names, params, n_copies=N)
    hidden = np.array([n * K1 + u for n in range(N) for u in hidden1])
    multi = DistributedVGMEngine(model, g["D"], g["A"], hidden=hidden, align=K1, theta_mode="proxy")
    multi.set_states_state_major(X_KT)
    multi.iteration()
    Xm = multi.gather_states()
    th_m = multi.get_theta()
    out = {"world": world, "n_trajectories": N, "T": T, "hidden_states": int(len(hidden)), "halo_empty": bool(multi.halo.empty)}
    if rank == 0:
        single = VGMEngine(model, g["D"], g["A"], hidden=hidden, theta_mode="proxy")
        single.set_states_state_major(torch.as_tensor(X_KT))
        single.iteration()
        Xs = single.X[:, :T]
        out["state_rel_err"] = ((Xs - Xm).abs().max() / Xs.abs().max()).item()
        ths = single.get_theta()
        out["theta_rel_err"] = float(np.max(np.abs(np.asarray(th_m) - np.asarray(ths))) / np.max(np.abs(ths)))
        del single
    multi.set_states_state_major(X_KT)
    multi.iteration()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        multi.iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        out["ms_per_iteration"] = round(ms.item(), 3)
        out["state_time_updates_per_s"] = len(hidden) * T / (ms.item() * 1e-3)
        out["ok"] = bool(out["state_rel_err"] < 1e-10 and out["theta_rel_err"] < 1e-10)
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
