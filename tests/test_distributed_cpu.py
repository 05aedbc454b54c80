"""Host-side logic of the multi-GPU sharding (no GPU needed): block localisation of the model
tables / sweep plan, and the halo exchange over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from variational_gradient_matching_for_dynamical_systems_b200 import OdeModel, build_host_plan
from variational_gradient_matching_for_dynamical_systems_b200.distributed import (HaloExchange, block_bounds, choose_shift,
                                                                                   live_edges, localize, owned_states, owner_of)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_block_bounds_alignment():
    b = block_bounds(100000, 8, align=2)
    assert b[0] == 0 and b[-1] == 100000 and all(x % 2 == 0 for x in b)
    assert max(np.diff(b)) - min(np.diff(b)) <= 2
    assert block_bounds(10, 3, align=2) == [0, 2, 6, 10]


def test_weighted_bounds_give_the_chain_owner_fewer_states():
    from variational_gradient_matching_for_dynamical_systems_b200.distributed import state_weights
    K, world = 100000, 8
    model = OdeModel.lorenz96(K)
    hp = build_host_plan(model, list(range(1, K, 2)))
    w = state_weights(model, hp, 1800)
    assert w.sum() == K + 1800 and w[K - 1] == 1801          # the wrap-around state carries its level
    b = block_bounds(K, world, align=2, weights=w)
    sizes = np.diff(b)
    assert b[0] == 0 and b[-1] == K and all(x % 2 == 0 for x in b)
    assert sizes[-1] < sizes[0] - 1500 and max(sizes[:-1]) - min(sizes[:-1]) <= 2
    tot = [w[b[g]:b[g + 1]].sum() for g in range(world)]
    assert max(tot) - min(tot) <= 4
    # an all-hidden chain (every level tiny) is not re-weighted; no weights = equal counts
    hp2 = build_host_plan(OdeModel.lorenz96(64), list(range(64)))
    assert np.all(state_weights(OdeModel.lorenz96(64), hp2, 1800) == 1.0)
    assert block_bounds(10, 3, align=2) == [0, 2, 6, 10]


@pytest.mark.parametrize("K,world,allhidden", [(24, 3, False), (20, 2, False), (12, 2, True), (30, 4, False)])
def test_localize_reproduces_global_structure(K, world, allhidden):
    model = OdeModel.lorenz96(K)
    hidden = list(range(K)) if allhidden else list(range(1, K, 2))
    hp = build_host_plan(model, hidden)
    bounds = block_bounds(K, world, align=2)
    rng = np.random.default_rng(0)
    Xg = rng.standard_normal((K, 7))
    seen_slots = []
    pair_slot_g = np.repeat(np.arange(hp.n_hidden), np.diff(hp.pair_ptr))
    for r in range(world):
        loc = localize(model, hp, bounds[r], bounds[r + 1])
        rows = np.concatenate([np.arange(loc.k0, loc.k1), loc.halo_states])
        Xl = Xg[rows]
        # theta-side tables: b_k = x_{k+1} x_{k-1} - x_{k-2} x_{k-1} - x_k evaluated through the local gather table
        s = Xl[loc.ode_idx]                                 # [n_odes, 4, T]
        b_loc = s[:, 0] * s[:, 2] - s[:, 1] * s[:, 2] - s[:, 3]
        k = np.arange(loc.k0, loc.k1)
        b_glob = Xg[(k + 1) % K] * Xg[(k - 1) % K] - Xg[(k - 2) % K] * Xg[(k - 1) % K] - Xg[k]
        assert np.allclose(b_loc[:loc.n_stats_odes], b_glob)
        assert np.array_equal(rows[loc.ode_state[:loc.n_stats_odes]], k)
        # halo width: Lorenz96 reaches 3 states to either side
        if world > 1:
            assert len(loc.halo_states) <= 6
        # sweep plan: same (state, ODE, code, self, live) tuples as the global plan restricted to the block
        lp = loc.plan
        g_sel = np.nonzero((hp.slot_state[pair_slot_g] >= loc.k0) & (hp.slot_state[pair_slot_g] < loc.k1))[0]
        l_slot = np.repeat(np.arange(lp.n_hidden), np.diff(lp.pair_ptr))
        assert np.array_equal(rows[lp.slot_state[l_slot]], hp.slot_state[pair_slot_g[g_sel]])
        assert np.array_equal(rows[loc.ode_state[lp.pair_ode]], model.ode_state[hp.pair_ode[g_sel]])
        assert np.array_equal(lp.pair_code, hp.pair_code[g_sel])
        assert np.array_equal(lp.pair_self, hp.pair_self[g_sel])
        assert np.array_equal(lp.pair_live >= 0, hp.pair_live[g_sel] >= 0)
        # live reads point at the right states and producer levels
        for e in np.nonzero(lp.pair_live >= 0)[0]:
            q = lp.pair_live[e]
            src_state_local = rows[lp.live_state[lp.live_pair_src[q]]]
            gq = hp.pair_live[g_sel[e]]
            assert src_state_local == hp.live_state[hp.live_pair_src[gq]]
            src = lp.live_pair_src[q]
            lev_src = np.searchsorted(lp.live_level_ptr, src, side="right") - 1
            lev_cons = lp.level[lp.live_pair_slot[q]]
            assert lev_src < lev_cons
        assert lp.n_levels == hp.n_levels and lp.level_ptr[-1] == lp.n_hidden
        seen_slots += rows[lp.slot_state].tolist()
    assert sorted(seen_slots) == sorted(hidden)


def _halo_worker(rank, world, port, K, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = OdeModel.lorenz96(K)
        hp = build_host_plan(model, list(range(1, K, 2)))
        bounds = block_bounds(K, world, align=2)
        loc = localize(model, hp, bounds[rank], bounds[rank + 1])
        hx = HaloExchange(loc, bounds)
        Xg = torch.arange(K * 5, dtype=torch.float64).reshape(K, 5) + 0.25
        X = torch.zeros((loc.n_rows, 5), dtype=torch.float64)
        X[:loc.n_owned] = Xg[loc.k0:loc.k1]
        hx(X)
        rows = np.concatenate([np.arange(loc.k0, loc.k1), loc.halo_states])
        ok = bool(torch.equal(X, Xg[torch.as_tensor(rows)]))
        # the parameter statistics all-reduce: sum of per-rank partial sums == global sum
        part = X[:loc.n_owned].sum().reshape(1).clone()
        dist.all_reduce(part)
        ok = ok and abs(part.item() - Xg.sum().item()) < 1e-9
        with open(os.path.join(outdir, "rank%d.txt" % rank), "w") as f:
            f.write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_halo_exchange_and_allreduce_gloo_world2(tmp_path):
    world, K = 2, 16
    port = _free_port()
    # spawn context only (no fork: forking a process that already runs BLAS threads can deadlock)
    mp.spawn(_halo_worker, args=(world, port, K, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d.txt" % r)).read() == "ok"


PT_LINES = ["-_k_1 * _S - _k_2 * _S * _R + _k_3 * _RS", "_k_1 * _S", "-_k_2 * _S * _R + _k_3 * _RS + _V * _Rpp / (0.3 + _Rpp)",
            "_k_2 * _S * _R - _k_3 * _RS - _k_4 * _RS", "_k_4 * _RS - _V * _Rpp / (0.3 + _Rpp)"]
PT_NAMES = ["_S", "_dS", "_R", "_RS", "_Rpp"]
PT_PARAMS = ["_k_1", "_k_2", "_k_3", "_k_4", "_V"]


def _batch_worker(rank, world, port, n_traj, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        K1 = len(PT_NAMES)
        model = OdeModel(PT_LINES, PT_NAMES, PT_PARAMS, n_copies=n_traj)
        hidden = [n * K1 + u for n in range(n_traj) for u in (0, 3)]        # _S and _RS of every trajectory
        hp = build_host_plan(model, hidden)
        bounds = block_bounds(model.K, world, align=K1)                       # whole trajectories per rank
        loc = localize(model, hp, bounds[rank], bounds[rank + 1])
        hx = HaloExchange(loc, bounds)
        ok = all(b % K1 == 0 for b in bounds) and len(loc.halo_states) == 0 and hx.empty
        X = torch.full((loc.n_rows, 4), float(rank + 1), dtype=torch.float64)
        hx(X)                                                                 # no exchange, nothing touched
        ok = ok and bool((X == float(rank + 1)).all())
        # every trajectory's ODE rows and hidden slots live on exactly one rank
        cnt = torch.tensor([loc.n_stats_odes, loc.plan.n_hidden], dtype=torch.int64)
        dist.all_reduce(cnt)
        ok = ok and cnt.tolist() == [model.K, len(hidden)]
        with open(os.path.join(outdir, "rank%d.txt" % rank), "w") as f:
            f.write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_trajectory_batch_partition_has_no_halo_gloo_world2(tmp_path):
    """Config C4's sharding: trajectories split across ranks, no cross-trajectory coupling, so the halo
    exchange is skipped and the only collective left is the all-reduce of the pooled statistics."""
    world = 2
    port = _free_port()
    mp.spawn(_batch_worker, args=(world, port, 6, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d.txt" % r)).read() == "ok"


@pytest.mark.parametrize("K,world", [(84, 3), (120, 4), (100000, 8)])
def test_shifted_blocks_keep_the_wrap_around_read_on_one_rank(K, world):
    """Lorenz96 with every second state observed has ONE live read, x_1 -> x_{K-1} (proxies...py:225); moving the
    blocks two states up the ring puts both on the last rank, so no rank waits for another between levels, and that
    rank's local plan keeps the tail-overlap tables (early = x_1, tail = x_{K-1})."""
    model = OdeModel.lorenz96(K)
    hidden = list(range(1, K, 2))
    hp = build_host_plan(model, hidden)
    src, dst = live_edges(hp)
    assert src.tolist() == [1] and dst.tolist() == [K - 1]
    assert hp.n_early == 1 and hp.slot_state[hp.early_slots].tolist() == [1]
    assert len(hp.bulk_slots) == hp.level_ptr[1] - 1 and 0 not in hp.bulk_slots.tolist()
    bounds = block_bounds(K, world, align=2)
    assert choose_shift(hp, bounds, K, align=2) == 2
    assert owner_of([1, K - 1, 2, 0], bounds, K, 2).tolist() == [world - 1, world - 1, 0, world - 1]
    seen, hidden_seen = [], []
    for r in range(world):
        own = owned_states(bounds, r, K, 2)
        loc = localize(model, hp, bounds[r], bounds[r + 1], owned=own)
        assert np.array_equal(loc.owned, own) and loc.n_owned == len(own)
        rows = np.concatenate([loc.owned, loc.halo_states])
        assert np.array_equal(loc.row_map[rows], np.arange(len(rows)))
        assert len(loc.halo_states) <= 6
        lp = loc.plan
        hidden_seen += rows[lp.slot_state].tolist()
        seen += own.tolist()
        if r == world - 1:
            assert lp.n_early == 1 and rows[lp.slot_state[lp.early_slots]].tolist() == [1]
            assert rows[lp.slot_state[lp.level_slots[lp.level_ptr[1]:]]].tolist() == [K - 1]
            assert sorted(lp.bulk_slots.tolist() + lp.early_slots.tolist()) == sorted(lp.level_slots[:lp.level_ptr[1]].tolist())
            # the live read is local: its source row is an OWNED row
            assert lp.n_live == 1 and lp.live_state[0] < loc.n_owned
        else:
            assert lp.n_early == 0 and lp.n_live_pairs == 0 and lp.level_ptr[1] == lp.n_hidden
    assert sorted(seen) == list(range(K)) and sorted(hidden_seen) == hidden


def test_all_hidden_chain_has_no_tail_tables_and_no_good_shift():
    K = 16
    hp = build_host_plan(OdeModel.lorenz96(K), list(range(K)))
    assert hp.n_early == 0 and hp.n_levels > 8
    assert choose_shift(hp, block_bounds(K, 2, align=2), K, align=2) is None


def _shift_worker(rank, world, port, K, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = OdeModel.lorenz96(K)
        hp = build_host_plan(model, list(range(1, K, 2)))
        bounds = block_bounds(K, world, align=2)
        sh = choose_shift(hp, bounds, K, align=2)
        loc = localize(model, hp, bounds[rank], bounds[rank + 1], owned=owned_states(bounds, rank, K, sh))
        hx = HaloExchange(loc, bounds, shift=sh, K=K)
        Xg = torch.arange(K * 5, dtype=torch.float64).reshape(K, 5) + 0.25
        X = torch.zeros((loc.n_rows, 5), dtype=torch.float64)
        X[:loc.n_owned] = Xg[torch.as_tensor(loc.owned)]
        hx(X)
        rows = np.concatenate([loc.owned, loc.halo_states])
        ok = sh == 2 and bool(torch.equal(X, Xg[torch.as_tensor(rows)]))
        # the same exchange as include/vgm_b200.h documents it for vgm_comm_halo_exchange (pack pub_rows padded to
        # n_pub -> all-gather -> X[halo_rows[i]] = gathered[halo_src[i]]), emulated with the tables HaloExchange
        # hands to the C entry point (gloo all_gather stands in for ncclAllGather)
        X2 = torch.zeros_like(X)
        X2[:loc.n_owned] = Xg[torch.as_tensor(loc.owned)]
        pub = torch.zeros((hx.n_pub, 5), dtype=torch.float64)
        if len(hx.pub_rows_local):
            pub[:len(hx.pub_rows_local)] = X2[torch.as_tensor(hx.pub_rows_local)]
        parts = [torch.zeros_like(pub) for _ in range(world)]
        dist.all_gather(parts, pub)
        gathered = torch.cat(parts)
        X2[torch.as_tensor(hx.halo_rows_local)] = gathered[torch.as_tensor(hx.src_flat)]
        ok = ok and bool(torch.equal(X2, X))
        with open(os.path.join(outdir, "rank%d.txt" % rank), "w") as f:
            f.write("ok" if ok else "bad")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_halo_exchange_with_shifted_blocks_gloo_world2(tmp_path):
    world, K = 2, 20
    port = _free_port()
    mp.spawn(_shift_worker, args=(world, port, K, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d.txt" % r)).read() == "ok"


def test_partition_properties_random_shapes():
    """Property check over random (K, world, align, shift): every state has exactly one owner, `owner_of` inverts
    `owned_states`, block boundaries sit on multiples of `align`, and block sizes differ by at most one aligned unit
    (the quantities the halo tables and the C-ABI halo exchange are built from)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 8), st.integers(1, 4), st.integers(0, 40), st.integers(0, 6))
    def check(world, align, extra, shift_units):
        K = world * align * 2 + extra * align          # at least two aligned units per rank, K a multiple of align
        bounds = block_bounds(K, world, align=align)
        assert bounds[0] == 0 and bounds[-1] == K and all(b % align == 0 for b in bounds)
        sizes = np.diff(bounds)
        assert sizes.min() > 0 and sizes.max() - sizes.min() <= align
        shift = shift_units * align
        seen = np.zeros(K, dtype=int)
        for r in range(world):
            own = owned_states(bounds, r, K, shift)
            assert len(own) == sizes[r]
            seen[own] += 1
            assert (owner_of(own, bounds, K, shift) == r).all()
        assert (seen == 1).all()

    check()
