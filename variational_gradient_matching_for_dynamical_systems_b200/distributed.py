"""Multi-GPU sharding of the VGM loop: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the B200 box, gloo in CPU tests) for the two exchanges the path really has
(the reference itself is single-process):

* **state blocks** (Lorenz96 K=100k): rank g owns a contiguous block of states; before a sweep
  and between dependency levels the few boundary rows of X that neighbours read (3 on each
  side for Lorenz96) are refreshed with one small all-gather;
* **trajectory batches**: trajectories are split across ranks, no halo at all;

and, in both cases, one all-reduce (sum) of the ODE-parameter sufficient statistics
``[m | S]`` (p + p^2 doubles) per iteration, after which every rank solves the p x p system
redundantly.  Every rank executes the *global* dependency levels in the reference's order, so
the result is the single-GPU result (up to the summation order of the all-reduce).

``comm="native"`` routes both exchanges through the library's own C-ABI wrappers (``vgm_comm_halo_exchange``: pack
kernel -> ``ncclAllGather`` -> unpack kernel on one stream; ``vgm_comm_allreduce_stats``), i.e. exactly what a
C / C++ caller without PyTorch would call; ``torch.distributed`` then only carries the 128-byte NCCL id at set-up.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .codegen import HostPlan, OdeModel, build_host_plan


class NativeComm:
    """A ``vgm_comm`` (include/vgm_b200.h): the library's own NCCL communicator over the ranks of a
    ``torch.distributed`` group, which is used once, to hand rank 0's 128-byte NCCL id to the others."""

    def __init__(self, group=None):
        import ctypes as C
        self.lib = _lib.load()                       # torch is imported above: the process's libnccl.so.2 is torch's
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if self.rank == 0:
            _lib.check(self.lib.vgm_comm_unique_id(ident))
        box = [ident.raw if self.rank == 0 else None]
        src = 0 if group is None else dist.get_global_rank(group, 0)
        dist.broadcast_object_list(box, src=src, group=group)
        handle = C.c_void_p()
        _lib.check(self.lib.vgm_comm_init(box[0], self.rank, self.world, C.byref(handle)))
        self.handle = handle

    def allreduce_stats(self, stats):
        """In-place sum over ranks of a contiguous FP64 device tensor, on the current stream."""
        _lib.check(self.lib.vgm_comm_allreduce_stats(self.handle, stats.data_ptr(), stats.numel(), _lib.stream_ptr()))

    def halo_ws(self, n_pub, ld, device):
        return torch.empty(max(self.lib.vgm_comm_halo_ws_bytes(self.world, n_pub, ld), 16), dtype=torch.uint8, device=device)

    def halo_exchange(self, X, pub_rows, n_pub, halo_rows, halo_src, ws):
        _lib.check(self.lib.vgm_comm_halo_exchange(self.handle, X.data_ptr(), X.shape[1], _lib.ptr(pub_rows),
                                                   0 if pub_rows is None else pub_rows.numel(), n_pub,
                                                   _lib.ptr(halo_rows), _lib.ptr(halo_src),
                                                   0 if halo_rows is None else halo_rows.numel(), ws.data_ptr(),
                                                   ws.numel(), _lib.stream_ptr()))

    def close(self):
        if self.handle is not None:
            self.lib.vgm_comm_free(self.handle)
            self.handle = None


def block_bounds(n_items, world, align=1, weights=None):
    """Contiguous blocks [b[g], b[g+1]) with boundaries on multiples of ``align``: nearly equal counts, or -- with
    per-item ``weights`` -- nearly equal total weight (boundaries at the weight quantiles)."""
    if weights is None:
        units = n_items // align
        return [(units * g // world) * align for g in range(world)] + [n_items]
    w = np.asarray(weights, dtype=np.float64)
    if len(w) != n_items or np.any(w < 0) or w.sum() <= 0:
        raise ValueError("weights must be one non-negative number per item")
    cum = np.concatenate([[0.0], np.cumsum(w)])
    b = [0]
    for g in range(1, world):
        k = int(np.searchsorted(cum, cum[-1] * g / world))
        k = int(round(k / align)) * align
        b.append(min(max(k, b[-1]), n_items))
    return b + [n_items]


def state_weights(model: OdeModel, hp: HostPlan, tail_level_cost_states):
    """Work per state in units of "one bulk system": 1 for every state, plus the cost of the dependency levels
    after the first for the states that are solved there.  Such a level holds a handful of systems (Lorenz96 with
    every second state observed: the wrap-around state alone) but costs a full sequence of launch-bound kernels,
    ``tail_level_cost_states`` bulk systems' worth, on the rank that owns it while the others wait."""
    w = np.ones(model.K, dtype=np.float64)
    level = np.asarray(hp.level)
    if hp.n_levels > 1 and tail_level_cost_states > 0 and hp.level_ptr[1] > 8 * (hp.n_hidden - hp.level_ptr[1]):
        counts = np.bincount(level, minlength=hp.n_levels).astype(np.float64)
        later = np.nonzero(level > 0)[0]
        w[np.asarray(hp.slot_state)[later]] += tail_level_cost_states / counts[level[later]]
    return w


class Localized:
    """Result of :func:`localize` (everything a rank needs, all host-side NumPy)."""


def owned_states(bounds, rank, K, shift=0):
    """States of rank ``rank``: the block [bounds[rank], bounds[rank+1]) moved ``shift`` states up the ring
    (mod K), in local row order.  A shift keeps a live read that wraps around the ring (Lorenz96: the last
    hidden state reads the updated first one, proxies...py:225) inside one rank."""
    return (np.arange(bounds[rank], bounds[rank + 1], dtype=np.int64) + shift) % K


def owner_of(states, bounds, K, shift=0):
    """Rank that owns each state under ``owned_states``."""
    s = (np.asarray(states, dtype=np.int64) - shift) % K
    return np.searchsorted(np.asarray(bounds), s, side="right") - 1


def live_edges(hp: HostPlan):
    """(source state, consumer state) of every live read of the plan: the consumer's system takes D.x of the
    source from the column updated earlier in the same sweep (proxies...py:185,225)."""
    if hp.n_live_pairs == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    slot_state = np.asarray(hp.slot_state, dtype=np.int64)
    src = np.asarray(hp.live_state, dtype=np.int64)[np.asarray(hp.live_pair_src, dtype=np.int64)]
    dst = slot_state[np.asarray(hp.live_pair_slot, dtype=np.int64)]
    return src, dst


def choose_shift(hp: HostPlan, bounds, K, align=1, max_tries=16):
    """Smallest block shift (a multiple of ``align``) under which no live read crosses a rank boundary, or None.
    Lorenz96 with every second state observed: the only live read is x_1 -> x_{K-1}; shift 2 puts both on the last
    rank, so no rank waits for another between dependency levels."""
    src, dst = live_edges(hp)
    for i in range(max_tries):
        sh = i * align
        if len(src) == 0 or np.array_equal(owner_of(src, bounds, K, sh), owner_of(dst, bounds, K, sh)):
            return sh
    return None


def localize(model: OdeModel, hp: HostPlan, k0: int, k1: int, owned=None) -> Localized:
    """Restrict a global model + sweep plan to the states owned by one rank: [k0, k1), or the explicit list
    ``owned`` (local row order; ``owned_states`` gives the cyclically shifted blocks).

    Local X rows: owned states first (local row = position in ``owned``), then the halo states (ascending
    global index) that owned ODE rows / owned hidden states read.  Local ODE rows: owned ODEs
    first (they alone enter the theta statistics), then the neighbours' ODEs that couple to an
    owned hidden state (needed for R_uk, r_uk and D x_k of those pairs)."""
    K = model.K
    owned = np.arange(k0, k1, dtype=np.int64) if owned is None else np.asarray(owned, dtype=np.int64)
    is_owned = np.zeros(K, dtype=bool)
    is_owned[owned] = True
    arity_of = np.array([c.arity for c in model.classes], dtype=np.int64)
    slot_state = np.asarray(hp.slot_state, dtype=np.int64)
    loc_slots = np.nonzero(is_owned[slot_state])[0]
    slot_map = np.full(hp.n_hidden, -1, dtype=np.int64)
    slot_map[loc_slots] = np.arange(len(loc_slots))
    pair_slot = np.repeat(np.arange(hp.n_hidden), np.diff(hp.pair_ptr))
    pair_sel = np.nonzero(slot_map[pair_slot] >= 0)[0]
    owned_odes = owned.copy()
    extra_odes = np.setdiff1d(np.unique(hp.pair_ode[pair_sel]).astype(np.int64), owned_odes)
    odes = np.concatenate([owned_odes, extra_odes])
    ode_map = np.full(K, -1, dtype=np.int64)
    ode_map[odes] = np.arange(len(odes))
    # states referenced by those ODE rows
    idx = model.ode_idx[odes].astype(np.int64)
    valid = np.arange(model.arity)[None, :] < arity_of[model.ode_class[odes]][:, None]
    # live sources read by local consumers
    live_e = pair_sel[hp.pair_live[pair_sel] >= 0]
    src_global_state = np.asarray(hp.live_state, dtype=np.int64)[hp.live_pair_src[hp.pair_live[live_e]]] \
        if len(live_e) else np.zeros(0, dtype=np.int64)
    needed = np.unique(np.concatenate([idx[valid], model.ode_state[odes].astype(np.int64), src_global_state]))
    halo = needed[~is_owned[needed]]
    n_owned = len(owned)
    row_map = np.full(K, -1, dtype=np.int64)
    row_map[owned] = np.arange(n_owned)
    row_map[halo] = n_owned + np.arange(len(halo))
    out = Localized()
    out.k0, out.k1 = k0, k1
    out.owned = owned
    out.row_map = row_map
    out.n_owned = n_owned
    out.halo_states = halo
    out.n_rows = n_owned + len(halo)
    out.ode_class = model.ode_class[odes].astype(np.int32)
    li = np.where(valid, row_map[idx], 0)
    assert (li >= 0).all()
    out.ode_idx = li.astype(np.int32)
    out.ode_state = row_map[model.ode_state[odes]].astype(np.int32)
    out.n_stats_odes = len(owned_odes)
    # ---- local sweep plan (global levels are kept so that all ranks step through them together)
    lp = HostPlan()
    nh = len(loc_slots)
    lp.n_hidden = nh
    lp.slot_state = row_map[slot_state[loc_slots]].astype(np.int32)
    counts = np.diff(hp.pair_ptr)[loc_slots]
    lp.pair_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    lp.n_pairs = len(pair_sel)
    lp.pair_ode = ode_map[hp.pair_ode[pair_sel]].astype(np.int32)
    lp.pair_code = hp.pair_code[pair_sel].astype(np.int32)
    lp.pair_self = hp.pair_self[pair_sel].astype(np.int32)
    lp.slot_has_self = hp.slot_has_self[loc_slots].astype(np.int32)
    lp.n_levels = hp.n_levels
    lp.level = hp.level[loc_slots]
    order = np.lexsort((np.arange(nh), lp.level))
    lp.level_slots = order.astype(np.int32)
    lptr = np.zeros(hp.n_levels + 1, dtype=np.int64)
    np.add.at(lptr, lp.level + 1, 1)
    lp.level_ptr = np.cumsum(lptr).astype(np.int32)
    # live pairs: keep the global ordering (by consumer level, slot, k)
    g_live_ord = hp.pair_live[pair_sel]                       # -1 or global live-pair ordinal
    is_live = g_live_ord >= 0
    live_local_idx = np.nonzero(is_live)[0]
    o = np.argsort(g_live_ord[live_local_idx], kind="stable")
    live_local_idx = live_local_idx[o]
    lp.pair_live = np.full(len(pair_sel), -1, dtype=np.int32)
    lp.pair_live[live_local_idx] = np.arange(len(live_local_idx), dtype=np.int32)
    lp.n_live_pairs = len(live_local_idx)
    g_ord = g_live_ord[live_local_idx]
    consumer_slot_local = slot_map[pair_slot[pair_sel[live_local_idx]]]
    lp.live_pair_slot = consumer_slot_local.astype(np.int32)
    g_src = hp.live_pair_src[g_ord] if len(g_ord) else np.zeros(0, dtype=np.int64)   # global live-buffer slot
    # producer level of every global live source
    g_src_level = np.zeros(hp.n_live, dtype=np.int64)
    for lev in range(hp.n_levels):
        g_src_level[hp.live_level_ptr[lev]:hp.live_level_ptr[lev + 1]] = lev
    used_src = np.unique(g_src)                               # ascending = grouped by producer level
    src_map = {int(s): i for i, s in enumerate(used_src.tolist())}
    lp.n_live = len(used_src)
    lp.live_state = row_map[np.asarray(hp.live_state, dtype=np.int64)[used_src]].astype(np.int32) \
        if len(used_src) else np.zeros(0, dtype=np.int32)
    llp = np.zeros(hp.n_levels + 1, dtype=np.int64)
    if len(used_src):
        np.add.at(llp, g_src_level[used_src] + 1, 1)
    lp.live_level_ptr = np.cumsum(llp).astype(np.int32)
    lp.live_pair_src = np.array([src_map[int(s)] for s in g_src.tolist()], dtype=np.int32)
    lpl = np.zeros(hp.n_levels + 1, dtype=np.int64)
    if lp.n_live_pairs:
        np.add.at(lpl, lp.level[consumer_slot_local] + 1, 1)
    lp.live_pair_level_ptr = np.cumsum(lpl).astype(np.int32)
    lp.ruu_const, lp.ruu_value = hp.ruu_const, hp.ruu_value
    # tail overlap (codegen.tail_overlap_tables): usable on this rank only if it owns the whole chain, i.e. every
    # early system of the global plan together with every system of the later levels
    n_early = int(getattr(hp, "n_early", 0))
    lp.n_early, lp.early_slots, lp.bulk_slots = 0, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32)
    if n_early:
        early_l = slot_map[np.asarray(hp.early_slots, dtype=np.int64)]
        tail_g = np.nonzero(np.asarray(hp.level) > 0)[0]
        if (early_l >= 0).all() and (slot_map[tail_g] >= 0).all():
            lvl0 = np.nonzero(lp.level == 0)[0]
            lp.n_early = n_early
            lp.early_slots = np.sort(early_l).astype(np.int32)
            lp.bulk_slots = np.setdiff1d(lvl0, early_l).astype(np.int32)
    out.plan = lp
    out.hidden_global = slot_state[loc_slots]
    return out


class HaloExchange:
    """Refresh of the halo rows of X: every rank contributes the owned rows that any other
    rank reads (padded to a common count) to one all-gather and then picks its halo rows."""

    def __init__(self, loc: Localized, bounds, group=None, shift=0, K=None, comm=None):
        self.group = group
        self.comm = comm                                # NativeComm: vgm_comm_halo_exchange instead of torch ops
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        halos = [None] * self.world
        dist.all_gather_object(halos, loc.halo_states.tolist(), group=group)
        bounds = np.asarray(bounds)
        K = int(bounds[-1]) if K is None else int(K)
        send = [[] for _ in range(self.world)]          # owned states each rank must publish
        for r, hs in enumerate(halos):
            for s, owner in zip(hs, owner_of(hs, bounds, K, shift).tolist() if len(hs) else []):
                send[owner].append(int(s))
        send = [sorted(set(v)) for v in send]
        # trajectory batches: nobody reads anybody's rows, so there is no exchange at all (same answer on every rank)
        self.empty = all(len(v) == 0 for v in send)
        self.n_pub = max(1, max(len(v) for v in send))
        self.pub_rows_local = np.array([loc.row_map[s] for s in send[self.rank]], dtype=np.int64)
        pos = [{s: i for i, s in enumerate(v)} for v in send]
        owners = owner_of(loc.halo_states, bounds, K, shift).tolist() if len(loc.halo_states) else []
        self.src_flat = np.array([o * self.n_pub + pos[o][int(s)] for o, s in zip(owners, loc.halo_states.tolist())],
                                 dtype=np.int64)
        self.halo_rows_local = loc.n_owned + np.arange(len(loc.halo_states), dtype=np.int64)
        self._buf = None

    def __call__(self, X):
        """X: [n_rows, ld] tensor of this rank (owned rows first); halo rows are overwritten."""
        if self.empty:
            return
        ld = X.shape[1]
        if self.comm is not None:
            if self._buf is None or self._buf[0] != ld or self._buf[1].device != X.device:
                as_i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32), device=X.device) if len(a) else None
                self._buf = (ld, self.comm.halo_ws(self.n_pub, ld, X.device), as_i32(self.pub_rows_local),
                             as_i32(self.halo_rows_local), as_i32(self.src_flat))
            _, ws, pub_idx, halo_idx, src_idx = self._buf
            self.comm.halo_exchange(X, pub_idx, self.n_pub, halo_idx, src_idx, ws)
            return
        if self._buf is None or self._buf[0].shape[1] != ld or self._buf[0].device != X.device:
            self._buf = (torch.zeros((self.n_pub, ld), dtype=X.dtype, device=X.device),
                         torch.zeros((self.world * self.n_pub, ld), dtype=X.dtype, device=X.device),
                         torch.as_tensor(self.pub_rows_local, device=X.device),
                         torch.as_tensor(self.src_flat, device=X.device),
                         torch.as_tensor(self.halo_rows_local, device=X.device))
        pub, gathered, pub_idx, src_idx, halo_idx = self._buf
        if len(self.pub_rows_local):
            pub[:len(self.pub_rows_local)] = X.index_select(0, pub_idx)
        dist.all_gather_into_tensor(gathered, pub, group=self.group)
        if len(self.src_flat):
            X.index_copy_(0, halo_idx, gathered.index_select(0, src_idx))


class DistributedVGMEngine:
    """State-block (or trajectory-batch) sharded VGM engine; same iteration as ``VGMEngine``."""

    # Optional load balancing: give the rank that owns the tiny dependency levels (the wrap-around state) fewer
    # states, ``tail_level_cost_states`` per such level (one costs ~0.5 ms of launch-bound corrections ~ 1800 bulk
    # systems at T = 1000).  Measured on 4 and 8 B200s it changes nothing (7.17 vs 7.14 ms per step at 8 GPUs: the
    # tail is not what the other ranks wait for), so it is off by default.
    TAIL_LEVEL_COST_STATES = 0

    def __init__(self, model: OdeModel, D, A=None, hidden=None, schedule="reference", align=2, group=None,
                 tail_level_cost_states=None, shift="auto", comm="torch", **kw):
        from .engine import LocalProblem, VGMEngine
        if comm not in ("torch", "native"):
            raise ValueError("comm must be 'torch' (torch.distributed collectives) or 'native' (vgm_comm_* C ABI)")
        self.group = group
        self.comm = NativeComm(group) if comm == "native" else None
        self._side = None
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.model = model
        if hidden is None:
            hidden = np.arange(model.K)
        hp = build_host_plan(model, hidden, schedule=schedule)
        self.global_plan = hp
        tail = self.TAIL_LEVEL_COST_STATES if tail_level_cost_states is None else tail_level_cost_states
        self.bounds = block_bounds(model.K, self.world, align=align,
                                   weights=state_weights(model, hp, tail) if tail > 0 else None)
        # blocks are moved up the ring by `shift` states when that keeps every live read inside one rank: then no
        # rank waits for another between dependency levels (and the tail overlap of the sweep stays available)
        if shift == "auto":
            sh = choose_shift(hp, self.bounds, model.K, align=align)
            self.shift = 0 if sh is None else int(sh)
        else:
            self.shift = int(shift)
        src, dst = live_edges(hp)
        self.cross_rank_live = bool(len(src)) and not np.array_equal(owner_of(src, self.bounds, model.K, self.shift),
                                                                     owner_of(dst, self.bounds, model.K, self.shift))
        k0, k1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.loc = localize(model, hp, k0, k1, owned=owned_states(self.bounds, self.rank, model.K, self.shift))
        prob = LocalProblem(self.loc.n_rows, self.loc.n_stats_odes, self.loc.ode_class, self.loc.ode_idx,
                            self.loc.ode_state, self.loc.plan, model.p)
        self.engine = VGMEngine(model, D, A, problem=prob, **kw)
        self.halo = HaloExchange(self.loc, self.bounds, group=group, shift=self.shift, K=model.K, comm=self.comm)
        self.n_hidden_global = hp.n_hidden
        self.n_levels = hp.n_levels

    # ---- data ------------------------------------------------------------------------------------
    def set_states_state_major(self, X_global_KT):
        """Every rank passes the same global [K, T] array (or at least its own block + halo)."""
        e, loc = self.engine, self.loc
        rows = np.concatenate([loc.owned, loc.halo_states]).astype(np.int64)
        Xg = torch.as_tensor(X_global_KT)
        idx = torch.as_tensor(rows, device=Xg.device)
        e.X.zero_()
        e.X[:, :e.T] = Xg.index_select(0, idx)[:, :e.T].to(e.dev, dtype=torch.float64)

    def owned_states(self):
        """[n_owned, T] tensor of this rank's block (local row order = ``self.loc.owned``)."""
        return self.engine.X[:self.loc.n_owned, :self.engine.T]

    def gather_states(self):
        """Global [K, T] tensor on every rank (for tests/reporting)."""
        e = self.engine
        K = self.model.K
        sizes = [self.bounds[g + 1] - self.bounds[g] for g in range(self.world)]
        mx = max(sizes)
        mine = torch.zeros((mx, e.T), dtype=torch.float64, device=e.dev)
        mine[:self.loc.n_owned] = self.owned_states()
        allb = torch.zeros((self.world * mx, e.T), dtype=torch.float64, device=e.dev)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        out = torch.empty((K, e.T), dtype=torch.float64, device=e.dev)
        for g in range(self.world):
            idx = torch.as_tensor(owned_states(self.bounds, g, K, self.shift), device=e.dev)
            out.index_copy_(0, idx, allb[g * mx:g * mx + sizes[g]])
        return out

    def gather_rows(self, states, source=None):
        """[len(states), ld] tensor, on every rank, of the rows of the GLOBAL states ``states``: every rank
        contributes the rows it owns (one all-reduce of a few rows; for parity checks and reporting).  ``source`` is
        a local [n_rows, ld] tensor laid out like the engine's X (default: the engine's X)."""
        e = self.engine
        src = e.X if source is None else source
        states = np.asarray(states, dtype=np.int64)
        mine = owner_of(states, self.bounds, self.model.K, self.shift) == self.rank
        out = torch.zeros((len(states), src.shape[1]), dtype=src.dtype, device=src.device)
        if mine.any():
            pos = torch.as_tensor(np.nonzero(mine)[0], device=src.device)
            rows = torch.as_tensor(self.loc.row_map[states[mine]], device=src.device)
            out.index_copy_(0, pos, src.index_select(0, rows))
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- one coordinate-ascent iteration ---------------------------------------------------------------
    def iteration(self, allow_unconverged=False):
        e = self.engine
        self.halo(e.X)                                   # neighbours' boundary rows (snapshot)
        e.compute_DX()
        e.param_stats()
        # With the reference's literal theta in the state update (proxies...py:211) the sweep does not read the new
        # theta: the all-reduce then runs beside the sweep and nobody waits for the slowest rank in mid-iteration.
        late_theta = e.theta_mode == "reference"
        if self.comm is None:
            work = dist.all_reduce(e.stats, op=dist.ReduceOp.SUM, group=self.group, async_op=late_theta)
        elif late_theta:
            if self._side is None:
                self._side = torch.cuda.Stream(device=e.dev)
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self.comm.allreduce_stats(e.stats)
        else:
            self.comm.allreduce_stats(e.stats)
        if not late_theta:
            e.param_solve()
        e.sweep_prepare()
        if not self.cross_rank_live:
            # every live read stays on its rank: the levels run back to back, nobody waits for anybody
            e.sweep_levels(allow_unconverged=allow_unconverged)
        else:
            for lev in range(self.n_levels):
                if lev > 0:
                    self.halo(e.X)                       # rows updated by their owners in level lev-1
                e.sweep_level(lev, allow_unconverged=allow_unconverged)
        if late_theta:
            if self.comm is None:
                work.wait()
            else:
                torch.cuda.current_stream().wait_stream(self._side)
            e.param_solve()
        return e.last_info

    def param_loss(self, theta, alpha=0.0):
        """Pooled loss and gradient of the 'minimizer' parameter update (``VGMEngine.param_loss``) over all ranks:
        every rank evaluates its own ODE rows, one all-reduce of 1 + p doubles.  D.X must be current on every rank
        (``engine.compute_DX`` after a halo refresh)."""
        e = self.engine
        loss, grad = e.param_loss(theta, 0.0)
        buf = torch.as_tensor(np.concatenate([[loss], grad]), device=e.dev)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        out = buf.cpu().numpy()
        th = np.asarray(theta, dtype=np.float64).ravel()
        return float(out[0] + alpha * np.sum(th ** 2)), out[1:] + 2.0 * alpha * th

    def get_theta(self):
        return self.engine.get_theta()
