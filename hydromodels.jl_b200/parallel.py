"""Multi-GPU sharding of the forward run: one process per GPU, basins / ensemble
members / grid cells are independent (`/root/reference/src/bucket.jl:229-233` broadcasts over nodes
with no cross-node term), so every rank runs the stepper on a contiguous slice of the node axis
and there is NO data-path collective.  The only exchange is an optional gather of small per-node
results (objective values, final states) over `torch.distributed` (NCCL on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n_nodes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced `[n0, n1)`: the first `n_nodes % world` ranks get one extra node."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_nodes), int(world))
    n0 = rank * base + min(rank, rem)
    return n0, n0 + base + (1 if rank < rem else 0)


def shard_params(params: Dict, n0: int, n1: int, n_nodes: int) -> Dict:
    """Slice per-node parameter vectors (length `n_nodes`) to `[n0, n1)`; scalars, shared
    (shorter) vectors and neural-network parameters are replicated."""
    out = {"params": {}}
    for k, v in params["params"].items():
        a = np.asarray(v)
        out["params"][k] = a[n0:n1] if (a.ndim == 1 and a.size == n_nodes) else v
    if "nns" in params:
        out["nns"] = params["nns"]
    return out


def shard_htypes(htypes: np.ndarray, n0: int, n1: int) -> np.ndarray:
    """Local HRU-type vector: per-node types (`1:N`) are renumbered to `1:(n1-n0)`; shared types are kept."""
    h = np.asarray(htypes)
    if np.array_equal(h, np.arange(1, len(h) + 1)):
        return np.arange(1, n1 - n0 + 1)
    return h[n0:n1]


def shard_initstates(initstates, n0: int, n1: int, n_nodes: int, n_states: int):
    if initstates is None:
        return None
    if isinstance(initstates, dict):
        return {k: (np.asarray(v)[n0:n1] if np.ndim(v) == 1 else v) for k, v in initstates.items()}
    v = np.asarray(initstates).reshape(n_states, n_nodes)       # state-major, node-fastest
    return np.ascontiguousarray(v[:, n0:n1]).reshape(-1)


def gather_nodes(local: np.ndarray, n_nodes: int, group=None, device: Optional[str] = None) -> np.ndarray:
    """All-gather a per-node vector (`[n1-n0]` or `[k, n1-n0]`, node axis last) from every rank into the
    full `[..., n_nodes]` array, in rank order.  Uses the default process group's backend: NCCL
    (tensors staged on `device`) or gloo (CPU)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    loc = np.ascontiguousarray(local, dtype=np.float64)
    lead = loc.shape[:-1]
    sizes = [shard_range(n_nodes, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    pad = np.zeros(lead + (width,), dtype=np.float64)
    pad[..., :loc.shape[-1]] = loc
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t, group=group)
    full = np.empty(lead + (n_nodes,), dtype=np.float64)
    for r, (a, b) in enumerate(sizes):
        full[..., a:b] = outs[r].cpu().numpy()[..., :b - a]
    return full


# ---------------------------------------------------------------------------------------------------
# The communicator INSIDE the library (C-ABI `hb_comm_*`, csrc/hb_comm.cu): what a Julia host would use.
# ---------------------------------------------------------------------------------------------------
class LibComm:
    """`hb_comm_t` of this rank.  The 128-byte NCCL id of rank 0 reaches the other ranks through whatever the host has —
    here `torch.distributed` (any backend); a Julia host would use MPI or a shared file.  `halo_mode`: "nccl" = pack
    kernel + one grouped ncclSend/ncclRecv per neighbour + unpack kernel, all on the compute stream; "peer" = the pack
    kernel stores the blocks straight into the neighbours' staging over NVLink (CUDA IPC peer pointers) and publishes an
    epoch flag the neighbour's unpack kernel waits for."""

    def __init__(self, rank: int, world: int, group=None, halo_mode: str = "nccl"):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from .engine import check, lib
        if halo_mode not in ("nccl", "peer"):
            raise ValueError("halo_mode must be 'nccl' or 'peer'")
        self.rank, self.world, self.halo_mode = int(rank), int(world), halo_mode
        buf = (C.c_char * 128)()
        if rank == 0:
            check(lib().hb_comm_unique_id(buf, 128))
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        h = C.c_void_p()
        check(lib().hb_comm_init(raw, self.rank, self.world, C.byref(h)))
        self.handle = h
        self._peer_bytes = 0

    def enable_peer(self, max_block_bytes: int):
        from .engine import check, lib
        if self._peer_bytes == 0:
            check(lib().hb_halo_enable_peer(self.handle, int(max_block_bytes)))
            self._peer_bytes = int(max_block_bytes)
        elif max_block_bytes > self._peer_bytes:
            raise ValueError("peer staging was reserved for smaller blocks")

    def halo_exchange(self, plane_ptrs: Sequence[int], pitch: int, edges, stream: Optional[int] = None):
        """`edges`: [(peer, s_row0, s_col0, r_row0, r_col0, nrows, ncols, peer_slot)]; collective, asynchronous on `stream`."""
        import ctypes as C
        from .engine import HaloEdge, check, lib
        n = len(plane_ptrs)
        planes = (C.c_void_p * n)(*[int(x) for x in plane_ptrs])
        earr = (HaloEdge * max(1, len(edges)))(*[HaloEdge(*[int(v) for v in e]) for e in edges])
        mode = 1 if self.halo_mode == "peer" else 0
        if mode == 1:
            self.enable_peer(max([n * e[5] * e[6] * 8 for e in edges] or [8]))
        check(lib().hb_halo_exchange_device(self.handle, planes, n, int(pitch), earr, len(edges), mode, stream))

    def all_gather(self, send_ptr: int, recv_ptr: int, nbytes: int, stream: Optional[int] = None):
        from .engine import check, lib
        check(lib().hb_comm_all_gather(self.handle, send_ptr, recv_ptr, int(nbytes), stream))

    def status(self, stream: Optional[int] = None):
        from .engine import check, lib
        check(lib().hb_halo_status(self.handle, stream))

    def free(self):
        from .engine import lib
        if getattr(self, "handle", None):
            lib().hb_comm_free(self.handle)
            self.handle = None


# ---------------------------------------------------------------------------------------------------
# Sharded routing: contiguous node slices + ghost outflows exchanged every step
# ---------------------------------------------------------------------------------------------------
class RoutePartition:
    """Split a routed node set (global upstream CSR) into contiguous per-rank slices.

    For rank r owning `[n0, n1)`, upstream nodes outside the slice become GHOST slots appended to the
    local outflow buffer (`q[N_local + g]`), sorted by global id — so the ghosts owned by one peer form
    one contiguous run, which is exactly what that peer sends (its boundary outflows, in id order).
    Every rank derives all send/recv lists from the global topology; nothing is negotiated at run time.
    On a D8 grid sharded into row strips the ghosts of a strip are (part of) the boundary rows of its
    two neighbours — 8 bytes per boundary cell per step."""

    def __init__(self, up_ptr: np.ndarray, up_idx: np.ndarray, n_nodes: int, world: int):
        self.n_nodes, self.world = int(n_nodes), int(world)
        self.up_ptr = np.asarray(up_ptr, dtype=np.int64)
        self.up_idx = np.asarray(up_idx, dtype=np.int64)[: int(self.up_ptr[-1])]
        self.ranges = [shard_range(n_nodes, r, world) for r in range(world)]
        self._starts = np.array([a for a, _ in self.ranges] + [n_nodes], dtype=np.int64)
        self._ghosts = [self._ghost_ids(r) for r in range(world)]

    def _ghost_ids(self, rank: int) -> np.ndarray:
        n0, n1 = self.ranges[rank]
        li = self.up_idx[self.up_ptr[n0]:self.up_ptr[n1]]
        return np.unique(li[(li < n0) | (li >= n1)])

    def owner(self, ids: np.ndarray) -> np.ndarray:
        return np.searchsorted(self._starts, ids, side="right") - 1

    def local_csr(self, rank: int):
        """(up_ptr int32 [Nl+1], up_idx int32 with ghost slots >= Nl, n_ghost)."""
        n0, n1 = self.ranges[rank]
        ptr = (self.up_ptr[n0:n1 + 1] - self.up_ptr[n0]).astype(np.int32)
        li = self.up_idx[self.up_ptr[n0]:self.up_ptr[n1]]
        gh = self._ghosts[rank]
        remote = (li < n0) | (li >= n1)
        idx = np.where(remote, (n1 - n0) + np.searchsorted(gh, li), li - n0).astype(np.int32)
        if idx.size == 0:
            idx = np.zeros(1, dtype=np.int32)
        return ptr, np.ascontiguousarray(idx), int(gh.size)

    def recv_plan(self, rank: int):
        """[(peer, ghost offset, count)]: peer's values land in ghost slots [offset, offset+count)."""
        gh = self._ghosts[rank]
        own = self.owner(gh)
        plan = []
        for p in np.unique(own):
            sel = np.nonzero(own == p)[0]
            plan.append((int(p), int(sel[0]), int(sel.size)))
        return plan

    def send_plan(self, rank: int):
        """[(peer, local indices)]: the outflows of these local nodes, in this order, go to `peer`."""
        n0, n1 = self.ranges[rank]
        plan = []
        for p in range(self.world):
            if p == rank:
                continue
            gh = self._ghosts[p]
            mine = gh[(gh >= n0) & (gh < n1)]
            if mine.size:
                plan.append((p, (mine - n0).astype(np.int64)))
        return plan


def exchange_ghosts(q, n_local: int, send_plan, recv_plan, group=None):
    """Fill the ghost slots of the outflow buffer `q` (torch tensor `[n_local + n_ghost]`) with the
    peers' boundary outflows: grouped point-to-point sends/receives (NCCL on GPUs, gloo in tests)."""
    import torch
    import torch.distributed as dist

    if not send_plan and not recv_plan:
        return
    ops, keep = [], []
    for peer, idx in send_plan:
        buf = q[:n_local].index_select(0, idx).contiguous()
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, peer, group=group))
    for peer, off, cnt in recv_plan:
        ops.append(dist.P2POp(dist.irecv, q[n_local + off:n_local + off + cnt], peer, group=group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class ShardedRoutedRun:
    """Multi-GPU run of a model ending in a HydroRoute: this rank owns nodes `[n0, n1)`.

    1. the per-node stepper runs all T steps on the local nodes (no communication);
    2. T+1 routing launches (`hb_route_step_device`); between consecutive launches the freshly computed
       outflows of the boundary nodes are exchanged so every rank's ghost slots hold its upstream
       neighbours' outflows for the next step — the only data-path collective of the whole run.
    Records stay sharded (`[T][N_local][R]` per rank)."""

    def __init__(self, eng, n_nodes: int, T: int, params, *, config=None, initstates=None, group=None):
        import torch
        import torch.distributed as dist
        from .components import normalize_config
        from .engine import DeviceRun, RouteStage, upstream_csr

        if eng.route is None or eng.lumped is None:
            raise ValueError("ShardedRoutedRun needs per-node components followed by a HydroRoute")
        self.group = group
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        self.part = RoutePartition(*upstream_csr(eng.route.aggr, n_nodes), n_nodes, self.world)
        self.n0, self.n1 = self.part.ranges[self.rank]
        Nl = self.n1 - self.n0
        cfg = normalize_config(config)
        rows = eng.component.var_names
        s_name = eng.route.state_names[0]
        init = None if initstates is None else {k: v for k, v in initstates.items() if k != s_name}
        rng = (self.n0, self.n1)
        self.stepper = DeviceRun(eng, Nl, T, params, config=cfg, initstates=init, rows=rows, node_range=rng)
        ptr, idx, self.n_ghost = self.part.local_csr(self.rank)
        self.stage = RouteStage(eng, Nl, T, params, cfg, rows, None if initstates is None else initstates.get(s_name),
                                node_range=rng, csr=(ptr, idx))
        self.N, self.T, self.n_rows = Nl, int(T), len(rows)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.s = [torch.zeros(Nl, dtype=torch.float64, device=dev) for _ in range(2)]
        self.q = [torch.zeros(Nl + self.n_ghost, dtype=torch.float64, device=dev) for _ in range(2)]
        self.send = [(p, torch.from_numpy(ix).to(dev)) for p, ix in self.part.send_plan(self.rank)]
        self.recv = self.part.recv_plan(self.rank)
        self.init_state = None
        if self.stage.dinit is not None:
            self.init_state = torch.zeros(Nl, dtype=torch.float64, device=dev)
            host = self.stage.dinit.to_host((Nl,))
            self.init_state.copy_(torch.from_numpy(host))
        self.in_bytes, self.out_bytes = self.stepper.in_bytes, T * Nl * len(rows) * 8

    def launch(self, input_ptr: int, rec_ptr: int):
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        self.stepper.launch(input_ptr, rec_ptr, stream)
        s, q = self.s, self.q
        if self.init_state is not None:
            s[0].copy_(self.init_state)
        else:
            s[0].zero_()
        # prologue: outflows of step 0 from the initial state, then their ghosts
        self.stage.step(rec_ptr, -1, s[0].data_ptr(), s[1].data_ptr(), q[1].data_ptr(), q[0].data_ptr(), stream)
        exchange_ghosts(q[0], self.N, self.send, self.recv, self.group)
        for t in range(self.T):
            cur = t & 1
            self.stage.step(rec_ptr, t, s[cur].data_ptr(), s[cur ^ 1].data_ptr(), q[cur].data_ptr(), q[cur ^ 1].data_ptr(), stream)
            if t + 1 < self.T:
                exchange_ghosts(q[cur ^ 1], self.N, self.send, self.recv, self.group)


# ---------------------------------------------------------------------------------------------------
# dense D8 grids with the single-pass wave kernel: row strips + ghost rows, states exchanged every H steps
# ---------------------------------------------------------------------------------------------------
class GridStripPlan:
    """Row-strip decomposition of a `rows x cols` grid for temporal blocking ("H-step temporal
    blocking").  Rank k owns rows `[r0, r1)` and computes rows `[g0, g1)` = its strip plus up to `halo` ghost
    rows on either side.  D8 routing moves water one cell per step (src/route.jl:387-403), so after h steps
    without communication exactly the outermost h ghost rows on each side may differ from a whole-grid run —
    the owned rows stay exact for `halo` steps, after which the ghost rows' STATES (bucket states + river
    state) are refreshed from the neighbouring ranks.  One exchange of `(S+1) * halo * cols * 8` bytes per
    edge every `halo` steps replaces the per-step outflow exchange of `ShardedRoutedRun`."""

    def __init__(self, rows: int, cols: int, rank: int, world: int, halo: int):
        self.rows, self.cols, self.rank, self.world, self.halo = rows, cols, rank, world, halo
        self.r0, self.r1 = shard_range(rows, rank, world)
        if world > 1 and min(shard_range(rows, k, world)[1] - shard_range(rows, k, world)[0] for k in range(world)) < halo:
            raise ValueError(f"every strip needs at least halo = {halo} rows")
        self.g0, self.g1 = max(0, self.r0 - halo), min(rows, self.r1 + halo)
        self.n_local = (self.g1 - self.g0) * cols
        self.out_first = (self.r0 - self.g0) * cols
        self.out_count = (self.r1 - self.r0) * cols

    def node_range(self) -> Tuple[int, int]:
        return self.g0 * self.cols, self.g1 * self.cols

    def exchanges(self) -> List[Tuple[int, Tuple[int, int], Tuple[int, int]]]:
        """[(peer, (send_lo, send_hi), (recv_lo, recv_hi))] in LOCAL node indices: owned rows next to the edge go
        out, the ghost rows beyond the edge come in."""
        C_, H, out = self.cols, self.halo, []
        if self.rank > 0:                                   # upper edge: global rows r0-H..r0 are the peer's, r0..r0+H mine
            out.append((self.rank - 1, ((self.r0 - self.g0) * C_, (self.r0 - self.g0 + H) * C_), (0, (self.r0 - self.g0) * C_)))
        if self.rank < self.world - 1:
            out.append((self.rank + 1, ((self.r1 - H - self.g0) * C_, (self.r1 - self.g0) * C_),
                        ((self.r1 - self.g0) * C_, (self.g1 - self.g0) * C_)))
        return out


def strip_halo_edges(plan: GridStripPlan):
    """The strip's neighbours as `hb_halo_edge` tuples (planes are `[rows_local][cols]`, pitch = cols)."""
    C_, out = plan.cols, []
    for peer, (s0, s1), (q0, q1) in plan.exchanges():
        # the peer lists its lower-rank neighbour first: I am its LAST edge when I am above it, its FIRST when below
        peer_slot = (1 if peer > 0 else 0) if peer < plan.rank else 0
        out.append((peer, s0 // C_, 0, q0 // C_, 0, (s1 - s0) // C_, C_, peer_slot))
    return out


def exchange_strip_states(states: Sequence, plan: GridStripPlan, group=None):
    """Refresh the ghost rows of every state array (`[..., n_local]` torch tensors, node axis last and contiguous)
    from the neighbouring ranks (grouped send/recv: NCCL on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    ops, keep = [], []
    for peer, (s0, s1), (q0, q1) in plan.exchanges():
        for st in states:
            flat = st.reshape(-1, st.shape[-1])
            for row in range(flat.shape[0]):
                snd = flat[row, s0:s1].contiguous()
                keep.append(snd)
                ops.append(dist.P2POp(dist.isend, snd, peer, group))
                ops.append(dist.P2POp(dist.irecv, flat[row, q0:q1], peer, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class ShardedGridRun:
    """Multi-GPU run of `[buckets…, HydroRoute]` on a dense D8 grid with the wave kernel: every rank runs its row
    strip plus `halo` ghost rows for `halo` steps at a time (`GridDeviceRun.launch_chunk`), then the ghost rows'
    states are exchanged.  Records stay sharded: `[T][owned cells][R]` per rank, ghost rows are not stored.
    The forcing passed to `launch` covers the rank's COMPUTED rows: `[T][plan.n_local][V]`."""

    def __init__(self, eng, rows: int, cols: int, T: int, params, *, config=None, initstates=None, halo: int = 16, group=None,
                 stages: int = 0, max_registers: int = 0, block_threads: int = 0, comm: Optional["LibComm"] = None):
        """`comm`: a `LibComm` — the ghost states then travel through the library's own exchange (`hb_halo_exchange_device`:
        NCCL or NVLink peer stores) on the compute stream instead of `torch.distributed` point-to-point calls."""
        import torch
        import torch.distributed as dist
        from .engine import GridDeviceRun

        self.group, self.comm = group, comm
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        self.plan = GridStripPlan(rows, cols, self.rank, self.world, halo)
        pl = self.plan
        self.run = GridDeviceRun(eng, pl.n_local, T, params, config=config, initstates=initstates, variant="wave",
                                 steps_per_chunk=halo, block_w=block_threads, stages=stages, max_registers=max_registers,
                                 row_range=(pl.g0, pl.g1))
        self.T, self.N, self.n_rows = int(T), pl.out_count, self.run.n_rows
        S = self.run.n_states
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bucket = [torch.zeros((max(S, 1), pl.n_local), dtype=torch.float64, device=dev) for _ in range(2)]
        self.river = [torch.zeros(pl.n_local, dtype=torch.float64, device=dev) for _ in range(2)]
        self.init_bucket = self.init_river = None
        if "init" in self.run.bufs:
            self.init_bucket = torch.from_numpy(self.run.bufs["init"].to_host((max(S, 1), pl.n_local))).to(dev)
        if "rinit" in self.run.bufs:
            self.init_river = torch.from_numpy(self.run.bufs["rinit"].to_host((pl.n_local,))).to(dev)
        self.n_inputs = len(self.run.ck.program.stepper.input_names)
        self.in_bytes, self.out_bytes = T * pl.n_local * self.n_inputs * 8, T * pl.out_count * self.n_rows * 8

    def launch(self, input_ptr: int, rec_ptr: int):
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        pl, H = self.plan, self.plan.halo
        cur = 0
        if self.init_bucket is not None:
            self.bucket[0].copy_(self.init_bucket)
        else:
            self.bucket[0].zero_()
        if self.init_river is not None:
            self.river[0].copy_(self.init_river)
        else:
            self.river[0].zero_()
        for t0 in range(0, self.T, H):
            ns = min(H, self.T - t0)
            self.run.launch_chunk(input_ptr + t0 * pl.n_local * self.n_inputs * 8, rec_ptr + t0 * pl.out_count * self.n_rows * 8, ns,
                                  self.bucket[cur].data_ptr(), self.river[cur].data_ptr(),
                                  self.bucket[cur ^ 1].data_ptr(), self.river[cur ^ 1].data_ptr(),
                                  pl.out_first, pl.out_count, stream)
            cur ^= 1
            if t0 + ns < self.T and self.world > 1:
                if self.comm is not None:
                    S = self.run.n_states
                    planes = [self.bucket[cur].data_ptr() + s * pl.n_local * 8 for s in range(S)] + [self.river[cur].data_ptr()]
                    self.comm.halo_exchange(planes, pl.cols, strip_halo_edges(pl), stream)
                else:
                    exchange_strip_states([self.bucket[cur], self.river[cur]], pl, self.group)

    def check(self) -> int:
        if self.comm is not None:
            self.comm.status()
        return self.run.check()


# ---------------------------------------------------------------------------------------------------
# column panels: the same temporal blocking with ghost COLUMNS.  The wave kernel runs a tall, narrow sub-grid markedly
# faster than a short, wide one (4096 x 544: 4.6 ms vs 544 x 4096: 6.7 ms per 60-step pass on a B200 — fewer resident
# rows are tied up in the T_c-deep dependency staircase), so a wide grid is better cut into panels than into strips.
# ---------------------------------------------------------------------------------------------------
class GridPanelPlan:
    """Rank k owns columns `[c0, c1)` of a `rows x cols` grid and computes columns `[g0, g1)` = its panel plus `halo`
    ghost columns on either side (clipped at the grid border).  After h steps without communication exactly the
    outermost h ghost columns may differ from a whole-grid run, so the ghost columns' states are refreshed every `halo`
    steps.  The local sub-grid is compact (`rows x (g1 - g0)`, row-major); `idx` maps its cells to global node ids."""

    def __init__(self, rows: int, cols: int, rank: int, world: int, halo: int):
        self.rows, self.cols, self.rank, self.world, self.halo = rows, cols, rank, world, halo
        self.c0, self.c1 = shard_range(cols, rank, world)
        if world > 1 and min(shard_range(cols, k, world)[1] - shard_range(cols, k, world)[0] for k in range(world)) < halo:
            raise ValueError(f"every panel needs at least halo = {halo} columns")
        self.g0, self.g1 = max(0, self.c0 - halo), min(cols, self.c1 + halo)
        self.local_cols = self.g1 - self.g0
        self.n_local = rows * self.local_cols
        self.out_window = (0, rows, self.c0 - self.g0, self.c1 - self.c0)
        self.out_count = rows * (self.c1 - self.c0)
        self.idx = (np.arange(rows, dtype=np.int64)[:, None] * cols + np.arange(self.g0, self.g1, dtype=np.int64)[None, :]).reshape(-1)

    def exchanges(self) -> List[Tuple[int, Tuple[int, int], Tuple[int, int]]]:
        """[(peer, (send_col_lo, send_col_hi), (recv_col_lo, recv_col_hi))] in LOCAL column indices."""
        H, out = self.halo, []
        if self.rank > 0:
            out.append((self.rank - 1, (self.c0 - self.g0, self.c0 - self.g0 + H), (0, self.c0 - self.g0)))
        if self.rank < self.world - 1:
            out.append((self.rank + 1, (self.c1 - H - self.g0, self.c1 - self.g0), (self.c1 - self.g0, self.g1 - self.g0)))
        return out


def panel_halo_edges(plan: GridPanelPlan):
    """The panel's neighbours as `hb_halo_edge` tuples (planes are `[rows][local_cols]`, pitch = local_cols)."""
    out = []
    for peer, (s0, s1), (q0, q1) in plan.exchanges():
        # edge lists are [left neighbour (if any), right neighbour (if any)]: seen from my LEFT neighbour I am its right
        # edge (slot 1, or 0 when it has no left neighbour); seen from my RIGHT neighbour I am its left edge (slot 0)
        peer_slot = (1 if peer > 0 else 0) if peer < plan.rank else 0
        out.append((peer, 0, s0, 0, q0, plan.rows, s1 - s0, peer_slot))
    return out


def exchange_panel_states(states: Sequence, plan: GridPanelPlan, group=None):
    """Refresh the ghost columns of every state array (`[..., rows * local_cols]` torch tensors, node axis last)."""
    import torch.distributed as dist
    ops, pending = [], []
    for peer, (s0, s1), (q0, q1) in plan.exchanges():
        for st in states:
            grid = st.reshape(-1, plan.rows, plan.local_cols)
            snd = grid[:, :, s0:s1].contiguous()
            rcv = grid.new_empty((grid.shape[0], plan.rows, q1 - q0))
            pending.append((grid, q0, q1, rcv, snd))
            ops.append(dist.P2POp(dist.isend, snd, peer, group))
            ops.append(dist.P2POp(dist.irecv, rcv, peer, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for grid, q0, q1, rcv, _ in pending:
        grid[:, :, q0:q1] = rcv


def submodel_for_cells(model, idx: np.ndarray, flw_local: np.ndarray):
    """The model restricted to the cells `idx` (global node ids, listed row-major over the local sub-grid `flw_local`):
    every component keeps its equations and gets `htypes[idx]` — parameter vectors stay GLOBAL and are gathered through
    those htypes — and the route aggregates over the local D8 grid."""
    import copy
    from .components import GridAggregation, HydroModel, HydroRoute
    R, Cc = flw_local.shape
    rr, cc = np.divmod(np.arange(R * Cc), Cc)
    comps = []
    for c in model.components:
        c2 = copy.copy(c)
        if c.htypes is None:
            raise ValueError("a distributed model needs htypes on every component")
        c2.htypes = np.asarray(c.htypes)[idx]
        if isinstance(c, HydroRoute):
            c2.aggr = GridAggregation(np.asarray(flw_local), np.stack([rr + 1, cc + 1], axis=1))
        comps.append(c2)
    return HydroModel(components=comps, name=model.name, input_names=model.input_names)


class ShardedPanelRun:
    """Multi-GPU run of `[buckets…, HydroRoute]` on a dense D8 grid cut into COLUMN panels (see `GridPanelPlan`).  Same
    contract as `ShardedGridRun`: forcing `[T][plan.n_local][V]` for the computed cells (local row-major order), records
    `[T][rows][owned columns][R]` for the owned cells."""

    def __init__(self, eng, rows: int, cols: int, T: int, params, *, config=None, initstates=None, halo: int = 32, group=None,
                 stages: int = 0, max_registers: int = 0, block_threads: int = 0, steps_per_launch: int = 0,
                 comm: Optional["LibComm"] = None):
        import torch
        import torch.distributed as dist
        from .engine import B200Model, GridDeviceRun

        self.group, self.comm = group, comm
        self.exchange_events = None          # [(start, end)] CUDA events around the exchanges of the last launch (set by `time_exchange`)
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
        self.plan = pl = GridPanelPlan(rows, cols, self.rank, self.world, halo)
        flw = np.asarray(eng.route.aggr.flwdir)[:, pl.g0:pl.g1]
        self.sub = submodel_for_cells(eng.component, pl.idx, flw)
        self.eng = B200Model(self.sub, fast=eng.fast)
        init = None if initstates is None else {k: np.broadcast_to(np.asarray(v, dtype=np.float64), (rows * cols,))[pl.idx] for k, v in initstates.items()}
        self.run = GridDeviceRun(self.eng, pl.n_local, T, params, config=config, initstates=init, variant="wave",
                                 steps_per_chunk=steps_per_launch, block_w=block_threads, stages=stages, max_registers=max_registers)
        self.T, self.N, self.n_rows = int(T), pl.out_count, self.run.n_rows
        S = self.run.n_states
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bucket = [torch.zeros((max(S, 1), pl.n_local), dtype=torch.float64, device=dev) for _ in range(2)]
        self.river = [torch.zeros(pl.n_local, dtype=torch.float64, device=dev) for _ in range(2)]
        self.init_bucket = self.init_river = None
        if "init" in self.run.bufs:
            self.init_bucket = torch.from_numpy(self.run.bufs["init"].to_host((max(S, 1), pl.n_local))).to(dev)
        if "rinit" in self.run.bufs:
            self.init_river = torch.from_numpy(self.run.bufs["rinit"].to_host((pl.n_local,))).to(dev)
        self.n_inputs = len(self.run.ck.program.stepper.input_names)
        self.in_bytes, self.out_bytes = T * pl.n_local * self.n_inputs * 8, T * pl.out_count * self.n_rows * 8

    def launch(self, input_ptr: int, rec_ptr: int):
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        pl, H = self.plan, self.plan.halo
        cur = 0
        self.bucket[0].copy_(self.init_bucket) if self.init_bucket is not None else self.bucket[0].zero_()
        self.river[0].copy_(self.init_river) if self.init_river is not None else self.river[0].zero_()
        for t0 in range(0, self.T, H):
            ns = min(H, self.T - t0)
            self.run.launch_chunk(input_ptr + t0 * pl.n_local * self.n_inputs * 8, rec_ptr + t0 * pl.out_count * self.n_rows * 8, ns,
                                  self.bucket[cur].data_ptr(), self.river[cur].data_ptr(),
                                  self.bucket[cur ^ 1].data_ptr(), self.river[cur ^ 1].data_ptr(), stream=stream,
                                  out_window=pl.out_window)
            cur ^= 1
            if t0 + ns < self.T and self.world > 1:
                ev = None
                if self.exchange_events is not None:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                if self.comm is not None:
                    S = self.run.n_states
                    planes = [self.bucket[cur].data_ptr() + s * pl.n_local * 8 for s in range(S)] + [self.river[cur].data_ptr()]
                    self.comm.halo_exchange(planes, pl.local_cols, panel_halo_edges(pl), stream)
                else:
                    exchange_panel_states([self.bucket[cur], self.river[cur]], pl, self.group)
                if ev is not None:
                    ev[1].record()
                    self.exchange_events.append(ev)

    def halo_bytes_per_exchange(self) -> int:
        """Bytes this rank SENDS per exchange (states of `halo` owned columns per neighbour)."""
        return sum(e[5] * e[6] for e in panel_halo_edges(self.plan)) * (self.run.n_states + 1) * 8

    def check(self) -> int:
        if self.comm is not None:
            self.comm.status()
        return self.run.check()


class PanelledGridRun:
    """ONE GPU, a wide dense D8 grid: the same column panels as `ShardedPanelRun`, run one after the other on the same device
    and IN PLACE on the whole grid's forcing and record arrays (pitched addressing of the wave kernel), ghost-column states
    copied between the panels every `halo` steps.  A 4096-column grid as 4 panels of 1024 + 2 x 32 columns runs ~1.3x faster
    than as one sub-grid (the wave kernel prefers tall, narrow sub-grids)."""

    def __init__(self, eng, rows: int, cols: int, T: int, params, *, config=None, initstates=None, panels: int = 0, halo: int = 32,
                 steps_per_launch: int = 0, stages: int = 0, max_registers: int = 0, block_threads: int = 0):
        import torch
        from .engine import B200Model, GridDeviceRun
        K = int(panels) if panels else max(1, int(round(cols / 1024)))
        while K > 1 and cols // K < (halo if panels else max(halo, 64)):     # every panel must own at least `halo` columns
            K -= 1
        self.rows, self.cols, self.T, self.K = rows, cols, int(T), K
        self.N = rows * cols
        self.plans = [GridPanelPlan(rows, cols, k, K, halo) for k in range(K)]
        flw = np.asarray(eng.route.aggr.flwdir)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.runs, self.bucket, self.river, self.init = [], [], [], []
        for pl in self.plans:
            sub = submodel_for_cells(eng.component, pl.idx, flw[:, pl.g0:pl.g1])
            e = B200Model(sub, fast=eng.fast)
            li = None if initstates is None else {k: np.broadcast_to(np.asarray(v, dtype=np.float64), (self.N,))[pl.idx] for k, v in initstates.items()}
            run = GridDeviceRun(e, pl.n_local, T, params, config=config, initstates=li, variant="wave", steps_per_chunk=steps_per_launch,
                                block_w=block_threads, stages=stages, max_registers=max_registers)
            S = run.n_states
            self.runs.append(run)
            self.bucket.append([torch.zeros((max(S, 1), pl.n_local), dtype=torch.float64, device=dev) for _ in range(2)])
            self.river.append([torch.zeros(pl.n_local, dtype=torch.float64, device=dev) for _ in range(2)])
            ib = torch.from_numpy(run.bufs["init"].to_host((max(S, 1), pl.n_local))).to(dev) if "init" in run.bufs else None
            ir = torch.from_numpy(run.bufs["rinit"].to_host((pl.n_local,))).to(dev) if "rinit" in run.bufs else None
            self.init.append((ib, ir))
        self.n_rows = self.runs[0].n_rows
        self.n_inputs = len(self.runs[0].ck.program.stepper.input_names)
        self.ck = self.runs[0].ck
        self.in_bytes, self.out_bytes = T * self.N * self.n_inputs * 8, T * self.N * self.n_rows * 8

    def _exchange(self, cur: int):
        H = self.plans[0].halo
        for k in range(self.K - 1):
            a, b = self.plans[k], self.plans[k + 1]
            for sa, sb in ((self.bucket[k][cur], self.bucket[k + 1][cur]), (self.river[k][cur], self.river[k + 1][cur])):
                ga = sa.reshape(-1, a.rows, a.local_cols)
                gb = sb.reshape(-1, b.rows, b.local_cols)
                # a's right ghost columns <- b's first owned columns; b's left ghost columns <- a's last owned columns
                ga[:, :, a.c1 - a.g0:] = gb[:, :, b.c0 - b.g0:b.c0 - b.g0 + (a.g1 - a.c1)]
                gb[:, :, :b.c0 - b.g0] = ga[:, :, a.c1 - a.g0 - (b.c0 - b.g0):a.c1 - a.g0]

    def launch(self, input_ptr: int, rec_ptr: int, stream: Optional[int] = None):
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        H = self.plans[0].halo
        for k in range(self.K):
            ib, ir = self.init[k]
            self.bucket[k][0].copy_(ib) if ib is not None else self.bucket[k][0].zero_()
            self.river[k][0].copy_(ir) if ir is not None else self.river[k][0].zero_()
        cur = 0
        for t0 in range(0, self.T, H):
            ns = min(H, self.T - t0)
            for k, (pl, run) in enumerate(zip(self.plans, self.runs)):
                run.launch_chunk(input_ptr + t0 * self.N * self.n_inputs * 8, rec_ptr + t0 * self.N * self.n_rows * 8, ns,
                                 self.bucket[k][cur].data_ptr(), self.river[k][cur].data_ptr(),
                                 self.bucket[k][cur ^ 1].data_ptr(), self.river[k][cur ^ 1].data_ptr(), stream=st,
                                 out_window=pl.out_window, in_pitched=(self.N, self.cols, pl.g0), out_pitched=(self.N, self.cols, pl.c0))
            cur ^= 1
            if t0 + ns < self.T and self.K > 1:
                self._exchange(cur)

    def check(self) -> int:
        return min(r.check() for r in self.runs)
