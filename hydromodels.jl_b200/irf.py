"""IRF river routing: host-side mirrors of the reference's
`RouteIRF`, `build_irf_kernels`, `SparseRouteKernel`, `aggregate_route_kernel`, `SparseRouteConvolution`
(`/root/reference/src/route.jl:417-610`).

Kernel construction (`aggregate_route_kernel`: topologically ordered composition of truncated node kernels,
O(nnz·H²), done once per network) stays on the host, exactly like the unit-hydrograph ordinates; the hot
operator — `SparseRouteConvolution(input)`: nnz × T × H multiply-adds — runs on the GPU
(`hb_sparse_route_conv_ordered_device`, `csrc/aot_kernels.cu`).  No CPU execution of the convolution exists here.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def balanced_destination_order(counts, window: int = 4096) -> np.ndarray:
    """Order in which `hb_sparse_conv_tiled` blocks (32 destinations each) should take the destinations: sorted by
    decreasing row count WITHIN windows of `window` consecutive nodes — equal counts inside a block (a block runs for as
    many rows as its longest destination) while a block's sources stay close in node order (river networks are local in
    node order; a global sort spreads them over the whole range and loses the L2 reuse between neighbouring blocks)."""
    counts = np.asarray(counts)
    out = np.empty(counts.size, dtype=np.int32)
    for a in range(0, counts.size, window):
        b = min(a + window, counts.size)
        out[a:b] = a + np.argsort(-counts[a:b], kind="stable")
    return out


def synthetic_river_tree(n_nodes: int, seed: int = 0, p_children=(0.3, 0.45, 0.25)) -> np.ndarray:
    """A river-like tree for benchmarks and tests: `down[i]` = the node that node `i` drains into (-1 for the outlet, node
    `n_nodes - 1`).  Built breadth-first from the outlet, every node receiving 1, 2 or 3 tributaries (reaches, confluences)
    with probabilities `p_children`, and numbered so that every node's id is smaller than its downstream neighbour's
    (ids are a topological order, headwaters first).  Mean depth ~ log2(n): 65 536 nodes give ~1 M (destination, source)
    kernel rows."""
    n = int(n_nodes)
    rng = np.random.default_rng(seed)
    c = rng.choice(np.array([1, 2, 3]), size=n, p=np.asarray(p_children, dtype=float))
    first_child = 1 + np.cumsum(c) - c                       # breadth-first index of a node's first tributary
    j = np.arange(1, n)
    parent = np.searchsorted(first_child, j, side="right") - 1
    down = np.full(n, -1, dtype=np.int64)
    down[n - 1 - j] = n - 1 - parent
    return down


class RouteIRF:
    """`RouteIRF(params, kernel_func)` (`src/route.jl:417-432`): `kernel_func(params, delta_t, horizon)`
    returns the node's impulse response; the call pads / truncates it to `horizon`."""

    def __init__(self, params: Sequence[str], kernel_func: Callable, name: Optional[str] = None):
        self.param_names = [str(p) for p in params]
        self.kernel_func = kernel_func
        self.name = name or "##route_irf"

    def __call__(self, params, *, delta_t: float = 1.0, horizon: int = 32) -> np.ndarray:
        if horizon <= 0:
            raise ValueError(f"horizon must be positive, got {horizon}")
        k = self.kernel_func(np.asarray(params, dtype=np.float64), float(delta_t), int(horizon))
        return _pad_route_kernel(k, int(horizon))


def _pad_route_kernel(kernel, horizon: int) -> np.ndarray:
    """`_pad_route_kernel` (`src/route.jl:497-504`)."""
    out = np.zeros(horizon)
    k = np.asarray(kernel, dtype=np.float64).reshape(-1)
    upper = min(len(k), horizon)
    out[:upper] = k[:upper]
    return out


def build_irf_kernels(irf: RouteIRF, params, *, delta_t: float = 1.0, horizon: int = 32) -> List[np.ndarray]:
    """`build_irf_kernels` (`src/route.jl:440-456`): `params` is `[nodes, n_params]` or a list of per-parameter
    vectors."""
    if isinstance(params, np.ndarray) and params.ndim == 2:
        return [irf(params[n], delta_t=delta_t, horizon=horizon) for n in range(params.shape[0])]
    cols = [np.asarray(p, dtype=np.float64) for p in params]
    nn = len(cols[0]) if cols else 0
    return [irf([c[n] for c in cols], delta_t=delta_t, horizon=horizon) for n in range(nn)]


class SparseRouteKernel:
    """`SparseRouteKernel(indices[nnz,2] (dest, source; 1-based), weights[nnz,H], n_out, n_in)`
    (`src/route.jl:458-480`)."""

    def __init__(self, indices, weights, n_out: int, n_in: int):
        indices = np.asarray(indices, dtype=np.int64)
        weights = np.asarray(weights, dtype=np.float64)
        if indices.ndim != 2 or indices.shape[1] != 2:
            raise ValueError("indices must have shape [nnz, 2].")
        if indices.shape[0] != weights.shape[0]:
            raise ValueError("indices and weights must agree on nnz.")
        if weights.ndim != 2 or weights.shape[1] <= 0:
            raise ValueError("weights must contain at least one lag.")
        self.indices, self.weights = indices, weights
        self.n_out, self.n_in, self.horizon = int(n_out), int(n_in), int(weights.shape[1])


def _edge_list(adjacency, nn: int):
    """(dest, upstream, weight) arrays, 0-based, of the non-zero entries of `adjacency[dest, upstream]` — a dense array, a
    scipy.sparse matrix, or an explicit `(dest, upstream[, weight])` tuple of 0-based arrays (networks too large for a dense
    matrix)."""
    if isinstance(adjacency, tuple):
        d, u = np.asarray(adjacency[0], dtype=np.int64), np.asarray(adjacency[1], dtype=np.int64)
        w = np.asarray(adjacency[2], dtype=np.float64) if len(adjacency) > 2 else np.ones(d.size)
        if d.size and (min(d.min(), u.min()) < 0 or max(d.max(), u.max()) >= nn):
            raise ValueError("adjacency and node_kernels size mismatch.")
    elif hasattr(adjacency, "tocoo"):
        if adjacency.shape != (nn, nn):
            raise ValueError("adjacency and node_kernels size mismatch.")
        coo = adjacency.tocoo()
        d, u, w = coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float64)
    else:
        A = np.asarray(adjacency, dtype=np.float64)
        if A.shape != (nn, nn):
            raise ValueError("adjacency and node_kernels size mismatch.")
        d, u = np.nonzero(A)
        w = A[d, u]
    keep = w != 0.0
    return d[keep], u[keep], w[keep]


def topological_order(adjacency, n_nodes: Optional[int] = None) -> List[int]:
    """Kahn's algorithm, 1-based node ids, FIFO over the initially free nodes in id order and then in the order their last
    upstream neighbour was processed (`_topological_order`, `src/route.jl:147-169`); `adjacency[node, upstream] != 0` means
    `upstream` drains into `node`.  Works on the sparse edge list: O(nodes + edges)."""
    if n_nodes is None:
        n_nodes = adjacency.shape[0]
    d, u, _ = _edge_list(adjacency, n_nodes)
    n = int(n_nodes)
    indeg = np.bincount(d, minlength=n).astype(np.int64)
    by_up = np.argsort(u, kind="stable")
    dn, ptr = d[by_up], np.concatenate([[0], np.cumsum(np.bincount(u, minlength=n))])
    order, ready, head = [], [int(i) for i in np.nonzero(indeg == 0)[0]], 0
    while head < len(ready):
        v = ready[head]
        head += 1
        order.append(v + 1)
        for t in np.sort(dn[ptr[v]:ptr[v + 1]]):          # the reference scans `downstream in 1:nn`
            indeg[t] -= 1
            if indeg[t] == 0:
                ready.append(int(t))
    if len(order) != n:
        raise ValueError("adjacency must define an acyclic upstream-to-downstream network.")
    return order


class CompositionPlan:
    """Row structure of `aggregate_route_kernel` for one network: which (destination, source) rows exist, in which LEVEL
    each is computed, and which finished upstream rows (with which edge weights, in which order) it accumulates."""

    def __init__(self, n_nodes, level_row_ptr, row_dest, row_src, contrib_ptr, contrib_row, contrib_ew, final_perm):
        self.n_nodes = n_nodes
        self.level_row_ptr = level_row_ptr          # [levels + 1] offsets into the plan's row order (level by level)
        self.row_dest, self.row_src = row_dest, row_src
        self.contrib_ptr, self.contrib_row, self.contrib_ew = contrib_ptr, contrib_row, contrib_ew
        self.final_perm = final_perm                # final row (dest ascending, source ascending) -> plan row

    @property
    def nnz(self) -> int:
        return int(self.row_dest.size)


def plan_route_composition(adjacency, n_nodes: int, topo_order=None) -> CompositionPlan:
    """Symbolic phase of `aggregate_route_kernel` (`/root/reference/src/route.jl:536-560`), vectorised per level.

    The reference visits the nodes in `order` (default `1:nn`) and lets a node see only what its upstream neighbours hold
    AT THAT TIME: an upstream neighbour that comes later in `order` contributes nothing.  So the edges that count are those
    whose upstream end precedes the node in `order`; on them, a node's level is one more than its deepest upstream
    neighbour's.  A row (node, source) accumulates one candidate per upstream neighbour that knows `source`, in ascending
    upstream id (the reference's `for upstream in 1:nn`)."""
    nn = int(n_nodes)
    d, u, w = _edge_list(adjacency, nn)
    pos = np.arange(nn) if topo_order is None else np.empty(nn, dtype=np.int64)
    if topo_order is not None:
        order0 = np.asarray(topo_order, dtype=np.int64) - 1
        if order0.size != nn or not np.array_equal(np.sort(order0), np.arange(nn)):
            raise ValueError("topo_order must be a permutation of 1:nn")
        pos[order0] = np.arange(nn)
    keep = pos[u] < pos[d]
    d, u, w = d[keep], u[keep], w[keep]
    # ---- levels (Kahn, frontier at a time) ----
    level = np.zeros(nn, dtype=np.int64)
    indeg = np.bincount(d, minlength=nn)
    by_up = np.argsort(u, kind="stable")
    up_ptr = np.concatenate([[0], np.cumsum(np.bincount(u, minlength=nn))])
    frontier = np.nonzero(indeg == 0)[0]
    lv = 0
    done = 0
    while frontier.size:
        level[frontier] = lv
        done += frontier.size
        cnt = up_ptr[frontier + 1] - up_ptr[frontier]
        if cnt.sum() == 0:
            break
        idx = np.repeat(up_ptr[frontier], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        targets = d[by_up[idx]]
        dec = np.bincount(targets, minlength=nn)
        indeg = indeg - dec
        frontier = np.nonzero((indeg == 0) & (dec > 0))[0]
        lv += 1
    if done != nn:
        raise ValueError("adjacency must define an acyclic upstream-to-downstream network.")
    n_levels = int(level.max()) + 1 if nn else 0
    # edges grouped by the level of their destination
    e_order = np.lexsort((u, d, level[d]))
    d, u, w = d[e_order], u[e_order], w[e_order]
    e_ptr = np.concatenate([[0], np.cumsum(np.bincount(level[d], minlength=n_levels))]) if d.size else np.zeros(n_levels + 1, dtype=np.int64)
    nodes_by_level = np.argsort(level, kind="stable")
    n_ptr = np.concatenate([[0], np.cumsum(np.bincount(level, minlength=n_levels))]) if nn else np.zeros(1, dtype=np.int64)
    start = np.zeros(nn, dtype=np.int64)            # first plan row of a node
    count = np.zeros(nn, dtype=np.int64)
    rows_d, rows_s, c_ptrs, c_rows, c_ews, lvl_ptr = [], [], [], [], [], [0]
    src_all = np.zeros(0, dtype=np.int64)           # plan row -> source node
    total_rows, total_contrib = 0, 0
    for L in range(n_levels):
        nodes = nodes_by_level[n_ptr[L]:n_ptr[L + 1]]
        ed, eu, ew = d[e_ptr[L]:e_ptr[L + 1]], u[e_ptr[L]:e_ptr[L + 1]], w[e_ptr[L]:e_ptr[L + 1]]
        cnt = count[eu]
        tot = int(cnt.sum())
        rep = np.repeat(np.arange(eu.size), cnt)
        within = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        up_row = start[eu][rep] + within                                   # finished rows of the upstream neighbours
        cand_dest = np.concatenate([nodes, ed[rep]])
        cand_src = np.concatenate([nodes, src_all[up_row]])
        cand_up = np.concatenate([np.full(nodes.size, -1, dtype=np.int64), eu[rep]])
        cand_row = np.concatenate([np.full(nodes.size, -1, dtype=np.int64), up_row])
        cand_ew = np.concatenate([np.zeros(nodes.size), ew[rep]])
        o = np.lexsort((cand_up, cand_src, cand_dest))
        cand_dest, cand_src, cand_up, cand_row, cand_ew = cand_dest[o], cand_src[o], cand_up[o], cand_row[o], cand_ew[o]
        first = np.ones(cand_dest.size, dtype=bool)
        first[1:] = (cand_dest[1:] != cand_dest[:-1]) | (cand_src[1:] != cand_src[:-1])
        row_of = np.cumsum(first) - 1                                      # level-local row of every candidate
        n_rows = int(first.sum())
        rd, rs = cand_dest[first], cand_src[first]
        real = cand_up >= 0                                                # the node's own entry carries no contribution
        c_count = np.bincount(row_of[real], minlength=n_rows)
        rows_d.append(rd)
        rows_s.append(rs)
        c_ptrs.append(total_contrib + np.cumsum(c_count) - c_count)
        c_rows.append(cand_row[real])
        c_ews.append(cand_ew[real])
        # rows are sorted by (dest, source): a node's rows are contiguous
        per_node = np.bincount(rd, minlength=nn)
        firsts = np.nonzero(np.concatenate([[True], rd[1:] != rd[:-1]]))[0] if n_rows else np.zeros(0, dtype=np.int64)
        start[rd[firsts]] = total_rows + firsts
        count[nodes] = per_node[nodes]
        src_all = np.concatenate([src_all, rs])
        total_rows += n_rows
        total_contrib += int(real.sum())
        lvl_ptr.append(total_rows)
    row_dest = np.concatenate(rows_d) if rows_d else np.zeros(0, dtype=np.int64)
    row_src = np.concatenate(rows_s) if rows_s else np.zeros(0, dtype=np.int64)
    contrib_ptr = np.concatenate(c_ptrs + [np.array([total_contrib])]) if rows_d else np.zeros(1, dtype=np.int64)
    contrib_row = np.concatenate(c_rows) if c_rows else np.zeros(0, dtype=np.int64)
    contrib_ew = np.concatenate(c_ews) if c_ews else np.zeros(0)
    final_perm = np.lexsort((row_src, row_dest))                           # `sort!(collect(keys(pairwise[dest])))`, dest ascending
    return CompositionPlan(nn, np.asarray(lvl_ptr, dtype=np.int64), row_dest.astype(np.int32), row_src.astype(np.int64),
                           contrib_ptr.astype(np.int64), contrib_row.astype(np.int64), contrib_ew.astype(np.float64), final_perm.astype(np.int64))


class DeviceRouteKernel:
    """An aggregated kernel whose weights stay on the device: `w_conv` is `[H][nnz]` in CSR-by-destination order (what
    `hb_sparse_route_conv_device` reads); `.weights` downloads the reference's `[nnz, H]` matrix on demand."""

    def __init__(self, indices, n_out, n_in, horizon, w_rows_buf, w_conv_buf):
        self.indices, self.n_out, self.n_in, self.horizon = indices, int(n_out), int(n_in), int(horizon)
        self._w_rows, self.w_conv = w_rows_buf, w_conv_buf
        self._weights = None

    @property
    def weights(self) -> np.ndarray:
        if self._weights is None:
            self._weights = self._w_rows.to_host((self.indices.shape[0], self.horizon))
        return self._weights


def aggregate_route_kernel(adjacency, node_kernels: Sequence, *, topo_order=None, horizon: int, keep_on_device: bool = False):
    """`aggregate_route_kernel` (`/root/reference/src/route.jl:523-577`): source -> destination kernels by composing the node
    kernels down the network, truncated to `horizon`.  The row structure is planned on the host from the sparse edge list
    (`plan_route_composition`); the arithmetic — one truncated `H x H` convolution per (row, upstream neighbour) — runs on
    the device level by level (`hb_irf_compose_device`), with the reference's operations in the reference's order: the
    weights are bit-identical to a literal evaluation of the reference's loops (tests: oracle.aggregate_route_kernel).
    `adjacency`: dense array, scipy.sparse matrix, or `(dest, upstream[, weight])` 0-based edge arrays; `None` pairs every
    node with itself (`:533-537`)."""
    from .engine import DeviceBuffer, check, lib
    nn = len(node_kernels)
    H = int(horizon)
    if H <= 0:
        raise ValueError(f"horizon must be positive, got {horizon}")
    padded = np.zeros((max(nn, 1), H))
    for i, k in enumerate(node_kernels):
        padded[i] = _pad_route_kernel(k, H)
    if adjacency is None:
        idx = np.stack([np.arange(1, nn + 1), np.arange(1, nn + 1)], axis=1)
        return SparseRouteKernel(idx, padded[:nn], nn, nn)
    plan = plan_route_composition(adjacency, nn, topo_order)
    nnz = plan.nnz
    indices = np.stack([plan.row_dest[plan.final_perm].astype(np.int64) + 1, plan.row_src[plan.final_perm] + 1], axis=1)
    if nnz == 0:
        return SparseRouteKernel(indices.reshape(0, 2), np.zeros((0, H)), nn, nn)
    bufs = [DeviceBuffer.from_host(a) for a in (plan.row_dest, plan.contrib_ptr, plan.contrib_row if plan.contrib_row.size else np.zeros(1, np.int64),
                                                plan.contrib_ew if plan.contrib_ew.size else np.zeros(1), padded, plan.final_perm)]
    w_plan, w_rows = DeviceBuffer(nnz * H * 8), DeviceBuffer(nnz * H * 8)
    L = lib()
    check(L.hb_irf_compose_device(len(plan.level_row_ptr) - 1, plan.level_row_ptr.ctypes.data_as(C.c_void_p), bufs[0].ptr, bufs[1].ptr,
                                  bufs[2].ptr, bufs[3].ptr, bufs[4].ptr, w_plan.ptr, H, None))
    check(L.hb_irf_gather_rows_device(w_plan.ptr, bufs[5].ptr, nnz, H, 0, w_rows.ptr, None))          # reference row order [nnz][H]
    if keep_on_device:
        w_conv = DeviceBuffer(nnz * H * 8)
        check(L.hb_irf_gather_rows_device(w_rows.ptr, None, nnz, H, 1, w_conv.ptr, None))             # [H][nnz] for the convolution
        check(L.hb_stream_synchronize(None))
        for b in bufs + [w_plan]:
            b.free()
        return DeviceRouteKernel(indices, nn, nn, H, w_rows, w_conv)
    weights = w_rows.to_host((nnz, H))
    for b in bufs + [w_plan, w_rows]:
        b.free()
    return SparseRouteKernel(indices, weights, nn, nn)


class SparseRouteConvolution:
    """`SparseRouteConvolution(kernel)` functor (`src/route.jl:482-495,593-610`): `out[n_out, T] = conv(input[n_in, T])`
    on the GPU."""

    def __init__(self, kernel, *, name: Optional[str] = None, inputs=(), outputs=(), params=(), segment_rows: Optional[int] = 64):
        """`kernel`: a `SparseRouteKernel` (host weights) or a `DeviceRouteKernel` (weights already on the device, as
        `aggregate_route_kernel(..., keep_on_device=True)` leaves them).  `segment_rows`: destinations with more kernel rows
        than this are split into segments convolved by different blocks and added up in order afterwards (river networks:
        the outlet has one row per node); `None` keeps every destination whole (the reference's single running sum)."""
        self.kernel = kernel
        self.name = name or "##sparse_route"
        self.input_names, self.output_names, self.param_names = list(inputs), list(outputs), list(params)
        if kernel.indices.size and (kernel.indices[:, 1].max() > kernel.n_in or kernel.indices.min() < 1 or
                                    kernel.indices[:, 0].max() > kernel.n_out):
            raise ValueError("kernel indices out of range")
        dest = kernel.indices[:, 0] - 1
        sorted_rows = bool(np.all(np.diff(dest) >= 0))
        order = np.arange(dest.size) if sorted_rows else np.argsort(dest, kind="stable")   # CSR by destination, row order kept
        self._src = np.ascontiguousarray((kernel.indices[order, 1] - 1).astype(np.int32))
        ptr = np.zeros(kernel.n_out + 1, dtype=np.int64)
        np.add.at(ptr, dest + 1, 1)
        row_ptr = np.cumsum(ptr)
        self._w_dev = None
        if isinstance(kernel, DeviceRouteKernel):
            if not sorted_rows:
                raise ValueError("a device-resident kernel must list its rows by destination")
            self._w, self._w_dev = None, kernel.w_conv
        else:
            self._w = np.ascontiguousarray(kernel.weights[order].T)   # [H][nnz]
        # long destinations -> segments ("virtual destinations"): same CSR rows, more row pointers
        counts = np.diff(row_ptr)
        self._seg_ptr = None
        self.n_virtual = kernel.n_out
        if segment_rows is not None and counts.size and counts.max() > int(segment_rows):
            S = int(segment_rows)
            nseg = np.maximum(1, -(-counts // S))
            seg_ptr = np.concatenate([[0], np.cumsum(nseg)])
            self.n_virtual = int(seg_ptr[-1])
            which = np.repeat(np.arange(kernel.n_out), nseg)                   # virtual -> real destination
            k_in = np.arange(self.n_virtual) - seg_ptr[which]                  # segment number inside its destination
            v_start = row_ptr[which] + k_in * S
            v_end = np.minimum(v_start + S, row_ptr[which + 1])
            row_ptr = np.concatenate([v_start, [v_end[-1]]])
            assert np.array_equal(row_ptr[1:], v_end)
            self._seg_ptr = seg_ptr.astype(np.int32)
            counts = np.diff(row_ptr)
        self._row_ptr = row_ptr.astype(np.int32)
        # destinations by decreasing row count: the 32 destinations of a kernel block then finish together
        self._dest_order = balanced_destination_order(counts) if counts.size and counts.max() != counts.min() else None
        self._static = None                                           # device copies of the CSR structure (first call)
        self._part = None                                             # [T][n_virtual] partial results (split destinations)

    def _device_structure(self):
        from .engine import DeviceBuffer
        if self._static is None:
            w = self._w_dev if self._w_dev is not None else DeviceBuffer.from_host(self._w if self._w.size else np.zeros(1))
            self._static = (DeviceBuffer.from_host(self._row_ptr), DeviceBuffer.from_host(self._src if self._src.size else np.zeros(1, np.int32)), w,
                            DeviceBuffer.from_host(self._dest_order) if self._dest_order is not None else None,
                            DeviceBuffer.from_host(self._seg_ptr) if self._seg_ptr is not None else None)
        return self._static

    def launch(self, in_ptr: int, out_ptr: int, T: int, stream: Optional[int] = None):
        """Device-resident call: `in_ptr` `[T][n_in]`, `out_ptr` `[T][n_out]` (asynchronous on `stream`)."""
        from .engine import DeviceBuffer, check, lib
        k = self.kernel
        L = lib()
        L.hb_sparse_route_conv_ordered_device.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64] + [C.c_void_p] * 7
        L.hb_sparse_route_conv_ordered_device.restype = C.c_int
        rp, src, w, order, seg = self._device_structure()
        nnz = len(self._src) if k.indices.size else 0
        if seg is None:
            check(L.hb_sparse_route_conv_ordered_device(k.n_out, k.n_in, int(T), k.horizon, nnz, rp.ptr, src.ptr, w.ptr, in_ptr,
                                                        order.ptr if order is not None else None, out_ptr, stream))
            return
        need = int(T) * self.n_virtual * 8
        if self._part is None or self._part.nbytes < need:
            if self._part is not None:
                self._part.free()
            self._part = DeviceBuffer(need)
        check(L.hb_sparse_route_conv_ordered_device(self.n_virtual, k.n_in, int(T), k.horizon, nnz, rp.ptr, src.ptr, w.ptr, in_ptr,
                                                    order.ptr if order is not None else None, self._part.ptr, stream))
        check(L.hb_irf_reduce_segments_device(k.n_out, int(T), self.n_virtual, seg.ptr, self._part.ptr, out_ptr, stream))

    def __call__(self, input, timing: Optional[dict] = None) -> np.ndarray:
        from .engine import DeviceBuffer, check, lib
        x = np.asarray(input, dtype=np.float64)
        vec = x.ndim == 1
        if vec:
            x = x.reshape(1, -1)
        k = self.kernel
        if x.shape[0] != k.n_in:
            raise ValueError(f"Input node count {x.shape[0]} does not match kernel.n_in {k.n_in}.")
        T = x.shape[1]
        xin = np.asfortranarray(x)                                     # memory [T][n_in]
        din = DeviceBuffer.from_host(xin.ravel(order="K") if xin.size else np.zeros(1))
        dout = DeviceBuffer(max(1, k.n_out * T) * 8)
        self.launch(din.ptr, dout.ptr, T, None)
        out = dout.to_host((k.n_out, T), order="F")
        din.free()
        dout.free()
        return out[0] if vec else out
