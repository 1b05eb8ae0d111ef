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


def _truncated_route_convolution(a, b, horizon: int) -> np.ndarray:
    """`_truncated_route_convolution` (`src/route.jl:506-521`)."""
    out = np.zeros(horizon)
    ua, ub = min(len(a), horizon), min(len(b), horizon)
    for ia in range(ua):
        av = float(a[ia])
        if av == 0.0:
            continue
        mb = min(ub, horizon - ia)
        out[ia:ia + mb] += av * np.asarray(b[:mb], dtype=np.float64)
    return out


def topological_order(adjacency) -> List[int]:
    """Kahn's algorithm, 1-based node ids (`_topological_order`, `src/route.jl:147-169`); `adjacency[node,
    upstream] != 0` means `upstream` drains into `node`."""
    A = np.asarray(adjacency) != 0
    n = A.shape[0]
    indeg = A.sum(axis=1).astype(int)
    order, ready = [], [i for i in range(n) if indeg[i] == 0]
    while ready:
        u = ready.pop(0)
        order.append(u + 1)
        for v in np.nonzero(A[:, u])[0]:
            indeg[v] -= 1
            if indeg[v] == 0:
                ready.append(int(v))
    if len(order) != n:
        raise ValueError("river network has a cycle")
    return order


def aggregate_route_kernel(adjacency, node_kernels: Sequence, *, topo_order=None, horizon: int) -> SparseRouteKernel:
    """`aggregate_route_kernel` (`src/route.jl:523-577`): source→destination kernels by composing node kernels
    down the network in topological order, truncated to `horizon`."""
    nn = len(node_kernels)
    H = int(horizon)
    if H <= 0:
        raise ValueError(f"horizon must be positive, got {horizon}")
    if adjacency is None:
        idx = np.stack([np.arange(1, nn + 1), np.arange(1, nn + 1)], axis=1)
        w = np.stack([_pad_route_kernel(k, H) for k in node_kernels]) if nn else np.zeros((0, H))
        return SparseRouteKernel(idx, w, nn, nn)
    A = np.asarray(adjacency, dtype=np.float64)
    if A.shape != (nn, nn):
        raise ValueError("adjacency and node_kernels size mismatch.")
    order = list(range(1, nn + 1)) if topo_order is None else [int(i) for i in topo_order]
    pairwise: List[Dict[int, np.ndarray]] = [dict() for _ in range(nn)]
    for node in order:
        i = node - 1
        selfk = _pad_route_kernel(node_kernels[i], H)
        pairwise[i][i] = selfk.copy()
        for up in range(nn):
            ew = A[i, up]
            if ew == 0.0:
                continue
            for source, upk in pairwise[up].items():
                cand = _truncated_route_convolution(selfk, upk, H)
                if ew != 1.0:
                    cand = cand * ew
                if source in pairwise[i]:
                    pairwise[i][source] = pairwise[i][source] + cand
                else:
                    pairwise[i][source] = cand
    rows_i, rows_w = [], []
    for dest in range(nn):
        for source in sorted(pairwise[dest]):
            rows_i.append((dest + 1, source + 1))
            rows_w.append(pairwise[dest][source])
    return SparseRouteKernel(np.asarray(rows_i, dtype=np.int64).reshape(-1, 2), np.asarray(rows_w).reshape(-1, H), nn, nn)


class SparseRouteConvolution:
    """`SparseRouteConvolution(kernel)` functor (`src/route.jl:482-495,593-610`): `out[n_out, T] = conv(input[n_in, T])`
    on the GPU."""

    def __init__(self, kernel: SparseRouteKernel, *, name: Optional[str] = None, inputs=(), outputs=(), params=()):
        self.kernel = kernel
        self.name = name or "##sparse_route"
        self.input_names, self.output_names, self.param_names = list(inputs), list(outputs), list(params)
        dest = kernel.indices[:, 0] - 1
        order = np.argsort(dest, kind="stable")                       # CSR by destination, row order kept
        self._src = np.ascontiguousarray((kernel.indices[order, 1] - 1).astype(np.int32))
        ptr = np.zeros(kernel.n_out + 1, dtype=np.int64)
        np.add.at(ptr, dest + 1, 1)
        self._row_ptr = np.cumsum(ptr).astype(np.int32)
        self._w = np.ascontiguousarray(kernel.weights[order].T)       # [H][nnz]
        # destinations by decreasing row count: the 32 destinations of a kernel block then finish together
        counts = np.diff(self._row_ptr)
        self._dest_order = balanced_destination_order(counts) if counts.size and counts.max() != counts.min() else None
        if kernel.indices.size and (kernel.indices[:, 1].max() > kernel.n_in or kernel.indices.min() < 1 or
                                    kernel.indices[:, 0].max() > kernel.n_out):
            raise ValueError("kernel indices out of range")

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
        L = lib()
        L.hb_sparse_route_conv_ordered_device.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64] + [C.c_void_p] * 7
        L.hb_sparse_route_conv_ordered_device.restype = C.c_int
        xin = np.asfortranarray(x)                                     # memory [T][n_in]
        bufs = [DeviceBuffer.from_host(self._row_ptr), DeviceBuffer.from_host(self._src if self._src.size else np.zeros(1, np.int32)),
                DeviceBuffer.from_host(self._w if self._w.size else np.zeros(1)), DeviceBuffer.from_host(xin.ravel(order="K") if xin.size else np.zeros(1))]
        dout = DeviceBuffer(max(1, k.n_out * T) * 8)
        if self._dest_order is not None:
            bufs.append(DeviceBuffer.from_host(self._dest_order))
        check(L.hb_sparse_route_conv_ordered_device(k.n_out, k.n_in, T, k.horizon, len(self._src) if k.indices.size else 0,
                                                    bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr,
                                                    bufs[4].ptr if self._dest_order is not None else None, dout.ptr, None))
        out = dout.to_host((k.n_out, T), order="F")
        for b in bufs + [dout]:
            b.free()
        return out[0] if vec else out
