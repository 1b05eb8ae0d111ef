"""CPU-only: IRF routing kernel construction (host side) against the oracle's literal restatement and
against first principles (a chain of linear reservoirs)."""
import os
import sys

import numpy as np

from hydromodels_jl_b200 import irf

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402


def _network(n, rng):
    """Random tree: node i drains into a random node j > i (adjacency[dest, upstream] = 1)."""
    A = np.zeros((n, n))
    for i in range(n - 1):
        A[rng.integers(i + 1, n), i] = 1.0
    return A


def test_truncated_convolution_and_padding():
    rng = np.random.default_rng(0)
    a, b = rng.random(7), rng.random(5)
    assert np.allclose(irf._truncated_route_convolution(a, b, 9), ho.truncated_route_convolution(a, b, 9), rtol=1e-15)
    assert np.allclose(irf._truncated_route_convolution(a, b, 9), np.convolve(a, b)[:9], rtol=1e-14)
    assert np.array_equal(irf._pad_route_kernel([1.0, 2.0], 4), [1.0, 2.0, 0.0, 0.0])
    assert np.array_equal(irf._pad_route_kernel([1.0, 2.0, 3.0], 2), [1.0, 2.0])


def test_aggregate_route_kernel_composes_down_the_network():
    rng = np.random.default_rng(1)
    n, H = 9, 12
    A = _network(n, rng)
    order = irf.topological_order(A)
    assert sorted(order) == list(range(1, n + 1))
    pos = {v: i for i, v in enumerate(order)}
    for d, u in zip(*np.nonzero(A)):
        assert pos[u + 1] < pos[d + 1]                           # upstream before downstream
    route = irf.RouteIRF(["k"], lambda p, dt, h: (1 - np.exp(-dt / p[0])) * np.exp(-np.arange(h) * dt / p[0]))
    kernels = irf.build_irf_kernels(route, rng.uniform(0.5, 3.0, (n, 1)), horizon=H)
    K = irf.aggregate_route_kernel(A, kernels, topo_order=order, horizon=H)
    assert K.n_out == K.n_in == n and K.horizon == H
    # every (dest, source) kernel = convolution of the node kernels along the unique path source -> dest
    down = {u: d for d, u in zip(*np.nonzero(A))}
    for (d1, s1), w in zip(K.indices, K.weights):
        path, cur = [s1 - 1], s1 - 1
        while cur != d1 - 1:
            cur = down[cur]
            path.append(cur)
        ref = np.zeros(H)
        ref[0] = 1.0
        for node in path:
            ref = np.convolve(ref, kernels[node])[:H]
        assert np.allclose(w, ref, rtol=1e-12, atol=1e-15)
    # no adjacency: identity pairing
    K0 = irf.aggregate_route_kernel(None, kernels, horizon=H)
    assert np.array_equal(K0.indices[:, 0], K0.indices[:, 1]) and np.allclose(K0.weights, np.stack(kernels))


def test_oracle_sparse_convolution_is_a_causal_convolution():
    rng = np.random.default_rng(2)
    idx = np.array([[1, 1], [2, 1], [2, 2], [3, 2]])
    w = rng.random((4, 5))
    x = rng.random((2, 40))
    out = ho.sparse_route_convolution(idx, w, 3, 2, x)
    ref = np.zeros((3, 40))
    for (d, s), k in zip(idx, w):
        ref[d - 1] += np.convolve(x[s - 1], k)[:40]
    assert np.allclose(out, ref, rtol=1e-13)


def test_balanced_destination_order_is_a_windowed_permutation():
    """`irf.balanced_destination_order`: a permutation that only moves destinations inside their window of consecutive
    nodes and sorts each window by decreasing row count (ties keep node order)."""
    from hydromodels_jl_b200.irf import balanced_destination_order
    rng = np.random.default_rng(0)
    counts = np.minimum(200, 1 + np.floor(3.0 * rng.pareto(1.4, 10_000))).astype(np.int64)
    for window in (64, 4096, 20_000):
        order = balanced_destination_order(counts, window)
        assert order.dtype == np.int32 and sorted(order.tolist()) == list(range(counts.size))
        for a in range(0, counts.size, window):
            seg = order[a:a + window]
            assert seg.min() == a and seg.max() == min(a + window, counts.size) - 1
            c = counts[seg]
            assert np.all(np.diff(c) <= 0)
            same = np.diff(c) == 0
            assert np.all(np.diff(seg)[same] > 0)
    # blocks of 32 destinations: padding (block max x 32 over the true row total) shrinks from ~10x to < 1.3x
    waste = lambda o: counts[o].reshape(-1, 32).max(axis=1).sum() * 32 / counts[o].sum()
    assert waste(balanced_destination_order(counts, 4096)[:9984]) < 1.3 < waste(np.arange(9984))
