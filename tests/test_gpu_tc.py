"""tcgen05 / TMEM dense-layer path of the Float32 mode (csrc/templates/hb_tc.cuh): the stand-alone probe against numpy,
then M50 in Float32 mode with the tensor path against the per-thread Float32 path and the Float64 oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200.engine import DeviceBuffer, check, lib

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (24, 32), (64, 64)])
@pytest.mark.parametrize("reps", [1, 4])
def test_tcgen05_dense_layer_probe_matches_numpy(K, N, reps):
    """D[128][N] = A[128][K] * W^T by tcgen05.mma kind::tf32 with the 3xTF32 operand split: as accurate as an fp32 dot
    product (error bound 2e-6 of sum |a||w|; a single TF32 pass would be ~5e-4), for every row / column of the tile,
    repeated through the same TMEM columns and mbarrier."""
    hm.engine.init(0)
    rng = np.random.default_rng(K * 100 + N)
    A = rng.normal(size=(128, K)).astype(np.float32)
    A[3] = 0.0
    A[5, :] = np.arange(K) == 2                                   # picks out one weight column exactly
    W = rng.normal(size=(N, K))                                   # W[n][k]; device layout n + N*k
    dA, dW = DeviceBuffer.from_host(A), DeviceBuffer.from_host(np.ascontiguousarray(W.T).ravel())
    dD = DeviceBuffer(128 * N * 4)
    check(lib().hb_probe_tc_layer(K, N, dA.ptr, dW.ptr, dD.ptr, reps, None))
    got = dD.to_host((128, N), dtype=np.float32).astype(np.float64)
    Wf = W.astype(np.float32).astype(np.float64)
    ref = A.astype(np.float64) @ Wf.T
    bound = 2e-6 * (np.abs(A).astype(np.float64) @ np.abs(Wf).T) + 1e-30
    assert np.all(np.abs(got - ref) <= bound), float(np.max(np.abs(got - ref) / bound))
    assert np.array_equal(got[3], np.zeros(N))
    assert np.allclose(got[5], Wf[:, 2], rtol=1e-6, atol=0)


def _m50_case(N, T, n0=0):
    from hydromodels_jl_b200 import synthetic as sy
    model = hm.models.m50(htypes=np.arange(1, N + 1))
    inp = sy.forcing("m50", n0, n0 + N, T)
    et, q = hm.models.m50_chains()
    p = {"params": sy.parameters("m50", n0, n0 + N), "nns": {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}}
    init = {k: np.full(N, v) for k, v in sy.INIT["m50"].items()}
    return model, inp, p, init


@pytest.mark.parametrize("N,T", [(256, 365), (300, 90), (1, 40), (129, 64)])
def test_m50_float32_tensor_path_meets_the_float32_tolerance(N, T):
    """M50 in the opt-in Float32 mode with its 16 -> 16 layers on tcgen05 (3xTF32): the Float32 mode's stated tolerance
    against the Float64 oracle (per node NSE(flow) >= 1 - 1e-6, every row's RMSE <= 1e-4 of its range + 1e-5), the same
    bound the per-thread Float32 path is held to, on full, ragged (shadow threads in the last CTA) and single-node shapes;
    and the two Float32 paths agree with each other far inside that tolerance."""
    model, inp, p, init = _m50_case(N, T)
    if N == 1:
        model = hm.models.m50()
        inp = inp[:, 0, :]
        p = {"params": {k: float(np.asarray(v)[0]) for k, v in p["params"].items()}, "nns": p["nns"]}
        init = {k: float(v[0]) for k, v in init.items()}
    ref = ho.run_model(model, inp, p, initstates=init)
    x32 = inp.astype(np.float32)
    tc = hm.B200Model(model, precision="f32", nn_tensor=True)
    got = tc(x32, p, initstates=init)
    assert tc.last_kernel.spec.nn_tc_bytes > 0 and "tcgen05.mma" in tc.last_kernel.source()
    ff = hm.B200Model(model, precision="f32", nn_tensor=False)(x32, p, initstates=init)
    assert got.dtype == np.float32 and got.shape == ref.shape and np.all(np.isfinite(got))
    g64, r64 = got.astype(np.float64).reshape(ref.shape[0], -1, T), ref.reshape(ref.shape[0], -1, T)
    flow32, flow64 = g64[-1], r64[-1]
    nse = 1 - np.sum((flow32 - flow64) ** 2, axis=1) / np.sum((flow64 - flow64.mean(axis=1, keepdims=True)) ** 2, axis=1)
    assert nse.min() >= 1 - 1e-6, nse.min()
    rmse = np.sqrt(np.mean((g64 - r64) ** 2, axis=2))
    span = r64.max(axis=2) - r64.min(axis=2)
    assert np.all(rmse <= 1e-4 * span + 1e-5), float((rmse - 1e-4 * span).max())
    d = np.abs(got.astype(np.float64) - ff.astype(np.float64)).reshape(ref.shape[0], -1, T)
    assert np.all(np.sqrt(np.mean(d ** 2, axis=2)) <= 2e-5 * span + 1e-5)
    assert np.array_equal(tc(x32, p, initstates=init), got)                       # deterministic across launches
