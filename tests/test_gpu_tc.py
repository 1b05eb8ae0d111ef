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
