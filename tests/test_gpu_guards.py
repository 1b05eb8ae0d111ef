"""Guard-band checks (compute-sanitizer is closed on this GPU pool, profiles/round2_sanitizer_unavailable.txt): every
kernel family runs on device buffers whose outputs are framed by canary regions and whose inputs are read back afterwards —
no byte outside the declared output may change, no input may be modified.  Shapes are small and ragged on purpose (the
per-thread I/O paths, partial warps, partial tiles, grid borders)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceBuffer, DeviceRun, GridDeviceRun, RoutedDeviceRun, check, lib

pytestmark = pytest.mark.gpu
GUARD = 4096                      # bytes on either side of every output
CANARY = -1.2345678912345678e+300


class Guarded:
    """`nbytes` of payload between two canary regions."""

    def __init__(self, nbytes):
        self.n = int(nbytes)
        self.buf = DeviceBuffer(self.n + 2 * GUARD)
        check(lib().hb_fill_device(self.buf.ptr, (self.n + 2 * GUARD) // 8, C.c_double(CANARY), None))
        check(lib().hb_stream_synchronize(None))
        self.ptr = self.buf.ptr + GUARD

    def payload(self, shape, dtype=np.float64):
        raw = self.buf.to_host(((self.n + 2 * GUARD) // 8,))
        lo, hi = raw[:GUARD // 8], raw[(GUARD + self.n) // 8:]
        assert np.all(lo == CANARY) and np.all(hi == CANARY), "a kernel wrote outside its output buffer"
        body = raw[GUARD // 8:(GUARD + self.n) // 8]
        return body.view(dtype).reshape(shape) if dtype != np.float64 else body.reshape(shape)


def _inputs(a):
    a = np.ascontiguousarray(a)
    return DeviceBuffer.from_host(a), a.copy()


def _unchanged(buf, host):
    assert np.array_equal(buf.to_host(host.shape, dtype=host.dtype), host), "a kernel modified one of its inputs"


@pytest.mark.parametrize("name,N,T", [("exphydro", 256, 23), ("exphydro", 77, 11), ("gr4j", 129, 17), ("m50", 65, 9), ("hbv", 1, 30)])
def test_stepper_writes_only_its_records(name, N, T):
    hm.engine.init(0)
    model = getattr(hm.models, name)(htypes=np.arange(1, N + 1))
    p = {"params": sy.parameters(name, 0, N)}
    if name == "m50":
        et, q = hm.models.m50_chains()
        p["nns"] = {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}
    init = {k: np.full(N, v) for k, v in sy.INIT[name].items()}
    inp = np.asfortranarray(sy.forcing(name, 0, N, T))
    eng = hm.B200Model(model)
    ref = eng(inp, p, initstates=init)
    run = DeviceRun(eng, N, T, p, initstates=init)
    d_in, h_in = _inputs(inp.ravel(order="K"))
    R = len(model.var_names)
    out = Guarded(T * N * R * 8)
    fin = Guarded(len(model.state_names) * N * 8)
    run.launch(d_in.ptr, out.ptr, None, final_states_ptr=fin.ptr)
    check(lib().hb_stream_synchronize(None))
    got = out.payload((T, N, R)).transpose(2, 1, 0)
    assert np.array_equal(got, ref)
    assert np.array_equal(fin.payload((len(model.state_names), N)), ref[:len(model.state_names), :, -1])
    _unchanged(d_in, h_in)
    for key in ("params", "init"):
        if key in run.bufs:
            pass                                    # staged by the run itself; their integrity shows in the equality above


@pytest.mark.parametrize("R,Cc,T", [(33, 70, 21), (40, 128, 12), (7, 5, 9)])
def test_grid_kernels_write_only_their_records(R, Cc, T):
    hm.engine.init(0)
    rng = np.random.default_rng(R + Cc)
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    rr, cc = np.divmod(np.arange(R * Cc), Cc)
    N = R * Cc
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))}
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
    inp = np.asfortranarray(sy.forcing("exphydro", 0, N, T))
    eng = hm.B200Model(model)
    ref = eng(inp, p, initstates=init)
    d_in, h_in = _inputs(inp.ravel(order="K"))
    for kind in ("wave", "generic3", "generic2"):
        out = Guarded(T * N * 13 * 8)
        if kind == "wave":
            run = GridDeviceRun(eng, N, T, p, initstates=init, variant="wave", steps_per_chunk=5)
            run.launch(d_in.ptr, out.ptr, None)
            run.check()
        else:
            run = RoutedDeviceRun(eng, N, T, p, initstates=init, legacy=(kind == "generic2"))
            run.launch(d_in.ptr, out.ptr, None)
            check(lib().hb_stream_synchronize(None))
        assert np.array_equal(out.payload((T, N, 13)).transpose(2, 1, 0), ref), kind
        _unchanged(d_in, h_in)


def test_irf_and_raster_kernels_write_only_their_outputs():
    hm.engine.init(0)
    from hydromodels_jl_b200 import irf
    rng = np.random.default_rng(5)
    n, H, T = 333, 16, 71
    down = irf.synthetic_river_tree(n, seed=2)
    up = np.nonzero(down >= 0)[0]
    kernels = [rng.random(H) / H for _ in range(n)]
    K = irf.aggregate_route_kernel((down[up], up), kernels, horizon=H)
    x = np.abs(rng.normal(size=(n, T)))
    conv = irf.SparseRouteConvolution(K, segment_rows=32)
    ref = conv(x)
    d_x, h_x = _inputs(np.asfortranarray(x).ravel(order="K"))
    out = Guarded(T * n * 8)
    conv.launch(d_x.ptr, out.ptr, T, None)
    check(lib().hb_stream_synchronize(None))
    assert np.array_equal(out.payload((T, n)).T, ref)
    _unchanged(d_x, h_x)
    # dense forcing pack + D8 on a ragged grid
    R, Cc, Tn, V = 37, 45, 6, 3
    cubes = [rng.normal(size=(Tn, R, Cc)) for _ in range(V)]
    dcs = [_inputs(c) for c in cubes]
    ptrs = DeviceBuffer.from_host(np.array([d.ptr for d, _ in dcs], dtype=np.int64))
    packed = Guarded(Tn * R * Cc * V * 8)
    check(lib().hb_pack_forcing_dense_device(ptrs.ptr, V, Tn, R * Cc, Cc, 1, R, Cc, packed.ptr, None))
    check(lib().hb_stream_synchronize(None))
    got = packed.payload((Tn, R * Cc, V))
    for v in range(V):
        assert np.array_equal(got[:, :, v], cubes[v].reshape(Tn, R * Cc))
        _unchanged(*dcs[v])
    dem = np.round(rng.normal(500, 50, (R, Cc)), 1)
    d_dem, h_dem = _inputs(dem)
    codes = Guarded(((R * Cc + 7) // 8) * 8)
    check(lib().hb_d8_flow_direction_device(d_dem.ptr, R, Cc, Cc, 1, 30.0, 30.0, codes.ptr, None))
    check(lib().hb_stream_synchronize(None))
    flat = codes.payload((((R * Cc + 7) // 8) * 8,), dtype=np.uint8)[:R * Cc]
    assert set(np.unique(flat)) <= {0, 1, 2, 4, 8, 16, 32, 64, 128}
    _unchanged(d_dem, h_dem)
