"""CPU-only: the C restatement of the reference's hot loops (oracle/hydro_oracle.c) agrees with the
numpy restatement (oracle/hydro_oracle.py) and with the reference's exact KATs."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402
import hydro_oracle_c as hoc  # noqa: E402


def close(a, b):
    # numpy's SIMD exp/tanh and glibc's differ by <= 1 ulp per call
    return a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["exphydro", "gr4j", "hbv", "m50"])
def test_c_oracle_matches_numpy_oracle_multinode(name):
    N, T = 300, 120      # more than one 256-node block, ragged
    ht = np.arange(N) % 11 + 1
    model = getattr(hm.models, name)(htypes=ht)
    inp = sy.forcing(name, 0, N, T)
    p = {"params": sy.parameters(name, 0, 11)}
    if name == "m50":
        et, q = hm.models.m50_chains()
        p["nns"] = {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}
    init = {k: np.full(N, v) for k, v in sy.INIT[name].items()}
    a = ho.run_model(model, inp, p, initstates=init)
    b = hoc.run_model(model, inp, p, initstates=init)
    assert close(a, b)
    assert np.array_equal(b, hoc.run_model(model, inp, p, initstates=init, n_threads=1))   # thread-count invariant


def test_c_oracle_published_exphydro_run(golden):
    d = golden("exphydro_01013500.npz")
    model = hm.models.exphydro(docs_variant=True)
    inp = np.stack([d["temp"], d["lday"], d["prcp"]])
    p = {"params": dict(f=0.0167, Smax=1709.46, Qmax=18.47, Df=2.674, Tmax=0.17, Tmin=-2.09)}
    out = hoc.run_model(model, inp, p, hm.HydroConfig(min_value=1e-300), initstates=dict(snowpack=0.0, soilwater=1303.0))
    obs, sim = d["flow"], out[-1]
    nse = 1 - np.sum((obs - sim) ** 2) / np.sum((obs - obs.mean()) ** 2)
    assert abs(nse - 0.7323810502829138) < 5e-15        # docs/src/get_start_en.md:466-468


def test_c_oracle_route_and_d8_kats():
    s = hm.Symbols("q q_routed s_river", "lag")
    flwdir = np.array([[1, 4, 8], [1, 4, 4], [1, 1, 2]])
    positions = [(i, j) for i in (1, 2, 3) for j in (1, 2, 3)]
    route = hm.HydroRoute(rfluxes=[hm.HydroFlux("q_routed ~ s_river / (1 + lag)", symbols=s)],
                          dfluxes=[hm.StateFlux("s_river ~ q - q_routed", symbols=s)],
                          aggr_func=hm.build_aggr_func(flwdir, positions), htypes=np.arange(1, 10))
    x = np.ones((1, 9, 20))
    p = {"params": dict(lag=np.full(9, 0.2))}
    a, b = ho.run_route(route, x, p), hoc.run_route(route, x, p)
    assert np.allclose(a, b, rtol=1e-15, atol=0)
    assert np.allclose(b[0, :, 1], [1.166667, 2.0, 1.166667, 1.166667, 3.666667, 1.166667, 1.166667, 2.833333, 2.833333], atol=1e-6)
    # exact D8 KATs through the C push (test/miscellaneous/run_d8_grid.jl:20-31)
    flw = np.ascontiguousarray(np.array([[0, 4, 4, 0], [1, 2, 4, 8], [0, 1, 4, 0], [0, 0, 1, 1]], dtype=np.int32))
    pos = np.array([(1, 2), (1, 3), (2, 1), (2, 2), (2, 3), (2, 4), (3, 2), (3, 3), (4, 3), (4, 4)]) - 1
    r0, c0 = np.ascontiguousarray(pos[:, 0].astype(np.int32)), np.ascontiguousarray(pos[:, 1].astype(np.int32))
    q, infl, work = np.ones(10), np.zeros(10), np.zeros(16 + 36)
    P = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    hoc.lib().ho_d8_push(4, 4, P(flw, C.c_int32), C.c_int64(10), P(r0, C.c_int32), P(c0, C.c_int32), P(q), P(infl), P(work))
    assert np.array_equal(infl, [0.0, 0.0, 0.0, 2.0, 1.0, 0.0, 0.0, 4.0, 1.0, 1.0])


def test_c_oracle_grid_model_matches_numpy():
    rng = np.random.default_rng(0)
    R, Cc = 5, 6
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc)]
    N, T = len(pos), 60
    model = hm.models.exphydro_grid(flw, pos)
    inp = sy.forcing("exphydro", 0, N, T)
    p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=np.full(N, 1.0), lag=np.full(N, 0.2))}
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=np.zeros(N))
    assert close(ho.run_model(model, inp, p, initstates=init), hoc.run_model(model, inp, p, initstates=init))
