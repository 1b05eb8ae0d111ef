"""CPU-only validation of the lowering: the generated CUDA-C body, compiled with g++ through the
test harness, must reproduce the oracle within the north-star tolerance (1e-10 relative in
Float64; absolute floor 1e-12 for values that the reference itself only resolves to ~1e-16
because `tanh(5x)+1` cancels)."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.dirname(__file__))
import hydro_oracle as ho  # noqa: E402
from cpu_harness import run_body_on_cpu  # noqa: E402

RTOL, ATOL = 1e-10, 1e-12


def assert_parity(got, ref):
    assert got.shape == ref.shape
    err = np.abs(got - ref) - (RTOL * np.abs(ref) + ATOL)
    assert np.all(np.isfinite(got)) and err.max() <= 0, f"max excess {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def _multinode_case(name, N, T, shared=False):
    ht = (np.arange(N) % 5 + 1) if shared else np.arange(1, N + 1)
    ntypes = int(ht.max())
    model = getattr(hm.models, name)(htypes=ht)
    inp = sy.forcing(name, 0, N, T)
    p = {"params": sy.parameters(name, 0, ntypes)}
    if name == "m50":
        et, q = hm.models.m50_chains()
        p["nns"] = {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}
    init = {k: np.full(N, v) for k, v in sy.INIT[name].items()}
    return model, inp, p, init


@pytest.mark.parametrize("name", ["exphydro", "gr4j", "hbv", "m50"])
@pytest.mark.parametrize("shared", [False, True])
def test_body_matches_oracle_multinode(name, shared):
    N, T = 24, 150
    model, inp, p, init = _multinode_case(name, N, T, shared)
    ref = ho.run_model(model, inp, p, initstates=init)
    got = run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init)
    assert_parity(got, ref)


@pytest.mark.parametrize("fast", [True, False])
def test_body_matches_oracle_single_basin_docs(golden, fast):
    d = golden("exphydro_01013500.npz")
    T = 2000
    inp = np.stack([d["temp"][:T], d["lday"][:T], d["prcp"][:T]])
    p = {"params": dict(f=0.0167, Smax=1709.46, Qmax=18.47, Df=2.674, Tmax=0.17, Tmin=-2.09)}
    init = dict(snowpack=0.0, soilwater=1303.0)
    model = hm.models.exphydro(docs_variant=True)
    ref = ho.run_model(model, inp, p, initstates=init)
    got = run_body_on_cpu(hm.B200Model(model, fast=fast), inp, p, initstates=init)
    assert_parity(got, ref)


def test_component_boundaries_are_honoured(golden):
    """Two-bucket ExpHydro and the one-bucket/two-state variant differ; the fused
    stepper must follow each, not merge them."""
    d = golden("exphydro_01013500.npz")
    T = 400
    inp = np.stack([d["temp"][:T], d["lday"][:T], d["prcp"][:T]])
    p = {"params": dict(f=0.0167, Smax=1709.46, Qmax=18.47, Df=2.674, Tmax=0.17, Tmin=-2.09)}
    init = dict(snowpack=0.0, soilwater=1303.0)
    for model in (hm.models.exphydro(), hm.models.exphydro_single_bucket()):
        assert_parity(run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init),
                      ho.run_model(model, inp, p, initstates=init))


def test_gr4j_single_node_fixture(golden):
    d = golden("gr4j_sample.npz")
    T = 365
    inp = np.stack([d["prec"][:T], d["pet"][:T]])
    p = {"params": dict(x1=320.11, x2=2.42, x3=69.63, x4=1.39)}
    init = dict(soilwater=235.97, routingstore=45.47)
    model = hm.models.gr4j()
    assert_parity(run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init),
                  ho.run_model(model, inp, p, initstates=init))


def test_row_subset_and_default_states():
    N, T = 8, 60
    model, inp, p, _ = _multinode_case("exphydro", N, T)
    ref = ho.run_model(model, inp, p)                       # default initstates = zeros
    got = run_body_on_cpu(hm.B200Model(model), inp, p, rows=["flow", "snowpack"])
    assert_parity(got, ref[[model.var_names.index("flow"), model.var_names.index("snowpack")]])


def test_standalone_components():
    s = hm.Symbols("a b c d", "p1 p2")
    flux = hm.HydroFlux(["c ~ a * p1 + b * p2", "d ~ a + p1 + b + p2"], symbols=s)
    inp = np.array([[2.0, 3.0], [3.0, 2.0], [4.0, 1.0]]).T
    got = run_body_on_cpu(hm.B200Model(flux), inp, {"params": dict(p1=3.0, p2=4.0)})
    assert np.array_equal(got, np.array([[18.0, 12.0], [17.0, 12.0], [16.0, 12.0]]).T)   # exact KAT
    g = hm.models.gr4j()
    uh = g.components[2]
    x = np.abs(np.random.default_rng(0).normal(size=(1, 50)))
    assert_parity(run_body_on_cpu(hm.B200Model(uh), x, {"params": dict(x4=2.3)}),
                  ho.run_uh(uh, x, {"params": dict(x4=2.3)}))
    assert np.array_equal(run_body_on_cpu(hm.B200Model(uh), x, {"params": dict(x4=0.0)}), x)  # K = 0 → identity


def test_objective_row_and_final_states():
    N, T = 6, 90
    model, inp, p, init = _multinode_case("hbv", N, T)
    ref = ho.run_model(model, inp, p, initstates=init)
    sim = run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init, objective=True)
    assert_parity(sim, ref[-1])
    out, final = run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init, return_final_states=True)
    assert np.array_equal(final, out[:5, :, -1])            # states are the first rows


def test_lowering_statistics_exphydro():
    """The optimisations the design relies on actually fire: 6 exp-class evaluations per step
    (the reference evaluates 12 tanh + 4 exp; step(Tmin-temp)/step(temp-Tmin) share one), 4 carried
    state-only values."""
    prog = hm.lower_stepper(hm.models.exphydro())
    assert prog.stats["exp_like_per_step"] == 6
    assert prog.n_carry >= 3
    assert prog.row_names == hm.models.exphydro().var_names


def test_unit_hydrograph_ordinates_evaluated_in_the_prologue():
    """`uh_on_device`: the kernel prologue computes the S-curve ordinates from the node's parameters (what the device-resident
    calibration needs) — same rows as with the host-built table, also when the reserved K exceeds a node's own K."""
    N, T = 9, 80
    model = hm.models.gr4j(htypes=np.arange(1, N + 1))
    inp = sy.forcing("gr4j", 0, N, T)
    p = {"params": sy.parameters("gr4j", 0, N)}
    p["params"]["x4"] = np.linspace(1.05, 2.9, N)                      # K from (2, 3) to (3, 6)
    init = {k: np.full(N, v) for k, v in sy.INIT["gr4j"].items()}
    eng = hm.B200Model(model)
    host = run_body_on_cpu(eng, inp, p, initstates=init)
    for kmax in ([3, 6], [4, 8]):
        dev = run_body_on_cpu(eng, inp, p, initstates=init, uh_on_device=True, uh_kmax=kmax)
        assert np.all(np.isfinite(dev))
        assert_parity(dev, host)
    assert_parity(host, ho.run_model(model, inp, p, initstates=init))


def test_per_component_configs_bake_their_state_floors():
    """`_normalize_component_configs` (`/root/reference/src/model.jl:140-152`): one config per component; here
    the floors differ (the snow bucket empties in summer, so its floor is active every year)."""
    model, inp, p, init = _multinode_case("exphydro", 5, 400)
    cfgs = (hm.HydroConfig(min_value=1e-2), hm.HydroConfig(min_value=1e-6))
    ref = ho.run_model(model, inp, p, cfgs, init)
    got = run_body_on_cpu(hm.B200Model(model), inp, p, cfgs, init)
    assert_parity(got, ref)
    assert np.min(ref[0]) == 1e-2 and not np.allclose(ref, ho.run_model(model, inp, p, cfgs[1], init))
    same = ho.run_model(model, inp, p, {"components": cfgs}, init)
    assert np.array_equal(same, ref)
    with pytest.raises(ValueError, match="Component configs length"):
        ho.run_model(model, inp, p, cfgs[:1], init)


def test_issue_slot_lowerings_fire():
    """Round-2 lowerings that trim issue slots of the compute-bound kernels: `max(0, x)` becomes the sign-bit test `hb_relu`, and a
    body that raises to a power asks for the table logarithm (`HB_LOG_TABLE`, 2 KB of shared memory) — others do not pay for it."""
    hbv = hm.lower_stepper(hm.models.hbv(htypes=[1, 2]))
    assert "hb_relu(" in hbv.body and "hb_max(0.0" not in hbv.body and "hb_max(hb_minv" in hbv.body
    assert "#define HB_LOG_TABLE 1" in hbv.body and "hb_pow(" in hbv.body
    exp = hm.lower_stepper(hm.models.exphydro(htypes=[1]))
    assert "HB_LOG_TABLE" not in exp.body and "hb_relu(" in exp.body
    f32 = hm.lower_stepper(hm.models.hbv(htypes=[1]), precision="f32")
    assert "HB_LOG_TABLE" not in f32.body
