"""GPU gradients through the C-ABI (`hb_stepper_spec.n_tangents`, `hb_run_desc.gradient`) against the complex-step oracle
(`oracle/hydro_oracle.py: loss_gradient`), the counterpart of the reference's Zygote example (`README.md:161-187`)."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402

pytestmark = pytest.mark.gpu


def _case(name, N, T, ntypes, seed=3):
    ht = np.arange(N) % ntypes + 1
    model = getattr(hm.models, name)(htypes=ht)
    inp = sy.forcing(name, 0, N, T)
    p = {"params": sy.parameters(name, 0, ntypes)}
    init = {k: np.full(N, v) for k, v in sy.INIT[name].items()}
    target = 0.5 + 5.0 * np.random.default_rng(seed).random(T)
    return model, inp, p, init, target


def _close(got, ref, rtol=1e-8):
    scale = max(np.abs(ref).max(), 1e-300)
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("metric", ["SSE", "MSE", "NSE", "KGE", "LogKGE"])
def test_exphydro_gradient_matches_the_oracle(metric):
    model, inp, p, init, target = _case("exphydro", 70, 200, 7)          # 10 nodes share each parameter type
    eng = hm.B200Model(model)
    loss, g = eng.gradient(inp, p, target, metric=metric, warm_up=20, initstates=init)
    ref_loss, ref_g = ho.loss_gradient(model, inp, p, target, metric, 20, None, init)
    _close(loss, ref_loss, 1e-10)
    for nm in model.param_names:
        assert g[nm].shape == (7,)
        _close(g[nm], ref_g[nm])
    # the value is the one the fused objective kernel returns (SSE is only in the gradient vocabulary of the oracle)
    if metric != "SSE":
        _close(loss, eng.objective(inp, p, target, metric=metric, warm_up=20, initstates=init), 1e-12)


def test_hbv_gradient_takes_several_launches_and_per_node_targets():
    model, inp, p, init, _ = _case("hbv", 40, 150, 40)
    target = 0.5 + 3.0 * np.random.default_rng(5).random((40, 150))
    eng = hm.B200Model(model)
    t = {}
    loss, g = eng.gradient(inp, p, target, metric="SSE", initstates=init, max_directions=5, timing=t)
    ref_loss, ref_g = ho.loss_gradient(model, inp, p, target, "sse", 1, None, init)
    _close(loss, ref_loss, 1e-10)
    assert set(g) == set(model.param_names) and len(model.param_names) > 10 and t["kernel_ms"] > 0
    for nm in model.param_names:
        _close(g[nm], ref_g[nm])
    sub = eng.gradient(inp, p, target, metric="SSE", initstates=init, wrt=["FC", "BETA"])[1]
    assert list(sub) == ["FC", "BETA"]
    _close(sub["FC"], g["FC"], 1e-11)                # another kernel (2 directions, other FMA contractions): not bitwise


def test_ensemble_gradient_equals_the_multinode_gradient():
    """A calibration ensemble (single-basin model, shared forcing, one parameter set per member) differentiates like N
    independent nodes."""
    N, T = 33, 180
    names = list(hm.models.exphydro().param_names)
    pv = sy.parameters("exphydro", 0, N)
    inp1 = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    target = 0.5 + 5.0 * np.random.default_rng(1).random(T)
    init = dict(sy.INIT["exphydro"])
    eng = hm.B200Model(hm.models.exphydro())
    loss, g = eng.gradient(inp1, {"params": pv}, target, metric="NSE", initstates=init, n_members=N)
    multi = hm.models.exphydro(htypes=np.arange(1, N + 1))
    inpN = np.repeat(inp1[:, None, :], N, axis=1)
    ref_loss, ref_g = ho.loss_gradient(multi, inpN, {"params": pv}, target, "nse", 1, None, {k: np.full(N, v) for k, v in init.items()})
    _close(loss, ref_loss, 1e-10)
    for nm in names:
        _close(g[nm], ref_g[nm])


def test_gr4j_gradient_through_the_unit_hydrographs_against_finite_differences():
    N, T = 6, 200
    model, inp, p, init, target = _case("gr4j", N, T, N)
    p["params"]["x4"] = np.array([1.37, 1.82, 2.41, 1.15, 2.73, 1.61])    # away from the integer breakpoints
    eng = hm.B200Model(model)
    loss, g = eng.gradient(inp, p, target, metric="SSE", initstates=init)

    def sse(pp):
        return ((ho.run_model(model, inp, pp, None, init)[-1] - target) ** 2).sum(axis=1)

    _close(loss, sse(p), 1e-10)
    for nm in ["x1", "x2", "x3", "x4"]:
        h = 1e-6 * np.abs(p["params"][nm])
        pp = {"params": dict(p["params"])}
        pm = {"params": dict(p["params"])}
        pp["params"][nm] = p["params"][nm] + h
        pm["params"][nm] = p["params"][nm] - h
        _close(g[nm], (sse(pp) - sse(pm)) / (2 * h), 2e-5)


def test_single_basin_gradient_and_errors():
    T = 365
    model = hm.models.exphydro()
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    p = {"params": {k: float(v[0]) for k, v in sy.parameters("exphydro", 0, 1).items()}}
    target = 0.5 + 5.0 * np.random.default_rng(2).random(T)
    loss, g = hm.B200Model(model).gradient(inp, p, target, initstates=dict(sy.INIT["exphydro"]))     # README loss: SSE
    multi = hm.models.exphydro(htypes=[1])
    ref_loss, ref_g = ho.loss_gradient(multi, inp[:, None, :], {"params": {k: np.array([v]) for k, v in p["params"].items()}},
                                       target, "sse", 1, None, {k: np.array([v]) for k, v in sy.INIT["exphydro"].items()})
    _close(loss, ref_loss, 1e-10)
    for nm, v in g.items():
        assert np.ndim(v) == 0
        _close(np.array([v]), ref_g[nm])
    # forward-mode tangents do not cover neural components (their hundreds of weights are what the reverse-mode adjoint is
    # for: `gradient()` picks it by default for such models, tests/test_gpu_adjoint.py)
    with pytest.raises(hm.lowering.LoweringError, match="neural"):
        m50 = hm.models.m50(htypes=[1, 2])
        et, q = hm.models.m50_chains()
        hm.B200Model(m50).gradient(sy.forcing("m50", 0, 2, 50), {"params": sy.parameters("m50", 0, 2),
                                   "nns": {"etnn": et.init_params(1), "qnn": q.init_params(2)}}, np.ones(50), wrt=["Tmin"], mode="forward")


def test_gradient_with_respect_to_initial_states():
    model, inp, p, init, target = _case("exphydro", 40, 120, 5)
    init = {"snowpack": np.linspace(0.0, 150.0, 40), "soilwater": np.linspace(300.0, 1600.0, 40)}
    eng = hm.B200Model(model)
    wrt = ["soilwater", "Smax", "snowpack", "f"]
    loss, g = eng.gradient(inp, p, target, metric="NSE", warm_up=10, initstates=init, wrt=wrt)
    ref_loss, ref_g = ho.loss_gradient(model, inp, p, target, "nse", 10, None, init, wrt=wrt)
    _close(loss, ref_loss, 1e-10)
    assert g["soilwater"].shape == (40,) and g["Smax"].shape == (5,)
    for nm in wrt:
        _close(g[nm], ref_g[nm])
    assert np.any(g["snowpack"] != 0) and np.any(g["soilwater"] != 0)
