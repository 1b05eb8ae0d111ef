"""Gradients (`/root/reference/README.md:161-187`): the complex-step oracle is pinned to the
forward oracle by finite differences; the tangent (hb_dual) lowering of the stepper body, compiled with g++ through the
test harness, must reproduce the oracle's derivatives."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.lowering import LoweringError, lower_stepper

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.dirname(__file__))
import hydro_oracle as ho  # noqa: E402
from cpu_harness import run_body_on_cpu  # noqa: E402


def _case(name, N, T, ntypes):
    ht = np.arange(N) % ntypes + 1
    model = getattr(hm.models, name)(htypes=ht)
    inp = sy.forcing(name, 0, N, T)
    p = {"params": sy.parameters(name, 0, ntypes)}
    init = {k: np.full(N, v) for k, v in sy.INIT[name].items()}
    target = 5.0 * np.random.default_rng(3).random(T)
    return model, ht, inp, p, init, target


def _sse(model, inp, p, init, target):
    sim = ho.run_model(model, inp, p, None, init)[-1]
    return ((sim - target) ** 2).sum()


def test_complex_step_oracle_agrees_with_finite_differences_of_the_forward_oracle():
    model, ht, inp, p, init, target = _case("exphydro", 6, 150, 3)
    loss, g = ho.loss_gradient(model, inp, p, target, "sse", 1, None, init)
    sim = ho.run_model(model, inp, p, None, init)[-1]
    np.testing.assert_allclose(loss, ((sim - target) ** 2).sum(axis=1), rtol=1e-13)
    for nm in model.param_names:
        for k in range(3):
            h = 1e-6 * abs(p["params"][nm][k])
            pp = {"params": {a: np.array(b, dtype=float) for a, b in p["params"].items()}}
            pm = {"params": {a: np.array(b, dtype=float) for a, b in p["params"].items()}}
            pp["params"][nm][k] += h
            pm["params"][nm][k] -= h
            fd = (_sse(model, inp, pp, init, target) - _sse(model, inp, pm, init, target)) / (2 * h)
            assert abs(g[nm][k] - fd) <= 1e-5 * abs(fd) + 1e-6, (nm, k, g[nm][k], fd)


@pytest.mark.parametrize("name", ["exphydro", "hbv"])
def test_tangent_body_matches_the_oracle_gradient(name):
    model, ht, inp, p, init, target = _case(name, 5, 120, 5)
    names = list(model.param_names)[:6]
    sim, dsim = run_body_on_cpu(hm.B200Model(model), inp, p, None, init, objective=True, tangents=names)
    loss, g = ho.loss_gradient(model, inp, p, target, "sse", 1, None, init, wrt=names)
    np.testing.assert_allclose(((sim - target) ** 2).sum(axis=1), loss, rtol=1e-10)
    for k, nm in enumerate(names):
        got = (2.0 * (sim - target) * dsim[k]).sum(axis=1)           # d SSE_n / d p_n, types are 1:1 with nodes here
        np.testing.assert_allclose(got, g[nm], rtol=1e-8, atol=1e-9 * np.abs(g[nm]).max())


def test_tangent_body_of_gr4j_differentiates_through_the_device_unit_hydrographs():
    N, T = 3, 150
    model, ht, inp, p, init, target = _case("gr4j", N, T, N)
    p["params"]["x4"] = np.array([1.37, 1.82, 2.41])                 # away from the integer breakpoints of the ordinates
    names = ["x1", "x2", "x3", "x4"]
    eng = hm.B200Model(model)
    kmax = eng.prepare(lower_stepper(model, objective=True), p, N)[5]
    sim, dsim = run_body_on_cpu(eng, inp, p, None, init, objective=True, tangents=names, uh_on_device=True, uh_kmax=kmax)
    np.testing.assert_allclose(sim, ho.run_model(model, inp, p, None, init)[-1], rtol=1e-10, atol=1e-12)
    for k, nm in enumerate(names):
        h = 1e-6 * np.abs(p["params"][nm])
        pp = {"params": dict(p["params"])}
        pm = {"params": dict(p["params"])}
        pp["params"][nm] = p["params"][nm] + h
        pm["params"][nm] = p["params"][nm] - h
        fd = (ho.run_model(model, inp, pp, None, init)[-1] - ho.run_model(model, inp, pm, None, init)[-1]) / (2 * h[:, None])
        np.testing.assert_allclose(dsim[k], fd, rtol=2e-5, atol=2e-6 * np.abs(fd).max())


def test_gradient_lowering_rejects_what_it_does_not_cover():
    et, q = hm.models.m50_chains()
    with pytest.raises(LoweringError, match="neural"):
        lower_stepper(hm.models.m50(htypes=[1, 2]), objective=True, tangents=["f"])
    with pytest.raises(LoweringError, match="does not exist"):
        lower_stepper(hm.models.exphydro(htypes=[1, 2]), objective=True, tangents=["nope"])
    with pytest.raises(LoweringError, match="objective run"):
        lower_stepper(hm.models.exphydro(htypes=[1, 2]), tangents=["f"])
    with pytest.raises(LoweringError, match="ordinates evaluated on the device"):
        lower_stepper(hm.models.gr4j(htypes=[1, 2]), objective=True, tangents=["x4"], uh_kmax=[2, 4])


def test_multi_start_adam_loop_with_the_oracle_as_the_gradient():
    """`solve(prob, MultiStartAdam())` host logic (projected Adam in box-normalised coordinates, best-so-far bookkeeping)
    driven by the oracle's gradient instead of the device's."""
    from hydromodels_jl_b200 import calibration as cal
    T, n = 150, 8
    names = list(hm.models.exphydro().get_param_names())
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    lb = np.array([-3.0, 0.0, 0.5, 500.0, 5.0, 0.005])
    ub = np.array([0.0, 3.0, 5.0, 2500.0, 40.0, 0.1])
    truth = dict(zip(names, [-2.09, 0.17, 2.674, 1709.46, 18.47, 0.0167]))
    init = dict(snowpack=0.0, soilwater=1303.0)
    one = hm.models.exphydro(htypes=[1])
    target = ho.run_model(one, inp[:, None, :], {"params": {k: np.array([v]) for k, v in truth.items()}}, None,
                          {k: np.array([v]) for k, v in init.items()})[-1, 0]
    prob = cal.OptimizationProblem.__new__(cal.OptimizationProblem)      # no engine: the loop only needs the box
    prob.calibrated, prob.fixed, prob.lb, prob.ub = names, {}, lb, ub
    calls = []

    def grad_fn(P):
        assert np.all(P >= lb) and np.all(P <= ub)
        calls.append(len(P))
        m = hm.models.exphydro(htypes=np.arange(1, len(P) + 1))
        f, g = ho.loss_gradient(m, np.repeat(inp[:, None, :], len(P), axis=1), {"params": {k: P[:, j].copy() for j, k in enumerate(names)}},
                                target, "nse", 30, None, {k: np.full(len(P), v) for k, v in init.items()})
        return f, np.stack([g[k] for k in names], axis=1)

    sol = cal._solve_adam(prob, cal.MultiStartAdam(n_starts=n, lr=0.03, seed=1), 12, grad_fn)
    assert calls == [n] * 13 and sol.history.shape == (13,)
    assert np.all(np.diff(sol.history) <= 0) and sol.history[-1] < sol.history[0] - 0.5
    assert sol.objective == sol.history[-1] == sol.fitness.min()
    f_check, _ = grad_fn(sol.u[None, :])
    assert abs(f_check[0] - sol.objective) < 1e-12
    assert set(sol.params) == set(names) and sol.population.shape == (n, 6)


def test_tangent_body_differentiates_with_respect_to_initial_states():
    model, ht, inp, p, init, target = _case("exphydro", 4, 90, 4)
    init = {"snowpack": np.array([0.0, 5.0, 40.0, 120.0]), "soilwater": np.array([1303.0, 800.0, 1500.0, 300.0])}
    names = ["soilwater", "Smax", "snowpack"]
    sim, dsim = run_body_on_cpu(hm.B200Model(model), inp, p, None, init, objective=True, tangents=names)
    loss, g = ho.loss_gradient(model, inp, p, target, "sse", 1, None, init, wrt=names)
    for k, nm in enumerate(names):
        got = (2.0 * (sim - target) * dsim[k]).sum(axis=1)
        np.testing.assert_allclose(got, g[nm], rtol=1e-8, atol=1e-9 * np.abs(g[nm]).max())
    # and the oracle's state derivative against finite differences of the forward oracle
    h = 1e-4
    up = {k: v.copy() for k, v in init.items()}
    dn = {k: v.copy() for k, v in init.items()}
    up["soilwater"] += h
    dn["soilwater"] -= h
    sse = lambda st: ((ho.run_model(model, inp, p, None, st)[-1] - target) ** 2).sum(axis=1)
    np.testing.assert_allclose(g["soilwater"], (sse(up) - sse(dn)) / (2 * h), rtol=1e-5, atol=1e-8)   # exact zeros: a soil store above Smax spills its excess in one step


def test_adjoint_bodies_lower_and_compile_without_a_device():
    """`lower_stepper(adjoint=True)` + NVRTC for sm_100a (no GPU needed to compile): bucket models with and without neural
    fluxes; unit hydrographs and gradient-unfriendly settings are refused with the reference-style error text."""
    import numpy as np
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200.engine import CompiledAdjoint
    from hydromodels_jl_b200.lowering import LoweringError, lower_stepper
    for name in ("exphydro", "hbv", "m50"):
        model = getattr(hm.models, name)(htypes=np.arange(1, 5))
        prog = lower_stepper(model, objective=True, adjoint=True)
        assert prog.state_names == model.state_names and "HB_SEED_STMT" in prog.body and "hb_lam[" in prog.body
        assert (prog.n_groups > 0) == bool(prog.nn_layout)
        ck = CompiledAdjoint(prog)
        assert "hb_adjoint" in ck.source() and len(ck.cubin()) > 1000
    m50 = lower_stepper(hm.models.m50(htypes=np.arange(1, 5)), objective=True, adjoint=True)
    assert m50.nn_words == 690 and "hb_mlp_bwd_0" in m50.body and "hb_outer_acc<16, 16>" in m50.body and "hb_outer_flush<8, 16>" in m50.body
    with pytest.raises(LoweringError, match="unit hydrographs"):
        lower_stepper(hm.models.gr4j(htypes=np.arange(1, 5)), objective=True, adjoint=True, uh_kmax=[3, 6])
    with pytest.raises(LoweringError, match="objective row"):
        lower_stepper(hm.models.exphydro(htypes=np.arange(1, 5)), objective=False, adjoint=True)
