"""GPU: on-device calibration — DE kernels bit-exact against their numpy restatement, the
problem's objective equal to the reference's closure evaluated with the oracle, and a planted-truth calibration."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import calibration as cal
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceBuffer, check, lib

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402

pytestmark = pytest.mark.gpu

LB = np.array([-3.0, 0.0, 0.5, 500.0, 5.0, 0.005])
UB = np.array([0.0, 3.0, 5.0, 2500.0, 40.0, 0.1])
TRUE = np.array([-2.09, 0.17, 2.674, 1709.46, 18.47, 0.0167])        # docs/src/get_start_en.md:288-301
INIT = dict(snowpack=0.0, soilwater=1303.0)


def _problem(T=730, metric="NSE", **kw):
    model = hm.models.exphydro()
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    p = {"params": dict(zip(model.get_param_names(), TRUE))}
    target = ho.run_model(model, inp, p, initstates=INIT)[-1]
    return model, inp, target, cal.OptimizationProblem(model, inp, target, metric=metric, warm_up=30, lb_pas=LB, ub_pas=UB,
                                                       default_initstates=INIT, **kw)


@pytest.mark.parametrize("N,D,offs", [(4, 1, [0]), (37, 3, [0, 37, 74]), (1000, 6, [0, 1000, 2001, 3001, 4001, 5002])])
def test_de_kernels_match_the_numpy_restatement_bit_for_bit(N, D, offs):
    rng = np.random.default_rng(N + D)
    lb, ub = -rng.random(D), 1 + rng.random(D)
    pop = lb + rng.random((N, D)) * (ub - lb)
    n_flat = offs[-1] + N + 3
    flat = rng.random(n_flat)                                           # values between the dimensions must survive
    for d in range(D):
        flat[offs[d]:offs[d] + N] = pop[:, d]
    d_pop, d_trial = DeviceBuffer.from_host(flat), DeviceBuffer.from_host(flat)
    d_off = DeviceBuffer.from_host(np.array(offs, dtype=np.int64))
    d_lb, d_ub = DeviceBuffer.from_host(lb), DeviceBuffer.from_host(ub)
    for gen in (1, 5):
        check(lib().hb_de_trial_device(d_pop.ptr, d_trial.ptr, d_off.ptr, D, N, d_lb.ptr, d_ub.ptr, 0.7, 0.9, 12345, gen, None))
        got = d_trial.to_host((n_flat,))
        ref, _ = ho.de_trial(pop, lb, ub, 0.7, 0.9, 12345, gen)
        assert np.array_equal(np.stack([got[o:o + N] for o in offs], axis=1), ref)
        mask = np.ones(n_flat, bool)
        for o in offs:
            mask[o:o + N] = False
        assert np.array_equal(got[mask], flat[mask])
    fit, tfit = rng.random(N), rng.random(N)
    tfit[::7] = np.nan
    fit[1::5] = np.nan                        # incumbents without a valid objective: replaced by any real-valued trial
    d_fit, d_tfit, d_hist = DeviceBuffer.from_host(fit), DeviceBuffer.from_host(tfit), DeviceBuffer.from_host(np.zeros(4))
    check(lib().hb_de_select_device(d_pop.ptr, d_trial.ptr, d_fit.ptr, d_tfit.ptr, d_off.ptr, D, N, d_hist.ptr, 1, None))
    new_pop, new_fit = ho.de_select(pop, ref, fit, tfit)
    got = d_pop.to_host((n_flat,))
    assert np.array_equal(np.stack([got[o:o + N] for o in offs], axis=1), new_pop)
    assert np.array_equal(d_fit.to_host((N,)), new_fit, equal_nan=True)
    hist = d_hist.to_host((4,))
    assert hist[2] == np.nanmin(new_fit) and int(hist[3]) == int(np.nanargmin(new_fit))
    check(lib().hb_de_best_device(d_fit.ptr, N, d_hist.ptr, 0, None))          # generation-0 bookkeeping entry point
    hist = d_hist.to_host((4,))
    assert hist[0] == np.nanmin(new_fit) and int(hist[1]) == int(np.nanargmin(new_fit))
    with pytest.raises(ValueError):                                            # aliased arguments are refused
        check(lib().hb_de_select_device(d_pop.ptr, d_trial.ptr, d_fit.ptr, d_fit.ptr, d_off.ptr, D, N, d_hist.ptr, 1, None))


@pytest.mark.parametrize("metric", ["NSE", "KGE", "LogKGE", "MSE"])
def test_problem_objective_is_the_references_closure(metric):
    model, inp, target, prob = _problem(T=400, metric=metric)
    rng = np.random.default_rng(1)
    P = LB + rng.random((5, 6)) * (UB - LB)
    got = prob.objective_batch(P)
    for k in range(5):
        sim = ho.run_model(model, inp, {"params": dict(zip(model.get_param_names(), P[k]))}, initstates=INIT)[-1]
        ref = ho.objective(metric.lower(), target, sim, warm_up=30)    # loss(target[warm_up:end], output[end, warm_up:end])
        if metric == "LogKGE" and np.std(np.log(sim[29:] + 1e-6)) < 1e-8:
            # a flow of ~1e-17 mm/d: log(sim + 1e-6) is constant to 11 digits and its correlation is ill-conditioned in
            # ANY arithmetic (the two-pass oracle keeps ~4 digits here); the kernel must stay finite and close
            assert got[k] == pytest.approx(ref, rel=1e-6)
            continue
        assert got[k] == pytest.approx(ref, rel=1e-9, abs=1e-12)
    assert prob.objective(P[2]) == got[2]
    assert prob.objective(TRUE) == pytest.approx(0.0 if metric == "MSE" else -1.0, abs=1e-9)


def test_device_de_recovers_planted_parameters_and_follows_the_oracle_loop():
    model, inp, target, prob = _problem()
    alg = cal.DifferentialEvolution(popsize=512, F=0.7, CR=0.9, seed=3)
    sol = cal.solve(prob, alg, maxiters=150)
    assert sol.objective < -0.999 and sol.n_evaluations == 512 * 151 and sol.gpu_launches == 3 + 4 * 150
    assert np.all(np.diff(sol.history) <= 0) and sol.history[-1] == sol.objective == sol.fitness.min()
    assert np.all(sol.population >= LB) and np.all(sol.population <= UB)
    assert prob.objective(sol.u) == pytest.approx(sol.objective, rel=1e-12)
    # the snow and flow parameters are identifiable from 2 years of synthetic flow
    assert abs(sol.params["Smax"] - TRUE[3]) / TRUE[3] < 0.1 and abs(sol.params["f"] - TRUE[5]) / TRUE[5] < 0.1
    # same loop on the host with the oracle kernels and the device objective: identical trajectory for a few generations
    rng = np.random.default_rng(alg.seed)
    pop = LB + rng.random((512, 6)) * (UB - LB)
    fit = prob.objective_batch(pop)
    hist = [fit.min()]
    for g in range(1, 6):
        trial, _ = ho.de_trial(pop, LB, UB, alg.F, alg.CR, alg.seed, g)
        pop, fit = ho.de_select(pop, trial, fit, prob.objective_batch(trial))
        hist.append(fit.min())
    assert np.array_equal(np.array(hist), sol.history[:6])


@pytest.mark.parametrize("N,D,offs,radius", [(8, 1, [0], 8), (37, 3, [0, 37, 74], 4), (1000, 6, [0, 1000, 2001, 3001, 4001, 5002], 8)])
def test_adaptive_de_kernels_match_the_numpy_restatement_bit_for_bit(N, D, offs, radius):
    """`hb_ade_*_device` (BlackBoxOptim's adaptive, radius-limited DE made generation-synchronous) against the oracle's
    restatement: per-member F / CR draws, windows, trials with random re-entry at the bounds, strict selection and the
    redraws of the members that failed — every value identical."""
    rng = np.random.default_rng(N + D)
    lb, ub = -rng.random(D), 1 + rng.random(D)
    pop = lb + rng.random((N, D)) * (ub - lb)
    n_flat = offs[-1] + N + 3
    flat = rng.random(n_flat)
    for d in range(D):
        flat[offs[d]:offs[d] + N] = pop[:, d]
    d_pop, d_trial = DeviceBuffer.from_host(flat), DeviceBuffer.from_host(flat)
    d_off = DeviceBuffer.from_host(np.array(offs, dtype=np.int64))
    d_lb, d_ub = DeviceBuffer.from_host(lb), DeviceBuffer.from_host(ub)
    d_f, d_cr = DeviceBuffer(N * 8), DeviceBuffer(N * 8)
    check(lib().hb_ade_init_device(d_f.ptr, d_cr.ptr, N, 777, None))
    f, cr = ho.ade_params(N, 777, 0)
    assert np.array_equal(d_f.to_host((N,)), f) and np.array_equal(d_cr.to_host((N,)), cr)
    assert f.min() >= 0.0 and f.max() <= 1.0 and cr.min() >= 0.0 and cr.max() <= 1.0
    for gen in (1, 2, 3):
        check(lib().hb_ade_trial_device(d_pop.ptr, d_trial.ptr, d_off.ptr, D, N, d_lb.ptr, d_ub.ptr, d_f.ptr, d_cr.ptr, radius, 777, gen, None))
        got = d_trial.to_host((n_flat,))
        ref, _ = ho.ade_trial(pop, lb, ub, f, cr, radius, 777, gen)
        assert np.array_equal(np.stack([got[o:o + N] for o in offs], axis=1), ref)
        assert np.all(ref >= lb) and np.all(ref <= ub)
        mask = np.ones(n_flat, bool)
        for o in offs:
            mask[o:o + N] = False
        assert np.array_equal(got[mask], flat[mask])
        fit, tfit = rng.random(N), rng.random(N)
        tfit[::7] = np.nan
        fit[1::5] = np.nan
        tfit[2::9] = fit[2::9]                                                 # ties keep the member (strict improvement)
        d_fit, d_tfit, d_hist = DeviceBuffer.from_host(fit), DeviceBuffer.from_host(tfit), DeviceBuffer.from_host(np.zeros(8))
        check(lib().hb_ade_select_device(d_pop.ptr, d_trial.ptr, d_fit.ptr, d_tfit.ptr, d_off.ptr, D, N, d_f.ptr, d_cr.ptr, 777, gen,
                                         d_hist.ptr, gen, None))
        pop, new_fit, f, cr = ho.ade_select(pop, ref, fit, tfit, f, cr, 777, gen)
        got = d_pop.to_host((n_flat,))
        assert np.array_equal(np.stack([got[o:o + N] for o in offs], axis=1), pop)
        assert np.array_equal(d_fit.to_host((N,)), new_fit, equal_nan=True)
        assert np.array_equal(d_f.to_host((N,)), f) and np.array_equal(d_cr.to_host((N,)), cr)
        hist = d_hist.to_host((8,))
        assert hist[2 * gen] == np.nanmin(new_fit) and int(hist[2 * gen + 1]) == int(np.nanargmin(new_fit))
    with pytest.raises(ValueError):
        check(lib().hb_ade_trial_device(d_pop.ptr, d_trial.ptr, d_off.ptr, D, N, d_lb.ptr, d_ub.ptr, d_f.ptr, d_cr.ptr, 3, 777, 1, None))


def test_adaptive_de_recovers_planted_parameters():
    """The optimiser of the reference's examples (`BBO_adaptive_de_rand_1_bin_radiuslimited`) on the device: planted ExpHydro
    parameters recovered, monotone history, population inside the box, and at least as good as plain DE on the same budget."""
    model, inp, target, prob = _problem()
    sol = cal.solve(prob, cal.AdaptiveDifferentialEvolution(popsize=512, radius=8, seed=3), maxiters=150)
    assert sol.objective < -0.999 and sol.n_evaluations == 512 * 151 and sol.gpu_launches == 4 + 4 * 150
    assert np.all(np.diff(sol.history) <= 0) and sol.history[-1] == sol.objective == sol.fitness.min()
    assert np.all(sol.population >= LB) and np.all(sol.population <= UB)
    assert prob.objective(sol.u) == pytest.approx(sol.objective, rel=1e-12)
    assert abs(sol.params["Smax"] - TRUE[3]) / TRUE[3] < 0.1 and abs(sol.params["f"] - TRUE[5]) / TRUE[5] < 0.1
    with pytest.raises(ValueError, match="radius"):
        cal.solve(prob, cal.AdaptiveDifferentialEvolution(popsize=6, radius=8), maxiters=1)


def test_fixed_parameters_and_unsupported_models():
    model, inp, target, _ = _problem(T=365)
    prob = cal.OptimizationProblem(model, inp, target, warm_up=30, lb_pas=LB[[0, 1, 2, 3]], ub_pas=UB[[0, 1, 2, 3]],
                                   fixed_params=dict(Qmax=TRUE[4], f=TRUE[5]), default_initstates=INIT)
    sol = cal.solve(prob, cal.DifferentialEvolution(popsize=128, seed=1), maxiters=60)
    assert sol.objective < -0.99 and sol.params["Qmax"] == TRUE[4] and sol.u.shape == (4,)


def test_device_de_on_a_unit_hydrograph_model():
    """GR4J + two unit hydrographs (BASELINE config 2's model): every generation changes x4, so the kernel prologue evaluates
    each member's S-curve ordinates itself.  The device loop's objective must equal the host-table objective of the same
    members, and the search must recover a planted x4."""
    model = hm.models.gr4j()
    T = 730
    inp = sy.forcing("gr4j", 0, 1, T)[:, 0, :]
    names = model.get_param_names()
    truth = dict(x1=350.0, x2=0.5, x3=90.0, x4=1.7)
    init = {k: float(v) for k, v in sy.INIT["gr4j"].items()}
    target = ho.run_model(model, inp, {"params": truth}, initstates=init)[-1]
    lb = np.array([dict(x1=100.0, x2=-5.0, x3=20.0, x4=1.1)[n] for n in names])
    ub = np.array([dict(x1=1200.0, x2=3.0, x3=300.0, x4=2.9)[n] for n in names])
    prob = cal.OptimizationProblem(model, inp, target, metric="NSE", warm_up=60, lb_pas=lb, ub_pas=ub, default_initstates=init)
    assert cal._uh_kmax_from_bounds(prob) == [3, 6]
    sol = cal.solve(prob, cal.DifferentialEvolution(popsize=256, seed=5), maxiters=80)
    assert np.all(np.diff(sol.history) <= 0) and sol.objective < -0.995
    # the loop's fitness (ordinates evaluated on the device) vs the ensemble objective with host-built ordinates
    host = prob.objective_batch(sol.population)
    assert np.allclose(sol.fitness, host, rtol=1e-9, atol=1e-12)
    assert abs(sol.params["x4"] - truth["x4"]) < 0.1


def test_gradient_batch_and_multi_start_adam_recover_planted_parameters():
    """`adtype=AutoForwardDiff()` + a gradient solver in the reference (`ext/HydroModelsOptimizationExt/README.md:107-120`):
    here one forward-mode launch differentiates the objective of every start."""
    T = 730
    model = hm.models.exphydro()
    names = list(model.get_param_names())
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    lb = np.array([-3.0, 0.0, 0.5, 500.0, 5.0, 0.005])
    ub = np.array([0.0, 3.0, 5.0, 2500.0, 40.0, 0.1])
    truth = np.array([-2.09, 0.17, 2.674, 1709.46, 18.47, 0.0167])
    init = dict(snowpack=0.0, soilwater=1303.0)
    target = model(inp, {"params": dict(zip(names, truth))}, initstates=init)[-1]
    prob = cal.OptimizationProblem(model, inp, target, metric="NSE", warm_up=61, lb_pas=lb[1:], ub_pas=ub[1:],
                                   default_initstates=init, fixed_params={"Tmin": -2.09})
    assert prob.calibrated == names[1:]
    P = lb[1:] + np.random.default_rng(0).random((50, 5)) * (ub[1:] - lb[1:])
    f, g = prob.gradient_batch(P)
    np.testing.assert_allclose(f, prob.objective_batch(P), rtol=1e-11, atol=1e-13)
    j, h = 2, 1e-4 * (ub[3] - lb[3])                                   # Smax by central differences of the objective kernel
    Pp, Pm = P.copy(), P.copy()
    Pp[:, j] += h
    Pm[:, j] -= h
    fd = (prob.objective_batch(Pp) - prob.objective_batch(Pm)) / (2 * h)
    np.testing.assert_allclose(g[:, j], fd, rtol=2e-4, atol=1e-6 * np.abs(fd).max())
    f1, g1 = prob.gradient(P[3])
    assert abs(f1 - f[3]) < 1e-12 and np.allclose(g1, g[3], rtol=1e-10)
    sol = cal.solve(prob, cal.MultiStartAdam(n_starts=256, lr=0.03, seed=2), maxiters=80)
    assert np.all(np.diff(sol.history) <= 0)
    assert sol.objective < -0.97, sol.objective                       # NSE > 0.97 from gradients alone
    assert abs(prob.objective(sol.u) - sol.objective) < 1e-10 and sol.params["Tmin"] == -2.09


def test_hybrid_training_with_adam_on_the_adjoint():
    """The reference's hybrid-model workflow: physical parameters fixed, the two MLPs of M50 trained by Adam on reverse-mode
    gradients (`OptimizationProblem(...; fixed_params=ComponentVector(params=…), adtype=AutoZygote())` + `solve(prob, Adam(lr))`).
    Target = the flow of planted weights; start = those weights perturbed."""
    T = 400
    model = hm.models.m50()
    et, q = hm.models.m50_chains()
    phys = {k: float(np.atleast_1d(v)[0]) for k, v in sy.parameters("m50", 0, 1).items()}
    truth = {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}
    inp = sy.forcing("m50", 0, 1, T)[:, 0, :]
    init = {k: float(v) for k, v in sy.INIT["m50"].items()}
    target = ho.run_model(model, inp, {"params": phys, "nns": truth}, initstates=init)[-1]
    rng = np.random.default_rng(0)
    start = {k: v + rng.normal(0, 0.05, v.size) for k, v in truth.items()}
    prob = cal.OptimizationProblem(model, inp, target, metric="MSE", warm_up=30, fixed_params={"params": phys},
                                   initial_params={"nns": start}, default_initstates=init)
    sol = cal.solve(prob, cal.Adam(lr=0.005), maxiters=120)
    assert sol.history[0] > 0 and sol.objective < 0.05 * sol.history[0]          # the loss falls by more than 20x
    assert sol.u.size == et.n_params + q.n_params and set(sol.nns) == {"etnn", "qnn"}
    # the returned weights reproduce the returned objective through the forward engine
    sim = hm.B200Model(model)(inp, {"params": phys, "nns": sol.nns}, initstates=init)[-1]
    assert np.mean((sim[29:] - target[29:]) ** 2) == pytest.approx(sol.objective, rel=1e-9)
    # one network fixed: only the other one moves
    prob2 = cal.OptimizationProblem(model, inp, target, metric="MSE", warm_up=30, fixed_params={"params": phys, "nns": {"qnn": truth["qnn"]}},
                                    initial_params={"nns": start}, default_initstates=init)
    sol2 = cal.solve(prob2, cal.Adam(lr=0.005), maxiters=20)
    assert sol2.u.size == et.n_params and np.array_equal(sol2.nns["qnn"], truth["qnn"]) and sol2.objective < sol2.history[0]
