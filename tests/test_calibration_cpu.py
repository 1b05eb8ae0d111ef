"""CPU-only: the numpy restatement of the device DE kernels (oracle/hydro_oracle.py: de_uniform, de_trial, de_select)
and the host-side checks of `calibration.OptimizationProblem` (keywords of the reference's
`ext/HydroModelsOptimizationExt/HydroModelsOptimizationExt.jl:72-199`)."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import calibration as cal

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402


def test_de_draws_are_uniform_and_counter_based():
    u = ho.de_uniform(42, 3, np.arange(200_000), 7)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert np.array_equal(u[:10], ho.de_uniform(42, 3, np.arange(10), 7))           # a function of the counter only
    assert not np.array_equal(u[:10], ho.de_uniform(42, 4, np.arange(10), 7))
    assert ho.de_uniform(0, 0, 0, 0) == (np.uint64(0xE220A8397B1DCDAF) >> np.uint64(11)) * 2.0 ** -53   # SplitMix64(0) known answer


@pytest.mark.parametrize("N,D", [(4, 1), (5, 3), (64, 6), (1000, 13)])
def test_de_trial_indices_are_distinct_and_trials_stay_in_bounds(N, D):
    rng = np.random.default_rng(N)
    lb, ub = -rng.random(D), 1 + rng.random(D)
    pop = lb + rng.random((N, D)) * (ub - lb)
    for gen in (1, 2, 77):
        trial, (r1, r2, r3) = ho.de_trial(pop, lb, ub, 0.9, 0.5, 7, gen)
        i = np.arange(N)
        assert np.all((r1 != i) & (r2 != i) & (r3 != i) & (r1 != r2) & (r1 != r3) & (r2 != r3))
        assert np.all((0 <= r1) & (r1 < N) & (0 <= r3) & (r3 < N))
        assert np.all(trial >= lb) and np.all(trial <= ub)
        assert np.all(np.any(trial != pop, axis=1) | (N < 6))                         # jrand: at least one mutated dimension
    keep_pop, keep_fit = ho.de_select(pop, trial, np.zeros(N), np.where(np.arange(N) % 2 == 0, -1.0, np.nan))
    assert np.array_equal(keep_pop[::2], trial[::2]) and np.array_equal(keep_pop[1::2], pop[1::2])   # NaN never wins
    assert np.array_equal(keep_fit, np.where(np.arange(N) % 2 == 0, -1.0, 0.0))
    # a member whose OWN objective is NaN (KGE of a constant series at generation 0) must not be stuck: any real trial replaces
    # it, a NaN trial does not
    fit = np.array([np.nan, np.nan, 0.5, 0.5][:N] + [0.5] * max(0, N - 4))
    tf = np.array([0.9, np.nan, np.nan, 0.4][:N] + [0.6] * max(0, N - 4))
    kp, kf = ho.de_select(pop, trial, fit, tf)
    assert np.array_equal(kp[0], trial[0]) and kf[0] == 0.9 and np.array_equal(kp[1], pop[1]) and np.isnan(kf[1])
    assert np.array_equal(kp[2], pop[2]) and kf[2] == 0.5 and np.array_equal(kp[3], trial[3]) and kf[3] == 0.4


def test_adaptive_de_restatement_has_blackboxoptims_operators():
    """`ade_params` / `ade_trial` / `ade_select` (oracle): the bimodal Cauchy mixtures BlackBoxOptim's adaptive DE draws F and
    CR from (modes at 0.65 | 1.0 and 0.1 | 0.95, scale 0.1, truncated to [0, 1] with mass piling up at 1), parents that are
    distinct, differ from the target and lie within the radius window, re-entry inside the box, strict selection."""
    N = 4000
    f, cr = ho.ade_params(N, 11, 0)
    assert f.min() >= 0 and f.max() <= 1 and cr.min() >= 0 and cr.max() <= 1
    # mixture modes: about half of F within 0.1 of 0.65 (a Cauchy has half its mass within one scale), a quarter+ exactly 1
    assert 0.18 < np.mean(np.abs(f - 0.65) < 0.1) < 0.32 and 0.2 < np.mean(f == 1.0) < 0.4
    assert 0.18 < np.mean(np.abs(cr - 0.1) < 0.1) < 0.36 and 0.18 < np.mean(np.abs(cr - 0.95) <= 0.1) < 0.45
    rng = np.random.default_rng(0)
    D, radius = 5, 8
    lb, ub = -rng.random(D), 1 + rng.random(D)
    pop = lb + rng.random((N, D)) * (ub - lb)
    trial, (r1, r2, r3, w) = ho.ade_trial(pop, lb, ub, f, cr, radius, 11, 1)
    i = np.arange(N)
    for r in (r1, r2, r3):
        assert np.all(r != i) and np.all(((r - (i - w)) % N) < radius)
    assert np.all(r1 != r2) and np.all(r1 != r3) and np.all(r2 != r3)
    assert np.all(trial >= lb) and np.all(trial <= ub) and np.any(trial != pop)
    assert np.all(np.any(trial != pop, axis=1))                               # jrand: every trial differs from its target
    fit, tfit = rng.random(N), rng.random(N)
    tfit[:10] = fit[:10]                                                      # ties: the member stays and redraws
    pop2, fit2, f2, cr2 = ho.ade_select(pop, trial, fit, tfit, f, cr, 11, 1)
    better = tfit < fit
    assert np.array_equal(pop2[better], trial[better]) and np.array_equal(pop2[~better], pop[~better])
    assert np.array_equal(f2[better], f[better]) and np.mean(f2[~better] != f[~better]) > 0.6
    assert np.array_equal(fit2, np.minimum(fit, tfit))


def test_optimization_problem_keywords():
    model = hm.models.exphydro()
    inp, tgt = np.zeros((3, 10)), np.zeros(10)
    lb, ub = [-3, 0, 0, 100, 10, 0], [0, 3, 5, 2000, 50, 0.1]
    prob = cal.OptimizationProblem(model, inp, tgt, metric="KGE", warm_up=2, lb_pas=lb, ub_pas=ub)
    assert prob.calibrated == model.get_param_names() and prob.fixed == {}
    prob = cal.OptimizationProblem(model, inp, tgt, lb_pas=dict(Tmin=-3, Tmax=0, Df=0, Smax=100), ub_pas=dict(Tmin=0, Tmax=3, Df=5, Smax=2000),
                                   fixed_params={"params": dict(Qmax=18.47, f=0.0167)})
    assert prob.calibrated == ["Tmin", "Tmax", "Df", "Smax"] and prob.fixed == dict(Qmax=18.47, f=0.0167)
    p = prob._params(np.array([[-1.0, 1.0, 2.0, 1500.0]]))["params"]
    assert p["Smax"][0] == 1500.0 and p["f"][0] == 0.0167
    with pytest.raises(ValueError, match="Unknown metric"):
        cal.OptimizationProblem(model, inp, tgt, metric="rmse", lb_pas=lb, ub_pas=ub)
    with pytest.raises(ValueError, match="lb_pas is required"):
        cal.OptimizationProblem(model, inp, tgt)
    with pytest.raises(KeyError, match="not parameters of the component"):
        cal.OptimizationProblem(model, inp, tgt, lb_pas=lb, ub_pas=ub, fixed_params=dict(nope=1.0))
    with pytest.raises(ValueError, match="entries"):
        cal.OptimizationProblem(model, inp, tgt, lb_pas=lb[:3], ub_pas=ub)


def test_optimization_problem_of_a_hybrid_model_carries_the_network_weights():
    """`default_pas = get_initial_params(component)` / `initial_params` and `fixed_params=ComponentVector(params=…)` of the
    reference (`…OptimizationExt.jl:94-131`): with every physical parameter fixed the optimisation vector is the weights."""
    model = hm.models.m50()
    et, q = hm.models.m50_chains()
    inp, tgt = np.zeros((3, 10)), np.zeros(10)
    phys = {nm: 1.0 for nm in model.get_param_names()}
    w = {"etnn": et.init_params(1), "qnn": q.init_params(2)}
    prob = cal.OptimizationProblem(model, inp, tgt, metric="MSE", fixed_params={"params": phys}, initial_params={"nns": w})
    assert prob.calibrated == [] and prob.nn_names == ["etnn", "qnn"] and prob.lb.size == 0
    assert np.array_equal(prob.nns0["etnn"], w["etnn"]) and prob.nns0["qnn"].size == q.n_params
    assert set(prob._params(np.zeros((1, 0)))) == {"params", "nns"}
    # without `initial_params` the chains' own initialisers are used (Lux's default in the reference)
    prob = cal.OptimizationProblem(model, inp, tgt, metric="MSE", fixed_params={"params": phys})
    assert prob.nns0["etnn"].shape == (et.n_params,) and np.any(prob.nns0["etnn"] != 0)
    # fixed networks stay out of the trained set
    prob = cal.OptimizationProblem(model, inp, tgt, metric="MSE", fixed_params={"params": phys, "nns": {"qnn": w["qnn"]}})
    assert prob.fixed_nns == {"qnn"} and np.array_equal(prob.nns0["qnn"], w["qnn"])
    assert cal._chains_of(model).keys() == {"etnn", "qnn"}
