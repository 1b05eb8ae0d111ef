"""Calibration on the device: the reference's `OptimizationProblem(component, input, target; …)`
(`/root/reference/ext/HydroModelsOptimizationExt/HydroModelsOptimizationExt.jl:72-199`) and a population optimiser
driving the ensemble objective kernel without host round trips.

The reference's objective evaluates ONE parameter vector per call on the host and is handed to Optimization.jl
(BlackBoxOptim's adaptive DE in every example, `docs/src/tutorials/optimimal_parameters.md:189-195`).  Here:

* `OptimizationProblem(...)` keeps the reference's keywords (`metric`, `loss`-less, `warm_up`, `lb_pas`, `ub_pas`,
  `fixed_params`, `default_initstates`) and offers `prob.objective(p)` — same value the reference's closure returns —
  and `prob.objective_batch(P)` for a whole population in one launch (`B200Model.objective`, ensemble mode);
* `solve(prob, DifferentialEvolution(...) | AdaptiveDifferentialEvolution(...), maxiters=…)` runs DE/rand/1/bin (fixed F and
  CR, or BlackBoxOptim's self-adapting, radius-limited variant) entirely on the GPU: per generation
  `hb_de_trial_device` (mutation + crossover + bounds) writes the trial population straight into the stepper's
  parameter buffer, `hb_run_device` in objective mode scores it, `hb_de_select_device` keeps the better member and
  records the generation's best — three kernels on one stream, nothing crosses PCIe until the end.

Unit hydrographs: a normal run reads host-built ordinates, but a search changes the unit-hydrograph parameters every
generation, so the device loop compiles the stepper with `uh_on_device` — the kernel prologue evaluates the S-curve
ordinates of each member from its own parameters (`UnitHydrograph.weights`), with the number of ordinates reserved
from the parameter bounds.

Gradient-based calibration: the reference enables it with `adtype=Optimization.AutoForwardDiff()` and a gradient solver
(`ext/HydroModelsOptimizationExt/README.md:107-120`, `…OptimizationExt.jl:193-197`).  Here `prob.gradient_batch(P)` returns
the objective AND its derivatives for a whole set of parameter vectors from one forward-mode launch per 6 parameters
(`B200Model.gradient`), and `solve(prob, MultiStartAdam(...))` runs projected Adam from many starting points at once — the
starts are the ensemble members of that launch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

from .components import Component


@dataclass
class DifferentialEvolution:
    """DE/rand/1/bin.  `popsize` members, differential weight `F`, crossover rate `CR`."""
    popsize: int = 1024
    F: float = 0.7
    CR: float = 0.9
    seed: int = 0


@dataclass
class AdaptiveDifferentialEvolution:
    """BlackBoxOptim.jl's `adaptive_de_rand_1_bin_radiuslimited` — what every example of the reference passes to `solve`
    (`/root/reference/ext/HydroModelsOptimizationExt/examples.jl:28`, `docs/src/tutorials/optimimal_parameters.md:189-195`) —
    run generation-synchronously on the device: per-member F and CR redrawn from bimodal Cauchy mixtures whenever a member's
    trial fails, parents from a window of `radius` neighbouring members, random re-entry at the bounds
    (`hb_ade_*_device`, csrc/aot_kernels.cu).  BlackBoxOptim's own defaults: popsize 50, radius 8."""
    popsize: int = 1024
    radius: int = 8
    seed: int = 0


@dataclass
class MultiStartAdam:
    """Projected Adam in box-normalised coordinates (x = (p - lb)/(ub - lb), clipped to [0, 1]) from `n_starts` uniformly
    drawn starting points, all advanced together: one gradient launch per iteration."""
    n_starts: int = 256
    lr: float = 0.02
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    seed: int = 0


@dataclass
class Adam:
    """`Optimization.solve(prob, Adam(lr); maxiters)` on the reverse-mode gradient — the reference trains the neural fluxes of
    a hybrid model this way (`adtype=AutoZygote()` through its ImmutableSolver, `/root/reference/README.md:161-187`,
    `src/solve.jl:62-79`).  One forward run (state trajectory) + one adjoint sweep per iteration (`B200Model.gradient_reverse`);
    the optimisation vector is [calibrated physical parameters, all network weights], as in the reference's
    `OptimizationProblem` (`…OptimizationExt.jl:94-104`).  Physical parameters with bounds are projected onto them."""
    lr: float = 0.01
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8


@dataclass
class Solution:
    u: np.ndarray                      # best calibrated vector (order = prob.calibrated)
    params: Dict[str, float]           # all parameters, fixed ones included
    objective: float
    history: np.ndarray                # best objective after every generation (index 0 = initial population)
    n_evaluations: int
    population: np.ndarray             # final population [popsize, D]
    fitness: np.ndarray
    gpu_launches: int = 0
    elapsed_s: float = 0.0
    extras: dict = field(default_factory=dict)
    nns: dict = field(default_factory=dict)   # trained network weights by chain name (hybrid models)


class OptimizationProblem:
    """`OptimizationProblem(component, input[V, T], target[T]; metric, warm_up, lb_pas, ub_pas, fixed_params,
    default_initstates)` for a single-node component.  Parameters named in `fixed_params` keep their value, all others
    are calibrated within `[lb_pas, ub_pas]` (sequences in the order of `calibrated`, or mappings by name)."""

    def __init__(self, component: Component, input, target, *, metric: str = "NSE", warm_up: int = 1, lb_pas=None, ub_pas=None,
                 fixed_params: Optional[Mapping] = None, default_initstates: Optional[Mapping] = None, config=None,
                 initial_params: Optional[Mapping] = None):
        from .engine import METRICS, B200Model
        if metric.lower() not in METRICS:
            raise ValueError(f"Unknown metric: {metric}. Supported metrics: KGE, NSE, LogKGE, MSE")
        # hybrid models: the network weights are part of the optimisation vector (`default_pas = get_initial_params(component)`
        # or the caller's `initial_params`, `…OptimizationExt.jl:99-104`); `fixed_params` may also fix them (`nns=…`)
        self.nn_names = list(component.get_nn_names())
        ip = dict(initial_params) if initial_params else {}
        nested = bool(fixed_params) and ("params" in fixed_params or "nns" in fixed_params)
        fixed_nns = dict(fixed_params.get("nns", {})) if nested else {}
        self.nns0 = {}
        if self.nn_names:
            chains = _chains_of(component)
            given = dict(ip.get("nns", {}))
            for nm in self.nn_names:
                w = fixed_nns.get(nm, given.get(nm))
                self.nns0[nm] = np.array(chains[nm].init_params() if w is None else w, dtype=np.float64)
        self.fixed_nns = set(fixed_nns)
        self.initial = {k: float(v) for k, v in dict(ip.get("params", {})).items()}
        if nested:
            fixed_params = fixed_params.get("params", {})
        self.component, self.metric, self.warm_up, self.config = component, metric, int(warm_up), config
        self.input = np.asarray(input, dtype=np.float64)
        self.target = np.asarray(target, dtype=np.float64)
        if self.input.ndim != 2 or self.target.shape != (self.input.shape[1],):
            raise ValueError("input must be [variables, time] and target [time]")
        fixed = dict(fixed_params.get("params", fixed_params)) if fixed_params else {}
        names = list(component.get_param_names())
        unknown = [k for k in fixed if k not in names]
        if unknown:
            raise KeyError(f"fixed parameters {unknown} are not parameters of the component")
        self.fixed = {k: float(v) for k, v in fixed.items()}
        self.calibrated: List[str] = [nm for nm in names if nm not in self.fixed]

        def bounds(b, what):
            if b is None and (not self.calibrated or self.nn_names):
                # nothing physical to calibrate, or a gradient run from `initial_params`: bounds are optional
                return np.full(len(self.calibrated), -np.inf if what == "lb_pas" else np.inf)
            if b is None:
                raise ValueError(f"{what} is required: the device optimiser samples inside the bounds")
            if isinstance(b, Mapping):
                b = [b[nm] for nm in self.calibrated]
            b = np.asarray(b, dtype=np.float64).reshape(-1)
            if b.size != len(self.calibrated):
                raise ValueError(f"{what} has {b.size} entries for {len(self.calibrated)} calibrated parameters")
            return b
        self.lb, self.ub = bounds(lb_pas, "lb_pas"), bounds(ub_pas, "ub_pas")
        if np.any(self.ub < self.lb):
            raise ValueError("ub_pas < lb_pas")
        self.initstates = None if default_initstates is None else {k: float(v) for k, v in default_initstates.items()}
        self.engine = B200Model(component)

    def _params(self, P: np.ndarray) -> Dict:
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        d = {nm: np.ascontiguousarray(P[:, j]) for j, nm in enumerate(self.calibrated)}
        d.update({k: np.array([v]) for k, v in self.fixed.items()})
        out = {"params": d}
        if self.nns0:
            out["nns"] = self.nns0
        return out

    def objective_batch(self, P) -> np.ndarray:
        """Objective of every row of `P[n, D]` (minimised: -NSE, -KGE, -LogKGE or MSE), one ensemble launch."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        n = P.shape[0]
        init = None if self.initstates is None else {k: np.full(n, v) for k, v in self.initstates.items()}
        return self.engine.objective(self.input, self._params(P), self.target, metric=self.metric, warm_up=self.warm_up,
                                     config=self.config, initstates=init, n_members=n)

    def gradient_batch(self, P):
        """`(objective[n], d objective / d P [n, D])` for every row of `P[n, D]` — the reference's
        `OptimizationFunction(objective, AutoForwardDiff())` evaluated for a whole set of vectors at once."""
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        n = P.shape[0]
        init = None if self.initstates is None else {k: np.full(n, v) for k, v in self.initstates.items()}
        obj, g = self.engine.gradient(self.input, self._params(P), self.target, metric=self.metric, warm_up=self.warm_up,
                                      config=self.config, initstates=init, n_members=n, wrt=self.calibrated)
        return obj, np.stack([g[nm] for nm in self.calibrated], axis=1)

    def gradient(self, p):
        obj, g = self.gradient_batch(np.asarray(p, dtype=np.float64)[None, :])
        return float(obj[0]), g[0]

    def objective(self, p) -> float:
        """The reference's closure `(p, c) -> loss(target[warm_up:end], component(input, p, c)[end, warm_up:end])`."""
        return float(self.objective_batch(np.asarray(p, dtype=np.float64)[None, :])[0])


def _uh_kmax_from_bounds(prob: OptimizationProblem) -> List[int]:
    """Ordinates to reserve per unit hydrograph: the largest `ceil(max_lag(p))` over the corners of the parameter box."""
    import itertools
    import math
    from .components import HydroModel, UnitHydrograph
    from .symbolic import evaluate
    comp = prob.component
    uhs = [c for c in (comp.components if isinstance(comp, HydroModel) else (comp,)) if isinstance(c, UnitHydrograph)]
    out = []
    for uh in uhs:
        ranges = []
        for nm in uh.param_names:
            if nm in prob.fixed:
                ranges.append((prob.fixed[nm],))
            else:
                j = prob.calibrated.index(nm)
                ranges.append((float(prob.lb[j]), float(prob.ub[j])))
        lag = max(evaluate(uh.uh_conds[0][0], dict(zip(uh.param_names, corner))) for corner in itertools.product(*ranges))
        out.append(max(1, int(math.ceil(lag))))
    return out


def _solve_adam(prob: OptimizationProblem, alg: MultiStartAdam, maxiters: int, grad_fn=None) -> Solution:
    """`grad_fn(P) -> (objective[n], gradient[n, D])` defaults to the device (`prob.gradient_batch`); the CPU tests pass
    the oracle's to exercise this loop without a GPU."""
    import time
    grad_fn = grad_fn or prob.gradient_batch
    N, D = int(alg.n_starts), len(prob.calibrated)
    if N < 1 or D == 0:
        raise ValueError("need at least one start and one calibrated parameter")
    span = prob.ub - prob.lb
    rng = np.random.default_rng(alg.seed)
    X = rng.random((N, D))
    m, v = np.zeros((N, D)), np.zeros((N, D))
    best_f, best_x = np.full(N, np.inf), X.copy()
    history = np.empty(maxiters + 1)
    t0 = time.perf_counter()
    for it in range(maxiters + 1):
        f, g = grad_fn(prob.lb + X * span)
        f = np.where(np.isfinite(f), f, np.inf)
        better = f < best_f
        best_f = np.where(better, f, best_f)
        best_x[better] = X[better]
        history[it] = best_f.min()
        if it == maxiters:
            break
        gx = np.where(np.isfinite(g), g, 0.0) * span                 # chain rule into the normalised coordinates
        m = alg.beta1 * m + (1 - alg.beta1) * gx
        v = alg.beta2 * v + (1 - alg.beta2) * gx * gx
        mh, vh = m / (1 - alg.beta1 ** (it + 1)), v / (1 - alg.beta2 ** (it + 1))
        X = np.clip(X - alg.lr * mh / (np.sqrt(vh) + alg.eps), 0.0, 1.0)
    elapsed = time.perf_counter() - t0
    population = prob.lb + best_x * span
    best = int(np.argmin(best_f))
    params = dict(prob.fixed)
    params.update({nm: float(population[best, j]) for j, nm in enumerate(prob.calibrated)})
    per_eval = -(-D // 6)
    return Solution(u=population[best].copy(), params=params, objective=float(best_f[best]), history=history,
                    n_evaluations=N * (maxiters + 1), population=population, fitness=best_f,
                    gpu_launches=per_eval * (maxiters + 1), elapsed_s=elapsed)


def _chains_of(component) -> Dict[str, object]:
    """Every Lux-style chain of a component tree by name (`NeuralFlux.chain`, the three networks of a `NeuralBucket`)."""
    out: Dict[str, object] = {}

    def walk(c):
        for attr in ("chain", "flux_network", "state_network", "output_network"):
            ch = getattr(c, attr, None)
            if ch is not None and hasattr(ch, "init_params"):
                out[ch.name] = ch
        for attr in ("components", "fluxes", "dfluxes"):
            for sub in getattr(c, attr, None) or ():
                walk(sub)
    walk(component)
    return out


def _solve_hybrid_adam(prob: OptimizationProblem, alg: Adam, maxiters: int) -> Solution:
    """Adam over [calibrated physical parameters, network weights] of a hybrid model, gradients from the adjoint stepper."""
    import time
    m = prob.metric.lower()
    if m not in ("sse", "mse", "nse"):
        raise NotImplementedError("hybrid training uses the adjoint run: metric SSE, MSE or NSE")
    trained = [nm for nm in prob.nn_names if nm not in prob.fixed_nns]
    D = len(prob.calibrated)
    missing = [nm for nm in prob.calibrated if nm not in prob.initial]
    if missing:
        raise ValueError(f"initial_params is missing the calibrated parameters {missing} (a gradient run starts from a point)")
    x = np.concatenate([np.array([prob.initial[nm] for nm in prob.calibrated], dtype=np.float64)] + [prob.nns0[nm].copy() for nm in trained])
    sizes = [prob.nns0[nm].size for nm in trained]
    lb = np.concatenate([prob.lb, np.full(sum(sizes), -np.inf)])
    ub = np.concatenate([prob.ub, np.full(sum(sizes), np.inf)])
    init = prob.initstates
    eng = prob.engine

    def unpack(v):
        pd = {nm: float(v[j]) for j, nm in enumerate(prob.calibrated)}
        pd.update(prob.fixed)
        nns, o = dict(prob.nns0), D
        for nm, n in zip(trained, sizes):
            nns[nm] = v[o:o + n]
            o += n
        return {"params": pd, "nns": nns}

    def value_and_grad(v):
        loss, g = eng.gradient_reverse(prob.input, unpack(v), prob.target, metric=prob.metric, warm_up=prob.warm_up, config=prob.config,
                                       initstates=init)
        gv = np.concatenate([np.array([np.sum(g[nm]) for nm in prob.calibrated], dtype=np.float64)] + [g["nns"][nm] for nm in trained])
        return float(np.sum(loss)), gv

    mom, vel = np.zeros_like(x), np.zeros_like(x)
    hist, best_x, best = [], x.copy(), np.inf
    t0 = time.perf_counter()
    for it in range(1, maxiters + 1):
        f, g = value_and_grad(x)
        hist.append(f)
        if f < best:
            best, best_x = f, x.copy()
        mom = alg.beta1 * mom + (1 - alg.beta1) * g
        vel = alg.beta2 * vel + (1 - alg.beta2) * g * g
        x = x - alg.lr * (mom / (1 - alg.beta1 ** it)) / (np.sqrt(vel / (1 - alg.beta2 ** it)) + alg.eps)
        x = np.minimum(np.maximum(x, lb), ub)
    f, _ = value_and_grad(x)
    hist.append(f)
    if f < best:
        best, best_x = f, x.copy()
    full = unpack(best_x)
    sol = Solution(u=best_x, params=dict(full["params"]), objective=best, history=np.array(hist), n_evaluations=maxiters + 1,
                   population=best_x[None, :], fitness=np.array([best]), gpu_launches=2 * (maxiters + 1),
                   elapsed_s=time.perf_counter() - t0)
    sol.nns = {k: np.array(v) for k, v in full["nns"].items()}
    return sol


def _solve_de_sharded(prob: OptimizationProblem, alg: DifferentialEvolution, maxiters: int, comm) -> Solution:
    """DE/rand/1/bin with the population sharded over the ranks of `comm` (`parallel.LibComm`, one process per GPU): every rank
    scores `popsize / world` members; one all-gather of the population per generation is the only communication
    (`hb_de_trial_sharded_device`).  Same seed -> the single-GPU trajectory, bit for bit; every rank returns the same Solution."""
    import time
    from .engine import DeviceBuffer, DeviceRun, check, lib
    world, rank = comm.world, comm.rank
    N, D = int(alg.popsize), len(prob.calibrated)
    if N % world or N // world < 1 or N < 4:
        raise ValueError("the sharded search needs popsize to be a multiple of the number of ranks (and >= 4)")
    if D == 0:
        raise ValueError("nothing to calibrate: every parameter is fixed")
    nl, m0 = N // world, rank * (N // world)
    eng = prob.engine
    uh_dev = eng._has_uh()
    kmax = _uh_kmax_from_bounds(prob) if uh_dev else None
    T = prob.input.shape[1]
    rng = np.random.default_rng(alg.seed)
    P0 = prob.lb + rng.random((N, D)) * (prob.ub - prob.lb)                     # the whole initial population, same on every rank
    mine = P0[m0:m0 + nl]
    init = None if prob.initstates is None else {k: np.full(nl, v) for k, v in prob.initstates.items()}
    run = DeviceRun(eng, nl, T, prob._params(mine), config=prob.config, initstates=init, objective=prob.metric,
                    target=prob.target, warm_up=prob.warm_up, ensemble=True, _kmax=kmax, uh_on_device=uh_dev)
    slots = run.ck.program.param_slots
    off_of = {nm: int(run._host[0][j]) for j, (nm, _) in enumerate(slots)}
    missing = [nm for nm in prob.calibrated if nm not in off_of]
    if missing:
        raise KeyError(f"calibrated parameters {missing} are not used by the model's kernels")
    dim_off = np.array([off_of[nm] for nm in prob.calibrated], dtype=np.int64)
    trial = run.bufs["params"]
    # pop_all[world][D][nl]: rank r's block = its members, parameter-major
    blocks = np.ascontiguousarray(P0.reshape(world, nl, D).transpose(0, 2, 1))
    pop_all = DeviceBuffer.from_host(blocks.ravel())
    blk_bytes = D * nl * 8
    my_block = pop_all.ptr + rank * blk_bytes
    d_off, d_lb, d_ub = DeviceBuffer.from_host(dim_off), DeviceBuffer.from_host(prob.lb), DeviceBuffer.from_host(prob.ub)
    fit, tfit = DeviceBuffer(nl * 8), DeviceBuffer(nl * 8)
    hist, hist_all = DeviceBuffer((maxiters + 1) * 16), DeviceBuffer(world * (maxiters + 1) * 16)
    fit_all = DeviceBuffer(N * 8)
    inp = prob.input if prob.input.flags.f_contiguous else np.asfortranarray(prob.input)
    d_in = DeviceBuffer.from_host(inp.ravel(order="K"))
    L = lib()
    t0 = time.perf_counter()
    run.launch(d_in.ptr, fit.ptr, None)
    check(L.hb_de_best_device(fit.ptr, nl, hist.ptr, 0, None))
    launches = 3
    for g in range(1, maxiters + 1):
        check(L.hb_de_trial_sharded_device(pop_all.ptr, trial.ptr, d_off.ptr, D, nl, N, m0, d_lb.ptr, d_ub.ptr, float(alg.F), float(alg.CR),
                                           C.c_uint64(alg.seed), g, None))
        run.launch(d_in.ptr, tfit.ptr, None)
        check(L.hb_de_select_sharded_device(my_block, trial.ptr, fit.ptr, tfit.ptr, d_off.ptr, D, nl, hist.ptr, g, None))
        comm.all_gather(my_block, pop_all.ptr, blk_bytes, None)
        launches += 5
    comm.all_gather(fit.ptr, fit_all.ptr, nl * 8, None)
    comm.all_gather(hist.ptr, hist_all.ptr, (maxiters + 1) * 16, None)
    check(L.hb_stream_synchronize(None))
    elapsed = time.perf_counter() - t0
    fitness = fit_all.to_host((N,))
    population = pop_all.to_host((world, D, nl)).transpose(0, 2, 1).reshape(N, D)
    history = hist_all.to_host((world, maxiters + 1, 2))[:, :, 0].min(axis=0)
    best = int(np.argmin(fitness))
    params = dict(prob.fixed)
    params.update({nm: float(population[best, j]) for j, nm in enumerate(prob.calibrated)})
    for b in (pop_all, d_off, d_lb, d_ub, fit, tfit, hist, hist_all, fit_all, d_in):
        b.free()
    return Solution(u=population[best].copy(), params=params, objective=float(fitness[best]), history=history.copy(),
                    n_evaluations=N * (maxiters + 1), population=population, fitness=fitness, gpu_launches=launches, elapsed_s=elapsed,
                    extras={"world": world, "members_per_rank": nl, "all_gather_bytes_per_generation": blk_bytes * world})


def solve(prob: OptimizationProblem, alg=None, *, maxiters: int = 100, comm=None) -> Solution:
    """`solve(prob, alg; maxiters)`: differential evolution on the device (default) — returns the best member after
    `maxiters` generations — or `MultiStartAdam` on the device gradients.  `comm` (`parallel.LibComm` of a multi-GPU job):
    the population of a `DifferentialEvolution` search is sharded over its ranks."""
    import time
    from .engine import DeviceBuffer, DeviceRun, check, lib
    if isinstance(alg, MultiStartAdam):
        return _solve_adam(prob, alg, maxiters)
    if isinstance(alg, Adam):
        return _solve_hybrid_adam(prob, alg, maxiters)
    alg = alg or DifferentialEvolution()
    if comm is not None and getattr(comm, "world", 1) > 1:
        if type(alg) is not DifferentialEvolution:
            raise NotImplementedError("the sharded search covers DifferentialEvolution")
        return _solve_de_sharded(prob, alg, maxiters, comm)
    adaptive = isinstance(alg, AdaptiveDifferentialEvolution)
    N, D = int(alg.popsize), len(prob.calibrated)
    if N < 4:
        raise ValueError("differential evolution needs at least 4 members")
    if adaptive and (alg.radius < 4 or N < alg.radius):
        raise ValueError("adaptive differential evolution needs radius >= 4 and at least `radius` members")
    if D == 0:
        raise ValueError("nothing to calibrate: every parameter is fixed")
    eng = prob.engine
    uh_dev = eng._has_uh()
    kmax = _uh_kmax_from_bounds(prob) if uh_dev else None
    T = prob.input.shape[1]
    rng = np.random.default_rng(alg.seed)
    P0 = prob.lb + rng.random((N, D)) * (prob.ub - prob.lb)
    init = None if prob.initstates is None else {k: np.full(N, v) for k, v in prob.initstates.items()}
    run = DeviceRun(eng, N, T, prob._params(P0), config=prob.config, initstates=init, objective=prob.metric,
                    target=prob.target, warm_up=prob.warm_up, ensemble=True, _kmax=kmax, uh_on_device=uh_dev)
    slots = run.ck.program.param_slots
    off_of = {nm: int(run._host[0][j]) for j, (nm, _) in enumerate(slots)}
    missing = [nm for nm in prob.calibrated if nm not in off_of]
    if missing:
        raise KeyError(f"calibrated parameters {missing} are not used by the model's kernels")
    dim_off = np.array([off_of[nm] for nm in prob.calibrated], dtype=np.int64)
    n_flat = run.n_param_values
    trial = run.bufs["params"]                                         # the stepper reads the TRIAL population
    flat0 = trial.to_host((n_flat,))
    pop = DeviceBuffer.from_host(flat0)
    d_off, d_lb, d_ub = DeviceBuffer.from_host(dim_off), DeviceBuffer.from_host(prob.lb), DeviceBuffer.from_host(prob.ub)
    fit, tfit, hist = DeviceBuffer(N * 8), DeviceBuffer(N * 8), DeviceBuffer((maxiters + 1) * 16)
    inp = prob.input if prob.input.flags.f_contiguous else np.asfortranarray(prob.input)
    d_in = DeviceBuffer.from_host(inp.ravel(order="K"))
    L = lib()
    t0 = time.perf_counter()
    run.launch(d_in.ptr, fit.ptr, None)                                # initial population (the trial buffer holds it)
    # generation 0 bookkeeping: only the best of the initial population is recorded (no selection)
    check(L.hb_de_best_device(fit.ptr, N, hist.ptr, 0, None))
    launches = 3
    extra = []
    if adaptive:
        d_f, d_cr = DeviceBuffer(N * 8), DeviceBuffer(N * 8)
        extra = [d_f, d_cr]
        check(L.hb_ade_init_device(d_f.ptr, d_cr.ptr, N, C.c_uint64(alg.seed), None))
        launches += 1
    for g in range(1, maxiters + 1):
        if adaptive:
            check(L.hb_ade_trial_device(pop.ptr, trial.ptr, d_off.ptr, D, N, d_lb.ptr, d_ub.ptr, d_f.ptr, d_cr.ptr, int(alg.radius),
                                        C.c_uint64(alg.seed), g, None))
        else:
            check(L.hb_de_trial_device(pop.ptr, trial.ptr, d_off.ptr, D, N, d_lb.ptr, d_ub.ptr, float(alg.F), float(alg.CR),
                                       C.c_uint64(alg.seed), g, None))
        run.launch(d_in.ptr, tfit.ptr, None)
        if adaptive:
            check(L.hb_ade_select_device(pop.ptr, trial.ptr, fit.ptr, tfit.ptr, d_off.ptr, D, N, d_f.ptr, d_cr.ptr, C.c_uint64(alg.seed), g,
                                         hist.ptr, g, None))
        else:
            check(L.hb_de_select_device(pop.ptr, trial.ptr, fit.ptr, tfit.ptr, d_off.ptr, D, N, hist.ptr, g, None))
        launches += 4
    check(L.hb_stream_synchronize(None))
    elapsed = time.perf_counter() - t0
    flat = pop.to_host((n_flat,))
    fitness = fit.to_host((N,))
    history = hist.to_host((maxiters + 1, 2))
    population = np.stack([flat[o:o + N] for o in dim_off], axis=1)
    best = int(np.argmin(fitness))
    params = dict(prob.fixed)
    params.update({nm: float(population[best, j]) for j, nm in enumerate(prob.calibrated)})
    for b in [pop, d_off, d_lb, d_ub, fit, tfit, hist, d_in] + extra:
        b.free()
    return Solution(u=population[best].copy(), params=params, objective=float(fitness[best]), history=history[:, 0].copy(),
                    n_evaluations=N * (maxiters + 1), population=population, fitness=fitness, gpu_launches=launches,
                    elapsed_s=elapsed)
