"""Host-side mirror of the reference's component types for the forward-run hot path.

Same names, argument meaning and error behaviour as the Julia types they stand in for:

* `HydroFlux` / `StateFlux`      — `/root/reference/src/flux.jl:22-86,225-250`
* `NeuralFlux` (+ `Dense`/`Chain`)— `/root/reference/src/nn.jl:32-165`
* `HydroBucket`                  — `/root/reference/src/bucket.jl:24-91,169-280`
* `UnitHydrograph`               — `/root/reference/src/uh.jl:21-88,187-296`
* `HydroRoute`, `build_aggr_func`— `/root/reference/src/route.jl:232-415`
* `HydroModel`                   — `/root/reference/src/model.jl:21-109,241-309`
* `HydroConfig`, `SolverType`    — `/root/reference/src/config.jl:30-52`, `src/HydroModels.jl:88-93`

These classes only *describe* a model (expression trees, row bookkeeping, topology).
Calling a component runs it on the GPU through `engine.py` (NVRTC-compiled kernels behind
the C-ABI `libhydrob200.so`); there is no CPU execution path in this package.
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

from .symbolic import (Call, Const, Equation, Expr, Sym, Symbols, as_expr, evaluate, free_symbols,
                       parse_equation, parse_expr)


# ---------------------------------------------------------------------------------------
# Enums / config
# ---------------------------------------------------------------------------------------
class SolverType(enum.Enum):
    """`@enum SolverType` (`/root/reference/src/HydroModels.jl:88-93`)."""
    MutableSolver = 0
    ImmutableSolver = 1
    ODESolver = 2
    DiscreteSolver = 3


MutableSolver = SolverType.MutableSolver
ImmutableSolver = SolverType.ImmutableSolver
ODESolver = SolverType.ODESolver
DiscreteSolver = SolverType.DiscreteSolver


class ConstantInterpolation:  # marker types, `/root/reference/src/interpolate.jl:45-50`
    pass


class LinearInterpolation:  # `/root/reference/src/interpolate.jl:91-103`
    pass


class ConfigurationError(ValueError):
    pass


class DimensionMismatchError(ValueError):
    pass


class ParameterError(KeyError):
    def __str__(self):  # KeyError quotes its message; keep the reference wording readable
        return self.args[0] if self.args else ""


@dataclass
class HydroConfig:
    """`HydroConfig` (`/root/reference/src/config.jl:30-52`).

    The engine consumes `solver`, `timeidx`, `min_value`; `device`/`parallel` are stored and
    ignored exactly as in the reference (`src/config.jl:36,85,116`)."""
    solver: SolverType = MutableSolver
    interpolator: type = ConstantInterpolation
    timeidx: Sequence[int] = field(default_factory=list)
    device: object = None
    min_value: float = 1e-6
    parallel: bool = False

    def __post_init__(self):
        if not (self.min_value > 0):
            # ArgumentError("min_value must be positive, got …") (`src/config.jl:46`)
            raise ValueError(f"min_value must be positive, got {self.min_value}")
        self.timeidx = [int(t) for t in self.timeidx]
        self.min_value = float(self.min_value)


def default_config() -> HydroConfig:
    return HydroConfig()


_CONFIG_KEYS = ("solver", "interpolator", "timeidx", "device", "min_value", "parallel")


def normalize_config(config) -> HydroConfig:
    """`normalize_config` (`/root/reference/src/config.jl:108-122`): accepts a `HydroConfig`,
    `None`, or a dict (the NamedTuple back-compat form, unknown keys dropped, nested
    `hydro_config` honoured)."""
    if config is None:
        return default_config()
    if isinstance(config, HydroConfig):
        return config
    if isinstance(config, dict):
        if "hydro_config" in config:
            return normalize_config(config["hydro_config"])
        return HydroConfig(**{k: v for k, v in config.items() if k in _CONFIG_KEYS})
    raise TypeError(f"Unsupported config type {type(config).__name__}")


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def _uniq(seq: Iterable[str]) -> List[str]:
    out, seen = [], set()
    for s in seq:
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def _as_equations(eqs, symbols: Optional[Symbols]) -> List[Equation]:
    if isinstance(eqs, (Equation, str)):
        eqs = [eqs]
    out = []
    for e in eqs:
        if isinstance(e, Equation):
            out.append(e)
        elif isinstance(e, str):
            if symbols is None:
                raise TypeError("formula strings need a `symbols=` namespace")
            out.append(parse_equation(e, symbols.table))
        elif isinstance(e, tuple) and len(e) == 2:
            out.append(Equation(e[0], as_expr(e[1])))
        else:
            raise TypeError(f"cannot interpret {e!r} as an equation")
    return out


def _check_htypes(htypes) -> Optional[np.ndarray]:
    if htypes is None:
        return None
    h = np.asarray(htypes)
    if h.ndim != 1 or not np.issubdtype(h.dtype, np.integer):
        raise TypeError("htypes must be a vector of integers (1-based, as in the reference)")
    return h.astype(np.int64)


class Component:
    """Common name getters (`get_input_names` & co. forwarded from HydroModelCore)."""
    name: str
    input_names: List[str]
    output_names: List[str]
    state_names: List[str]
    param_names: List[str]
    nn_names: List[str]
    htypes: Optional[np.ndarray]

    def get_input_names(self): return list(self.input_names)
    def get_output_names(self): return list(self.output_names)
    def get_state_names(self): return list(self.state_names)
    def get_param_names(self): return list(self.param_names)
    def get_nn_names(self): return list(self.nn_names)

    @property
    def is_multinode(self) -> bool:
        return self.htypes is not None

    # The functor: `component(input, params, config; initstates)`.
    def __call__(self, input, params, config=None, *, initstates=None, **kwargs):
        from .engine import run_component
        return run_component(self, input, params, config, initstates=initstates, **kwargs)


# ---------------------------------------------------------------------------------------
# Fluxes
# ---------------------------------------------------------------------------------------
class HydroFlux(Component):
    """`@hydroflux` — one or more `output ~ expr` equations without state
    (`/root/reference/src/flux.jl:22-86`)."""

    def __init__(self, eqs, *, symbols: Optional[Symbols] = None, htypes=None, name: Optional[str] = None):
        self.eqs = _as_equations(eqs, symbols)
        if not self.eqs:
            raise ValueError("HydroFlux needs at least one equation")
        self.output_names = [e.lhs.name for e in self.eqs]
        vs, ps = [], []
        for e in self.eqs:
            v, p = free_symbols(e.rhs)
            vs += v
            ps += p
        self.input_names = [v for v in _uniq(vs) if v not in self.output_names]
        self.param_names = _uniq(ps)
        self.state_names = []
        self.nn_names = []
        self.htypes = _check_htypes(htypes)
        self.name = name or f"##hydro_flux#{abs(hash(tuple(self.output_names))):x}"

    @property
    def exprs(self) -> List[Expr]:
        return [e.rhs for e in self.eqs]


class StateFlux(Component):
    """`@stateflux state ~ expr` (`/root/reference/src/flux.jl:225-250`): the state's
    per-step increment."""

    def __init__(self, eq, *, symbols: Optional[Symbols] = None, name: Optional[str] = None):
        eqs = _as_equations(eq, symbols)
        if len(eqs) != 1:
            raise ValueError("StateFlux takes exactly one equation")
        self.eq = eqs[0]
        self.state_names = [self.eq.lhs.name]
        vs, ps = free_symbols(self.eq.rhs)
        self.input_names = [v for v in vs if v != self.eq.lhs.name]
        self.param_names = list(ps)
        self.output_names = []
        self.nn_names = []
        self.htypes = None
        self.name = name or f"##state_flux#{self.eq.lhs.name}"

    @property
    def state(self) -> str:
        return self.eq.lhs.name

    @property
    def expr(self) -> Expr:
        return self.eq.rhs


_ACTIVATIONS = ("identity", "tanh", "leakyrelu", "relu", "sigmoid")


@dataclass(frozen=True)
class Dense:
    """`Lux.Dense(in => out, act)`: `act.(W*x .+ b)`; `leakyrelu` slope 0.01 (NNlib)."""
    n_in: int
    n_out: int
    activation: str = "identity"

    def __post_init__(self):
        if self.activation not in _ACTIVATIONS:
            raise ValueError(f"Unknown activation {self.activation!r}; supported: {_ACTIVATIONS}")

    @property
    def n_params(self) -> int:
        return self.n_out * self.n_in + self.n_out


@dataclass(frozen=True)
class Chain:
    """`Lux.Chain(layers...; name)`.  Flat parameter layout = `ComponentVector(
    LuxCore.initialparameters(rng, chain))`: per layer `weight[out,in]` column-major
    (out index fastest) followed by `bias[out]`."""
    layers: Tuple[Dense, ...]
    name: str

    def __post_init__(self):
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if a.n_out != b.n_in:
                raise ValueError("Chain layer sizes do not match")

    @property
    def n_in(self): return self.layers[0].n_in
    @property
    def n_out(self): return self.layers[-1].n_out
    @property
    def n_params(self): return sum(l.n_params for l in self.layers)

    def init_params(self, seed: int = 42) -> np.ndarray:
        """Glorot-uniform weights / zero bias like Lux's default initialiser (values differ
        from `StableRNG(42)` — not reproducible without Julia)."""
        rng = np.random.default_rng(seed)
        parts = []
        for l in self.layers:
            lim = math.sqrt(6.0 / (l.n_in + l.n_out))
            w = rng.uniform(-lim, lim, size=(l.n_out, l.n_in)).astype(np.float32).astype(np.float64)
            parts.append(w.reshape(-1, order="F"))
            parts.append(np.zeros(l.n_out))
        return np.concatenate(parts)


class NeuralFlux(Component):
    """`@neuralflux out ~ chain([in1, in2, …])` / `NeuralFlux(inputs, outputs, chain; norm)`
    (`/root/reference/src/nn.jl:32-70,111-132`).  Usable inside a `HydroBucket`, as a component of a `HydroModel`, or
    stand-alone: its functor takes 2D `[inputs, time]` or 3D `[inputs, nodes, time]` arrays (`:140-165` — there is no
    `htypes`, the chain's parameters are shared by all nodes) and returns `chain(norm(input))`.

    `norm` (`:48`, default identity) normalises the inputs before the chain.  In the reference it is a Julia function of
    the whole input array; here it is a function of the list of input SYMBOLS returning one expression per input
    (e.g. `lambda x: [(x[0] - 3.0) / 2.0, x[1], x[2]]`), so it lowers into the same kernel."""

    def __init__(self, inputs: Sequence[Sym], outputs: Union[Sym, Sequence[Sym]], chain: Chain, *,
                 name: Optional[str] = None, norm=None):
        if isinstance(outputs, Sym):
            outputs = [outputs]
        self.chain = chain
        self.input_names = [s.name for s in inputs]
        self.output_names = [s.name for s in outputs]
        if len(self.input_names) != chain.n_in or len(self.output_names) != chain.n_out:
            raise DimensionMismatchError(
                f"chain {chain.name} maps {chain.n_in}=>{chain.n_out} but flux has "
                f"{len(self.input_names)} inputs / {len(self.output_names)} outputs")
        self.param_names = []
        self.state_names = []
        self.nn_names = [chain.name]
        self.htypes = None
        self.name = name or f"##neural_flux#{chain.name}"
        self.norm_exprs: Optional[List[Expr]] = None
        if norm is not None:
            from .symbolic import as_expr, free_symbols
            ex = [as_expr(e) for e in norm([Sym(n) for n in self.input_names])]
            if len(ex) != len(self.input_names):
                raise DimensionMismatchError("norm must return one expression per input")
            for e in ex:
                extra = [v for group in free_symbols(e) for v in group if v not in self.input_names]
                if extra:
                    raise ValueError(f"norm may only use the flux's own inputs, found {extra}")
            self.norm_exprs = ex


class NeuralBucket(Component):
    """`NeuralBucket(; name, flux_network, state_network, output_network, n_inputs, n_states, n_outputs, inputs, states,
    outputs, htypes)` (`/root/reference/src/nn.jl:195-262`): an RNN-style bucket,

        flux_t  = flux_network(vcat(S_{t-1}, x_t))
        S_t     = S_{t-1} + state_network(vcat(S_{t-1}, flux_t))     (residual update — no `min_value` clamp)
        y_t     = output_network(flux_t)

    returning `[states; outputs]` like a `HydroBucket` (`:308-326`).  The three chains' parameters are `params[:nns]`
    entries named after the chains and are shared by all nodes (`:329-358`)."""

    def __init__(self, *, name: str, flux_network: "Chain", state_network: "Chain", output_network: "Chain", n_inputs: int,
                 n_states: int, n_outputs: int, inputs: Sequence[str], states: Sequence[str], outputs: Sequence[str], htypes=None):
        if len(inputs) != n_inputs:
            raise AssertionError("Number of input names must match n_inputs")
        if len(states) != n_states:
            raise AssertionError("Number of state names must match n_states")
        if len(outputs) != n_outputs:
            raise AssertionError("Number of output names must match n_outputs")
        n_fluxes = flux_network.n_out
        if flux_network.n_in != n_states + n_inputs:
            raise DimensionMismatchError(f"flux_network takes {flux_network.n_in} inputs, expected n_states + n_inputs = {n_states + n_inputs}")
        if state_network.n_in != n_states + n_fluxes or state_network.n_out != n_states:
            raise DimensionMismatchError(f"state_network must map n_states + n_fluxes = {n_states + n_fluxes} => n_states = {n_states}")
        if output_network.n_in != n_fluxes or output_network.n_out != n_outputs:
            raise DimensionMismatchError(f"output_network must map n_fluxes = {n_fluxes} => n_outputs = {n_outputs}")
        self.name = name
        self.flux_network, self.state_network, self.output_network = flux_network, state_network, output_network
        self.n_inputs, self.n_states, self.n_outputs, self.n_fluxes = n_inputs, n_states, n_outputs, n_fluxes
        self.input_names = [str(getattr(v, "name", v)) for v in inputs]
        self.state_names = [str(getattr(v, "name", v)) for v in states]
        self.output_names = [str(getattr(v, "name", v)) for v in outputs]
        self.param_names = []
        self.nn_names = [flux_network.name, state_network.name, output_network.name]
        self.htypes = None if htypes is None else np.asarray(htypes, dtype=np.int64)

    @property
    def has_states(self) -> bool:
        return True


# ---------------------------------------------------------------------------------------
# Bucket
# ---------------------------------------------------------------------------------------
class HydroBucket(Component):
    """`@hydrobucket` (`/root/reference/src/bucket.jl:24-65`).

    `htypes=None` → single-node, 2D input `[variables, time]`; `htypes=[…]` (1-based HRU
    type per node) → multi-node, 3D input `[variables, nodes, time]` with parameter sharing
    `p[htypes]` (`src/utils.jl:174-201`)."""

    def __init__(self, *, fluxes: Sequence[Component], dfluxes: Sequence[StateFlux] = (),
                 htypes=None, name: Optional[str] = None):
        self.fluxes = list(fluxes)
        self.dfluxes = list(dfluxes)
        for f in self.fluxes:
            if not isinstance(f, (HydroFlux, NeuralFlux)):
                raise TypeError("fluxes must be HydroFlux / NeuralFlux components")
        for d in self.dfluxes:
            if not isinstance(d, StateFlux):
                raise TypeError("dfluxes must be StateFlux components")
        self.state_names = [d.state for d in self.dfluxes]
        self.output_names = _uniq(o for f in self.fluxes for o in f.output_names)
        produced = set(self.output_names) | set(self.state_names)
        self.input_names = [v for v in _uniq(
            [i for f in self.fluxes for i in f.input_names] + [i for d in self.dfluxes for i in d.input_names])
            if v not in produced]
        self.param_names = _uniq([p for f in self.fluxes for p in f.param_names] +
                                 [p for d in self.dfluxes for p in d.param_names])
        self.nn_names = _uniq(n for f in self.fluxes for n in f.nn_names)
        self.neural_fluxes = [f for f in self.fluxes if isinstance(f, NeuralFlux)]
        self.htypes = _check_htypes(htypes)
        self.name = name or f"##bucket#{abs(hash((tuple(self.input_names), tuple(self.output_names)))):x}"
        self._check_order()

    @property
    def has_states(self) -> bool:
        return bool(self.state_names)

    def _check_order(self):
        """Fluxes are evaluated in declaration order (the reference never calls
        `sort_fluxes`, `src/bucket.jl:45-58`); a flux reading a later flux's output cannot be
        evaluated — reject instead of guessing."""
        known = set(self.input_names) | set(self.state_names)
        for f in self.fluxes:
            for i in f.input_names:
                if i not in known:
                    raise ValueError(
                        f"flux {f.output_names} of bucket {self.name} reads '{i}' before it is defined; "
                        "declare fluxes in dependency order")
            known |= set(f.output_names)


# ---------------------------------------------------------------------------------------
# Unit hydrograph
# ---------------------------------------------------------------------------------------
class UnitHydrograph(Component):
    """`@unithydro` (`/root/reference/src/uh.jl:21-88,100-181`).

    `uh_conds` = list of `(threshold_expr, value_expr)` in the time symbol `t`, **largest
    threshold first** (`max_lag` = first threshold, `src/uh.jl:65,172`).  Semantics of the
    HydroModelCore-generated `uh_func`/`max_lag` (standard GR4J S-curves):
    `K = ceil(max_lag(p))`; `S(k)` = value of the condition with the smallest threshold
    `>= k`, and `1` beyond the largest threshold; weights `w_k = S(k) - S(k-1)`, normalised.
    """

    def __init__(self, input: Sym, output: Sym, uh_conds: Sequence[Tuple[object, object]], *,
                 symbols: Optional[Symbols] = None, tvar: str = "t", htypes=None, name: Optional[str] = None):
        if not uh_conds:
            raise ValueError("Missing uh_func in unit hydrograph definition")
        self.tvar = tvar
        table = dict(symbols.table) if symbols is not None else {}
        table.setdefault(tvar, Sym(tvar, False))

        def ex(x):
            return parse_expr(x, table) if isinstance(x, str) else as_expr(x)

        self.uh_conds = [(ex(c), ex(v)) for c, v in uh_conds]
        self.input_names = [input.name]
        self.output_names = [output.name]
        ps = []
        for c, v in self.uh_conds:
            ps += free_symbols(c)[1] + free_symbols(v)[1]
        self.param_names = _uniq(ps)
        self.state_names = []
        self.nn_names = []
        self.htypes = _check_htypes(htypes)
        self.name = name or f"##uh#{input.name}_{output.name}"

    def max_lag(self, p: Dict[str, float]) -> int:
        return int(math.ceil(evaluate(self.uh_conds[0][0], p)))

    def uh_func(self, k: int, p: Dict[str, float]) -> float:
        env = dict(p)
        env[self.tvar] = float(k)
        conds = sorted(((evaluate(c, p), v) for c, v in self.uh_conds), key=lambda cv: cv[0])
        for thr, v in conds:
            if k <= thr:
                return evaluate(v, env)
        return 1.0

    def weights(self, p: Dict[str, float]) -> np.ndarray:
        """Normalised UH ordinates for one node (`src/uh.jl:215-229` + `hydro_conv` `:188`)."""
        K = self.max_lag(p)
        if K <= 0:
            return np.zeros(0)
        S = np.array([self.uh_func(k, p) for k in range(1, K + 1)])
        w = np.empty(K)
        w[0] = S[0]
        w[1:] = S[1:] - S[:-1]
        return w / w.sum()


# ---------------------------------------------------------------------------------------
# Routing
# ---------------------------------------------------------------------------------------
D8_CODES = (1, 2, 4, 8, 16, 32, 64, 128)
# code -> (d_row, d_col); derived from the pad table of `build_aggr_func`
# (`/root/reference/src/route.jl:343-347`; same table `ext/HydroModelsRastersExt/…jl:40-52`)
D8_OFFSETS = {1: (0, 1), 2: (1, 1), 4: (1, 0), 8: (1, -1), 16: (0, -1), 32: (-1, -1), 64: (-1, 0), 128: (-1, 1)}


@dataclass
class GridAggregation:
    """`build_aggr_func(flwdir, positions)` (`/root/reference/src/route.jl:342-365`).
    `positions[n] = (row, col)`, 1-based as in the reference; node order = list order."""
    flwdir: np.ndarray
    positions: np.ndarray  # [N, 2] 1-based

    def downstream(self) -> np.ndarray:
        """0-based index of the node each node drains to, or -1 (off-grid / unlisted cell /
        code not in the D8 table → flow is dropped)."""
        R, C = self.flwdir.shape
        pos0 = self.positions - 1
        cell = -np.ones((R, C), dtype=np.int64)
        cell[pos0[:, 0], pos0[:, 1]] = np.arange(len(pos0))
        codes = self.flwdir[pos0[:, 0], pos0[:, 1]]
        dr = np.zeros(len(pos0), dtype=np.int64)
        dc = np.zeros(len(pos0), dtype=np.int64)
        valid = np.zeros(len(pos0), dtype=bool)
        for code, (a, b) in D8_OFFSETS.items():
            m = codes == code
            dr[m], dc[m], valid[m] = a, b, True
        r = pos0[:, 0] + dr
        c = pos0[:, 1] + dc
        valid &= (r >= 0) & (r < R) & (c >= 0) & (c < C)
        down = -np.ones(len(pos0), dtype=np.int64)
        down[valid] = cell[r[valid], c[valid]]
        return down


@dataclass
class GraphAggregation:
    """`build_aggr_func(graph::DiGraph)` = `adjacency' * outflow`
    (`/root/reference/src/route.jl:336-340`).  `edges` = 1-based `(src, dst)` pairs."""
    n_nodes: int
    edges: np.ndarray  # [E, 2] 1-based


def build_aggr_func(*args):
    """`build_aggr_func(flwdir, positions)` or `build_aggr_func(n_nodes, edges)`."""
    a, b = args
    if np.ndim(a) == 2:
        flw = np.asarray(a).astype(np.int64)
        pos = (np.asarray(b, dtype=np.int64) if isinstance(b, np.ndarray)
               else np.asarray([tuple(p) for p in b], dtype=np.int64)).reshape(-1, 2)
        R, C = flw.shape
        if pos.size and (pos.min() < 1 or pos[:, 0].max() > R or pos[:, 1].max() > C):
            raise IndexError("positions outside flwdir")  # BoundsError in the reference
        return GridAggregation(flw, pos)
    edges = np.asarray(b, dtype=np.int64).reshape(-1, 2)
    n = int(a)
    if edges.size and (edges.min() < 1 or edges.max() > n):
        raise IndexError("edge endpoint outside 1..n_nodes")
    return GraphAggregation(n, edges)


class HydroRoute(Component):
    """`@hydroroute` (`/root/reference/src/route.jl:237-334,367-404`): per step every node's
    outflow is computed from its *previous* state, pushed one hop downstream and added to the
    downstream node's increment in the same step (Jacobi).  Restricted to
    one rflux output (the outflow) and one state, as in every reference example."""

    def __init__(self, *, rfluxes: Sequence[HydroFlux], dfluxes: Sequence[StateFlux], aggr_func,
                 htypes, name: Optional[str] = None):
        if htypes is None or len(htypes) == 0:
            raise AssertionError("htypes must be specified for HydroRoute")
        if not isinstance(aggr_func, (GridAggregation, GraphAggregation)):
            raise TypeError("aggr_func must come from build_aggr_func(...)")
        self.rfluxes = list(rfluxes)
        self.dfluxes = list(dfluxes)
        self.aggr = aggr_func
        self.state_names = [d.state for d in self.dfluxes]
        self.output_names = _uniq(o for f in self.rfluxes for o in f.output_names)
        if len(self.output_names) != 1 or len(self.state_names) != 1:
            raise NotImplementedError("HydroRoute supports exactly one rflux output and one state "
                                      "(layout for more is ambiguous in the reference, SURVEY.md App. C.4)")
        produced = set(self.output_names) | set(self.state_names)
        self.input_names = [v for v in _uniq(
            [i for f in self.rfluxes for i in f.input_names] + [i for d in self.dfluxes for i in d.input_names])
            if v not in produced]
        self.param_names = _uniq([p for f in self.rfluxes for p in f.param_names] +
                                 [p for d in self.dfluxes for p in d.param_names])
        self.nn_names = []
        self.htypes = _check_htypes(htypes)
        self.name = name or f"##route#{self.output_names[0]}"


# ---------------------------------------------------------------------------------------
# Model
# ---------------------------------------------------------------------------------------
class HydroModel(Component):
    """`@hydromodel` (`/root/reference/src/model.jl:21-109`): an ordered tuple of components.

    Row bookkeeping follows `_prepare_indices` (`src/model.jl:74-109`): component *k* finds
    its inputs by name among `[model inputs; states+outputs of components < k]` (first match
    wins); the result rows are `[all states (component order); all outputs (component
    order)]`.  `input_names=` pins the model-input row order (in the reference it is whatever
    `get_input_names(model)` returns; the caller always consults it)."""

    def __init__(self, *, components: Sequence[Component], name: Optional[str] = None,
                 input_names: Optional[Sequence[str]] = None):
        self.components = tuple(components)
        if not self.components:
            raise ValueError("HydroModel needs at least one component")
        produced: List[str] = []
        inputs: List[str] = []
        for c in self.components:
            for i in c.input_names:
                if i not in produced and i not in inputs:
                    inputs.append(i)
            produced += c.state_names + c.output_names
        if input_names is not None:
            if set(input_names) != set(inputs):
                raise ValueError(f"input_names must be a permutation of {inputs}")
            inputs = list(input_names)
        self.input_names = inputs
        self.state_names = _uniq(s for c in self.components for s in c.state_names)
        self.output_names = _uniq(o for c in self.components for o in c.output_names)
        self.param_names = _uniq(p for c in self.components for p in c.param_names)
        self.nn_names = _uniq(n for c in self.components for n in c.nn_names)
        self.name = name or "##model"
        multi = [c.is_multinode for c in self.components if not isinstance(c, NeuralFlux)]   # a NeuralFlux takes 2D or 3D
        if any(multi) and not all(multi):
            raise ValueError("mixing single-node and multi-node (htypes) components in one model")
        self.htypes = self.components[0].htypes if all(multi) else None
        self._varindices, self._outputindices = self._prepare_indices()

    def _prepare_indices(self):
        available = list(self.input_names)
        input_idx = []
        for c in self.components:
            idx = []
            for nm in c.input_names:
                if nm in available:
                    idx.append(available.index(nm))
                # missing inputs only warn in the reference (`src/model.jl:90-93`); here they
                # cannot occur because model inputs are derived from the components
            input_idx.append(idx)
            available += c.state_names + c.output_names
        out_idx = [available.index(nm) for nm in self.state_names + self.output_names if nm in available]
        return input_idx, out_idx

    @property
    def var_names(self) -> List[str]:
        """Row names of the result: `vcat(get_state_names, get_output_names)`."""
        return self.state_names + self.output_names
