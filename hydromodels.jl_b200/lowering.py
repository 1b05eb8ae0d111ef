"""Lowering: HydroModel expression trees → the CUDA C body spliced into `templates/stepper.cu`.

This is the host-side half of the engine (the north-star's "Julia host code lowers the model's
Symbolics flux and state expressions to a CUDA C body"; a Julia twin of this file is sketched in
INTEGRATION.md).  It replaces HydroModelCore's `build_bucket_func` / `build_flux_func` /
`build_uh_func` code generators (call sites `/root/reference/src/bucket.jl:58`, `src/flux.jl:52`,
`src/uh.jl:73`) — but instead of one Julia closure per component evaluated over the whole
series, it emits ONE straight-line step for the whole model:

    for each component, in declaration order (`/root/reference/src/model.jl:255-268`):
        bucket with states : fluxes(pre-state) → dS → u ← max(min_value, u + dS) → fluxes(post-state)
                             (`src/solve.jl:47-52`, `src/bucket.jl:184-192`: every state-dependent
                             flux is evaluated twice per step)
        stateless bucket / flux : one evaluation (`src/bucket.jl:196-204`, `src/flux.jl:157-166`)
        unit hydrograph : y_t = Σ_k w_k x_{t-k+1} on a register ring (`src/uh.jl:187-237`)
        neural flux : dense MLP on the listed inputs (`src/nn.jl:56-58`)

Work the straight-line form makes removable (all value-preserving to ≤ 1e-15 relative):

* hash-consing CSE across the pre-/post-state evaluations: anything that does not depend on the
  bucket's own states (e.g. `pet`, `snowfall`) is computed once per step instead of twice;
* parameter-only subexpressions (and reciprocals of parameter-only divisors) are hoisted to the
  prologue (class P);
* state-only subexpressions (e.g. `step_func(soilwater)`, `exp(-f*max(0,Smax-soilwater))`): the
  post-update value at step t IS the pre-update value at step t+1 → carried in a register
  (`hb_c[k]`) instead of being recomputed (class S);
* `step_func` → logistic form (one exp + one reciprocal), integer / half-integer powers →
  multiplications and `sqrt`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .components import (Component, HydroBucket, HydroFlux, HydroModel, HydroRoute, NeuralBucket, NeuralFlux, StateFlux,
                         UnitHydrograph)
from .symbolic import Call, Const, Expr, Sym, evaluate, walk

M_X, M_S, M_U = 1, 2, 4      # dependency bits: forcing, states, unit-hydrograph history

_COST = {"add": 1, "sub": 1, "mul": 1, "neg": 0, "abs": 0, "min": 1, "max": 1, "clamp": 2, "div": 10, "rcp": 10,
         "sqrt": 12, "exp": 25, "log": 30, "tanh": 40, "step": 36, "pow": 80, "fma": 1, "leakyrelu": 2}


class LoweringError(ValueError):
    pass


def _c_double(v: float) -> str:
    if math.isnan(v):
        return "(0.0/0.0)"
    if math.isinf(v):
        return "(1.0/0.0)" if v > 0 else "(-1.0/0.0)"
    s = repr(float(v))
    if "e" not in s and "." not in s:
        s += ".0"
    return s


@dataclass
class Val:
    """One SSA value of the per-step program."""
    id: int
    op: str                  # 'const' 'in' 'param' 'state' 'uh' 'nn' or an arithmetic op
    args: Tuple[int, ...] = ()
    payload: object = None
    mask: int = 0
    cost: int = 0            # cumulative cost of the subtree (for the carry heuristic)
    erased: object = None    # structural key with state versions erased
    vers: frozenset = frozenset()   # state versions appearing below


@dataclass
class StepperProgram:
    body: str
    input_names: List[str]
    row_names: List[str]            # rows written per step, in order
    all_row_names: List[str]
    state_names: List[str]          # order of hb_u / initstates / final_states
    param_slots: List[Tuple[str, Optional[int]]]   # hb_p[j] = (parameter name, htypes-set index | None)
    htype_sets: List[np.ndarray]    # distinct 1-based htypes vectors
    uh: List[UnitHydrograph]
    uh_kmax: List[int]
    uh_htype_set: List[Optional[int]]
    nn_layout: List[Tuple[str, int, int]]          # (chain name, offset, n_params) in the smem block
    nn_words: int
    n_carry: int
    objective: bool
    stats: Dict[str, int] = field(default_factory=dict)
    nn_tensor: bool = False         # some dense chain is lowered to FP64 MMA fragments
    precision: str = "f64"          # "f64" (reference arithmetic) | "f32" (opt-in)
    uh_on_device: bool = False      # the prologue evaluates the unit-hydrograph ordinates (host table ignored)
    tangents: List[str] = field(default_factory=list)   # gradient run: parameter name of each direction
    nn_tc_bytes: int = 0            # Float32 tensor path (tcgen05): shared-memory bytes of the A and B tiles (0 = not used)
    aux_names: List[str] = field(default_factory=list)   # rows copied from the auxiliary input stream (hb_aux[k])


class _Builder:
    def __init__(self, fast: bool):
        self.fast = fast
        self.vals: List[Val] = []
        self.memo: Dict[object, int] = {}
        self.nn_calls: List[Tuple[int, Tuple[int, ...], List[int]]] = []   # (net index, input ids, output val ids)

    # ---- value construction -------------------------------------------------------------------
    def _new(self, key, op, args=(), payload=None, mask=0, cost=0, erased=None, vers=frozenset()) -> int:
        if key in self.memo:
            return self.memo[key]
        v = Val(len(self.vals), op, tuple(args), payload, mask, cost, erased if erased is not None else key, vers)
        self.vals.append(v)
        self.memo[key] = v.id
        return v.id

    def const(self, x: float) -> int:
        x = float(x)
        return self._new(("const", x.hex()), "const", payload=x)

    def leaf_in(self, idx: int) -> int:
        return self._new(("in", idx), "in", payload=idx, mask=M_X)

    def leaf_param(self, idx: int) -> int:
        return self._new(("param", idx), "param", payload=idx)

    def leaf_state(self, idx: int, ver: int) -> int:
        return self._new(("state", idx, ver), "state", payload=(idx, ver), mask=M_S, erased=("state", idx),
                         vers=frozenset([ver]))

    def opaque(self, kind: str, payload, mask: int, cost: int = 1000) -> int:
        return self._new((kind, payload), kind, payload=payload, mask=mask, cost=cost)

    def is_const(self, i: int) -> bool:
        return self.vals[i].op == "const"

    def cval(self, i: int) -> float:
        return self.vals[i].payload

    def op(self, name: str, *args: int) -> int:
        a = list(args)
        if all(self.is_const(i) for i in a):           # constant folding (IEEE double, libm)
            return self.const(_fold(name, [self.cval(i) for i in a]))
        if name in ("add", "mul", "min", "max") and a[0] > a[1]:
            a = [a[1], a[0]]                            # commutative normal form → more CSE hits
        key = (name,) + tuple(a)
        mask = 0
        cost = _COST.get(name, 1)
        vers = frozenset()
        for i in a:
            mask |= self.vals[i].mask
            cost += self.vals[i].cost
            vers |= self.vals[i].vers
        er = [self.vals[i].erased for i in a]
        if name in ("add", "mul", "min", "max"):
            er.sort(key=repr)                           # pre/post twins must not depend on value ids
        erased = (name,) + tuple(er)
        return self._new(key, name, a, mask=mask, cost=cost, erased=erased, vers=vers)

    # ---- expression trees ---------------------------------------------------------------------
    def lower_expr(self, e: Expr, env: Dict[str, int], params: Dict[str, int]) -> int:
        memo: Dict[object, int] = {}
        for n in walk(e):
            k = n.key()
            if isinstance(n, Const):
                memo[k] = self.const(n.value)
            elif isinstance(n, Sym):
                table = params if n.is_param else env
                if n.name not in table:
                    what = "Parameter" if n.is_param else "Variable"
                    raise LoweringError(f"{what} {n.name} is not defined where it is used")
                memo[k] = table[n.name]
            else:
                memo[k] = self._lower_call(n.op, [memo[x.key()] for x in n.args])
        return memo[e.key()]

    def _lower_call(self, op: str, a: List[int]) -> int:
        if op == "div":
            if self.fast and self.vals[a[1]].mask == 0 and not self.is_const(a[0]):
                return self.op("mul", a[0], self.op("rcp", a[1]))   # divisor is parameter-only → hoisted 1/b
            return self.op("div", a[0], a[1])
        if op == "pow":
            return self._lower_pow(a[0], a[1])
        if op == "clamp":
            return self.op("min", self.op("max", a[0], a[1]), a[2])
        if op == "step" and self.fast:
            # step_func(a - b) + step_func(b - a) == 1 (logistic symmetry): the snow/rain split of every
            # reference model evaluates both — the second one costs a subtraction instead of exp + rcp
            x = self.vals[a[0]]
            if x.op == "sub":
                twin = self.memo.get(("sub", x.args[1], x.args[0]))
                if twin is not None and ("step", twin) in self.memo:
                    return self.op("sub", self.const(1.0), self.memo[("step", twin)])
        return self.op(op, *a)

    def _lower_pow(self, base: int, ex: int) -> int:
        if self.is_const(ex):
            p = self.cval(ex)
            if p == int(p) and abs(p) <= 16:
                n = int(abs(p))
                if n == 0:
                    return self.const(1.0)
                r = self._ipow(base, n)
                return r if p > 0 else self.op("div", self.const(1.0), r)
            if self.fast and p * 2 == int(p * 2) and 0 < p <= 8.5:     # half-integer: x^n * sqrt(x)
                n = int(p)
                s = self.op("sqrt", base)
                return s if n == 0 else self.op("mul", self._ipow(base, n), s)
        return self.op("pow", base, ex)

    def _ipow(self, base: int, n: int) -> int:
        result = None
        sq = base
        while n:
            if n & 1:
                result = sq if result is None else self.op("mul", result, sq)
            n >>= 1
            if n:
                sq = self.op("mul", sq, sq)
        return result


def _fold(name: str, a: List[float]) -> float:
    try:
        if name == "add": return a[0] + a[1]
        if name == "sub": return a[0] - a[1]
        if name == "mul": return a[0] * a[1]
        if name == "div": return a[0] / a[1] if a[1] != 0 else math.copysign(math.inf, a[0]) * math.copysign(1.0, a[1])
        if name == "rcp": return 1.0 / a[0] if a[0] != 0 else math.copysign(math.inf, a[0])
        if name == "neg": return -a[0]
        if name == "pow": return math.pow(a[0], a[1])
        if name == "exp": return math.exp(a[0])
        if name == "log": return math.log(a[0])
        if name == "sqrt": return math.sqrt(a[0])
        if name == "abs": return abs(a[0])
        if name == "tanh": return math.tanh(a[0])
        if name == "step": return (math.tanh(5.0 * a[0]) + 1.0) * 0.5
        if name == "min": return min(a[0], a[1])
        if name == "max": return max(a[0], a[1])
    except (OverflowError, ValueError):
        return math.nan
    raise LoweringError(f"cannot fold {name}")


# ---------------------------------------------------------------------------------------------
# model → program
# ---------------------------------------------------------------------------------------------
def _flatten(component: Component) -> Tuple[List[Component], List[str], List[str]]:
    if isinstance(component, HydroModel):
        return list(component.components), list(component.input_names), list(component.var_names)
    return [component], list(component.input_names), list(component.state_names) + list(component.output_names)


def lower_stepper(component: Component, *, rows: Optional[Sequence[str]] = None, objective: bool = False,
                  uh_kmax: Optional[Sequence[int]] = None, fast: bool = True,
                  holes: Sequence[str] = (), nn_tensor: bool = False, precision: str = "f64",
                  uh_on_device: bool = False,
                  component_min_values: Optional[Sequence[Optional[float]]] = None,
                  tangents: Optional[Sequence[str]] = None, aux_rows: Sequence[str] = (), adjoint: bool = False):
    """Lower a per-node model (buckets, fluxes, neural fluxes, unit hydrographs) to one stepper
    body.  `rows`: subset/order of result rows to write (default: all = the reference result);
    `objective`: emit the last output row as `hb_sim` for the fused metric reduction instead of
    rows; `uh_kmax[u]`: ordinates per unit hydrograph (from the run's parameters);
    `component_min_values[c]`: the state floor of component c when the call carries per-component configs
    (`/root/reference/src/model.jl:140-152`) — baked in as a literal; None = the run's `min_value` argument;
    `tangents`: parameter names to differentiate the objective by (gradient run: the body's temporaries become
    `auto` / `hb_dual`, see hb_math.cuh; objective mode, Float64, no neural components);
    `aux_rows`: rows (a subset of `holes`) whose values arrive through the kernel's auxiliary input stream
    (`aux_input[T][N][len(aux_rows)]`, in this order) and are copied into the records instead of 0.0."""
    comps, input_names, all_rows = _flatten(component)
    aux_rows = list(aux_rows)
    if any(r not in holes for r in aux_rows):
        raise LoweringError("aux_rows must be rows another kernel fills (holes)")
    if component_min_values is not None and len(component_min_values) != len(comps):
        raise LoweringError("Component configs length must equal components length")
    for c in comps:
        if isinstance(c, HydroRoute):
            raise LoweringError("HydroRoute is a grid-coupled component; it is lowered by lower_route")
        if not isinstance(c, (HydroBucket, HydroFlux, UnitHydrograph, NeuralBucket, NeuralFlux)):
            raise LoweringError(f"cannot lower component of type {type(c).__name__}")
    b = _Builder(fast)

    # ---- bookkeeping tables -------------------------------------------------------------------
    htype_sets: List[np.ndarray] = []

    def htype_set(c: Component) -> Optional[int]:
        if c.htypes is None:
            return None
        for i, h in enumerate(htype_sets):
            if h.shape == c.htypes.shape and np.array_equal(h, c.htypes):
                return i
        htype_sets.append(np.asarray(c.htypes))
        return len(htype_sets) - 1

    param_slots: List[Tuple[str, Optional[int]]] = []

    def param_table(c: Component) -> Dict[str, int]:
        hs = htype_set(c)
        out = {}
        for nm in c.param_names:
            slot = (nm, hs)
            if slot not in param_slots:
                param_slots.append(slot)
            out[nm] = b.leaf_param(param_slots.index(slot))
        return out

    state_names: List[str] = [s for c in comps for s in c.state_names]
    if len(set(state_names)) != len(state_names):
        raise LoweringError("duplicate state names across components")
    uhs = [c for c in comps if isinstance(c, UnitHydrograph)]
    if uh_kmax is None:
        uh_kmax = [1] * len(uhs)
    if len(uh_kmax) != len(uhs):
        raise LoweringError("uh_kmax must have one entry per unit hydrograph")
    uh_kmax = [max(1, int(k)) for k in uh_kmax]
    uh_w_off = np.concatenate([[0], np.cumsum(uh_kmax)]).astype(int)
    uh_h_off = np.concatenate([[0], np.cumsum([k - 1 for k in uh_kmax])]).astype(int)
    nets: List = []           # distinct chains
    nn_layout: List[Tuple[str, int, int]] = []

    def net_index(chain) -> int:
        for i, ch in enumerate(nets):
            if ch.name == chain.name:
                if ch != chain:
                    raise LoweringError(f"two different chains share the name {chain.name}")
                return i
        off = sum(n for _, _, n in nn_layout)
        nets.append(chain)
        nn_layout.append((chain.name, off, chain.n_params))
        return len(nets) - 1

    # ---- walk the components in order, building the step DAG ------------------------------------
    env: Dict[str, int] = {nm: b.leaf_in(i) for i, nm in enumerate(input_names)}
    state_post: Dict[int, int] = {}           # state index → val id of the post-update value
    uh_steps: List[Tuple[int, int]] = []      # (uh index, val id of x_t)
    uh_device: List[Tuple[int, List[int], List[List[int]]]] = []   # (uh index, threshold vals, value vals per k)
    statements: List[Tuple[str, object]] = []  # ordered side effects ('update', (sidx, val)), …

    raw_states = set()                        # states updated without the min_value clamp (NeuralBucket)
    state_floor: Dict[int, float] = {}        # state index → literal floor (per-component configs)

    def nn_outputs(chain, ins: Tuple[int, ...]) -> List[int]:
        ni = net_index(chain)
        mask = 0
        for i in ins:
            mask |= b.vals[i].mask
        out = []
        for j in range(chain.n_out):
            vid = b.opaque("nn", (ni, ins, j), mask)
            # erased key / versions so that pre/post twins of an NN output can be carried
            v = b.vals[vid]
            v.erased = ("nn", ni, tuple(b.vals[i].erased for i in ins), j)
            v.vers = frozenset().union(*[b.vals[i].vers for i in ins]) if ins else frozenset()
            v.args = ins
            out.append(vid)
        return out

    def neural_flux_outputs(f: NeuralFlux, local: Dict[str, int]) -> List[int]:
        """`chain(norm(inputs))` (`/root/reference/src/nn.jl:56-58,148,161`)."""
        if f.norm_exprs is not None:
            ins = tuple(b.lower_expr(e, {i: local[i] for i in f.input_names}, {}) for e in f.norm_exprs)
        else:
            ins = tuple(local[i] for i in f.input_names)
        return nn_outputs(f.chain, ins)

    def eval_fluxes(bucket: HydroBucket, local: Dict[str, int], ptab: Dict[str, int]) -> Dict[str, int]:
        outs = {}
        for f in bucket.fluxes:
            if isinstance(f, NeuralFlux):
                for o, vid in zip(f.output_names, neural_flux_outputs(f, local)):
                    local[o] = outs[o] = vid
            else:
                for eq in f.eqs:
                    vid = b.lower_expr(eq.rhs, local, ptab)
                    local[eq.lhs.name] = outs[eq.lhs.name] = vid
        return outs

    for ci, c in enumerate(comps):
        missing = [i for i in c.input_names if i not in env]
        if missing:
            raise LoweringError(f"Input variables {missing} of component {c.name} not found")
        if isinstance(c, HydroBucket):
            ptab = param_table(c)
            if not c.has_states:
                local = {i: env[i] for i in c.input_names}
                env.update(eval_fluxes(c, local, ptab))
                continue
            sidx = [state_names.index(s) for s in c.state_names]
            if component_min_values is not None and component_min_values[ci] is not None:
                for si in sidx:
                    state_floor[si] = float(component_min_values[ci])
            # pass 1: pre-update states
            local = {i: env[i] for i in c.input_names}
            for s, si in zip(c.state_names, sidx):
                local[s] = b.leaf_state(si, 0)
            eval_fluxes(c, local, ptab)
            dus = [b.lower_expr(d.expr, local, ptab) for d in c.dfluxes]
            # update: u ← max(min_value, u + du)
            for s, si, du in zip(c.state_names, sidx, dus):
                post = b.leaf_state(si, 1)
                state_post[si] = post
                statements.append(("update", (si, b.leaf_state(si, 0), du)))
            # pass 2: post-update states → rows
            local = {i: env[i] for i in c.input_names}
            for s, si in zip(c.state_names, sidx):
                local[s] = state_post[si]
            outs = eval_fluxes(c, local, ptab)
            for s, si in zip(c.state_names, sidx):
                env[s] = state_post[si]
            env.update(outs)
        elif isinstance(c, NeuralBucket):
            # /root/reference/src/nn.jl:270-291: fluxes from the PREVIOUS state, residual state update (no clamp),
            # outputs from those same fluxes; rows = [new states; outputs] (:308-326)
            htype_set(c)
            sidx = [state_names.index(s) for s in c.state_names]
            pre = tuple(b.leaf_state(si, 0) for si in sidx)
            xs = tuple(env[i] for i in c.input_names)
            flux = tuple(nn_outputs(c.flux_network, pre + xs))
            delta = nn_outputs(c.state_network, pre + flux)
            for si, du in zip(sidx, delta):
                state_post[si] = b.leaf_state(si, 1)
                raw_states.add(si)
                statements.append(("update", (si, b.leaf_state(si, 0), du)))
            for o, vid in zip(c.output_names, nn_outputs(c.output_network, flux)):
                env[o] = vid
            for s, si in zip(c.state_names, sidx):
                env[s] = state_post[si]
        elif isinstance(c, NeuralFlux):
            # stand-alone / model-level neural flux (`/root/reference/src/nn.jl:140-165`): stateless, one evaluation per step
            for o, vid in zip(c.output_names, neural_flux_outputs(c, {i: env[i] for i in c.input_names})):
                env[o] = vid
        elif isinstance(c, HydroFlux):
            ptab = param_table(c)
            local = {i: env[i] for i in c.input_names}
            for eq in c.eqs:
                vid = b.lower_expr(eq.rhs, local, ptab)
                local[eq.lhs.name] = vid
                env[eq.lhs.name] = vid
        else:  # UnitHydrograph
            u = uhs.index(c)
            if uh_on_device:
                ptab = param_table(c)
                thr = [b.lower_expr(cnd, {}, ptab) for cnd, _ in c.uh_conds]
                vals = [[b.lower_expr(v, {c.tvar: b.const(float(k))}, ptab) for _, v in c.uh_conds] for k in range(1, uh_kmax[u] + 1)]
                uh_device.append((u, thr, vals))
            x = env[c.input_names[0]]
            vid = b.opaque("uh", (u, x), b.vals[x].mask | M_U, cost=2 * uh_kmax[u])
            b.vals[vid].args = (x,)
            uh_steps.append((u, x))
            env[c.output_names[0]] = vid

    # ---- outputs ------------------------------------------------------------------------------------
    if objective:
        row_names = [all_rows[-1]]           # `[end, :]` of the result (`…OptimizationExt.jl:163,173`)
    else:
        row_names = list(all_rows) if rows is None else list(rows)
    zero = b.const(0.0)
    for r in row_names:
        if r not in env and r not in holes:
            raise LoweringError(f"row {r} is not produced by the model")
    # `holes`: rows another kernel fills in later (the route's state / outflow rows) — written as 0.0
    out_vals = [env[r] if r in env else zero for r in row_names]

    if adjoint:
        # reverse-mode program of the same step DAG (templates/adjoint.cu): see _emit_adjoint
        if not objective or precision != "f64" or not fast:
            raise LoweringError("the adjoint run differentiates the objective row, in Float64, with the hb_math helpers")
        if uhs or any(isinstance(c, NeuralBucket) for c in comps):
            raise LoweringError("the adjoint run covers HydroBucket / HydroFlux / NeuralFlux components (no unit hydrographs, no NeuralBucket)")
        return _emit_adjoint(b, out_vals[0], statements, param_slots, nets, nn_layout, raw_states, state_floor, input_names,
                             state_names, htype_sets)

    # ---- carry detection (class S) --------------------------------------------------------------------
    by_erased: Dict[object, Dict[int, int]] = {}
    for v in b.vals:
        if v.mask == M_S and len(v.vers) == 1 and v.op not in ("state",):
            by_erased.setdefault(v.erased, {})[next(iter(v.vers))] = v.id
    users: Dict[int, List[int]] = {v.id: [] for v in b.vals}
    for v in b.vals:
        for a in v.args:
            users[a].append(v.id)
    root_uses = set(out_vals) | {du for k, p in statements if k == "update" for du in (p[2],)} | {x for _, x in uh_steps}
    carried: Dict[int, Tuple[int, int]] = {}    # pre val id → (carry slot, post val id)
    if fast:
        for erased, d in by_erased.items():
            if 0 in d and 1 in d:
                pre, post = d[0], d[1]
                if b.vals[pre].cost < 6:
                    continue
                # maximal: used by something that is not itself a pure pre-state expression, or a root
                maximal = pre in root_uses or any(
                    not (b.vals[u].mask == M_S and b.vals[u].vers == frozenset([0])) for u in users[pre])
                # if a user is itself carried, this one is not needed separately
                if maximal:
                    carried[pre] = (len(carried), post)

    # ---- liveness (dead-code elimination) from the roots ----------------------------------------------
    live = set()
    stack = list(out_vals) + [p[2] for k, p in statements if k == "update"] + [x for _, x in uh_steps] + \
        [post for _, post in carried.values()]

    def visit(i):
        st = [i]
        while st:
            j = st.pop()
            if j in live:
                continue
            live.add(j)
            if j in carried:
                continue                     # value comes from hb_c[], its subtree is not needed
            st.extend(b.vals[j].args)
    for i in stack:
        visit(i)
    # pre-state subtrees needed only to initialise carries are emitted in the prologue
    # ---- emission ---------------------------------------------------------------------------------------
    pro: List[str] = []
    step: List[str] = []
    pro_names: Dict[int, str] = {}

    # dense chains whose shape fits the FP64 tensor-core fragments are contracted with DMMA
    if precision not in ("f64", "f32"):
        raise LoweringError("precision must be 'f64' or 'f32'")
    net_mma = [bool(nn_tensor and fast and precision == "f64" and _mma_eligible(ch)) for ch in nets]
    # Float32 mode: dense layers of tensor-core shape go through tcgen05.mma (kind::tf32, accumulator in TMEM), the CTA's
    # 128 basins being the M dimension (templates/hb_tc.cuh)
    net_tc = [bool(nn_tensor and fast and precision == "f32" and _tc_layers(ch)) for ch in nets]

    def nn_call(ni: int, args: List[str], arr: str) -> str:
        pre = "sNN, hb_scr, lane, " if net_mma[ni] else ("hb_tc, " if net_tc[ni] else "sNN, ")
        return f"hb_mlp_{ni}({pre}{', '.join(args)}, {arr});"

    def nn_var(ni: int, ins: Tuple[int, ...], tag: str) -> str:
        return f"nn{tag}{ni}_" + "_".join(str(k) for k in ins)

    def pro_ref(i: int) -> str:
        """Name of val i evaluated in the prologue: states read the (unclamped) initial values."""
        v = b.vals[i]
        if v.op == "const": return _c_double(v.payload)
        if v.op == "param": return f"hb_p[{v.payload}]"
        if v.op == "state": return f"hb_u[{v.payload[0]}]"
        if v.op in ("in", "uh") or (v.mask & (M_X | M_U)):
            raise LoweringError("internal: prologue value depends on forcing")
        if i in pro_names:
            return pro_names[i]
        args = [pro_ref(a) for a in v.args]
        nm = f"h{i}"
        if v.op == "nn":
            ni, ins, j = v.payload
            arr = nn_var(ni, ins, "p")
            if arr not in pro_names.values():
                pro.append(f"double {arr}[{nets[ni].n_out}]; " + nn_call(ni, args, arr))
                pro_names[-len(pro_names) - 1] = arr
            pro.append(f"const double {nm} = {arr}[{j}];")
        else:
            pro.append(f"const double {nm} = {_expr_with(v, args, fast)};")
        pro_names[i] = nm
        return nm

    def step_ref(i: int) -> str:
        v = b.vals[i]
        if v.op == "const": return _c_double(v.payload)
        if v.op == "in": return f"hb_x[{v.payload}]"
        if v.op == "param": return f"hb_p[{v.payload}]"
        if v.op == "state":
            si, ver = v.payload
            return f"hb_u[{si}]" if ver == 0 else f"hb_un{si}"
        if i in carried: return f"hb_c[{carried[i][0]}]"
        if v.mask == 0: return pro_ref(i)          # class P: hoisted
        return f"v{i}"

    # carries: initial value = the pre-state expression at the (unclamped) initial states
    for pre, (slot, post) in sorted(carried.items(), key=lambda kv: kv[1][0]):
        pro.append(f"hb_c[{slot}] = {pro_ref(pre)};")
    # unit-hydrograph ordinates evaluated on the device (UnitHydrograph.weights: K = ceil(max_lag), S(k) = value of the
    # condition with the smallest threshold >= k, 1 beyond the largest; w_k = S(k) - S(k-1), normalised; K == 0: identity)
    for u, thr, vals in uh_device:
        K = uh_kmax[u]
        th = [pro_ref(i) for i in thr]
        vnames = [[pro_ref(vi) for vi in row] for row in vals]     # materialise every operand at prologue scope first
        # the ordinate COUNT is piecewise constant in the parameters: taken from the value (hb_val), no tangent
        pro.append(f"{{ const double ulag = hb_val({th[0]});")
        pro.append(f"  const int uK = ulag > 0.0 ? (int)hb_min({float(K)}, ceil(ulag)) : 0;")
        pro.append("  double us_prev = 0.0, usum = 0.0;")
        for k in range(1, K + 1):
            pro.append("  { double uval = 1.0, ubest = 1.0e300;")
            for ci, vn in enumerate(vnames[k - 1]):
                pro.append(f"    if ({float(k)} <= {th[ci]} && {th[ci]} < ubest) {{ uval = {vn}; ubest = {th[ci]}; }}")
            pro.append(f"    const double uw = {k} <= uK ? uval - us_prev : 0.0; us_prev = uval; usum += uw; hb_uhw[{uh_w_off[u] + k - 1}] = uw; }}")
        pro.append(f"  if (uK > 0) {{ const double uinv = 1.0 / usum; " +
                   " ".join(f"hb_uhw[{uh_w_off[u] + k}] *= uinv;" for k in range(K)) + " }")
        pro.append(f"  else {{ hb_uhw[{uh_w_off[u]}] = 1.0; }} }}")

    pending_updates = {p[0]: p for k, p in statements if k == "update"}
    nn_done = set()
    for v in b.vals:                                  # construction order is topological
        i = v.id
        if v.op == "state" and v.payload[1] == 1:
            si = v.payload[0]                         # all increments of its bucket precede this leaf
            if si in raw_states:
                step.append(f"const double hb_un{si} = hb_u[{si}] + {step_ref(pending_updates[si][2])};")
            else:
                floor = _c_double(state_floor[si]) if si in state_floor else "hb_minv"
                step.append(f"const double hb_un{si} = hb_max({floor}, hb_u[{si}] + {step_ref(pending_updates[si][2])});")
            continue
        if i not in live or v.mask == 0 or v.op in ("const", "in", "param", "state") or i in carried:
            continue
        if v.op == "uh":
            u, x = v.payload
            step.append(f"double v{i} = hb_uhw[{uh_w_off[u]}] * {step_ref(x)};")
            for k in range(1, uh_kmax[u]):
                step.append(f"v{i} = fma(hb_uhw[{uh_w_off[u] + k}], hb_uhh[{uh_h_off[u] + k - 1}], v{i});")
        elif v.op == "nn":
            ni, ins, j = v.payload
            arr = nn_var(ni, ins, "s")
            if arr not in nn_done:
                nn_done.add(arr)
                step.append(f"double {arr}[{nets[ni].n_out}]; " + nn_call(ni, [step_ref(k) for k in ins], arr))
            step.append(f"const double v{i} = {arr}[{j}];")
        else:
            step.append(f"const double v{i} = {_expr_with(v, [step_ref(a) for a in v.args], fast)};")
    # end of step: rows, carries, states, UH history shift
    if objective:
        step.append(f"hb_sim = {step_ref(out_vals[0])};")
    else:
        for r, vid in enumerate(out_vals):
            if row_names[r] in aux_rows and row_names[r] not in env:
                step.append(f"hb_out[{r}] = hb_aux[{aux_rows.index(row_names[r])}];")
            else:
                step.append(f"hb_out[{r}] = {step_ref(vid)};")
    for pre, (slot, post) in sorted(carried.items(), key=lambda kv: kv[1][0]):
        step.append(f"hb_c[{slot}] = {step_ref(post)};")
    for si in sorted(state_post):
        step.append(f"hb_u[{si}] = hb_un{si};")
    for u, x in uh_steps:
        K = uh_kmax[u]
        for k in range(K - 2, 0, -1):
            step.append(f"hb_uhh[{uh_h_off[u] + k}] = hb_uhh[{uh_h_off[u] + k - 1}];")
        if K >= 2:
            step.append(f"hb_uhh[{uh_h_off[u]}] = {step_ref(x)};")

    tc_defs, tc_bytes = "", 0
    if any(net_tc):
        helpers, b_off, stage, kmax, nmax = [], 0, [], 8, 16
        for i, (ch, (_, off, _)) in enumerate(zip(nets, nn_layout)):
            if not net_tc[i]:
                helpers.append(_mlp_helper(i, ch, off, fast))
                continue
            text, b_off, st, km, nm = _mlp_helper_tc(i, ch, off, b_off)
            helpers.append(text)
            stage += st
            kmax, nmax = max(kmax, km), max(nmax, nm)
        a_floats = (kmax // 8) * 1024                          # one [128 x kmax] tile (hi); the lo tile follows
        tc_bytes = (2 * a_floats + b_off) * 4
        tc_defs = (f"#define HB_NN_TC 1\n#define HB_TC_A_FLOATS {a_floats}\n#define HB_TC_B_FLOATS {b_off}\n"
                   f"#define HB_TC_COLS {32 if nmax <= 32 else 64}\n"
                   f"#define HB_TC_STAGE(sB, nn, tid, nthr) {' '.join(stage)}\n")
    else:
        helpers = [(_mlp_helper_mma if net_mma[i] else _mlp_helper)(i, ch, off, fast)
                   for i, (ch, (_, off, _)) in enumerate(zip(nets, nn_layout))]
    body = "//@@HB:DEFINES\n" + f"#define HB_NC {len(carried)}\n" + tc_defs + \
           "//@@HB:HELPERS\n" + "\n".join(helpers) + "\n" + \
           "//@@HB:PROLOGUE\n" + "\n".join("    " + s for s in pro) + "\n" + \
           "//@@HB:STEP\n" + "\n".join("        " + s for s in step) + "\n"
    if precision == "f32":
        body = _body_to_f32(body)
    elif fast and "hb_pow(" in body:
        # bodies that raise to a power get the table form of the logarithm (hb_math.cuh: 8 instead of 24 FP64 operations, 2 KB more
        # shared memory per block)
        body = body.replace("//@@HB:DEFINES\n", "//@@HB:DEFINES\n#define HB_LOG_TABLE 1\n", 1)
    tangents = list(tangents) if tangents else []
    if tangents:
        if not objective or precision != "f64" or not fast:
            raise LoweringError("a gradient run is an objective run in Float64 with the hb_math helpers")
        if nets:
            raise LoweringError("forward-mode gradients are for the physical parameters of bucket models; models with neural "
                                "components have hundreds of weights and need the reverse-mode adjoint (not built)")
        if uhs and not uh_on_device:
            raise LoweringError("gradient runs of models with unit hydrographs need the ordinates evaluated on the device")
        known = [nm for nm, _ in param_slots]
        for nm in tangents:                      # a direction is a parameter name or the INITIAL value of a state
            if nm not in known and nm not in state_names:
                raise LoweringError(f"Parameter {nm} does not exist in the model")
        if len(set(tangents)) != len(tangents):
            raise LoweringError("duplicate parameter in the gradient directions")
        tdir = ", ".join(str(tangents.index(nm)) if nm in tangents else "-1" for nm in known) or "-1"
        sdir = ", ".join(str(tangents.index(nm)) if nm in tangents else "-1" for nm in state_names) or "-1"
        body = _body_to_tangent(body).replace("//@@HB:DEFINES\n", f"//@@HB:DEFINES\n#define HB_TDIR {{{tdir}}}\n#define HB_TSDIR {{{sdir}}}\n", 1)
    n_ops = sum(1 for v in b.vals if v.id in live and v.mask != 0 and v.op not in ("const", "in", "param", "state")
                and v.id not in carried)
    stats = {"step_values": n_ops, "prologue_values": len(pro), "carried": len(carried),
             "exp_like_per_step": sum(1 for v in b.vals if v.id in live and v.mask != 0 and v.id not in carried
                                      and v.op in ("exp", "step", "tanh", "log", "pow"))}
    return StepperProgram(body=body, input_names=input_names, row_names=row_names, all_row_names=list(all_rows),
                          state_names=state_names, param_slots=param_slots, htype_sets=htype_sets, uh=uhs,
                          uh_kmax=list(uh_kmax), uh_htype_set=[htype_set(u) for u in uhs],
                          nn_layout=nn_layout, nn_words=sum(n for _, _, n in nn_layout), n_carry=len(carried),
                          objective=objective, stats=stats, nn_tensor=any(net_mma) or any(net_tc), precision=precision,
                          uh_on_device=bool(uh_on_device and uhs), tangents=tangents, nn_tc_bytes=tc_bytes,
                          aux_names=[r for r in aux_rows if r in row_names] if not objective else [])


# ---------------------------------------------------------------------------------------------
# reverse mode: the adjoint of one step (templates/adjoint.cu) — what Zygote does through the reference's
# ImmutableSolver (/root/reference/README.md:161-187, src/solve.jl:62-79), here generated from the same step DAG
# ---------------------------------------------------------------------------------------------
@dataclass
class AdjointProgram:
    body: str
    input_names: List[str]
    state_names: List[str]
    param_slots: List[Tuple[str, Optional[int]]]
    htype_sets: List[np.ndarray]
    nn_layout: List[Tuple[str, int, int]]
    nn_words: int
    nn_group_offset: List[int]      # offset (doubles) of every chain's accumulator fragments in the per-warp shared block
    n_groups: int                   # accumulator doubles per warp (all chains)
    uh: List = field(default_factory=list)            # (no unit hydrographs in an adjoint run: what B200Model.prepare expects)
    uh_kmax: List[int] = field(default_factory=list)
    uh_htype_set: List = field(default_factory=list)


def _act_derivative(act: str, z: str, a: str) -> str:
    """d act / d z given the pre-activation `z` and the activation value `a` (selects mirror hb_max: ties take the identity branch)."""
    if act == "identity": return "1.0"
    if act == "tanh": return f"fma(-{a}, {a}, 1.0)"
    if act == "leakyrelu": return f"((0.01 * {z} > {z}) ? 0.01 : 1.0)"
    if act == "relu": return f"((0.0 > {z}) ? 0.0 : 1.0)"
    if act == "sigmoid": return f"({a} * (1.0 - {a}))"
    raise LoweringError(f"no derivative for activation {act}")


def _mlp_helper_bwd(idx: int, chain, offset: int) -> str:
    """Forward recomputation + backward pass of one chain for ONE node, and the accumulation of the weight gradient over the
    32 nodes of the warp: per layer `dW += delta (x) input`, summed over the nodes, as FP64 tensor-core MMAs with the nodes as
    the K dimension (hb_outer_acc; accumulator fragments in shared memory at `gw`, transposition scratch `scr`)."""
    acts = _ACT_C
    args = ", ".join(f"const double x{i}" for i in range(chain.n_in))
    dys = ", ".join(f"const double dy{o}" for o in range(chain.n_out))
    L = [f"__device__ __forceinline__ void hb_mlp_bwd_{idx}(const int lane, {args}, {dys}, double* __restrict__ dx, double* __restrict__ scr, "
         f"double* __restrict__ gw) {{"]
    prev = [f"x{i}" for i in range(chain.n_in)]
    acts_in = [prev]
    zs, off, offs = [], offset, []
    for li, l in enumerate(chain.layers):
        bias = off + l.n_out * l.n_in
        z = [f"z{li + 1}_{o}" for o in range(l.n_out)]
        a = [f"a{li + 1}_{o}" for o in range(l.n_out)]
        for o in range(l.n_out):
            L.append(f"    double {z[o]} = HB_NNW({bias + o});")
        for i in range(l.n_in):
            for o in range(l.n_out):
                L.append(f"    {z[o]} = fma(HB_NNW({off + i * l.n_out + o}), {prev[i]}, {z[o]});")
        for o in range(l.n_out):
            L.append(f"    const double {a[o]} = {acts[l.activation].format(z[o])};")
        zs.append(z)
        offs.append(off)
        off += l.n_params
        prev = a
        acts_in.append(a)
    # backward: delta of the last layer from dy, then down
    g_out = [f"dy{o}" for o in range(chain.n_out)]
    deltas = [None] * len(chain.layers)
    for li in range(len(chain.layers) - 1, -1, -1):
        l = chain.layers[li]
        d = [f"d{li + 1}_{o}" for o in range(l.n_out)]
        for o in range(l.n_out):
            L.append(f"    const double {d[o]} = {g_out[o]} * {_act_derivative(l.activation, zs[li][o], acts_in[li + 1][o])};")
        deltas[li] = d
        g_in = [f"g{li}_{i}" for i in range(l.n_in)]
        for i in range(l.n_in):
            terms = [f"HB_NNW({offs[li] + i * l.n_out + o}) * {d[o]}" for o in range(l.n_out)]
            L.append(f"    double {g_in[i]} = {terms[0]};")
            for o in range(1, l.n_out):
                L.append(f"    {g_in[i]} = fma(HB_NNW({offs[li] + i * l.n_out + o}), {d[o]}, {g_in[i]});")
        g_out = g_in
    for i in range(chain.n_in):
        L.append(f"    dx[{i}] = {g_out[i]};")
    # weight gradient: per layer one warp-wide outer-product accumulation on the FP64 tensor pipe (templates/adjoint.cu:
    # hb_outer_acc), delta and input vectors zero-padded to multiples of 8
    layers, bias_off, _ = _adjoint_acc_layout(chain, offset)
    for li, l in enumerate(chain.layers):
        acc_off, NO, NI, _, _, _, slot = layers[li]
        if NO > 16 or NI > 16:
            raise LoweringError("the adjoint run handles dense layers up to 16 x 16 (scratch pitch)")
        dd = deltas[li] + ["0.0"] * (NO - l.n_out)
        aa = acts_in[li] + ["0.0"] * (NI - l.n_in)
        L.append(f"    {{ const double dd[{NO}] = {{{', '.join(dd)}}}; const double aa[{NI}] = {{{', '.join(aa)}}};")
        L.append(f"      hb_outer_acc<{NO}, {NI}>(scr, gw + {acc_off}, gw + {bias_off}, {slot}, lane, dd, aa); }}")
    L.append("}")
    return "\n".join(L)


def _adjoint_acc_layout(chain, offset: int):
    """[(acc offset in doubles, NO, NI, flat parameter offset, n_out, n_in, bias slot)] per layer, the offset of the chain's shared
    bias tile and the chain's accumulator doubles.  The bias gradients of all layers share one 8 x 8 tile: output tile `ot` of a
    layer owns column `slot + ot` (templates/adjoint.cu: hb_outer_acc)."""
    out, acc_off, off, slot = [], 0, offset, 0
    for l in chain.layers:
        NO, NI = -(-l.n_out // 8) * 8, -(-l.n_in // 8) * 8
        out.append((acc_off, NO, NI, off, l.n_out, l.n_in, slot))
        acc_off += (NO // 8) * (NI // 8) * 64
        slot += NO // 8
        off += l.n_params
    if slot > 8:
        raise LoweringError("the adjoint run packs the bias gradients of a network into one 8-column tile: at most 8 output tiles per chain")
    bias_off = acc_off
    return out, bias_off, acc_off + 64


def _emit_adjoint(b: "_Builder", sim: int, statements, param_slots, nets, nn_layout, raw_states, state_floor, input_names,
                  state_names, htype_sets) -> AdjointProgram:
    """Body of templates/adjoint.cu for one time step, walked backwards in time by the kernel:

        forward   recompute every value of the step from the PRE-update states hb_u[] (read from the stored trajectory),
                  the forcing hb_x[] and the parameters (parameter-only values hoisted to the prologue as in the forward body);
        seed      adjoint of the simulated value += hb_seed (d loss / d sim_t), adjoint of the post-update states = hb_lam[]
                  (what the later steps passed back);
        backward  the vector-Jacobian product of every value in reverse order: adjoints of the pre-update states -> hb_lam[]
                  for step t-1, of the parameters -> hb_gp[], of parameter-only values -> persistent accumulators swept once
                  in the epilogue, of neural-flux inputs / weights through hb_mlp_bwd_k.
    Selects (min / max / the state floor) pass the adjoint to the branch the forward pass took — the same convention as the
    forward-mode duals in hb_math.cuh."""
    updates = {p[0]: p for k, p in statements if k == "update"}       # state index -> (si, pre leaf, du val)
    vals = b.vals
    live = set()
    stack = [sim]
    while stack:
        j = stack.pop()
        if j in live:
            continue
        live.add(j)
        v = vals[j]
        stack.extend(v.args)
        if v.op == "state" and v.payload[1] == 1:
            stack.append(updates[v.payload[0]][2])
    for si in range(len(state_names)):                               # every state is stepped, whether the objective sees it or not
        for j in (b.leaf_state(si, 0), b.leaf_state(si, 1), updates[si][2]) if si in updates else ():
            stack.append(j)
    while stack:
        j = stack.pop()
        if j in live:
            continue
        live.add(j)
        stack.extend(vals[j].args)
        if vals[j].op == "state" and vals[j].payload[1] == 1:
            stack.append(updates[vals[j].payload[0]][2])
    vals = b.vals                                                     # leaf_state above may have appended leaves
    pclass = lambda i: vals[i].mask == 0 and vals[i].op not in ("const", "param")
    # values that depend on something differentiable (a parameter, a state, a neural flux): forcing-only subexpressions
    # (ExpHydro's whole PET chain) carry no adjoint
    diff = set()
    for v in vals:
        if v.op in ("param", "state", "nn") or any(k in diff for k in v.args):
            diff.add(v.id)

    def ref(i: int) -> str:
        v = vals[i]
        if v.op == "const": return _c_double(v.payload)
        if v.op == "in": return f"hb_x[{v.payload}]"
        if v.op == "param": return f"hb_p[{v.payload}]"
        if v.op == "state": return f"hb_u[{v.payload[0]}]" if v.payload[1] == 0 else f"hb_un{v.payload[0]}"
        return f"h{i}" if v.mask == 0 else f"v{i}"

    def adj(i: int) -> Optional[str]:
        v = vals[i]
        if v.op in ("const", "in") or i not in diff: return None
        if v.op == "param": return f"hb_gp[{v.payload}]"
        if v.op == "state": return f"as{v.payload[0]}" if v.payload[1] == 0 else f"an{v.payload[0]}"
        return f"ah{i}" if v.mask == 0 else f"a{i}"

    def vjp(v: Val, g: str) -> List[str]:
        """statements adding the contributions of `g` = adjoint of v to the adjoints of its arguments"""
        a = [ref(k) for k in v.args]
        t = [adj(k) for k in v.args]
        me = ref(v.id)
        out: List[str] = []

        def add(k, expr):
            if t[k] is not None:
                out.append(f"{t[k]} += {expr};")
        op = v.op
        if op == "add": add(0, g); add(1, g)
        elif op == "sub": add(0, g); add(1, f"-{g}")
        elif op == "mul": add(0, f"{g} * {a[1]}"); add(1, f"{g} * {a[0]}")
        elif op == "div": add(0, f"{g} / {a[1]}"); add(1, f"-{g} * {me} / {a[1]}")
        elif op == "rcp": add(0, f"-{g} * {me} * {me}")
        elif op == "neg": add(0, f"-{g}")
        elif op == "exp": add(0, f"{g} * {me}")
        elif op == "log": add(0, f"{g} / {a[0]}")
        elif op == "sqrt": add(0, f"{g} * 0.5 / {me}")
        elif op == "abs": add(0, f"({a[0]} < 0.0 ? -{g} : {g})")
        elif op == "tanh": add(0, f"{g} * fma(-{me}, {me}, 1.0)")
        elif op == "step": add(0, f"{g} * 10.0 * {me} * (1.0 - {me})")
        elif op == "min":
            add(0, f"({a[0]} < {a[1]} ? {g} : 0.0)"); add(1, f"({a[0]} < {a[1]} ? 0.0 : {g})")
        elif op == "max":
            add(0, f"({a[0]} > {a[1]} ? {g} : 0.0)"); add(1, f"({a[0]} > {a[1]} ? 0.0 : {g})")
        elif op == "pow":
            add(0, f"{g} * {a[1]} * hb_pow({a[0]}, {a[1]} - 1.0)")
            add(1, f"({a[0]} > 0.0 ? {g} * {me} * log({a[0]}) : 0.0)")
        else:
            raise LoweringError(f"no adjoint rule for {op}")
        return out

    pro: List[str] = []
    step: List[str] = []
    epi: List[str] = []
    # ---- prologue: parameter-only values and their persistent adjoints ----
    for v in vals:
        if v.id in live and pclass(v.id):
            if v.op == "nn":
                raise LoweringError("a neural flux whose inputs are parameters only is not covered by the adjoint run")
            pro.append(f"const double h{v.id} = {_expr_with(v, [ref(k) for k in v.args], True)};")
            if v.id in diff:
                pro.append(f"double ah{v.id} = 0.0;")
    # ---- step, forward ----
    nn_groups: Dict[Tuple[int, Tuple[int, ...]], List[int]] = {}
    for v in vals:
        i = v.id
        if v.op == "state" and v.payload[1] == 1 and i in live:
            si = v.payload[0]
            step.append(f"const double hb_raw{si} = hb_u[{si}] + {ref(updates[si][2])};")
            if si in raw_states:
                step.append(f"const double hb_un{si} = hb_raw{si};")
            else:
                floor = _c_double(state_floor[si]) if si in state_floor else "hb_minv"
                step.append(f"const double hb_un{si} = hb_max({floor}, hb_raw{si});")
            continue
        if i not in live or v.mask == 0 or v.op in ("const", "in", "param", "state"):
            continue
        if v.op == "nn":
            ni, ins, j = v.payload
            key = (ni, ins)
            if key not in nn_groups:
                nn_groups[key] = [None] * nets[ni].n_out
                arr = f"nnf{ni}_" + "_".join(str(k) for k in ins)
                step.append(f"double {arr}[{nets[ni].n_out}]; hb_mlp_{ni}(sNN, {', '.join(ref(k) for k in ins)}, {arr});")
            nn_groups[key][j] = i
            arr = f"nnf{ni}_" + "_".join(str(k) for k in ins)
            step.append(f"const double v{i} = {arr}[{j}];")
        else:
            step.append(f"const double v{i} = {_expr_with(v, [ref(k) for k in v.args], True)};")
    step.append(f"hb_sim = {ref(sim)};")
    step.append("HB_SEED_STMT")                                       # template macro: hb_seed = d loss / d sim_t, loss accumulation
    # ---- step, adjoint declarations and seeds ----
    for v in vals:
        if v.id in live and v.id in diff and v.mask != 0 and v.op not in ("const", "in", "param", "state"):
            step.append(f"double a{v.id} = 0.0;")
    for si in range(len(state_names)):
        step.append(f"double as{si} = 0.0, an{si} = hb_lam[{si}];")
    t_sim = adj(sim)
    if t_sim is not None:
        step.append(f"{t_sim} += hb_seed;")
    # ---- step, backward ----
    done_groups = set()
    for v in reversed(vals):
        i = v.id
        if i not in live or v.op in ("const", "in", "param"):
            continue
        if v.op == "state":
            si, ver = v.payload
            if ver == 1:
                du_t = adj(updates[si][2])
                body = f"as{si} += an{si};" + (f" {du_t} += an{si};" if du_t is not None else "")
                if si in raw_states:
                    step.append(body)
                else:
                    floor = _c_double(state_floor[si]) if si in state_floor else "hb_minv"
                    step.append(f"if (!({floor} > hb_raw{si})) {{ {body} }}")
            continue
        if v.mask == 0 or i not in diff:
            continue                                                   # parameter-only: swept once in the epilogue; forcing-only: no adjoint
        if v.op == "nn":
            ni, ins, j = v.payload
            key = (ni, ins)
            if key in done_groups:
                continue
            done_groups.add(key)
            dys = [f"a{vid}" if vid is not None else "0.0" for vid in nn_groups[key]]
            arr = f"nnb{ni}_" + "_".join(str(k) for k in ins)
            step.append(f"double {arr}[{nets[ni].n_in}]; hb_mlp_bwd_{ni}(lane, {', '.join(ref(k) for k in ins)}, {', '.join(dys)}, {arr}, "
                        f"hb_scr, hb_gw + HB_GOFF_{ni});")
            for k, src in enumerate(ins):
                t = adj(src)
                if t is not None:
                    step.append(f"{t} += {arr}[{k}];")
            continue
        step.extend(vjp(v, f"a{i}"))
    for si in range(len(state_names)):
        step.append(f"hb_lam[{si}] = as{si};")
    # ---- epilogue: parameter-only values, once ----
    for v in reversed(vals):
        if v.id in live and pclass(v.id) and v.id in diff:
            epi.extend(vjp(v, f"ah{v.id}"))
    goffs, acc_total, flush = [], 0, []
    for ch, (_, off, _) in zip(nets, nn_layout):
        layers, bias_off, n_acc = _adjoint_acc_layout(ch, off)
        goffs.append(acc_total)
        for a_off, NO, NI, p_off, n_out, n_in, slot in layers:
            flush.append(f"hb_outer_flush<{NO}, {NI}>(hb_gw + {acc_total + a_off}, hb_gw + {acc_total + bias_off}, {slot}, lane, "
                         f"a.grad_nn + {p_off}, {n_out}, {n_in});")
        acc_total += n_acc
    helpers = [_mlp_helper(i, ch, off, True) + "\n" + _mlp_helper_bwd(i, ch, off) for i, (ch, (_, off, _)) in enumerate(zip(nets, nn_layout))]
    defs = f"#define HB_ACC_DOUBLES {acc_total}\n" + "".join(f"#define HB_GOFF_{i} {g}\n" for i, g in enumerate(goffs))
    if any("hb_pow(" in x for x in pro + step + epi):
        defs += "#define HB_LOG_TABLE 1\n"
    body = "//@@HB:DEFINES\n" + defs + "//@@HB:HELPERS\n" + "\n".join(helpers) + "\n" + \
           "//@@HB:PROLOGUE\n" + "\n".join("    " + x for x in pro) + "\n" + \
           "//@@HB:STEP\n" + "\n".join("        " + x for x in step) + "\n" + \
           "//@@HB:EPILOGUE\n" + "\n".join("    " + x for x in epi) + "\n" + \
           "//@@HB:FLUSH\n" + "\n".join("    " + x for x in flush) + "\n"
    ng = acc_total
    return AdjointProgram(body=body, input_names=list(input_names), state_names=list(state_names), param_slots=list(param_slots),
                          htype_sets=list(htype_sets), nn_layout=list(nn_layout), nn_words=sum(n for _, _, n in nn_layout),
                          nn_group_offset=goffs, n_groups=ng)


def _body_to_tangent(body: str) -> str:
    """Gradient run: every temporary takes the type of its initialiser (`auto`: plain double for forcing-only
    subexpressions, hb_dual for anything downstream of a parameter or a state), accumulators become hb_dual."""
    import re
    out = []
    for line in body.split("\n"):
        if line.startswith("//@@HB:") or line.startswith("#define"):
            out.append(line)
            continue
        if "ulag" in line:                       # plain-double bookkeeping of the unit-hydrograph prologue
            out.append(line)
            continue
        line = re.sub(r"\bconst double uw\b", "const hb_dual uw", line)   # `cond ? dual : 0.0`
        line = re.sub(r"\bconst double\b", "const auto", line)
        out.append(re.sub(r"\bdouble\b", "hb_dual", line))
    return "\n".join(out)


def _expr_with(v: Val, a: List[str], fast: bool) -> str:
    op = v.op
    if op == "add": return f"{a[0]} + {a[1]}"
    if op == "sub": return f"{a[0]} - {a[1]}"
    if op == "mul": return f"{a[0]} * {a[1]}"
    if op == "div": return f"{a[0]} / {a[1]}"
    if op == "rcp": return f"1.0 / {a[0]}"
    if op == "neg": return f"-{a[0]}"
    if op == "pow": return f"hb_pow({a[0]}, {a[1]})" if fast else f"pow({a[0]}, {a[1]})"
    if op == "exp": return f"hb_exp({a[0]})" if fast else f"exp({a[0]})"
    if op == "log": return f"log({a[0]})"
    if op == "sqrt": return f"sqrt({a[0]})"
    if op == "abs": return f"fabs({a[0]})"
    if op == "tanh": return f"hb_tanh({a[0]})" if fast else f"tanh({a[0]})"
    if op == "step": return f"hb_step({a[0]})" if fast else f"(tanh(5.0 * {a[0]}) + 1.0) * 0.5"
    if op == "min": return f"hb_min({a[0]}, {a[1]})" if fast else f"fmin({a[0]}, {a[1]})"
    if op == "max":
        if fast and "0.0" in (a[0], a[1]) and a[0] != a[1]:
            # max(0, x): an integer test of x's sign bit (ISETP + 2 SEL) instead of an FP64 compare, which holds the issue port
            # for two clocks (DESIGN.md, M50 paragraph); same value for every x except the sign of a zero result
            return f"hb_relu({a[1] if a[0] == '0.0' else a[0]})"
        return f"hb_max({a[0]}, {a[1]})" if fast else f"fmax({a[0]}, {a[1]})"
    raise LoweringError(f"no emitter for {op}")


_ACT_C = {"identity": "{}", "tanh": "hb_tanh({})", "leakyrelu": "hb_leakyrelu({})", "relu": "hb_max(0.0, {})",
          "sigmoid": "hb_rcp(1.0 + hb_exp_narrow(hb_clamp_abs(-({}), HB_HI_700)))"}
_ACT_C_EXACT = dict(_ACT_C, tanh="tanh({})", sigmoid="(1.0 / (1.0 + exp(-({}))))")


def _mlp_helper(idx: int, chain, offset: int, fast: bool) -> str:
    """Register-resident matvec chain as straight-line code (every weight index is a compile-time
    constant, so `HB_NNW(i)` becomes a constant-bank operand of the DFMA — no load instruction, no
    register).  Layout per layer: weight[out,in] column-major then bias[out]
    (`/root/reference/src/nn.jl:52-58`)."""
    acts = _ACT_C if fast else _ACT_C_EXACT
    args = ", ".join(f"const double x{i}" for i in range(chain.n_in))
    lines = [f"__device__ __forceinline__ void hb_mlp_{idx}(const double* __restrict__ sNN, {args}, double* __restrict__ y) {{"]
    prev = [f"x{i}" for i in range(chain.n_in)]
    off = offset
    for li, l in enumerate(chain.layers):
        cur = []
        bias = off + l.n_out * l.n_in
        for o in range(l.n_out):
            nm = f"a{li + 1}_{o}"
            lines.append(f"    double {nm} = HB_NNW({bias + o});")
            cur.append(nm)
        # input-major order: 16 independent accumulator chains advance together (ILP)
        for i in range(l.n_in):
            for o in range(l.n_out):
                lines.append(f"    {cur[o]} = fma(HB_NNW({off + i * l.n_out + o}), {prev[i]}, {cur[o]});")
        for o in range(l.n_out):
            if l.activation != "identity":
                lines.append(f"    {cur[o]} = {acts[l.activation].format(cur[o])};")
        off += l.n_params
        prev = cur
    for o in range(chain.n_out):
        lines.append(f"    y[{o}] = {prev[o]};")
    lines.append("}")
    return "\n".join(lines)


# ---------------------------------------------------------------------------------------------
# HydroRoute → body of templates/route.cu
# ---------------------------------------------------------------------------------------------
@dataclass
class RouteProgram:
    body: str
    input_names: List[str]
    state_name: str
    output_name: str
    param_slots: List[Tuple[str, Optional[int]]]
    htype_sets: List[np.ndarray]


def _emit_scalar_function(b: "_Builder", root: int, name: str, signature: str, refs: Dict[int, str], fast: bool) -> str:
    """Straight-line device function computing value `root`; `refs` names the leaves."""
    lines = [f"__device__ __forceinline__ double {name}({signature}) {{"]
    names: Dict[int, str] = dict(refs)

    def ref(i: int) -> str:
        v = b.vals[i]
        if i in names:
            return names[i]
        if v.op == "const":
            return _c_double(v.payload)
        if v.op == "in":
            return f"hb_x[{v.payload}]"
        if v.op == "param":
            return f"hb_p[{v.payload}]"
        args = [ref(a) for a in v.args]
        names[i] = f"r{i}"
        lines.append(f"    const double r{i} = {_expr_with(v, args, fast)};")
        return names[i]

    out = ref(root)
    lines.append(f"    return {out};")
    lines.append("}")
    return "\n".join(lines)


def lower_route(route: HydroRoute, *, fast: bool = True) -> RouteProgram:
    """Lower a `HydroRoute`'s outflow flux and state increment
    (`/root/reference/src/route.jl:387-403`) to the two device functions `route.cu` calls."""
    b = _Builder(fast)
    htype_sets = [np.asarray(route.htypes)]
    param_slots: List[Tuple[str, Optional[int]]] = [(nm, 0) for nm in route.param_names]
    ptab = {nm: b.leaf_param(i) for i, nm in enumerate(route.param_names)}
    env = {nm: b.leaf_in(i) for i, nm in enumerate(route.input_names)}
    s_leaf = b.opaque("routestate", 0, M_S, cost=0)
    q_leaf = b.opaque("routeout", 0, M_S, cost=0)
    sname, oname = route.state_names[0], route.output_names[0]
    env_flux = dict(env)
    env_flux[sname] = s_leaf
    q_val = None
    for f in route.rfluxes:
        for eq in f.eqs:
            q_val = b.lower_expr(eq.rhs, env_flux, ptab)
            env_flux[eq.lhs.name] = q_val
    env_ds = dict(env)
    env_ds[sname] = s_leaf
    env_ds[oname] = q_leaf
    ds_val = b.lower_expr(route.dfluxes[0].expr, env_ds, ptab)
    f1 = _emit_scalar_function(b, q_val, "hb_rflux", "const double* hb_x, const double hb_s, const double* hb_p",
                               {s_leaf: "hb_s"}, fast)
    f2 = _emit_scalar_function(b, ds_val, "hb_rds", "const double* hb_x, const double hb_s, const double hb_q, const double* hb_p",
                               {s_leaf: "hb_s", q_leaf: "hb_q"}, fast)
    body = "//@@HB:HELPERS\n" + f1 + "\n" + f2 + "\n"
    return RouteProgram(body=body, input_names=list(route.input_names), state_name=sname, output_name=oname,
                        param_slots=param_slots, htype_sets=htype_sets)


def _tc_layers(chain) -> List[int]:
    """Layers of `chain` that go through tcgen05.mma in Float32 mode: K = n_in a multiple of 8, N = n_out a multiple of 16,
    both at most 64 (M50's 16 -> 16 hidden layers, `test/base/run_single_lumped.jl:216-229`).  The other layers (3 -> 16,
    16 -> 1) are too thin for an MMA and stay per-thread FFMA."""
    return [i for i, l in enumerate(chain.layers) if l.n_in % 8 == 0 and l.n_out % 16 == 0 and l.n_in <= 64 and l.n_out <= 64]


def _mlp_helper_tc(idx: int, chain, offset: int, b_off: int):
    """Float32 tensor path of one chain: thin layers as per-thread FFMA straight-line code, tensor-shaped layers through
    `hb_tc_layer<K, N>` (all 128 threads of the CTA call it at the same point).  Returns (text, next B offset in floats,
    staging statements, max K, max N)."""
    acts = _ACT_C
    tcl = set(_tc_layers(chain))
    args = ", ".join(f"const double x{i}" for i in range(chain.n_in))
    lines = [f"__device__ __forceinline__ void hb_mlp_{idx}(HbTc& hb_tc, {args}, double* __restrict__ y) {{"]
    prev = [f"x{i}" for i in range(chain.n_in)]
    off, stage, kmax, nmax = offset, [], 8, 16
    for li, l in enumerate(chain.layers):
        bias = off + l.n_out * l.n_in
        cur = [f"a{li + 1}_{o}" for o in range(l.n_out)]
        if li in tcl:
            K, N = l.n_in, l.n_out
            lines.append(f"    double t{li + 1}[{N}];")
            lines.append(f"    {{ const double xin[{K}] = {{{', '.join(prev)}}}; hb_tc_layer<{K}, {N}>(hb_tc, xin, {b_off}, t{li + 1}); }}")
            for o in range(N):
                lines.append(f"    double {cur[o]} = t{li + 1}[{o}] + HB_NNW({bias + o});")
            stage.append(f"hb_tc_stage_b<{K}, {N}>((sB) + {b_off}, (nn) + {off}, tid, nthr);")
            b_off += 2 * K * N
            kmax, nmax = max(kmax, K), max(nmax, N)
        else:
            for o in range(l.n_out):
                lines.append(f"    double {cur[o]} = HB_NNW({bias + o});")
            for i in range(l.n_in):
                for o in range(l.n_out):
                    lines.append(f"    {cur[o]} = fma(HB_NNW({off + i * l.n_out + o}), {prev[i]}, {cur[o]});")
        for o in range(l.n_out):
            if l.activation != "identity":
                lines.append(f"    {cur[o]} = {acts[l.activation].format(cur[o])};")
        off += l.n_params
        prev = cur
    for o in range(chain.n_out):
        lines.append(f"    y[{o}] = {prev[o]};")
    lines.append("}")
    return "\n".join(lines), b_off, stage, kmax, nmax


def _mma_eligible(chain) -> bool:
    """Shapes the DMMA lowering handles: `n_in <= 4 -> H (-> H')* -> 1` with every hidden width a
    multiple of 8 and at most 32 (M50: 3-16-16-1 and 2-16-16-1, `test/base/run_single_lumped.jl:216-229`)."""
    L = chain.layers
    if len(L) < 2 or L[0].n_in > 4 or L[-1].n_out != 1:
        return False
    return all(l.n_out % 8 == 0 and 8 <= l.n_out <= 32 for l in L[:-1])


def _mlp_helper_mma(idx: int, chain, offset: int, fast: bool) -> str:
    """Warp-cooperative dense chain on the FP64 tensor pipe (`mma.sync.m8n8k4.f64`, SASS DMMA).

    The warp's 32 basins are the M dimension (4 tiles of 8), neurons the N dimension, layer inputs K:
    `D[basin][neuron] = sum_k A[basin][k] * W[neuron][k] + bias`.  Lane (q = lane/4, j = lane%4) holds,
    for every M tile mt, the accumulator pair of neurons `8*nt + 2j + {0,1}` of basin `8*mt + q`.  The
    K index of the next layer is only summed over, so its order is free: chunk `c = 2*nt + e` of the
    next layer's A fragment is DEFINED as the neuron this very lane already holds (`8*nt + 2j + e`)
    and the weight fragment is gathered with the same permutation — activations never move between
    lanes from one layer to the next.  Only the layer-0 inputs (owned by lane = basin) and the final
    scalar output go through a 1 KB per-warp shared scratch."""
    acts = _ACT_C
    L = chain.layers
    args = ", ".join(f"const double x{i}" for i in range(chain.n_in))
    o = [f"__device__ __forceinline__ void hb_mlp_{idx}(const double* __restrict__ sW, double* __restrict__ scr, const int lane, {args}, double* __restrict__ y) {{",
         "    const int q = lane >> 2, j = lane & 3;",
         "    constexpr int HB_MT = HB_NPW / 8;     // M tiles of 8 basins (HB_NPW == 16: lanes 16..31 mirror lanes 0..15, only their tiles are skipped)"]
    for k in range(4):
        o.append(f"    scr[lane * 4 + {k}] = {'x%d' % k if k < chain.n_in else '0.0'};")
    o += ["    __syncwarp();", "    double afr[HB_MT];", "#pragma unroll",
          "    for (int mt = 0; mt < HB_MT; ++mt) afr[mt] = scr[(mt * 8 + q) * 4 + j];", "    __syncwarp();"]
    off = offset
    # layer 0: K = 4 (inputs zero-padded), one chunk
    l0 = L[0]
    nt0 = l0.n_out // 8
    bias = off + l0.n_out * l0.n_in
    o.append(f"    double h0[HB_MT][{nt0}][2];")
    o.append("#pragma unroll")
    o.append(f"    for (int nt = 0; nt < {nt0}; ++nt) {{")
    o.append(f"        const double b = (j < {l0.n_in}) ? sW[{off} + j * {l0.n_out} + nt * 8 + q] : 0.0;")
    o.append(f"        const double c0 = sW[{bias} + nt * 8 + 2 * j], c1 = sW[{bias} + nt * 8 + 2 * j + 1];")
    o.append("#pragma unroll")
    o.append("        for (int mt = 0; mt < HB_MT; ++mt) hb_dmma_c(h0[mt][nt][0], h0[mt][nt][1], afr[mt], b, c0, c1);")
    o.append("    }")
    if l0.activation != "identity":
        o.append("#pragma unroll")
        o.append(f"    for (int mt = 0; mt < HB_MT; ++mt) for (int nt = 0; nt < {nt0}; ++nt) for (int e = 0; e < 2; ++e) h0[mt][nt][e] = {acts[l0.activation].format('h0[mt][nt][e]')};")
    off += l0.n_params
    prev, prev_nt = "h0", nt0
    for li in range(1, len(L) - 1):
        l = L[li]
        ntl = l.n_out // 8
        cur = f"h{li}"
        bias = off + l.n_out * l.n_in
        o.append(f"    double {cur}[HB_MT][{ntl}][2];")
        o.append("#pragma unroll")
        o.append(f"    for (int nt = 0; nt < {ntl}; ++nt) {{")
        o.append(f"        const double c0 = sW[{bias} + nt * 8 + 2 * j], c1 = sW[{bias} + nt * 8 + 2 * j + 1];")
        o.append("#pragma unroll")
        o.append(f"        for (int c = 0; c < {l.n_in // 4}; ++c) {{")
        o.append(f"            const double b = sW[{off} + ((c >> 1) * 8 + 2 * j + (c & 1)) * {l.n_out} + nt * 8 + q];")
        o.append("#pragma unroll")
        o.append(f"            for (int mt = 0; mt < HB_MT; ++mt) {{")
        o.append(f"                if (c == 0) hb_dmma_c({cur}[mt][nt][0], {cur}[mt][nt][1], {prev}[mt][0][0], b, c0, c1);")
        o.append(f"                else hb_dmma({cur}[mt][nt][0], {cur}[mt][nt][1], {prev}[mt][c >> 1][c & 1], b);")
        o.append("            }")
        o.append("        }")
        o.append("    }")
        if l.activation != "identity":
            o.append("#pragma unroll")
            o.append(f"    for (int mt = 0; mt < HB_MT; ++mt) for (int nt = 0; nt < {ntl}; ++nt) for (int e = 0; e < 2; ++e) {cur}[mt][nt][e] = {acts[l.activation].format(cur + '[mt][nt][e]')};")
        off += l.n_params
        prev, prev_nt = cur, ntl
    # last layer: H -> 1, per-lane partial dot + quad reduction
    ll = L[-1]
    o.append("    double yv[HB_MT];")
    o.append("#pragma unroll")
    o.append("    for (int mt = 0; mt < HB_MT; ++mt) {")
    o.append("        double part = 0.0;")
    o.append("#pragma unroll")
    o.append(f"        for (int nt = 0; nt < {prev_nt}; ++nt) for (int e = 0; e < 2; ++e) part = fma(sW[{off} + nt * 8 + 2 * j + e], {prev}[mt][nt][e], part);")
    o.append("        part += __shfl_xor_sync(0xffffffffu, part, 1);")
    o.append("        part += __shfl_xor_sync(0xffffffffu, part, 2);")
    z = f"(part + sW[{off + ll.n_in}])"
    o.append(f"        yv[mt] = {acts[ll.activation].format(z) if ll.activation != 'identity' else z};")
    o.append("    }")
    o.append("    if (j == 0) {")
    o.append("#pragma unroll")
    o.append("        for (int mt = 0; mt < HB_MT; ++mt) scr[mt * 8 + q] = yv[mt];")
    o.append("    }")
    o += ["    __syncwarp();", "    y[0] = scr[lane % HB_NPW];", "    __syncwarp();", "}"]
    return "\n".join(o)


_FLOAT_LIT = None


def _body_to_f32(body: str) -> str:
    """Opt-in Float32 mode: the generated text is type-generic except for the spelled-out `double` and the
    floating literals — retype both (`real`, `…f`); the math helpers are overloaded on float."""
    import re
    global _FLOAT_LIT
    if _FLOAT_LIT is None:
        _FLOAT_LIT = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
    out = []
    for line in body.split("\n"):
        if line.startswith("//@@HB:") or line.startswith("#define"):
            out.append(line)
            continue
        line = re.sub(r"\bdouble\b", "real", line)
        out.append(_FLOAT_LIT.sub(lambda m: m.group(1) + "f", line))
    return "\n".join(out)


# ---------------------------------------------------------------------------------------------
# per-cell components + HydroRoute on a dense grid → body of templates/grid.cu
# ---------------------------------------------------------------------------------------------
@dataclass
class GridProgram:
    body: str
    stepper: StepperProgram          # the per-cell part (rows = all model rows, the route's are holes)
    route: RouteProgram
    row_names: List[str]
    s_row: int
    o_row: int


def lower_grid(lumped: HydroModel, route: HydroRoute, row_names: Sequence[str], *, fast: bool = True) -> GridProgram:
    """Fuse the per-cell step (`lower_stepper`) and the route's two device functions (`lower_route`) into
    one body: the route's inputs are taken from the rows the per-cell step has just produced."""
    holes = route.state_names + route.output_names
    sp = lower_stepper(lumped, rows=list(row_names), fast=fast, holes=holes)
    if sp.uh or sp.nn_layout:
        raise LoweringError("the fused grid kernel supports bucket / flux components only")
    rp = lower_route(route, fast=fast)
    missing = [nm for nm in rp.input_names if nm not in row_names]
    if missing:
        raise LoweringError(f"route inputs {missing} are not rows of the model")
    sel = "0"
    for v in reversed(range(len(rp.input_names))):
        sel = f"((v) == {v} ? hb_out[{list(row_names).index(rp.input_names[v])}] : {sel})"
    secs: Dict[str, List[str]] = {}
    cur = None
    for src in (sp.body, rp.body):
        for line in src.split("\n"):
            if line.startswith("//@@HB:"):
                cur = line[7:].strip()
                secs.setdefault(cur, [])
            elif cur is not None:
                secs[cur].append(line)
    secs.setdefault("DEFINES", []).append(f"#define HB_ROUTE_IN(v) {sel}")
    body = "".join(f"//@@HB:{k}\n" + "\n".join(v) + "\n" for k, v in secs.items())
    return GridProgram(body=body, stepper=sp, route=rp, row_names=list(row_names),
                       s_row=list(row_names).index(route.state_names[0]),
                       o_row=list(row_names).index(route.output_names[0]))
