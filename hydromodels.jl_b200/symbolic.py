"""Symbolic expression layer: the host-side stand-in for Symbolics.jl `Num` trees.

The reference builds every flux from `@variables` / `@parameters` and `lhs ~ rhs`
equations whose right-hand sides are Symbolics expression trees
(`/root/reference/src/flux.jl:22-86`, `HydroFlux.exprs`), with the closed function
vocabulary of `/root/reference/ext/HydroModelsYAMLExt/expression_parser.jl:113-143`
(`+ - * / ^ max min exp log sqrt abs tanh step_func`) plus `clamp` (docs HBV,
`docs/src/tutorials/neuralnetwork_embeding.md:180-182`).  This module provides the same
thing in Python: `Variable`/`Parameter` leaves, operator overloading, the function
vocabulary, and a parser for the reference's `"lhs ~ rhs"` formula strings (the YAML
schema's `formula:` field, `docs/examples/exphydro.yaml:57`).

Trees are kept exactly as written (no re-association, no constant folding); the oracle
evaluates them as written, the CUDA lowering (`lowering.py`) is free to rewrite within the
stated 1e-10 tolerance.
"""
from __future__ import annotations

import ast
import math
from typing import Dict, Iterable, List, Sequence, Tuple, Union

Number = Union[int, float]

# op name -> arity (None = variadic >= 2)
OPS = {
    "add": 2, "sub": 2, "mul": 2, "div": 2, "pow": 2, "neg": 1,
    "exp": 1, "log": 1, "sqrt": 1, "abs": 1, "tanh": 1, "step": 1,
    "min": 2, "max": 2, "clamp": 3,
}


class Expr:
    """Base class of expression nodes (immutable, hash-consed by structure)."""

    __slots__ = ("_key", "_hash")

    # -- structure -----------------------------------------------------------------
    def key(self):
        return self._key

    def __hash__(self):
        return self._hash

    def same(self, other: "Expr") -> bool:
        return isinstance(other, Expr) and self._key == other._key

    # -- operators -----------------------------------------------------------------
    def __add__(self, o): return Call("add", (self, as_expr(o)))
    def __radd__(self, o): return Call("add", (as_expr(o), self))
    def __sub__(self, o): return Call("sub", (self, as_expr(o)))
    def __rsub__(self, o): return Call("sub", (as_expr(o), self))
    def __mul__(self, o): return Call("mul", (self, as_expr(o)))
    def __rmul__(self, o): return Call("mul", (as_expr(o), self))
    def __truediv__(self, o): return Call("div", (self, as_expr(o)))
    def __rtruediv__(self, o): return Call("div", (as_expr(o), self))
    def __pow__(self, o): return Call("pow", (self, as_expr(o)))
    def __rpow__(self, o): return Call("pow", (as_expr(o), self))
    def __neg__(self): return Call("neg", (self,))
    def __pos__(self): return self
    def __abs__(self): return Call("abs", (self,))

    # `lhs ~ rhs` has no Python spelling; `Eq(lhs, rhs)` / `lhs.eq(rhs)` stands in.
    def eq(self, rhs) -> "Equation":
        return Equation(self, as_expr(rhs))


class Const(Expr):
    __slots__ = ("value",)

    def __init__(self, value: Number):
        self.value = float(value)
        self._key = ("c", self.value.hex())
        self._hash = hash(self._key)

    def __repr__(self):
        return repr(self.value)


class Sym(Expr):
    """A named leaf: a variable (input / state / flux) or a parameter."""

    __slots__ = ("name", "is_param")

    def __init__(self, name: str, is_param: bool = False):
        self.name = str(name)
        self.is_param = bool(is_param)
        self._key = ("p" if is_param else "v", self.name)
        self._hash = hash(self._key)

    def __repr__(self):
        return self.name


class Call(Expr):
    __slots__ = ("op", "args")

    def __init__(self, op: str, args: Sequence[Expr]):
        if op not in OPS:
            raise ValueError(f"Unknown function: {op}")
        args = tuple(as_expr(a) for a in args)
        if len(args) != OPS[op]:
            raise ValueError(f"{op} expects {OPS[op]} arguments, got {len(args)}")
        self.op = op
        self.args = args
        self._key = (op,) + tuple(a._key for a in args)
        self._hash = hash(self._key)

    def __repr__(self):
        a = self.args
        if self.op in ("add", "sub", "mul", "div", "pow"):
            s = {"add": "+", "sub": "-", "mul": "*", "div": "/", "pow": "^"}[self.op]
            return f"({a[0]!r} {s} {a[1]!r})"
        if self.op == "neg":
            return f"(-{a[0]!r})"
        name = "step_func" if self.op == "step" else self.op
        return f"{name}({', '.join(repr(x) for x in a)})"


class Equation:
    """`lhs ~ rhs` (one output variable per equation)."""

    __slots__ = ("lhs", "rhs")

    def __init__(self, lhs: Sym, rhs):
        if not isinstance(lhs, Sym) or lhs.is_param:
            raise TypeError("left-hand side of `~` must be a variable")
        self.lhs = lhs
        self.rhs = as_expr(rhs)

    def __repr__(self):
        return f"{self.lhs!r} ~ {self.rhs!r}"


def Eq(lhs, rhs) -> Equation:
    return Equation(lhs, rhs)


def as_expr(x) -> Expr:
    if isinstance(x, Expr):
        return x
    if isinstance(x, bool):
        raise TypeError("booleans are not expressions")
    if isinstance(x, (int, float)):
        return Const(x)
    try:  # numpy scalars
        return Const(float(x))
    except Exception as e:  # pragma: no cover
        raise TypeError(f"cannot use {type(x).__name__} in an expression") from e


# -- function vocabulary -------------------------------------------------------------
def exp(x): return Call("exp", (x,))
def log(x): return Call("log", (x,))
def sqrt(x): return Call("sqrt", (x,))
def tanh(x): return Call("tanh", (x,))
def hmin(a, b): return Call("min", (a, b))
def hmax(a, b): return Call("max", (a, b))
def clamp(x, lo, hi): return Call("clamp", (x, lo, hi))


def step_func(x):
    """`step_func(x) = (tanh(5x) + 1) / 2` (`/root/reference/src/HydroModels.jl:204`).

    Kept as one primitive so each back end chooses its evaluation: the oracle evaluates the
    literal tanh form, the CUDA lowering the algebraically identical logistic form."""
    return Call("step", (x,))


def variables(names: Union[str, Iterable[str]]) -> List[Sym]:
    """`@variables a b c` (HydroModelCore macro re-exported by the reference)."""
    if isinstance(names, str):
        names = names.replace(",", " ").split()
    return [Sym(n, False) for n in names]


def parameters(names: Union[str, Iterable[str]]) -> List[Sym]:
    """`@parameters p1 p2`."""
    if isinstance(names, str):
        names = names.replace(",", " ").split()
    return [Sym(n, True) for n in names]


# -- traversal helpers ---------------------------------------------------------------
def walk(e: Expr):
    """Post-order traversal (children before parents), each distinct node once."""
    seen = set()
    out = []

    def rec(n):
        if n._key in seen:
            return
        if isinstance(n, Call):
            for a in n.args:
                rec(a)
        seen.add(n._key)
        out.append(n)

    rec(e)
    return out


def free_symbols(e: Expr) -> Tuple[List[str], List[str]]:
    """(variable names, parameter names) in first-appearance order."""
    vs, ps = [], []
    for n in walk(e):
        if isinstance(n, Sym):
            (ps if n.is_param else vs).append(n.name)
    return vs, ps


def substitute(e: Expr, mapping: Dict[str, Expr]) -> Expr:
    """Replace variables by expressions (by name)."""
    memo = {}

    def rec(n):
        k = n._key
        if k in memo:
            return memo[k]
        if isinstance(n, Sym) and not n.is_param and n.name in mapping:
            r = mapping[n.name]
        elif isinstance(n, Call):
            r = Call(n.op, tuple(rec(a) for a in n.args))
        else:
            r = n
        memo[k] = r
        return r

    return rec(e)


# -- formula-string parser (the YAML schema's `formula:` strings) ---------------------
_FUNCS = {
    "exp": exp, "log": log, "sqrt": sqrt, "tanh": tanh, "abs": lambda x: Call("abs", (x,)),
    "min": hmin, "max": hmax, "clamp": clamp, "step_func": step_func,
}


class FormulaError(ValueError):
    pass


def parse_expr(text: str, symbols: Dict[str, Sym]) -> Expr:
    """Parse the reference's formula vocabulary into an expression tree.

    Grammar = Julia arithmetic restricted to the closed vocabulary of
    `ext/HydroModelsYAMLExt/expression_parser.jl:113-143`; `^` is power; implicit
    numeric-literal multiplication (`2x4`) as used in `@unithydro` blocks
    (`test/base/run_single_lumped.jl:132`) is accepted."""
    import re

    src = text.strip()
    src = re.sub(r"(?<![\w.])(\d+(?:\.\d+)?)([A-Za-z_])", r"\1*\2", src)  # 2x4 -> 2*x4
    src = src.replace("^", "**")
    try:
        tree = ast.parse(src, mode="eval").body
    except SyntaxError as e:
        raise FormulaError(f"Invalid formula {text!r}: {e.msg}") from e

    def rec(n) -> Expr:
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
            return Const(n.value)
        if isinstance(n, ast.Name):
            if n.id not in symbols:
                raise FormulaError(f"Undefined variable '{n.id}' detected in expression: `{text}`")
            return symbols[n.id]
        if isinstance(n, ast.UnaryOp):
            if isinstance(n.op, ast.USub):
                v = rec(n.operand)
                return Const(-v.value) if isinstance(v, Const) else -v
            if isinstance(n.op, ast.UAdd):
                return rec(n.operand)
        if isinstance(n, ast.BinOp):
            ops = {ast.Add: "add", ast.Sub: "sub", ast.Mult: "mul", ast.Div: "div", ast.Pow: "pow"}
            for k, v in ops.items():
                if isinstance(n.op, k):
                    return Call(v, (rec(n.left), rec(n.right)))
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and not n.keywords:
            if n.func.id not in _FUNCS:
                raise FormulaError(f"Unknown function: {n.func.id}")
            return _FUNCS[n.func.id](*[rec(a) for a in n.args])
        raise FormulaError(f"Unsupported syntax in formula {text!r}")

    return rec(tree)


def parse_equation(text: str, symbols: Dict[str, Sym]) -> Equation:
    if "~" not in text:
        raise FormulaError(f"Expected `lhs ~ rhs`, got {text!r}")
    lhs, rhs = text.split("~", 1)
    lhs = lhs.strip()
    if lhs not in symbols:
        raise FormulaError(f"Undefined variable '{lhs}' detected in expression: `{text}`")
    return Equation(symbols[lhs], parse_expr(rhs, symbols))


class Symbols:
    """A small namespace standing in for a Julia scope with `@variables`/`@parameters`."""

    def __init__(self, variables: Union[str, Iterable[str]] = "", parameters: Union[str, Iterable[str]] = ""):
        self.table: Dict[str, Sym] = {}
        self.add_variables(variables)
        self.add_parameters(parameters)

    def add_variables(self, names):
        for s in variables(names):
            self.table[s.name] = s
        return self

    def add_parameters(self, names):
        for s in parameters(names):
            self.table[s.name] = s
        return self

    def __getitem__(self, name: str) -> Sym:
        return self.table[name]

    def __getattr__(self, name: str) -> Sym:
        try:
            return self.__dict__["table"][name]
        except KeyError:
            raise AttributeError(name)

    def expr(self, text: str) -> Expr:
        return parse_expr(text, self.table)

    def eq(self, text: str) -> Equation:
        return parse_equation(text, self.table)


# -- scalar evaluation (used for UH weights on the host and by tests) ------------------
def evaluate(e: Expr, env: Dict[str, float]) -> float:
    """Evaluate with Python floats (IEEE double), literal operation order."""
    memo = {}

    def rec(n):
        k = n._key
        if k in memo:
            return memo[k]
        if isinstance(n, Const):
            r = n.value
        elif isinstance(n, Sym):
            r = float(env[n.name])
        else:
            a = [rec(x) for x in n.args]
            op = n.op
            if op == "add": r = a[0] + a[1]
            elif op == "sub": r = a[0] - a[1]
            elif op == "mul": r = a[0] * a[1]
            elif op == "div": r = a[0] / a[1]
            elif op == "neg": r = -a[0]
            elif op == "pow": r = math.pow(a[0], a[1])
            elif op == "exp": r = math.exp(a[0])
            elif op == "log": r = math.log(a[0])
            elif op == "sqrt": r = math.sqrt(a[0])
            elif op == "abs": r = abs(a[0])
            elif op == "tanh": r = math.tanh(a[0])
            elif op == "step": r = (math.tanh(5.0 * a[0]) + 1.0) * 0.5
            elif op == "min": r = min(a[0], a[1])
            elif op == "max": r = max(a[0], a[1])
            elif op == "clamp": r = min(max(a[0], a[1]), a[2])
            else:  # pragma: no cover
                raise ValueError(op)
        memo[k] = r
        return r

    return rec(e)


def evaluate_np(e: Expr, env):
    """Vectorised (numpy) evaluation, literal operation order — host-side helper for O(N·K)
    preparation work such as unit-hydrograph ordinates (never used for the time stepping)."""
    import numpy as np

    memo = {}
    for n in walk(e):
        k = n._key
        if isinstance(n, Const):
            memo[k] = np.float64(n.value)
        elif isinstance(n, Sym):
            memo[k] = env[n.name]
        else:
            a = [memo[x._key] for x in n.args]
            op = n.op
            with np.errstate(all="ignore"):
                if op == "add": r = a[0] + a[1]
                elif op == "sub": r = a[0] - a[1]
                elif op == "mul": r = a[0] * a[1]
                elif op == "div": r = a[0] / a[1]
                elif op == "neg": r = -a[0]
                elif op == "pow": r = np.power(a[0], a[1])
                elif op == "exp": r = np.exp(a[0])
                elif op == "log": r = np.log(a[0])
                elif op == "sqrt": r = np.sqrt(a[0])
                elif op == "abs": r = np.abs(a[0])
                elif op == "tanh": r = np.tanh(a[0])
                elif op == "step": r = (np.tanh(5.0 * a[0]) + 1.0) * 0.5
                elif op == "min": r = np.minimum(a[0], a[1])
                elif op == "max": r = np.maximum(a[0], a[1])
                elif op == "clamp": r = np.minimum(np.maximum(a[0], a[1]), a[2])
                else:  # pragma: no cover
                    raise ValueError(op)
            memo[k] = r
    return memo[e._key]
