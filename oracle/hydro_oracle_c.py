"""ctypes driver of the C oracle (`hydro_oracle.c`): same component semantics and model
composition as `hydro_oracle.py`, with the per-component hot loops (two-pass bucket, UH
convolution, Jacobi routing) executed by the C restatement — multithreaded over node blocks
(OpenMP).  TEST INFRASTRUCTURE ONLY (see the header of hydro_oracle.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List

import numpy as np

import hydro_oracle as ho
from hydromodels_jl_b200.components import (GridAggregation, HydroBucket, HydroFlux, HydroModel, HydroRoute, NeuralFlux,
                                            UnitHydrograph, normalize_config)
from hydromodels_jl_b200.symbolic import Call, Const, Sym, walk

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libhydro_oracle.so")
OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "neg": 4, "pow": 5, "exp": 6, "log": 7, "sqrt": 8, "abs": 9, "tanh": 10,
       "step": 11, "min": 12, "max": 13}
ACTS = {"identity": 0, "tanh": 1, "leakyrelu": 2, "relu": 3, "sigmoid": 4}


class Instr(C.Structure):
    _fields_ = [("op", C.c_int32), ("dst", C.c_int32), ("a", C.c_int32), ("b", C.c_int32)]


class NN(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("n_in", C.c_int32 * 8), ("n_out", C.c_int32 * 8), ("act", C.c_int32 * 8),
                ("offset", C.c_int32), ("in_regs", C.c_int32 * 16), ("out_regs", C.c_int32 * 16)]


class Prog(C.Structure):
    _fields_ = [("n_instr", C.c_int32), ("n_reg", C.c_int32), ("instr", C.POINTER(Instr)), ("n_const", C.c_int32),
                ("const_regs", C.POINTER(C.c_int32)), ("const_vals", C.POINTER(C.c_double)), ("n_nn", C.c_int32),
                ("nns", C.POINTER(NN)), ("nn_params", C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hydro_oracle.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB)
    return _lib


def max_threads() -> int:
    return lib().ho_max_threads()


class _Compiled:
    """Flux list + state derivatives → VM program.  Registers: inputs, states, params, consts, temps."""

    def __init__(self, input_names, state_names, param_names, fluxes, dexprs, nn_params: Dict[str, np.ndarray]):
        self.reg: Dict[str, int] = {}
        for nm in list(input_names) + list(state_names):
            self.reg[nm] = len(self.reg)
        self.preg = {nm: len(self.reg) + i for i, nm in enumerate(param_names)}
        self.n = len(self.reg) + len(param_names)
        self.consts: Dict[str, int] = {}
        self.const_vals: List[float] = []
        self.instr: List[tuple] = []
        self.nns: List[NN] = []
        nn_flat, nn_off = [], {}
        self.flux_regs = []
        for f in fluxes:
            if isinstance(f, NeuralFlux):
                ch = f.chain
                if ch.name not in nn_off:
                    nn_off[ch.name] = sum(len(v) for v in nn_flat)
                    v = np.asarray(nn_params[ch.name], dtype=np.float64).reshape(-1)
                    if v.size != ch.n_params:
                        raise ValueError(f"chain {ch.name} expects {ch.n_params} parameters, got {v.size}")
                    nn_flat.append(v)
                nn = NN()
                nn.n_layers = len(ch.layers)
                for i, l in enumerate(ch.layers):
                    nn.n_in[i], nn.n_out[i], nn.act[i] = l.n_in, l.n_out, ACTS[l.activation]
                nn.offset = nn_off[ch.name]
                for i, nm in enumerate(f.input_names):
                    nn.in_regs[i] = self.reg[nm]
                for i, nm in enumerate(f.output_names):
                    r = self.n
                    self.n += 1
                    self.reg[nm] = r
                    nn.out_regs[i] = r
                    self.flux_regs.append(r)
                self.instr.append((14, 0, len(self.nns), -1))
                self.nns.append(nn)
            else:
                for eq in f.eqs:
                    r = self.expr(eq.rhs)
                    self.reg[eq.lhs.name] = r
                    self.flux_regs.append(r)
        self.dstate_regs = [self.expr(e) for e in dexprs]
        self.nn_flat = np.ascontiguousarray(np.concatenate(nn_flat)) if nn_flat else np.zeros(1)

    def expr(self, e) -> int:
        memo = {}
        for n in walk(e):
            k = n.key()
            if isinstance(n, Const):
                hx = float(n.value).hex()
                if hx not in self.consts:
                    self.consts[hx] = self.n
                    self.const_vals.append(float(n.value))
                    self.n += 1
                memo[k] = self.consts[hx]
            elif isinstance(n, Sym):
                table = self.preg if n.is_param else self.reg
                if n.name not in table:
                    raise KeyError(f"{'Parameter' if n.is_param else 'Variable'} {n.name} does not exist")
                memo[k] = table[n.name]
            else:
                a = [memo[x.key()] for x in n.args]
                if n.op == "clamp":
                    t = self._emit(OPS["max"], a[0], a[1])
                    memo[k] = self._emit(OPS["min"], t, a[2])
                else:
                    memo[k] = self._emit(OPS[n.op], a[0], a[1] if len(a) > 1 else -1)
        return memo[e.key()]

    def _emit(self, op, a, b) -> int:
        r = self.n
        self.n += 1
        self.instr.append((op, r, a, b))
        return r

    def prog(self):
        arr = (Instr * max(1, len(self.instr)))(*[Instr(*i) for i in self.instr])
        cregs = np.array(list(self.consts.values()) or [0], dtype=np.int32)
        cvals = np.array(self.const_vals or [0.0], dtype=np.float64)
        nns = (NN * max(1, len(self.nns)))(*self.nns)
        p = Prog(len(self.instr), self.n, arr, len(self.const_vals), cregs.ctypes.data_as(C.POINTER(C.c_int32)),
                 cvals.ctypes.data_as(C.POINTER(C.c_double)), len(self.nns), nns,
                 self.nn_flat.ctypes.data_as(C.POINTER(C.c_double)))
        p._keep = (arr, cregs, cvals, nns, self.nn_flat)
        return p


def _ptr(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _to_tnv(x3):
    """[V, N, T] (any layout) → C-contiguous [T][N][V]."""
    return np.ascontiguousarray(np.transpose(x3, (2, 1, 0)))


def run_bucket(bucket: HydroBucket, input, params, config=None, initstates=None, n_threads: int = 0):
    params = ho._as_componentvector(params)
    cfg = normalize_config(config)
    ho._solver_ok(cfg)
    x = np.asarray(input, dtype=np.float64)
    multi = bucket.htypes is not None
    if multi != (x.ndim == 3):
        return ho.run_bucket(bucket, input, params, config, initstates)   # raises the reference's error
    x3 = x if multi else x[:, None, :]
    V, N, T = x3.shape
    if V != len(bucket.input_names):
        raise AssertionError("Input variables length mismatch")
    if multi:
        pe = ho.expand_component_params(params, bucket.param_names, bucket.htypes)
    else:
        pe = {k: np.array([v]) for k, v in ho._scalar_params(params, bucket.param_names).items()}
    pmat = np.ascontiguousarray(np.stack([pe[k] for k in bucket.param_names])) if bucket.param_names else np.zeros((1, N))
    S, O = len(bucket.state_names), len(bucket.output_names)
    ho._timeidx(cfg, T)
    comp = _Compiled(bucket.input_names, bucket.state_names, bucket.param_names, bucket.fluxes,
                     [d.expr for d in bucket.dfluxes], params.get("nns", {}))
    prog = comp.prog()
    init = None
    if S:
        init = np.ascontiguousarray(ho._init_states(initstates, bucket.state_names, N if multi else None).reshape(S, N))
    out = np.zeros((T, N, S + O))
    fr = np.array(comp.flux_regs or [0], dtype=np.int32)
    dr = np.array(comp.dstate_regs or [0], dtype=np.int32)
    rc = lib().ho_run_bucket(C.c_int64(N), C.c_int64(T), V, S, len(bucket.param_names), O, _ptr(_to_tnv(x3)), _ptr(pmat),
                             _ptr(init), C.c_double(cfg.min_value), C.byref(prog), _ptr(fr, C.c_int32),
                             _ptr(dr, C.c_int32), _ptr(out), int(n_threads))
    if rc:
        raise MemoryError("ho_run_bucket failed")
    res = np.transpose(out, (2, 1, 0))
    return res if multi else res[:, 0, :]


def run_uh(uh: UnitHydrograph, input, params, config=None, n_threads: int = 0):
    params = ho._as_componentvector(params)
    x = np.asarray(input, dtype=np.float64)
    multi = uh.htypes is not None
    if multi != (x.ndim == 3):
        return ho.run_uh(uh, input, params, config)
    x3 = x if multi else x[:, None, :]
    N, T = x3.shape[1], x3.shape[2]
    if multi:
        pe = ho.expand_component_params(params, uh.param_names, uh.htypes)
    else:
        pe = {k: np.array([float(v)]) for k, v in ho._scalar_params(params, uh.param_names).items()}
    ws = [ho._uh_raw_weights(uh, {k: float(v[n]) for k, v in pe.items()}) for n in range(N)]
    K = np.array([len(w) for w in ws], dtype=np.int32)
    kmax = max(1, int(K.max()))
    W = np.zeros((kmax, N))
    for n, w in enumerate(ws):
        W[:len(w), n] = w
    xin = np.ascontiguousarray(x3[0].T)
    y = np.zeros((T, N))
    lib().ho_uh_conv(C.c_int64(N), C.c_int64(T), kmax, _ptr(xin), _ptr(W), _ptr(K, C.c_int32), _ptr(y), int(n_threads))
    res = y.T[None, :, :]
    if not multi:
        return x.copy() if K[0] == 0 else res[:, 0, :]
    return res


def run_route(route: HydroRoute, input, params, config=None, initstates=None):
    params = ho._as_componentvector(params)
    cfg = normalize_config(config)
    x = np.asarray(input, dtype=np.float64)
    if x.ndim != 3:
        return ho.run_route(route, input, params, config, initstates)
    V, N, T = x.shape
    pe = ho.expand_component_params(params, route.param_names, route.htypes)
    pmat = np.ascontiguousarray(np.stack([pe[k] for k in route.param_names])) if route.param_names else np.zeros((1, N))
    ho._timeidx(cfg, T)
    comp = _Compiled(route.input_names, route.state_names, route.param_names, route.rfluxes,
                     [route.dfluxes[0].expr], {})
    prog = comp.prog()
    init = np.ascontiguousarray(ho._init_states(initstates, route.state_names, N)[0])
    out = np.zeros((T, N, 2))
    if isinstance(route.aggr, GridAggregation):
        flw = np.ascontiguousarray(route.aggr.flwdir.astype(np.int32))
        r0 = np.ascontiguousarray((route.aggr.positions[:, 0] - 1).astype(np.int32))
        c0 = np.ascontiguousarray((route.aggr.positions[:, 1] - 1).astype(np.int32))
        R, Cc = flw.shape
        rc = lib().ho_run_route(C.c_int64(N), C.c_int64(T), V, len(route.param_names), _ptr(_to_tnv(x)), _ptr(pmat), _ptr(init),
                                C.c_double(cfg.min_value), C.byref(prog), comp.flux_regs[0], comp.dstate_regs[0], R, Cc,
                                _ptr(flw, C.c_int32), _ptr(r0, C.c_int32), _ptr(c0, C.c_int32), None, _ptr(out))
    else:
        down = -np.ones(N, dtype=np.int32)
        e = np.asarray(route.aggr.edges).reshape(-1, 2) - 1
        if len(set(e[:, 0].tolist())) != len(e):
            return ho.run_route(route, input, params, config, initstates)    # node with several outlets: numpy path
        down[e[:, 0]] = e[:, 1]
        rc = lib().ho_run_route(C.c_int64(N), C.c_int64(T), V, len(route.param_names), _ptr(_to_tnv(x)), _ptr(pmat), _ptr(init),
                                C.c_double(cfg.min_value), C.byref(prog), comp.flux_regs[0], comp.dstate_regs[0], 0, 0,
                                None, None, None, _ptr(down, C.c_int32), _ptr(out))
    if rc:
        raise MemoryError("ho_run_route failed")
    return np.transpose(out, (2, 1, 0))


def run_component(comp, input, params, config=None, initstates=None, n_threads: int = 0):
    if isinstance(comp, HydroModel):
        return run_model(comp, input, params, config, initstates, n_threads)
    if isinstance(comp, HydroBucket):
        return run_bucket(comp, input, params, config, initstates, n_threads)
    if isinstance(comp, UnitHydrograph):
        return run_uh(comp, input, params, config, n_threads)
    if isinstance(comp, HydroRoute):
        return run_route(comp, input, params, config, initstates)
    return ho.run_flux(comp, input, params, config)


def _collect_rows(blocks, indices, n_threads=0):
    """`_collect_rows` (`/root/reference/src/model.jl:178-192`) on Fortran-ordered blocks, C gather."""
    picks = []
    for idx in indices:
        r = idx
        for b in blocks:
            if r < b.shape[0]:
                picks.append((b, r))
                break
            r -= b.shape[0]
        else:
            raise IndexError(f"Row index {idx} out of bounds for concatenated blocks")
    k = len(picks)
    shape = blocks[0].shape[1:]
    out = np.empty((k,) + shape, dtype=np.float64, order="F")
    if k == 0:
        return out
    srcs = [np.asfortranarray(b) for b, _ in picks]
    NT = int(np.prod(shape))
    ptrs = (C.c_void_p * k)(*[b.ctypes.data for b in srcs])
    nrows = np.array([b.shape[0] for b in srcs], dtype=np.int32)
    rows = np.array([r for _, r in picks], dtype=np.int32)
    lib().ho_collect_rows(C.c_int64(NT), k, ptrs, _ptr(nrows, C.c_int32), _ptr(rows, C.c_int32),
                          out.ctypes.data_as(C.POINTER(C.c_double)), int(n_threads))
    return out


def run_model(model: HydroModel, input, params, config=None, initstates=None, n_threads: int = 0):
    """`HydroModel` functor (`/root/reference/src/model.jl:241-309`): sequential components, row
    plumbing by `_collect_rows` — identical to `hydro_oracle.run_model`, C hot loops."""
    params = ho._as_componentvector(params)
    x = np.asarray(input, dtype=np.float64)
    ncomp = len(model.components)
    cfgs = [normalize_config(c) for c in config] if isinstance(config, (tuple, list)) else [normalize_config(config)] * ncomp
    if x.shape[0] != len(model.input_names):
        raise AssertionError("Input variables length mismatch")
    N = x.shape[1] if x.ndim == 3 else None
    blocks = [np.asfortranarray(x)]
    for idx, comp, cfg in zip(model._varindices, model.components, cfgs):
        tmp = _collect_rows(blocks, idx, n_threads)
        st = ho._extract_component_states(initstates, comp.state_names, model.state_names, N) if comp.state_names else None
        blocks.append(np.asfortranarray(run_component(comp, tmp, params, cfg, st, n_threads)))
    return _collect_rows(blocks, model._outputindices, n_threads)
