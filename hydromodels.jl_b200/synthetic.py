"""Seeded synthetic forcings / parameters for the BASELINE configs.

`temp = 5 + 15 sin(2π(t-100)/365.25) + 5 z`, `lday = 0.5 + 0.15 sin(2π(t-80)/365.25)` (day
fraction as in `data/exphydro/01013500.csv`), `prcp = -6 ln(u')` on 35 % of days, Hamon-type
`pet/ep`; parameters uniform in the YAML bounds (`docs/examples/exphydro.yaml:14-49`) / around
`test/test_helpers.jl:152-185`.  Node `n` at step `t` depends only on `(seed, n, t)` — a slice
`[n0, n1)` generated on one rank equals the same slice of the full array (counter-based
Philox stream per node block), so shards regenerate their own basins locally.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

SEED = 20260119
_BLOCK = 4096   # nodes per Philox stream


def _node_blocks(n0: int, n1: int):
    b = n0 // _BLOCK
    while b * _BLOCK < n1:
        lo, hi = max(n0, b * _BLOCK), min(n1, (b + 1) * _BLOCK)
        yield b, lo - b * _BLOCK, hi - b * _BLOCK, lo - n0, hi - n0
        b += 1


def _uniforms(seed: int, stream: int, n0: int, n1: int, T: int, k: int) -> np.ndarray:
    """[k, n1-n0, T] uniforms; block `b` of 4096 nodes uses Philox key (seed, stream, b)."""
    out = np.empty((k, n1 - n0, T))
    for b, a, z, oa, oz in _node_blocks(n0, n1):
        g = np.random.Generator(np.random.Philox(key=[(seed << 20) + stream, b]))
        u = g.random((k, _BLOCK, T))
        out[:, oa:oz] = u[:, a:z]
    return out


def forcing(model: str, n0: int, n1: int, T: int, seed: int = SEED) -> np.ndarray:
    """Fortran-ordered `[V, n1-n0, T]` forcing for `model` ∈ exphydro | gr4j | hbv | m50
    (row order = the model's `input_names`)."""
    t = np.arange(1, T + 1, dtype=np.float64)
    u = _uniforms(seed, 1, n0, n1, T, 3)
    z = np.sqrt(-2.0 * np.log(1.0 - u[0])) * np.cos(2 * np.pi * u[1])          # Box–Muller
    temp = 5.0 + 15.0 * np.sin(2 * np.pi * (t - 100.0) / 365.25) + 5.0 * z
    lday = np.broadcast_to(0.5 + 0.15 * np.sin(2 * np.pi * (t - 80.0) / 365.25), temp.shape)
    w = u[2]
    with np.errstate(all="ignore"):
        prcp = np.where(w < 0.35, -6.0 * np.log(1.0 - w / 0.35 * (1 - 1e-12)), 0.0)
    pet = 29.8 * lday * 24 * 0.611 * np.exp(17.3 * temp / (temp + 237.3)) / (temp + 273.2)
    rows = {"exphydro": [temp, lday, prcp], "m50": [prcp, temp, lday], "gr4j": [prcp, pet],
            "hbv": [prcp, temp, pet]}[model]
    out = np.empty((len(rows), n1 - n0, T), order="F")
    for i, r in enumerate(rows):
        out[i] = r
    return out


_BOUNDS = {
    "exphydro": dict(Tmin=(-3.0, 0.0), Tmax=(0.0, 3.0), Df=(0.5, 5.0), Smax=(100.0, 2000.0), Qmax=(10.0, 50.0),
                     f=(0.005, 0.1)),
    "gr4j": dict(x1=(100.0, 1200.0), x2=(-5.0, 3.0), x3=(60.0, 300.0), x4=(1.1, 2.9)),
    "hbv": dict(TT=(-1.5, 1.5), CFMAX=(1.0, 8.0), CFR=(0.0, 0.1), CWH=(0.0, 0.2), LP=(0.3, 1.0), FC=(50.0, 500.0),
                BETA=(0.2, 1.0), PPERC=(0.01, 0.2), UZL=(5.0, 60.0), k0=(0.05, 0.5), k1=(0.01, 0.3), k2=(0.001, 0.1)),
    "m50": dict(Tmin=(-3.0, 0.0), Tmax=(0.0, 3.0), Df=(0.5, 5.0)),
}
INIT = {
    "exphydro": dict(snowpack=0.0, soilwater=1303.0),
    "gr4j": dict(soilwater=235.97, routingstore=45.47),
    "hbv": dict(snowpack=0.0, meltwater=0.0, soilwater=100.0, suz=3.0, slz=10.0),
    "m50": dict(snowpack=0.0, soilwater=1303.0),
}


def parameters(model: str, n0: int, n1: int, seed: int = SEED) -> Dict[str, np.ndarray]:
    """One parameter set per node, uniform in the bounds above (same slicing property)."""
    names = list(_BOUNDS[model])
    u = _uniforms(seed, 2, n0, n1, 1, len(names))[:, :, 0]
    p = {nm: lo + (hi - lo) * u[i] for i, (nm, (lo, hi)) in enumerate(_BOUNDS[model].items())}
    if model == "m50":
        p.update(snowpack_mean=np.full(n1 - n0, 20.0), snowpack_std=np.full(n1 - n0, 40.0),
                 soilwater_mean=np.full(n1 - n0, 1300.0), soilwater_std=np.full(n1 - n0, 150.0),
                 prcp_mean=np.full(n1 - n0, 2.1), prcp_std=np.full(n1 - n0, 4.6),
                 temp_mean=np.full(n1 - n0, 5.0), temp_std=np.full(n1 - n0, 11.0))
    return p
