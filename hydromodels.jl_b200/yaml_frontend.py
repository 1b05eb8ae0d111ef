"""Model files in the reference's YAML schema → `HydroModel` (host-side convenience only).

The schema is the reference's own (`/root/reference/docs/examples/exphydro.yaml`, loader
`ext/HydroModelsYAMLExt/yaml_loader.jl` + `component_builder.jl`): `parameters:` (names with
metadata), `components:` (`type: HydroBucket`, `fluxes:` / `state_fluxes:` lists of `formula:`
strings in the closed vocabulary of `expression_parser.jl:113-143`), `model.components` (order) and
`config`.  Only the pieces on the forward-run path are read; the CSV executor of the extension is out
of scope."""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

import yaml

from .components import HydroBucket, HydroConfig, HydroFlux, HydroModel, SolverType, StateFlux
from .symbolic import Symbols

_FUNCS = {"exp", "log", "sqrt", "abs", "tanh", "min", "max", "clamp", "step_func"}


def _names(formula: str) -> List[str]:
    return [t for t in re.findall(r"[A-Za-z_][A-Za-z_0-9]*", formula) if t not in _FUNCS]


def load_model_from_yaml(text_or_path: str, htypes=None, input_names=None) -> Tuple[HydroModel, Dict[str, float], HydroConfig]:
    """Returns `(model, default parameter values, config)`."""
    text = text_or_path
    if "\n" not in text_or_path and text_or_path.endswith((".yaml", ".yml")):
        with open(text_or_path) as f:
            text = f.read()
    doc = yaml.safe_load(text)
    if doc.get("schema", "hydromodels") != "hydromodels":
        raise ValueError(f"Unsupported schema: {doc.get('schema')}")
    pdefs = doc.get("parameters", {}) or {}
    params = list(pdefs)
    formulas = [f["formula"] for c in doc["components"] for f in (c.get("fluxes") or []) + (c.get("state_fluxes") or [])]
    variables = []
    for fo in formulas:
        for nm in _names(fo):
            if nm not in params and nm not in variables:
                variables.append(nm)
    sym = Symbols(variables, params)
    comps = {}
    for c in doc["components"]:
        if c.get("type") != "HydroBucket":
            raise NotImplementedError(f"component type {c.get('type')} is not supported by this loader")
        fluxes = [HydroFlux(f["formula"], symbols=sym) for f in c.get("fluxes") or []]
        dfluxes = [StateFlux(f["formula"], symbols=sym) for f in c.get("state_fluxes") or []]
        comps[c["name"]] = HydroBucket(name=c["name"], fluxes=fluxes, dfluxes=dfluxes, htypes=htypes)
    order = (doc.get("model") or {}).get("components") or list(comps)
    model = HydroModel(name=(doc.get("model") or {}).get("name"), components=tuple(comps[n] for n in order),
                       input_names=input_names)
    cfgd = doc.get("config") or {}
    cfg = HydroConfig(solver=SolverType[cfgd.get("solver", "MutableSolver")], min_value=float(cfgd.get("min_value", 1e-6)))
    defaults = {k: float(v["default"]) for k, v in pdefs.items() if isinstance(v, dict) and "default" in v}
    return model, defaults, cfg
