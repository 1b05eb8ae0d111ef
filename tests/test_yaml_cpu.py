"""CPU-only: a model file in the reference's YAML schema builds the same model as the Python DSL."""
import os
import sys

import numpy as np

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200.yaml_frontend import load_model_from_yaml

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.dirname(__file__))
import hydro_oracle as ho  # noqa: E402
from cpu_harness import run_body_on_cpu  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def test_yaml_model_equals_dsl_model(golden):
    model, defaults, cfg = load_model_from_yaml(os.path.join(HERE, "golden", "exphydro_model.yaml"),
                                                input_names=["temp", "lday", "prcp"])
    ref_model = hm.models.exphydro(docs_variant=True)
    assert model.var_names == ref_model.var_names and model.param_names == ["Tmin", "Tmax", "Df", "Smax", "Qmax", "f"]
    assert cfg.min_value == 1e-6 and defaults["Smax"] == 1709.46
    d = golden("exphydro_01013500.npz")
    inp = np.stack([d["temp"][:500], d["lday"][:500], d["prcp"][:500]])
    init = dict(snowpack=0.0, soilwater=1303.0)
    a = ho.run_model(model, inp, {"params": defaults}, cfg, initstates=init)
    b = ho.run_model(ref_model, inp, {"params": defaults}, cfg, initstates=init)
    assert np.array_equal(a, b)
    got = run_body_on_cpu(hm.B200Model(model), inp, {"params": defaults}, cfg, initstates=init)
    assert np.allclose(got, a, rtol=1e-10, atol=1e-12)
