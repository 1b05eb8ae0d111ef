"""Stand-alone `NeuralFlux` (`/root/reference/src/nn.jl:140-165`) and `norm`: the oracle against a hand-written dense
chain, and the generated body (g++ harness) against the oracle, on the shapes of the reference's own differential test
(`/root/reference/test/base/run_neural_flux.jl:17-24,41-48`: 2D, and 3D == the 2D result replicated over the nodes)."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200.lowering import LoweringError, lower_stepper

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.dirname(__file__))
import hydro_oracle as ho  # noqa: E402
from cpu_harness import run_body_on_cpu  # noqa: E402

S = hm.Symbols("a b c d e", "")
X2 = np.array([[1, 3, 3], [2, 2, 2], [1, 2, 1], [3, 1, 2]], dtype=float).T          # run_neural_flux.jl:17


def chain(n_out):
    return hm.Chain((hm.Dense(3, 16, "leakyrelu"), hm.Dense(16, 16, "leakyrelu"), hm.Dense(16, n_out, "leakyrelu")), name="testnn")


def by_hand(ch, x, flat):
    """act.(W*x .+ b) layer by layer with explicit loops (Lux.Dense; weight[out,in] column-major, then bias)."""
    h, off = np.array(x, dtype=float), 0
    for l in ch.layers:
        W = np.empty((l.n_out, l.n_in))
        for j in range(l.n_in):
            for i in range(l.n_out):
                W[i, j] = flat[off + j * l.n_out + i]
        off += l.n_in * l.n_out
        bias = flat[off:off + l.n_out]
        off += l.n_out
        z = np.array([[sum(W[i, k] * h[k, t] for k in range(l.n_in)) + bias[i] for t in range(h.shape[1])] for i in range(l.n_out)])
        h = {"leakyrelu": np.maximum(0.01 * z, z), "tanh": np.tanh(z), "identity": z}[l.activation]
    return h


@pytest.mark.parametrize("n_out", [1, 2])
def test_neural_flux_names_oracle_and_body(n_out):
    ch = chain(n_out)
    outs = [S.d, S.e][:n_out]
    f = hm.NeuralFlux([S.a, S.b, S.c], outs, ch)
    assert set(f.get_input_names()) == {"a", "b", "c"} and f.get_param_names() == [] and f.get_nn_names() == ["testnn"]
    assert f.get_output_names() == [o.name for o in outs]
    ps = ch.init_params(42) + np.random.default_rng(1).normal(0, 0.05, ch.n_params)
    pas = {"nns": {"testnn": ps}}                                  # ComponentVector(nns=(testnn=nn_ps,)): no `params` entry
    ref = ho.run_component(f, X2, pas)
    assert ref.shape == (n_out, 4) and np.allclose(ref, by_hand(ch, X2, ps), rtol=1e-13, atol=1e-15)
    got = run_body_on_cpu(hm.B200Model(f), X2, pas)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12)
    # multiple nodes: the 2D result replicated (run_neural_flux.jl:22-24)
    X3 = np.asfortranarray(np.repeat(X2[:, None, :], 10, axis=1))
    ref3 = ho.run_component(f, X3, pas)
    assert np.array_equal(ref3, np.repeat(ref[:, None, :], 10, axis=1))
    got3 = run_body_on_cpu(hm.B200Model(f), X3, pas)
    assert got3.shape == (n_out, 10, 4) and np.allclose(got3, ref3, rtol=1e-10, atol=1e-12)


def test_norm_function_is_applied_before_the_chain():
    ch = chain(1)
    ps = ch.init_params(7)
    f = hm.NeuralFlux([S.a, S.b, S.c], S.d, ch, norm=lambda v: [(v[0] - 2.0) / 3.0, v[1], hm.symbolic.tanh(v[2])])
    pas = {"nns": {"testnn": ps}}
    xn = np.stack([(X2[0] - 2.0) / 3.0, X2[1], np.tanh(X2[2])])
    ref = ho.run_component(f, X2, pas)
    assert np.allclose(ref, by_hand(ch, xn, ps), rtol=1e-13, atol=1e-15)
    assert np.allclose(run_body_on_cpu(hm.B200Model(f), X2, pas), ref, rtol=1e-10, atol=1e-12)
    with pytest.raises(Exception, match="one expression per input"):
        hm.NeuralFlux([S.a, S.b, S.c], S.d, ch, norm=lambda v: v[:2])
    with pytest.raises(ValueError, match="own inputs"):
        hm.NeuralFlux([S.a, S.b, S.c], S.d, ch, norm=lambda v: [v[0] + S.e, v[1], v[2]])


def test_m50_with_three_components_lowers_and_matches_the_oracle():
    """`/root/reference/test/models/m50.jl`: snow bucket, soil bucket (two neural fluxes), `flow ~ exp(log_flow)`."""
    from hydromodels_jl_b200 import synthetic as sy
    N, T = 6, 80
    model = hm.models.m50_three_components(htypes=np.arange(1, N + 1))
    assert model.get_nn_names() == ["epnn", "qnn"] and model.var_names[-1] == "flow" and len(model.components) == 3
    inp = sy.forcing("m50", 0, N, T)
    et, q = hm.models.m50_chains()
    p = {"params": sy.parameters("m50", 0, N), "nns": {"epnn": et.init_params(1) * 0.3, "qnn": q.init_params(2) * 0.3}}
    init = {k: np.full(N, v) for k, v in sy.INIT["m50"].items()}
    ref = ho.run_model(model, inp, p, initstates=init)
    got = run_body_on_cpu(hm.B200Model(model), inp, p, initstates=init)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12)
    assert np.allclose(ref[-1], np.exp(ref[model.var_names.index("log_flow")]))


def test_model_with_a_neural_flux_component():
    """A NeuralFlux as a component of a HydroModel (between a flux and a flux), single node and multi-node."""
    s = hm.Symbols("x y u v w z", "k")
    ch = hm.Chain((hm.Dense(2, 4, "tanh"), hm.Dense(4, 1, "identity")), name="mid")
    for ht in (None, np.array([1, 2, 1])):
        pre = hm.HydroFlux(["u ~ x * k", "v ~ y + k"], symbols=s, htypes=ht)
        nf = hm.NeuralFlux([s.u, s.v], s.w, ch)
        post = hm.HydroFlux("z ~ exp(w) + u", symbols=s, htypes=ht)
        model = hm.HydroModel(name="sandwich", components=(pre, nf, post), input_names=["x", "y"])
        rng = np.random.default_rng(3)
        shape = (2, 20) if ht is None else (2, 3, 20)
        inp = np.asfortranarray(rng.normal(size=shape))
        p = {"params": dict(k=1.5 if ht is None else np.array([1.5, 0.5])), "nns": {"mid": ch.init_params(2)}}
        ref = ho.run_model(model, inp, p)
        got = run_body_on_cpu(hm.B200Model(model), inp, p)
        assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-10, atol=1e-12)
