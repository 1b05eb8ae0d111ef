"""CPU-only: NeuralBucket (`/root/reference/src/nn.jl:195-371`) — the oracle restatement worked by hand
on a tiny network, constructor checks of the host mirror, and the generated stepper body run through the g++ harness."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200.components import Chain, Dense

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hydro_oracle as ho  # noqa: E402


def tiny_bucket(htypes=None):
    flux = Chain((Dense(3, 2, "tanh"),), "flux_net")            # (1 state + 2 inputs) -> 2 fluxes
    state = Chain((Dense(3, 1, "identity"),), "state_net")      # (1 state + 2 fluxes) -> 1 delta
    outp = Chain((Dense(2, 1, "identity"),), "output_net")      # 2 fluxes -> 1 output
    return hm.NeuralBucket(name="nb", flux_network=flux, state_network=state, output_network=outp, n_inputs=2, n_states=1,
                           n_outputs=1, inputs=["prcp", "pet"], states=["storage"], outputs=["q"], htypes=htypes)


def test_neural_bucket_oracle_by_hand():
    nb = tiny_bucket()
    # flux_net: W[out=2, in=3] column-major then bias; state_net W[1,3]; output_net W[1,2]
    Wf = np.array([[0.5, -0.2, 0.1], [0.3, 0.4, -0.6]]); bf = np.array([0.05, -0.1])
    Ws = np.array([[-0.1, 0.7, -0.3]]); bs = np.array([0.02])
    Wo = np.array([[1.5, -2.0]]); bo = np.array([0.25])
    p = {"params": {}, "nns": dict(flux_net=np.concatenate([Wf.reshape(-1, order="F"), bf]), state_net=np.concatenate([Ws.reshape(-1, order="F"), bs]),
                     output_net=np.concatenate([Wo.reshape(-1, order="F"), bo]))}
    x = np.array([[1.0, 0.0, 2.0], [0.3, 0.2, 0.1]])
    out = ho.run_neural_bucket(nb, x, p, initstates=dict(storage=0.4))
    S, rows = 0.4, []
    for t in range(3):
        f = np.tanh(Wf @ np.array([S, x[0, t], x[1, t]]) + bf)            # fluxes from the PREVIOUS state
        S = S + (Ws @ np.array([S, f[0], f[1]]) + bs).item()                # residual update, no clamp
        rows.append((S, (Wo @ f + bo).item()))
    assert np.allclose(out, np.array(rows).T, rtol=0, atol=1e-15) and out.shape == (2, 3)
    assert out[0].min() < 0.4                                             # nothing keeps a NeuralBucket state positive
    # multi-node = replicated single node (`test/base/run_multi_bucket.jl` style metamorphic check)
    nb3 = tiny_bucket(htypes=[1, 1, 1])
    x3 = np.stack([x, x * 0.5, x], axis=1)
    out3 = ho.run_neural_bucket(nb3, x3, p, initstates=dict(storage=np.array([0.4, 0.0, 0.4])))
    assert np.array_equal(out3[:, 0], out) and np.array_equal(out3[:, 2], out) and not np.array_equal(out3[:, 1], out)
    assert np.array_equal(ho.run_neural_bucket(nb, x, p)[:, 0], ho.run_neural_bucket(nb, x[:, :1], p, initstates=dict(storage=0.0))[:, 0])


def test_neural_bucket_constructor_checks_and_names():
    nb = tiny_bucket()
    assert nb.get_input_names() == ["prcp", "pet"] and nb.get_state_names() == ["storage"] and nb.get_output_names() == ["q"]
    assert nb.get_nn_names() == ["flux_net", "state_net", "output_net"] and nb.get_param_names() == []
    with pytest.raises(AssertionError, match="Number of input names must match n_inputs"):
        hm.NeuralBucket(name="x", flux_network=nb.flux_network, state_network=nb.state_network, output_network=nb.output_network,
                        n_inputs=3, n_states=1, n_outputs=1, inputs=["a", "b"], states=["s"], outputs=["q"])
    with pytest.raises(Exception, match="state_network must map"):
        hm.NeuralBucket(name="x", flux_network=nb.flux_network, state_network=Chain((Dense(2, 1),), "s"), output_network=nb.output_network,
                        n_inputs=2, n_states=1, n_outputs=1, inputs=["a", "b"], states=["s"], outputs=["q"])


def test_neural_bucket_body_through_the_cpu_harness():
    sys.path.insert(0, os.path.dirname(__file__))
    from cpu_harness import run_body_on_cpu
    from hydromodels_jl_b200.lowering import lower_stepper
    rng = np.random.default_rng(5)
    flux = Chain((Dense(5, 8, "tanh"), Dense(8, 4, "leakyrelu")), "flux_net")
    state = Chain((Dense(6, 8, "tanh"), Dense(8, 2, "identity")), "state_net")
    outp = Chain((Dense(4, 3, "identity"),), "output_net")
    N, T = 7, 25
    nb = hm.NeuralBucket(name="nb", flux_network=flux, state_network=state, output_network=outp, n_inputs=3, n_states=2, n_outputs=3,
                         inputs=["prcp", "temp", "lday"], states=["s1", "s2"], outputs=["et", "q", "aux"], htypes=np.ones(N, dtype=int))
    p = {"params": {}, "nns": {c.name: c.init_params(i) * 0.7 + rng.normal(0, 0.05, c.n_params) for i, c in enumerate((flux, state, outp))}}
    x = rng.normal(size=(3, N, T))
    init = dict(s1=rng.normal(size=N), s2=rng.normal(size=N))
    prog = lower_stepper(nb)
    assert prog.row_names == ["s1", "s2", "et", "q", "aux"] and "hb_max(hb_minv" not in prog.body
    got = run_body_on_cpu(hm.B200Model(nb), x, p, initstates=init)
    ref = ho.run_neural_bucket(nb, x, p, initstates=init)
    assert np.allclose(got, ref, rtol=1e-10, atol=1e-12)
