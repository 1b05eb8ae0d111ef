"""BASELINE.json's configurations at FULL size on one B200, checked through size-independent
properties (no oracle can run 3.65e8 basin-steps in a test): every row finite, states >= min_value,
algebraic row identities, and a random sample of nodes compared with the CPU oracle run on exactly
those nodes (nodes are independent, so a sampled node must match a stand-alone run of it)."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceRun, RoutedDeviceRun

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import hydro_oracle as ho  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-10, 1e-12


def assert_parity(got, ref):
    err = np.abs(got - ref) - (RTOL * np.abs(ref) + ATOL)
    assert np.all(np.isfinite(got)) and err.max() <= 0, f"max excess {err.max():.3e}"


def _device_case(name, N, T, rows):
    import torch
    import bench
    hm.engine.init(0)
    model, p, init = bench.build_problem(hm, sy, name, 0, N, T)
    d_in = bench.device_forcing(torch, name, N, T, 1234)
    d_out = torch.empty((T, N, rows), device="cuda", dtype=torch.float64)
    return torch, model, p, init, d_in, d_out


def _sample_check(torch, name, model, p, init, d_in, d_out, N, nsamp=48, chains=None):
    idx = np.sort(np.random.default_rng(3).choice(N, nsamp, replace=False))
    ti = torch.from_numpy(idx).cuda()
    got = d_out.index_select(1, ti).cpu().numpy().transpose(2, 1, 0)              # [rows, nsamp, T]
    inp = np.asfortranarray(d_in.index_select(1, ti).cpu().numpy().transpose(2, 1, 0))
    sub = getattr(hm.models, name)(htypes=np.arange(1, nsamp + 1))
    ps = {"params": {k: np.asarray(v)[idx] for k, v in p["params"].items()}}
    if "nns" in p:
        ps["nns"] = p["nns"]
    isub = {k: v[idx] for k, v in init.items()}
    assert_parity(got, ho.run_model(sub, inp, ps, initstates=isub))


def test_exphydro_1M_basins_365d():
    N, T = 1_000_000, 365
    torch, model, p, init, d_in, d_out = _device_case("exphydro", N, T, 10)
    DeviceRun(hm.B200Model(model), N, T, p, initstates=init).launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d_out).all())
    assert float(d_out[:, :, :2].min()) >= 1e-6                                   # states >= min_value
    assert bool(torch.equal(d_out[:, :, 9], d_out[:, :, 7] + d_out[:, :, 8]))     # flow = baseflow + surfaceflow
    assert float(d_out[:, :, 2:4].min()) >= 0.0                                   # snowfall, rainfall >= 0
    _sample_check(torch, "exphydro", model, p, init, d_in, d_out, N)


def test_gr4j_uh_100k_basins_10y():
    N, T = 100_000, 3650
    torch, model, p, init, d_in, d_out = _device_case("gr4j", N, T, 15)
    DeviceRun(hm.B200Model(model), N, T, p, initstates=init).launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d_out).all()) and float(d_out[:, :, :2].min()) >= 1e-6
    r = {n: i for i, n in enumerate(model.var_names)}
    # unit hydrographs conserve mass: sum over time of routed flow <= sum of inflow (weights sum to 1, causal)
    sf, sfr = d_out[:, :, r["slowflow"]].sum(0), d_out[:, :, r["slowflow_routed"]].sum(0)
    assert bool((sfr <= sf * (1 + 1e-12) + 1e-9).all()) and bool((sfr >= 0.9 * sf - 1e-6).all())
    assert bool(torch.allclose(d_out[:, :, r["slowflow"]] * (1.0 / 9.0), d_out[:, :, r["fastflow"]], rtol=1e-13, atol=1e-15))
    _sample_check(torch, "gr4j", model, p, init, d_in, d_out, N, nsamp=32)


def test_m50_50k_basins_10y():
    N, T = 50_000, 3650
    torch, model, p, init, d_in, d_out = _device_case("m50", N, T, 12)
    DeviceRun(hm.B200Model(model), N, T, p, initstates=init).launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d_out).all()) and float(d_out[:, :, :2].min()) >= 1e-6
    _sample_check(torch, "m50", model, p, init, d_in, d_out, N, nsamp=16)


def test_hbv_ensemble_1M_members_20y():
    """1 M parameter sets x 1 basin x 7305 d, fused -NSE; members sampled against the oracle; a member
    whose parameters generated the target scores exactly -1."""
    import torch
    hm.engine.init(0)
    M, T, warm = 1_000_000, 7305, 366
    model = hm.models.hbv()
    pm = sy.parameters("hbv", 0, M)
    init = dict(snowpack=0.0, meltwater=0.0, soilwater=100.0, suz=3.0, slz=10.0)
    inp = np.asfortranarray(sy.forcing("hbv", 0, 1, T)[:, 0, :])
    truth = ho.run_model(model, inp, {"params": {k: v[123456] for k, v in pm.items()}}, initstates=init)[-1]
    obj = hm.B200Model(model).objective(inp, {"params": pm}, truth, metric="NSE", warm_up=warm, initstates=init, n_members=M)
    assert obj.shape == (M,) and np.all(np.isfinite(obj)) and obj.min() >= -1.0 - 1e-12
    assert abs(obj[123456] + 1.0) < 1e-9 and np.argmin(obj) == 123456
    for i in (0, 999_999, 500_001):
        sim = ho.run_model(model, inp, {"params": {k: v[i] for k, v in pm.items()}}, initstates=init)[-1]
        assert abs(obj[i] - ho.objective("NSE", truth, sim, warm_up=warm)) < 1e-9 * max(1.0, abs(obj[i]))


def test_grid_2048x2048_d8_routing():
    """Distributed ExpHydro + D8 route on a 2048 x 2048 grid (one GPU's share of BASELINE config 5):
    row identities, positivity, outlet accumulation, and a 96 x 96 corner window — closed under
    'upstream of' by construction — re-run alone and compared with the oracle."""
    import torch
    import bench
    hm.engine.init(0)
    side, T = 2048, 20
    N = side * side
    rng = np.random.default_rng(0)
    flw = rng.choice(np.array([1, 2, 4]), size=(side, side))                    # E, SE, S only: flow never enters the top-left window from outside
    rr, cc = np.divmod(np.arange(N), side)
    pos = np.stack([rr + 1, cc + 1], axis=1)
    model = hm.models.exphydro_grid(flw, pos)
    pp = dict(sy.parameters("exphydro", 0, N), area_coef=np.full(N, 1.0), lag=rng.uniform(0.1, 2.0, N))
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=np.zeros(N))
    d_in = bench.device_forcing(torch, "exphydro", N, T, 99)
    d_out = torch.empty((T, N, 13), device="cuda", dtype=torch.float64)
    RoutedDeviceRun(hm.B200Model(model), N, T, {"params": pp}, initstates=init).launch(
        d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d_out).all()) and float(d_out[:, :, :3].min()) >= 1e-6
    assert bool(torch.equal(d_out[:, :, 11], d_out[:, :, 10]))                  # q = flow * 1.0
    grid_q = d_out[:, :, 12].mean(0).view(side, side)
    assert float(grid_q[-64:, -64:].mean()) > float(grid_q[:64, :64].mean())    # flow accumulates downstream
    w = 96
    win = (rr < w) & (cc < w)
    idx = np.nonzero(win)[0]
    ti = torch.from_numpy(idx).cuda()
    got = d_out.index_select(1, ti).cpu().numpy().transpose(2, 1, 0)
    inp = np.asfortranarray(d_in.index_select(1, ti).cpu().numpy().transpose(2, 1, 0))
    sub = hm.models.exphydro_grid(flw[:w, :w], np.stack([rr[idx] + 1, cc[idx] + 1], axis=1))
    ps = {"params": {k: v[idx] for k, v in pp.items()}}
    ref = ho.run_model(sub, inp, ps, initstates={k: v[idx] for k, v in init.items()})
    assert_parity(got, ref)


def test_grid_2048x2048_wave_kernel_is_bit_identical_to_the_two_pass_path():
    """BASELINE config 5 at one GPU's size, all eight D8 directions + sinks: the single-pass wave kernel (32 768 blocks, four
    launches of <= 16 steps, 4 blocks per SM exchanging outflows through L2) against the generic stepper + per-step route path —
    every one of the 4.19 M x 50 x 13 values, plus a relaunch (tickets / tags reset) and the strip + ghost-row decomposition."""
    import torch
    import bench
    from hydromodels_jl_b200.engine import GridDeviceRun
    from hydromodels_jl_b200 import parallel as par
    hm.engine.init(0)
    side, T = 2048, 50
    N = side * side
    rng = np.random.default_rng(1)
    flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128, 0]), size=(side, side), p=[.25, .2, .25, .05, .05, .05, .05, .05, .05])
    rr, cc = np.divmod(np.arange(N), side)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 2.0, N))}
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 1, N))
    d_in = bench.device_forcing(torch, "exphydro", N, T, 5)
    st = torch.cuda.current_stream().cuda_stream
    eng = hm.B200Model(model)
    ref = torch.empty((T, N, 13), device="cuda", dtype=torch.float64)
    RoutedDeviceRun(eng, N, T, p, initstates=init).launch(d_in.data_ptr(), ref.data_ptr(), st)
    got = torch.zeros((T, N, 13), device="cuda", dtype=torch.float64)
    run = GridDeviceRun(eng, N, T, p, initstates=init, variant="wave", steps_per_chunk=16)
    for rep in range(2):
        got.zero_()
        run.launch(d_in.data_ptr(), got.data_ptr(), st)
        assert run.check() == 16
        if not bool(torch.equal(got, ref)):
            bad = (got != ref).nonzero()
            t, n, r = [int(x) for x in bad[0]]
            raise AssertionError(f"rep {rep}: {bad.shape[0]} values differ; rows {torch.unique(bad[:, 2]).tolist()}, steps "
                                 f"{torch.unique(bad[:, 0]).tolist()[:12]}, first (t, n, row) = {(t, n, r)} at grid ({n // side}, {n % side}): "
                                 f"got {got[t, n, r].item()!r} ref {ref[t, n, r].item()!r}; cells {torch.unique(bad[:, 1]).numel()}")
    assert float(ref[:, :, 12].max()) > float(ref[:, :, 11].max())           # routing did accumulate flow
    # the middle strip of a 4-way decomposition, 16 ghost rows either side, first 16 steps from the initial states
    pl = par.GridStripPlan(side, side, 1, 4, 16)
    n_lo, n_hi = pl.node_range()
    strip = GridDeviceRun(eng, pl.n_local, T, p, initstates=init, variant="wave", steps_per_chunk=16, row_range=(pl.g0, pl.g1))
    s_in = d_in[:, n_lo:n_hi, :].contiguous()
    s_out = torch.zeros((16, pl.out_count, 13), device="cuda", dtype=torch.float64)
    b0 = torch.from_numpy(strip.bufs["init"].to_host((2, pl.n_local))).cuda()
    r0 = torch.from_numpy(strip.bufs["rinit"].to_host((pl.n_local,))).cuda()
    b1, r1 = torch.zeros_like(b0), torch.zeros_like(r0)
    strip.launch_chunk(s_in.data_ptr(), s_out.data_ptr(), 16, b0.data_ptr(), r0.data_ptr(), b1.data_ptr(), r1.data_ptr(),
                       pl.out_first, pl.out_count, st)
    strip.check()
    assert bool(torch.equal(s_out, ref[:16, pl.r0 * side:pl.r1 * side, :]))
