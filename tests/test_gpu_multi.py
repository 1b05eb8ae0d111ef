"""Multi-GPU (needs >= 2 devices; skipped on a 1-GPU box): sharded basins and sharded D8 routing with
per-step ghost exchange over NCCL reproduce the single-GPU result bit for bit."""
import os
import sys

import numpy as np
import pytest

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        return hm.device_count()
    except Exception:
        return 0


def _grid_problem(R, Cc, T):
    rng = np.random.default_rng(7)
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    rr, cc = np.divmod(np.arange(R * Cc), Cc)
    pos = np.stack([rr + 1, cc + 1], axis=1)
    N = R * Cc
    model = hm.models.exphydro_grid(flw, pos)
    pp = dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
    return model, {"params": pp}, init, N


def _worker(rank, world, port, R, Cc, T, q):
    import torch
    import torch.distributed as dist
    from hydromodels_jl_b200 import parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    hm.engine.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, p, init, N = _grid_problem(R, Cc, T)
    eng = hm.B200Model(model)
    run = par.ShardedRoutedRun(eng, N, T, p, initstates=init)
    inp = sy.forcing("exphydro", run.n0, run.n1, T)                 # local forcing slice, [V, Nl, T] Fortran
    d_in = torch.from_numpy(np.ascontiguousarray(inp.transpose(2, 1, 0))).cuda()
    rec = torch.zeros((T, run.N, run.n_rows), dtype=torch.float64, device="cuda")
    run.launch(d_in.data_ptr(), rec.data_ptr())
    run.launch(d_in.data_ptr(), rec.data_ptr())                     # relaunch must reproduce (buffers reset)
    torch.cuda.synchronize()
    local = rec.cpu().numpy().transpose(2, 0, 1).reshape(run.n_rows * T, run.N)       # node axis last
    full = par.gather_nodes(local, N)
    if rank == 0:
        q.put(full.reshape(run.n_rows, T, N).transpose(0, 2, 1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_sharded_grid_routing_matches_single_gpu():
    import torch.multiprocessing as mp
    R, Cc, T, world = 40, 33, 50, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, Cc, T, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    model, p, init, N = _grid_problem(R, Cc, T)
    single = model(sy.forcing("exphydro", 0, N, T), p, initstates=init)
    assert np.array_equal(got, single)


def _lib_comm(par, rank, world, exchange):
    """`exchange`: "torch" = torch.distributed point-to-point calls from Python; "nccl" / "peer" = the library's own exchange
    (`hb_comm_*` + `hb_halo_exchange_device`: grouped NCCL send/recv, or NVLink peer stores with epoch flags)."""
    return None if exchange == "torch" else par.LibComm(rank, world, halo_mode=exchange)


def _strip_worker(rank, world, port, R, Cc, T, halo, q, exchange="torch"):
    import torch
    import torch.distributed as dist
    from hydromodels_jl_b200 import parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    hm.engine.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, p, init, N = _grid_problem(R, Cc, T)
    eng = hm.B200Model(model)
    comm = _lib_comm(par, rank, world, exchange)
    run = par.ShardedGridRun(eng, R, Cc, T, p, initstates=init, halo=halo, comm=comm)
    n_lo, n_hi = run.plan.node_range()
    inp = sy.forcing("exphydro", n_lo, n_hi, T)                     # forcing of the computed rows (ghost rows included)
    d_in = torch.from_numpy(np.ascontiguousarray(inp.transpose(2, 1, 0))).cuda()
    rec = torch.zeros((T, run.N, run.n_rows), dtype=torch.float64, device="cuda")
    run.launch(d_in.data_ptr(), rec.data_ptr())
    run.launch(d_in.data_ptr(), rec.data_ptr())                     # relaunch must reproduce (states reset)
    run.check()
    q.put((run.plan.r0 * Cc, run.plan.r1 * Cc, rec.cpu().numpy().transpose(2, 1, 0)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["torch", "nccl", "peer"])
@pytest.mark.parametrize("R,Cc,T,halo", [(40, 33, 50, 4), (96, 64, 40, 16)])
def test_sharded_wave_grid_matches_single_gpu(R, Cc, T, halo, exchange):
    """Row strips + ghost rows, states exchanged every `halo` steps (temporal blocking): bit-identical to one GPU.  Uses
    up to 4 GPUs when the box has them (interior ranks exchange with two neighbours)."""
    import torch.multiprocessing as mp
    world = min(_ngpu(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 2000) + halo + 100 * ["torch", "nccl", "peer"].index(exchange)
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, R, Cc, T, halo, q, exchange)) for r in range(world)]
    for pr in procs:
        pr.start()
    model, p, init, N = _grid_problem(R, Cc, T)
    got = np.zeros((13, N, T))
    for _ in range(world):
        a, b, part = q.get(timeout=300)
        got[:, a:b, :] = part
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    single = model(sy.forcing("exphydro", 0, N, T), p, initstates=init)
    assert np.array_equal(got, single)


def test_strip_of_a_grid_on_one_gpu_matches_the_whole_grid():
    """The pieces of the sharded wave run that one GPU can check: a strip with ghost rows, chunked launches from explicit
    states, records of the owned rows only."""
    import torch
    from hydromodels_jl_b200 import parallel as par
    from hydromodels_jl_b200.engine import GridDeviceRun
    R, Cc, T, halo = 50, 40, 23, 5
    model, p, init, N = _grid_problem(R, Cc, T)
    whole = model(sy.forcing("exphydro", 0, N, T), p, initstates=init)
    eng = hm.B200Model(model)
    for rank, world in [(0, 3), (1, 3), (2, 3)]:
        pl = par.GridStripPlan(R, Cc, rank, world, halo)
        n_lo, n_hi = pl.node_range()
        run = GridDeviceRun(eng, pl.n_local, T, p, initstates=init, variant="wave", steps_per_chunk=halo, row_range=(pl.g0, pl.g1))
        d_in = torch.from_numpy(np.ascontiguousarray(sy.forcing("exphydro", n_lo, n_hi, T).transpose(2, 1, 0))).cuda()
        rec = torch.zeros((T, pl.out_count, 13), dtype=torch.float64, device="cuda")
        # one chunk of `halo` steps from the initial states: the owned rows are exact (ghost rows may not be)
        b0 = torch.from_numpy(run.bufs["init"].to_host((2, pl.n_local))).cuda()
        r0 = torch.from_numpy(run.bufs["rinit"].to_host((pl.n_local,))).cuda()
        b1, r1 = torch.zeros_like(b0), torch.zeros_like(r0)
        run.launch_chunk(d_in.data_ptr(), rec.data_ptr(), halo, b0.data_ptr(), r0.data_ptr(), b1.data_ptr(), r1.data_ptr(),
                         pl.out_first, pl.out_count, None)
        run.check()
        got = rec.cpu().numpy().transpose(2, 1, 0)[:, :, :halo]
        ref = whole[:, pl.r0 * Cc:pl.r1 * Cc, :halo]
        if not np.array_equal(got, ref):
            bad = np.argwhere(got != ref)
            raise AssertionError(f"strip {rank}/{world}: {len(bad)} values differ; rows {np.unique(bad[:, 0])}, steps {np.unique(bad[:, 2])}, "
                                 f"cells {np.unique(bad[:, 1])[:16]}, first {bad[0]}: got {got[tuple(bad[0])]!r} ref {ref[tuple(bad[0])]!r}")
        own = slice(pl.out_first, pl.out_first + pl.out_count)
        assert np.array_equal(b1.cpu().numpy()[:, own], whole[:2, pl.r0 * Cc:pl.r1 * Cc, halo - 1])
        assert np.array_equal(r1.cpu().numpy()[own], whole[2, pl.r0 * Cc:pl.r1 * Cc, halo - 1])


def _panel_worker(rank, world, port, R, Cc, T, halo, q, exchange="torch"):
    import torch
    import torch.distributed as dist
    from hydromodels_jl_b200 import parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    hm.engine.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model, p, init, N = _grid_problem(R, Cc, T)
    run = par.ShardedPanelRun(hm.B200Model(model), R, Cc, T, p, initstates=init, halo=halo, comm=_lib_comm(par, rank, world, exchange))
    inp = sy.forcing("exphydro", 0, N, T)[:, run.plan.idx, :]       # forcing of the computed cells, local row-major order
    d_in = torch.from_numpy(np.ascontiguousarray(inp.transpose(2, 1, 0))).cuda()
    rec = torch.zeros((T, run.N, run.n_rows), dtype=torch.float64, device="cuda")
    run.launch(d_in.data_ptr(), rec.data_ptr())
    run.launch(d_in.data_ptr(), rec.data_ptr())
    run.check()
    q.put((run.plan.c0, run.plan.c1, rec.cpu().numpy().transpose(2, 1, 0)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange", ["torch", "nccl", "peer"])
@pytest.mark.parametrize("R,Cc,T,halo", [(30, 41, 25, 4), (64, 256, 70, 32)])
def test_sharded_column_panels_match_single_gpu(R, Cc, T, halo, exchange):
    """Column panels + ghost columns (sub-model restricted to the panel's cells, global parameter vectors gathered through
    htypes), states exchanged every `halo` steps: bit-identical to one GPU."""
    import torch.multiprocessing as mp
    world = min(_ngpu(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 36500 + (os.getpid() % 2000) + halo + 100 * ["torch", "nccl", "peer"].index(exchange)
    procs = [ctx.Process(target=_panel_worker, args=(r, world, port, R, Cc, T, halo, q, exchange)) for r in range(world)]
    for pr in procs:
        pr.start()
    model, p, init, N = _grid_problem(R, Cc, T)
    got = np.zeros((13, R, Cc, T))
    for _ in range(world):
        a, b, part = q.get(timeout=300)
        got[:, :, a:b, :] = part.reshape(13, R, b - a, T)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    single = model(sy.forcing("exphydro", 0, N, T), p, initstates=init)
    assert np.array_equal(got.reshape(13, N, T), single)


def test_panel_of_a_grid_on_one_gpu_matches_the_whole_grid():
    """One GPU's view of the column-panel decomposition: the sub-model of a panel with ghost columns, one chunk of `halo`
    steps from the initial states, records of the owned window only (TMA-aligned 32-column ghosts and odd ones)."""
    import torch
    from hydromodels_jl_b200 import parallel as par
    from hydromodels_jl_b200.engine import GridDeviceRun
    for R, Cc, T, halo in [(40, 192, 33, 32), (25, 70, 12, 5)]:
        model, p, init, N = _grid_problem(R, Cc, T)
        whole = model(sy.forcing("exphydro", 0, N, T), p, initstates=init).reshape(13, R, Cc, T)
        for rank, world in [(0, 3), (1, 3), (2, 3)]:
            pl = par.GridPanelPlan(R, Cc, rank, world, halo)
            sub = par.submodel_for_cells(model, pl.idx, model.components[-1].aggr.flwdir[:, pl.g0:pl.g1])
            eng = hm.B200Model(sub)
            li = {k: v[pl.idx] for k, v in init.items()}
            run = GridDeviceRun(eng, pl.n_local, T, p, initstates=li, variant="wave", steps_per_chunk=16)
            d_in = torch.from_numpy(np.ascontiguousarray(sy.forcing("exphydro", 0, N, T)[:, pl.idx, :].transpose(2, 1, 0))).cuda()
            rec = torch.zeros((T, pl.out_count, 13), dtype=torch.float64, device="cuda")
            b0 = torch.from_numpy(run.bufs["init"].to_host((2, pl.n_local))).cuda()
            r0 = torch.from_numpy(run.bufs["rinit"].to_host((pl.n_local,))).cuda()
            b1, r1 = torch.zeros_like(b0), torch.zeros_like(r0)
            ns = min(halo, T)
            run.launch_chunk(d_in.data_ptr(), rec.data_ptr(), ns, b0.data_ptr(), r0.data_ptr(), b1.data_ptr(), r1.data_ptr(),
                             out_window=pl.out_window)
            run.check()
            got = rec.cpu().numpy().transpose(2, 1, 0).reshape(13, R, pl.c1 - pl.c0, T)[:, :, :, :ns]
            assert np.array_equal(got, whole[:, :, pl.c0:pl.c1, :ns])


@pytest.mark.parametrize("R,Cc,T,K,halo", [(40, 256, 70, 3, 32), (25, 70, 20, 2, 5), (64, 192, 45, 1, 32)])
def test_panelled_run_on_one_gpu_matches_the_single_kernel(R, Cc, T, K, halo):
    """A wide grid run as K column panels one after the other on ONE GPU, in place on the whole grid's forcing and record arrays
    (pitched addressing), ghost-column states copied between panels every `halo` steps: every value equals the one-kernel run."""
    import torch
    from hydromodels_jl_b200 import parallel as par
    model, p, init, N = _grid_problem(R, Cc, T)
    inp = sy.forcing("exphydro", 0, N, T)
    whole = model(inp, p, initstates=init)
    run = par.PanelledGridRun(hm.B200Model(model), R, Cc, T, p, initstates=init, panels=K, halo=halo)
    assert run.K == K
    d_in = torch.from_numpy(np.ascontiguousarray(inp.transpose(2, 1, 0))).cuda()
    rec = torch.zeros((T, N, 13), dtype=torch.float64, device="cuda")
    for rep in range(2):
        rec.zero_()
        run.launch(d_in.data_ptr(), rec.data_ptr())
        run.check()
        got = rec.cpu().numpy().transpose(2, 1, 0)
        if not np.array_equal(got, whole):
            bad = np.argwhere(got != whole)
            cells = np.unique(bad[:, 1])
            raise AssertionError(f"rep {rep}: {len(bad)} values differ; rows {np.unique(bad[:, 0])}, steps {np.unique(bad[:, 2])[:10]}, grid columns "
                                 f"{np.unique(cells % Cc)[:20]}, grid rows {np.unique(cells // Cc)[:10]}, first {bad[0]}: {got[tuple(bad[0])]!r} vs {whole[tuple(bad[0])]!r}")


def _de_problem():
    from hydromodels_jl_b200 import calibration as cal
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import hydro_oracle as ho
    T = 500
    model = hm.models.exphydro()
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    truth = np.array([-2.1, 0.2, 2.7, 1500.0, 18.0, 0.017])
    lb, ub = np.array([-3.0, 0.0, 0.5, 500.0, 5.0, 0.005]), np.array([0.0, 3.0, 5.0, 2500.0, 40.0, 0.05])
    init = dict(snowpack=0.0, soilwater=1303.0)
    target = ho.run_model(model, inp, {"params": dict(zip(model.get_param_names(), truth))}, initstates=init)[-1]
    return cal, cal.OptimizationProblem(model, inp, target, metric="NSE", warm_up=30, lb_pas=lb, ub_pas=ub, default_initstates=init)


def _de_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from hydromodels_jl_b200 import parallel as par
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    hm.engine.init(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    comm = par.LibComm(rank, world)
    cal, prob = _de_problem()
    sol = cal.solve(prob, cal.DifferentialEvolution(popsize=512, seed=9), maxiters=25, comm=comm)
    q.put((rank, sol.history, sol.population, sol.fitness, sol.extras))
    dist.barrier()
    comm.free()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_sharded_differential_evolution_follows_the_single_gpu_trajectory():
    """Population sharded over the ranks (`hb_de_trial_sharded_device` + one all-gather per generation): the history, the final
    population and the objectives equal the single-GPU search with the same seed bit for bit, on every rank."""
    import torch.multiprocessing as mp
    world = min(_ngpu(), 4)
    while 512 % world:
        world -= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_de_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=600) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    cal, prob = _de_problem()
    single = cal.solve(prob, cal.DifferentialEvolution(popsize=512, seed=9), maxiters=25)
    for rank, hist, pop, fit, extras in got:
        assert np.array_equal(hist, single.history), rank
        assert np.array_equal(pop, single.population) and np.array_equal(fit, single.fitness)
        assert extras["world"] == world and extras["all_gather_bytes_per_generation"] == 512 * 6 * 8
