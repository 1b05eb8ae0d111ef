"""CPU-only (gloo, world_size 2): the multi-GPU host logic — shard ranges cover the node axis,
per-shard runs reproduce the full run (checked with the oracle standing in for the kernel, since
this box has no GPU), and the gather returns node order."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import parallel as par
from hydromodels_jl_b200 import synthetic as sy

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


def test_shard_ranges_partition_the_node_axis():
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (4096 * 4096, 8), (5, 1)]:
        r = [par.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, N, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hydro_oracle as ho
    n0, n1 = par.shard_range(N, rank, world)
    ht = np.arange(1, N + 1)
    p = {"params": sy.parameters("exphydro", 0, N)}
    init = np.concatenate([np.zeros(N), np.full(N, 1303.0)])
    model = hm.models.exphydro(htypes=par.shard_htypes(ht, n0, n1))
    inp = sy.forcing("exphydro", n0, n1, T)                   # shards regenerate their own forcing
    out = ho.run_model(model, inp, par.shard_params(p, n0, n1, N), initstates=par.shard_initstates(init, n0, n1, N, 2))
    obj = np.array([ho.objective("nse", out[-1, i] * 1.1 + 0.01, out[-1, i]) for i in range(n1 - n0)])
    full_obj = par.gather_nodes(obj, N)
    full_final = par.gather_nodes(out[:2, :, -1], N)
    if rank == 0:
        q.put((full_obj, full_final))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_rank():
    import hydro_oracle as ho
    N, T, world = 37, 40, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, T, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    full_obj, full_final = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    model = hm.models.exphydro(htypes=np.arange(1, N + 1))
    ref = ho.run_model(model, sy.forcing("exphydro", 0, N, T), {"params": sy.parameters("exphydro", 0, N)},
                       initstates=np.concatenate([np.zeros(N), np.full(N, 1303.0)]))
    assert np.array_equal(full_final, ref[:2, :, -1])
    ref_obj = np.array([ho.objective("nse", ref[-1, i] * 1.1 + 0.01, ref[-1, i]) for i in range(N)])
    assert np.array_equal(full_obj, ref_obj)


# ---- sharded routing: partition + ghost exchange (gloo), numpy Jacobi step standing in for the kernel ----
def _route_worker(rank, world, port, R, Cc, T, q_in, lag, flw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hydromodels_jl_b200.engine import upstream_csr
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc)]
    N = len(pos)
    aggr = hm.build_aggr_func(flw, pos)
    part = par.RoutePartition(*upstream_csr(aggr, N), N, world)
    n0, n1 = part.ranges[rank]
    Nl = n1 - n0
    ptr, idx, ng = part.local_csr(rank)
    send = [(p, torch.from_numpy(ix)) for p, ix in part.send_plan(rank)]
    recv = part.recv_plan(rank)
    s = np.zeros(Nl)
    qb = [torch.zeros(Nl + ng, dtype=torch.float64) for _ in range(2)]
    lg = lag[n0:n1]
    rflux = lambda x, st: st / (1 + lg) + x
    out = np.zeros((2, Nl, T))
    qb[0][:Nl] = torch.from_numpy(rflux(q_in[n0:n1, 0], s))
    par.exchange_ghosts(qb[0], Nl, send, recv)
    for t in range(T):
        cur = t & 1
        qc = qb[cur].numpy()
        inflow = np.array([qc[idx[ptr[n]:ptr[n + 1]]].sum() for n in range(Nl)])
        x = q_in[n0:n1, t]
        s = np.maximum(1e-6, s + ((x - qc[:Nl]) + inflow))
        out[0, :, t], out[1, :, t] = s, rflux(x, s)
        if t + 1 < T:
            qb[cur ^ 1][:Nl] = torch.from_numpy(rflux(q_in[n0:n1, t + 1], s))
            par.exchange_ghosts(qb[cur ^ 1], Nl, send, recv)
    full = par.gather_nodes(out.reshape(2 * T, Nl) if False else np.ascontiguousarray(out.transpose(0, 2, 1)).reshape(2 * T, Nl), N)
    if rank == 0:
        q.put(full.reshape(2, T, N).transpose(0, 2, 1))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_route_partition_and_ghost_exchange(world):
    import hydro_oracle as ho
    rng = np.random.default_rng(11)
    R, Cc, T = 9, 7, 12
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    N = R * Cc
    q_in = np.abs(rng.normal(size=(N, T)))
    lag = rng.uniform(0.1, 2.0, N)
    ctx = mp.get_context("spawn")
    qu = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_route_worker, args=(r, world, port, R, Cc, T, q_in, lag, flw, qu)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = qu.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    s = hm.Symbols("q q_routed s_river", "lag")
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc)]
    route = hm.HydroRoute(rfluxes=[hm.HydroFlux("q_routed ~ s_river / (1 + lag) + q", symbols=s)],
                          dfluxes=[hm.StateFlux("s_river ~ q - q_routed", symbols=s)],
                          aggr_func=hm.build_aggr_func(flw, pos), htypes=np.arange(1, N + 1))
    ref = ho.run_route(route, q_in[None, :, :], {"params": dict(lag=lag)})
    assert np.allclose(got, ref, rtol=1e-13, atol=1e-14)


def test_route_partition_plans_are_consistent():
    from hydromodels_jl_b200.engine import upstream_csr
    rng = np.random.default_rng(5)
    R, Cc = 12, 10
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128], size=(R, Cc))
    pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc)]
    N = len(pos)
    part = par.RoutePartition(*upstream_csr(hm.build_aggr_func(flw, pos), N), N, 4)
    for r in range(4):
        for peer, off, cnt in part.recv_plan(r):
            sends = dict(part.send_plan(peer))
            assert r in sends and len(sends[r]) == cnt          # what r expects is what peer sends
        ptr, idx, ng = part.local_csr(r)
        n0, n1 = part.ranges[r]
        assert ptr[-1] == part.up_ptr[n1] - part.up_ptr[n0] and idx.max(initial=0) < (n1 - n0) + ng


# ---- temporal blocking for the wave grid kernel: row strips + ghost rows, states exchanged every `halo` steps ----
def _strip_worker(rank, world, port, R, Cc, T, halo, flw, qu):
    """The oracle stands in for `GridDeviceRun.launch_chunk`: each rank runs ExpHydro + D8 route on its strip
    plus ghost rows for `halo` steps from explicit states, then refreshes the ghost rows' states."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hydro_oracle as ho
    N = R * Cc
    pl = par.GridStripPlan(R, Cc, rank, world, halo)
    n_lo, n_hi = pl.node_range()
    rows_l = pl.g1 - pl.g0
    rr, cc = np.divmod(np.arange(pl.n_local), Cc)
    model = hm.models.exphydro_grid(flw[pl.g0:pl.g1], np.stack([rr + 1, cc + 1], axis=1))
    rng = np.random.default_rng(3)
    pg = dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))
    p = {"params": {k: v[n_lo:n_hi] for k, v in pg.items()}}
    inp = sy.forcing("exphydro", 0, N, T)[:, n_lo:n_hi, :]
    states = [torch.zeros(pl.n_local, dtype=torch.float64), torch.full((pl.n_local,), 1303.0, dtype=torch.float64),
              torch.from_numpy(rng.uniform(0, 2, N)[n_lo:n_hi].copy())]           # snowpack, soilwater, s_river
    out = np.zeros((13, pl.out_count, T))
    for t0 in range(0, T, halo):
        ns = min(halo, T - t0)
        init = dict(snowpack=states[0].numpy(), soilwater=states[1].numpy(), s_river=states[2].numpy())
        full = ho.run_model(model, inp[:, :, t0:t0 + ns], p, initstates=init)
        out[:, :, t0:t0 + ns] = full[:, pl.out_first:pl.out_first + pl.out_count, :]
        states = [torch.from_numpy(np.ascontiguousarray(full[i, :, -1])) for i in range(3)]
        if t0 + ns < T:
            par.exchange_strip_states(states, pl)
    qu.put((pl.r0 * Cc, pl.r1 * Cc, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, 3), (3, 2)])
def test_strip_plan_with_ghost_rows_reproduces_the_whole_grid(world, halo):
    import hydro_oracle as ho
    rng = np.random.default_rng(17)
    R, Cc, T = 11, 6, 10
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    N = R * Cc
    ctx = mp.get_context("spawn")
    qu = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_strip_worker, args=(r, world, port, R, Cc, T, halo, flw, qu)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = np.zeros((13, N, T))
    for _ in range(world):
        a, b, part = qu.get(timeout=180)
        got[:, a:b, :] = part
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    rr, cc = np.divmod(np.arange(N), Cc)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    rng = np.random.default_rng(3)
    pg = dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
    ref = ho.run_model(model, sy.forcing("exphydro", 0, N, T), {"params": pg}, initstates=init)
    assert np.array_equal(got, ref)          # same arithmetic on every owned cell: bit-identical


def test_strip_plan_geometry():
    for rows, cols, world, halo in [(4096, 4096, 8, 16), (11, 6, 3, 2), (64, 64, 1, 16)]:
        plans = [par.GridStripPlan(rows, cols, k, world, halo) for k in range(world)]
        assert plans[0].r0 == 0 and plans[-1].r1 == rows
        for a, b in zip(plans[:-1], plans[1:]):
            assert a.r1 == b.r0
            (pa, sa, qa), = [e for e in a.exchanges() if e[0] == b.rank]
            (pb, sb, qb), = [e for e in b.exchanges() if e[0] == a.rank]
            assert sa[1] - sa[0] == qb[1] - qb[0] == halo * cols and sb[1] - sb[0] == qa[1] - qa[0] == halo * cols
            # what a sends are its last owned rows = b's upper ghost rows, in global numbering
            assert a.g0 * cols + sa[0] == b.g0 * cols + qb[0] and b.g0 * cols + sb[0] == a.g0 * cols + qa[0]
        for pl in plans:
            assert pl.out_first + pl.out_count <= pl.n_local and pl.out_count == (pl.r1 - pl.r0) * cols


# ---- column panels: ghost COLUMNS, sub-model restricted to the panel's cells -----------------------------------------
def _panel_worker(rank, world, port, R, Cc, T, halo, flw, qu):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hydro_oracle as ho
    N = R * Cc
    pl = par.GridPanelPlan(R, Cc, rank, world, halo)
    rr, cc = np.divmod(np.arange(N), Cc)
    whole = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1), htypes=np.arange(1, N + 1))
    sub = par.submodel_for_cells(whole, pl.idx, flw[:, pl.g0:pl.g1])
    rng = np.random.default_rng(3)
    pg = dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))
    inp = sy.forcing("exphydro", 0, N, T)[:, pl.idx, :]
    s_river0 = rng.uniform(0, 2, N)
    states = [torch.zeros(pl.n_local, dtype=torch.float64), torch.full((pl.n_local,), 1303.0, dtype=torch.float64),
              torch.from_numpy(s_river0[pl.idx].copy())]
    r0, nr, c0, nc = pl.out_window
    own = (np.arange(r0, r0 + nr)[:, None] * pl.local_cols + np.arange(c0, c0 + nc)[None, :]).reshape(-1)
    out = np.zeros((13, pl.out_count, T))
    for t0 in range(0, T, halo):
        ns = min(halo, T - t0)
        init = dict(snowpack=states[0].numpy(), soilwater=states[1].numpy(), s_river=states[2].numpy())
        full = ho.run_model(sub, inp[:, :, t0:t0 + ns], {"params": pg}, initstates=init)       # GLOBAL parameter vectors
        out[:, :, t0:t0 + ns] = full[:, own, :]
        states = [torch.from_numpy(np.ascontiguousarray(full[i, :, -1])) for i in range(3)]
        if t0 + ns < T:
            par.exchange_panel_states(states, pl)
    qu.put((pl.c0, pl.c1, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(2, 3), (3, 2)])
def test_panel_plan_with_ghost_columns_reproduces_the_whole_grid(world, halo):
    import hydro_oracle as ho
    rng = np.random.default_rng(23)
    R, Cc, T = 6, 13, 9
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0], size=(R, Cc))
    N = R * Cc
    ctx = mp.get_context("spawn")
    qu = ctx.Queue()
    port = 35500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_panel_worker, args=(r, world, port, R, Cc, T, halo, flw, qu)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = np.zeros((13, R, Cc, T))
    for _ in range(world):
        a, b, part = qu.get(timeout=180)
        got[:, :, a:b, :] = part.reshape(13, R, b - a, T)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    rr, cc = np.divmod(np.arange(N), Cc)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    rng = np.random.default_rng(3)
    pg = dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
    ref = ho.run_model(model, sy.forcing("exphydro", 0, N, T), {"params": pg}, initstates=init)
    assert np.array_equal(got.reshape(13, N, T), ref)


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_library_halo_edges_pair_up_between_neighbouring_ranks(world):
    """`hb_halo_edge` lists of the strips / panels (what `LibComm.halo_exchange` hands to `hb_halo_exchange_device`): every
    edge's `peer_slot` must name the matching edge of the peer, with the same block shape, and the block I send must be,
    in GLOBAL cell coordinates, exactly the ghost block the peer receives."""
    from hydromodels_jl_b200.parallel import GridPanelPlan, GridStripPlan, panel_halo_edges, strip_halo_edges
    rows, cols, halo = 64 * world + 3, 40 * world + 7, 16
    panels = [GridPanelPlan(rows, cols, r, world, halo) for r in range(world)]
    strips = [GridStripPlan(rows, cols, r, world, halo) for r in range(world)]
    for plans, edges_of, origin in ((panels, panel_halo_edges, lambda pl: (0, pl.g0)), (strips, strip_halo_edges, lambda pl: (pl.g0, 0))):
        all_edges = [edges_of(pl) for pl in plans]
        for r, edges in enumerate(all_edges):
            assert len(edges) == (r > 0) + (r < world - 1)
            for k, (peer, sr, sc, rr, rc, nr, nc, slot) in enumerate(edges):
                back = all_edges[peer][slot]
                assert back[0] == r and back[7] == k and (back[5], back[6]) == (nr, nc)
                o_me, o_peer = origin(plans[r]), origin(plans[peer])
                assert (sr + o_me[0], sc + o_me[1]) == (back[3] + o_peer[0], back[4] + o_peer[1])
