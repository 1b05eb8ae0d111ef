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
