#!/usr/bin/env python
"""bench.py — basin-timesteps/s of the fused forward-run stepper (BASELINE.json's metric).

A "step" is ONE pass of the hot path over the whole synthetic batch: the workload's model run
for all N basins × T time steps (one kernel launch).  Default workload at one GPU: ExpHydro,
1 M basins × 365 d, FP64, full 10-row output — the configuration the north star quotes its
target on (≥0.6 of HBM roofline on 1M-basin ExpHydro).  With N GPUs every rank owns its own
1 M basins (weak scaling; basins are independent, no data-path collective — SURVEY.md §8e).

`value`  : device-resident throughput (forcing and results in HBM; CUDA events; max over ranks).
`e2e`    : same workload through the public functor `model(input, params, config; initstates)`
           with pinned HOST buffers (H2D of the forcing + D2H of all rows inside the timed region).
`roofline`: algorithmic bytes (8·(V + rows) per basin-timestep) ÷ kernel time ÷
           measured HBM copy bandwidth (MEASURED_PEAKS.json).
`cpu_baseline` / `--impl reference`: the CPU oracle (a port: the Julia reference cannot run
           here) timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, basins per GPU, steps, description)
    "exphydro_1M_basins_x_365d": ("exphydro", 1_000_000, 365),
    "gr4j_uh_100k_basins_x_3650d": ("gr4j", 100_000, 3650),
    "hbv_100k_basins_x_730d": ("hbv", 100_000, 730),
    "m50_50k_basins_x_3650d": ("m50", 50_000, 3650),
    # the opt-in Float32 mode of the same workload: --precision f32 [--nn-tensor 0|1] (per-thread FFMA | tcgen05 3xTF32)
    "exphydro_64k_basins_x_365d": ("exphydro", 65_536, 365),
    # BASELINE config 3: calibration ensemble, shared forcing, fused objective (no rows written)
    "hbv_ensemble_1M_members_x_7305d": ("hbv_ensemble", 1_000_000, 7305),
    # BASELINE config 5 (single-GPU slice): distributed ExpHydro + D8 routing on a 2048 x 2048 grid
    "exphydro_grid_2048x2048_x_60d": ("exphydro_grid", 2048 * 2048, 60),
    "exphydro_grid_512x512_x_60d": ("exphydro_grid", 512 * 512, 60),
    # BASELINE config 5 proper: ONE 4096 x 4096 grid, sharded over the ranks (strong scaling; 60 steps fit a single GPU too)
    "exphydro_grid_4096x4096_total_x_60d": ("exphydro_grid", 4096 * 4096, 60),
}
DEFAULT_WORKLOAD = "exphydro_1M_basins_x_365d"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (recipe's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.15] or [r for _, r in self.rows]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(rows)}


def build_problem(hm, sy, model_name, n0, N, T):
    ht = np.arange(1, N + 1)
    model = getattr(hm.models, model_name)(htypes=ht)
    p = {"params": sy.parameters(model_name, n0, n0 + N)}
    if model_name == "m50":
        et, q = hm.models.m50_chains()
        p["nns"] = {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}
    init = {k: np.full(N, v) for k, v in sy.INIT[model_name].items()}
    return model, p, init


def device_forcing(torch, model_name, N, T, seed):
    """Synthetic forcing generated ON the device, layout [T][N][V] (= Julia input[V,N,T])."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    t = torch.arange(1, T + 1, device="cuda", dtype=torch.float64).view(T, 1)
    temp = 5.0 + 15.0 * torch.sin(2 * np.pi * (t - 100.0) / 365.25) + 5.0 * torch.randn(T, N, generator=g, device="cuda", dtype=torch.float64)
    lday = (0.5 + 0.15 * torch.sin(2 * np.pi * (t - 80.0) / 365.25)).expand(T, N)
    u = torch.rand(T, N, generator=g, device="cuda", dtype=torch.float64)
    prcp = torch.where(u < 0.35, -6.0 * torch.log1p(-u / 0.35 * (1 - 1e-12)), torch.zeros_like(u))
    rows = {"exphydro": (temp, lday, prcp), "m50": (prcp, temp, lday)}
    if model_name in ("gr4j", "hbv"):
        pet = 29.8 * lday * 24 * 0.611 * torch.exp(17.3 * temp / (temp + 237.3)) / (temp + 273.2)
        rows.update(gr4j=(prcp, pet), hbv=(prcp, temp, pet))
    out = torch.stack(rows[model_name], dim=2).contiguous()
    del temp, u, prcp
    return out


class PinnedArray:
    """A pinned host array owned through the C-ABI (`hb_malloc_host`), viewed as a numpy array."""

    def __init__(self, shape, dtype=np.float64):
        import ctypes as C
        from hydromodels_jl_b200.engine import check, lib
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib().hb_malloc_host(C.byref(p), self.nbytes))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        from hydromodels_jl_b200.engine import lib
        if self.ptr:
            self.array = None
            lib().hb_free_host(self.ptr)
            self.ptr = None


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


def pcie_probe(torch, dist, world, barrier, nbytes=1 << 30, reps=6):
    """`hb_probe_pcie_gbs` on every rank AT THE SAME TIME (barrier first): per-rank and summed GB/s of plain pinned
    cudaMemcpyAsync, H2D / D2H alone and both directions at once — the ceiling of the host-buffer path on this box."""
    import ctypes as C
    from hydromodels_jl_b200.engine import check, lib
    g = (C.c_double * 4)()
    barrier()
    check(lib().hb_probe_pcie_gbs(nbytes, reps, g))
    mine = torch.tensor(list(g), device="cuda", dtype=torch.float64)
    if dist is not None:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per = torch.stack(allr).cpu().numpy()
    else:
        per = mine.cpu().numpy()[None, :]
    keys = ["h2d_alone", "d2h_alone", "h2d_while_d2h", "d2h_while_h2d"]
    return {"bytes_per_copy": nbytes, "copies": reps, "concurrent_ranks": world,
            "sum_GBps": {k: float(per[:, i].sum()) for i, k in enumerate(keys)},
            "min_rank_GBps": {k: float(per[:, i].min()) for i, k in enumerate(keys)}}


def e2e_leg(torch, dist, hm, eng, p, init, d_in, T, N, v_of, rows_of, world, rank, local_rank, args, unit, barrier):
    in_bytes, out_bytes = T * N * v_of * 8, T * N * rows_of * 8
    avail = mem_available_bytes()
    need = world * (in_bytes + out_bytes)
    # All T steps at every world size.  When the ranks' host arrays do not fit the box's memory together, the pass runs as
    # `n_win` consecutive calls on windows of the time axis that reuse ONE pinned window per rank (states carried from call
    # to call, as a consumer that processes a window and moves on would do) — same bytes over PCIe, same kernels.
    n_win = max(1, int(os.environ.get("HB_E2E_WINDOWS", "1")))      # (override for testing the windowed path)
    if avail is not None:
        while need / n_win > 0.6 * avail and n_win < T:
            n_win += 1
    Tw = -(-T // n_win)
    h_in, h_out = PinnedArray((Tw, N, v_of)), PinnedArray((Tw, N, rows_of))
    torch.from_numpy(h_in.array).copy_(d_in[:Tw])
    torch.cuda.synchronize()
    np_in = h_in.array.transpose(2, 1, 0)      # [V, N, T] Fortran-ordered view, zero copy
    np_out = h_out.array.transpose(2, 1, 0)
    assert np_in.flags.f_contiguous and np_out.flags.f_contiguous
    small = int(sum(np.asarray(v).nbytes for v in p["params"].values()) + 8 * N * len(init))

    def one_pass(x, y):
        if n_win == 1:
            eng(x, p, initstates=init, out=y)
            return
        st, done = init, 0
        while done < T:
            tw = min(Tw, T - done)
            xs = x if tw == Tw else np.asfortranarray(x[:, :, :tw])
            ys = y if tw == Tw else None
            _, fin = eng(xs, p, initstates=st, out=ys, return_final_states=True)
            st, done = fin.reshape(-1), done + tw

    def timed(x, y, steps):
        one_pass(x, y)                          # warm-up (library workspace; first touch of pageable pages)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_pass(x, y)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / steps

    s_pinned = timed(np_in, np_out, args.e2e_steps)
    probe = pcie_probe(torch, dist, world, barrier)
    # the pipeline moves H2D and D2H at the same time: its floor is the slower of the two streams at the box's concurrent rates
    floor = max(world * in_bytes / (probe["sum_GBps"]["h2d_while_d2h"] * 1e9), world * out_bytes / (probe["sum_GBps"]["d2h_while_h2d"] * 1e9))
    floor_min = max(in_bytes / (probe["min_rank_GBps"]["h2d_while_d2h"] * 1e9), out_bytes / (probe["min_rank_GBps"]["d2h_while_h2d"] * 1e9))
    e2e = {"value": world * N * T / s_pinned, "unit": unit, "timesteps": T, "h2d_bytes_per_step": in_bytes + small,
           "d2h_bytes_per_step": out_bytes, "steps": args.e2e_steps, "s_per_step": s_pinned,
           "api": "B200Model.__call__(input, params, config; initstates) with pinned host arrays (hb_malloc_host)",
           "host_windows": n_win, "host_memory": {"MemAvailable_GiB": None if avail is None else avail / 2**30, "needed_GiB": need / 2**30},
           "aggregate_GBps": {"h2d": world * in_bytes / s_pinned / 1e9, "d2h": world * out_bytes / s_pinned / 1e9},
           "pcie_probe": probe, "pcie_floor_s_per_step": floor, "pcie_frac": floor / s_pinned,
           # the timed region ends when the SLOWEST rank is done, and the box does not share its PCIe bandwidth evenly between
           # the ranks: the same floor from the slowest rank's measured rates
           "pcie_floor_slowest_rank_s_per_step": floor_min, "pcie_frac_slowest_rank": floor_min / s_pinned,
           "_launches": (args.e2e_steps + 1) * min(T, -(-out_bytes // (256 << 20)))}
    if not args.no_pageable:
        # same call, ordinary (pageable) numpy arrays: the library stages them through its pinned ring with host threads
        try:
            if avail is not None and 2 * need / n_win > 0.9 * avail:
                raise MemoryError("not enough host memory for a second, pageable copy of the arrays")
            pg_in = np.empty((Tw, N, v_of)); pg_in[...] = h_in.array
            pg_out = np.empty((Tw, N, rows_of))
            s_page = timed(pg_in.transpose(2, 1, 0), pg_out.transpose(2, 1, 0), 1)
            import ctypes as C
            nt = C.c_int(0)
            hm.engine.lib().hb_host_threads(C.byref(nt))
            same = bool(np.array_equal(pg_out[::37], h_out.array[::37]))
            e2e["pageable"] = {"value": world * N * T / s_page, "s_per_step": s_page, "host_threads_per_rank": nt.value,
                               "identical_to_pinned_run": same,
                               "api": "same call with ordinary numpy arrays (pageable, as a Julia Array): staged through the library's pinned ring"}
            e2e["_launches"] += 2 * min(T, -(-out_bytes // (64 << 20)))
            del pg_in, pg_out
        except Exception as ex:
            e2e["pageable"] = {"value": None, "error": f"{type(ex).__name__}: {ex}"[:200]}
    h_in.free(); h_out.free()
    return e2e


def bench_grid_total(torch, dist, hm, sy, args, rank, world, steps, barrier, halo_exchange):
    """BASELINE config 5 proper as a SECONDARY line: ONE 4096 x 4096 D8 grid (ExpHydro buckets + HydroRoute), 60 daily
    steps, cut into `world` column panels with 32 ghost columns; bucket and river states of the ghost columns are
    refreshed every 32 steps through the library's halo exchange (`hb_halo_exchange_device`).  Strong scaling: the
    grid is fixed, `value` = all cells x steps / max-over-ranks device time."""
    from hydromodels_jl_b200.parallel import LibComm, PanelledGridRun, ShardedPanelRun
    side, T, V, R = 4096, 60, 3, 13
    Ng = side * side
    rng = np.random.default_rng(sy.SEED)
    flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128]), size=(side, side), p=[.3, .2, .3, .05, .03, .02, .05, .05])
    rr, cc = np.divmod(np.arange(Ng), side)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    pp = sy.parameters("exphydro", 0, Ng)
    pp.update(area_coef=np.full(Ng, 1.0), lag=rng.uniform(0.1, 2.0, Ng))
    init = dict(snowpack=np.zeros(Ng), soilwater=np.full(Ng, 1303.0), s_river=np.zeros(Ng))
    eng = hm.B200Model(model)
    out = {"workload": "exphydro_grid_4096x4096_total_x_60d", "metric": "cell-timesteps/sec", "unit": "cell-timesteps/s",
           "scaling": "strong", "grid": [side, side], "timesteps": T, "n_gpus": world, "steps": steps}
    comm = None
    if world > 1:
        mode = halo_exchange
        if mode != "torch":
            try:
                comm = LibComm(rank, world, halo_mode=mode)
                if mode == "peer":
                    comm.enable_peer(side * 32 * 3 * 8)
            except Exception as ex:      # no CUDA IPC between the ranks' devices: the NCCL exchange of the library still works
                out["halo_exchange_fallback"] = f"{type(ex).__name__}: {ex}"[:200]
                if comm is not None and mode == "peer":
                    comm.halo_mode = mode = "nccl"
                else:
                    comm, mode = None, "torch"
        run = ShardedPanelRun(eng, side, side, T, {"params": pp}, initstates=init, halo=32, comm=comm)
        n_local, n_owned = run.plan.n_local, run.N
        out.update(grid_split=f"{world} column panels, 32 ghost columns, states exchanged every 32 steps",
                   halo_exchange={"torch": "torch.distributed point-to-point from Python", "nccl": "library: pack kernel + grouped ncclSend/ncclRecv + unpack kernel",
                                  "peer": "library: pack kernel stores into the neighbours' staging over NVLink (CUDA IPC) + epoch flag, unpack kernel waits"}[mode],
                   halo_bytes_sent_per_exchange_per_rank=run.halo_bytes_per_exchange(), exchanges_per_pass=(T - 1) // 32)
    else:
        run = PanelledGridRun(eng, side, side, T, {"params": pp}, initstates=init, panels=0)
        n_local = n_owned = Ng
        out.update(grid_split=f"one GPU: {run.K} column panels run one after the other in place")
    del flw, rr, cc, pp, init
    d_in = device_forcing(torch, "exphydro", n_local, T, sy.SEED + 17 * rank)
    d_out = torch.empty((T, n_owned, R), device="cuda", dtype=torch.float64)
    stream = torch.cuda.current_stream()
    for _ in range(2):
        run.launch(d_in.data_ptr(), d_out.data_ptr())
    barrier()
    ex_ms = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        run.launch(d_in.data_ptr(), d_out.data_ptr())
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    tc = run.check()
    if world > 1:
        # exchange share: one more pass with events around every exchange (kept out of the timed passes)
        run.exchange_events = []
        run.launch(d_in.data_ptr(), d_out.data_ptr())
        torch.cuda.synchronize()
        ex_ms = [a.elapsed_time(b) for a, b in run.exchange_events]
        run.exchange_events = None
    t = torch.tensor([ms, float(sum(ex_ms))], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, ex_max = float(t[0].item()), float(t[1].item())
    peak, _ = peaks()
    out.update(value=Ng * T / (ms_max * 1e-3), ms_per_step=ms_max, steps_per_launch=tc,
               roofline={"bound": "hbm", "kernel": "hb_grid_wave", "algorithmic_bytes_per_cell_timestep": 128,
                         "achieved_per_gpu": 128 * (Ng / world) * T / (ms_max * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac_per_gpu": 128 * (Ng / world) * T / (ms_max * 1e-3) / 1e9 / peak,
                         "ghost_work": n_local / n_owned - 1.0},
               exchange_ms_per_pass=ex_max, exchange_share=(ex_max / ms_max if ms_max > 0 else None),
               gpu_launches=(steps + 2) * (-(-T // 32)) * (3 if world > 1 else 1))
    if comm is not None:
        comm.free()
    del d_in, d_out
    torch.cuda.empty_cache()
    return out


def cpu_reference_rate(model_name, T, target_seconds=12.0):
    """Time the CPU oracle on a bounded sample of the workload: the C restatement of the
    reference's two-pass / per-component algorithm (oracle/hydro_oracle.c), OpenMP over node
    blocks on all host cores.  Returns (rate, seconds, sample description, cores, 1-thread rate)."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hydro_oracle_c as hoc
    cores = hoc.max_threads()
    Ts = min(T, 365)
    n = 4096
    model, p, init = build_problem(hm, sy, model_name, 0, n, Ts)
    inp = sy.forcing(model_name, 0, n, Ts)
    hoc.run_model(model, inp, p, initstates=init)                      # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    hoc.run_model(model, inp, p, initstates=init, n_threads=1)
    rate1 = n * Ts / (time.perf_counter() - t0)
    # size the multithreaded sample for ~target_seconds / 2 of CPU wall time
    t0 = time.perf_counter()
    hoc.run_model(model, inp, p, initstates=init)
    est = n * Ts / (time.perf_counter() - t0)
    n = int(min(262144, max(4096, est * target_seconds / 2 / Ts)))
    model, p, init = build_problem(hm, sy, model_name, 0, n, Ts)
    inp = sy.forcing(model_name, 0, n, Ts)
    hoc.run_model(model, inp, p, initstates=init)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        hoc.run_model(model, inp, p, initstates=init)
        best = min(best, time.perf_counter() - t0)
    return (n * Ts / best, best, f"{n} basins x {Ts} steps of the same synthetic workload, C oracle "
            f"(oracle/hydro_oracle.c, OpenMP over node blocks, {cores} threads; best of 2)", cores, rate1)


def bench_raster_ingest(args):
    """D8 flow directions of a 4096 x 4096 DEM and the gather of 3 forcing rasters
    (30 steps x 4096 x 4096) into the stepper's [T][N][V] layout, both HBM-bound; device-resident, CUDA events."""
    import torch
    import ctypes as C
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200.engine import check, lib
    hm.engine.init(0)
    peak, peak_src = peaks()
    n, T, V = 4096, 30, 3
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    dem = torch.randn(n, n, generator=g, device="cuda", dtype=torch.float64) * 50 + 500
    flw = torch.empty((n, n), device="cuda", dtype=torch.uint8)
    cubes = [torch.randn(T, n, n, generator=g, device="cuda", dtype=torch.float64) for _ in range(V)]
    rr, cc = torch.meshgrid(torch.arange(n, device="cuda", dtype=torch.int32), torch.arange(n, device="cuda", dtype=torch.int32), indexing="ij")
    d_r, d_c = rr.reshape(-1).contiguous(), cc.reshape(-1).contiguous()
    ptrs = torch.tensor([c.data_ptr() for c in cubes], dtype=torch.int64, device="cuda")
    out = torch.empty((T, n * n, V), device="cuda", dtype=torch.float64)
    res = {}
    for name, fn, nbytes in [
            ("d8_flow_direction", lambda: check(lib().hb_d8_flow_direction_device(dem.data_ptr(), n, n, n, 1, 30.0, 30.0, flw.data_ptr(), None)), n * n * 9),
            ("pack_forcing_trc", lambda: check(lib().hb_pack_forcing_device(ptrs.data_ptr(), V, T, n * n, n, 1, d_r.data_ptr(), d_c.data_ptr(), n * n, out.data_ptr(), None)),
             T * n * n * V * 16),
            ("pack_forcing_tcr", lambda: check(lib().hb_pack_forcing_device(ptrs.data_ptr(), V, T, n * n, 1, n, d_r.data_ptr(), d_c.data_ptr(), n * n, out.data_ptr(), None)),
             T * n * n * V * 16),
            ("pack_forcing_dense_trc", lambda: check(lib().hb_pack_forcing_dense_device(ptrs.data_ptr(), V, T, n * n, n, 1, n, n, out.data_ptr(), None)),
             T * n * n * V * 16),
            ("pack_forcing_dense_tcr", lambda: check(lib().hb_pack_forcing_dense_device(ptrs.data_ptr(), V, T, n * n, 1, n, n, n, out.data_ptr(), None)),
             T * n * n * V * 16)]:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[name] = {"ms": ms, "algorithmic_GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
    # host rasters -> packed device forcing, streamed (hb_ingest_forcing): pinned sources, then ordinary numpy arrays
    import ctypes as C
    from hydromodels_jl_b200 import rasters as rs
    host_bytes = T * n * n * 8
    probe = pcie_probe(torch, None, 1, lambda: None)
    pinned = [PinnedArray((T, n, n)) for _ in range(V)]
    for k, pa in enumerate(pinned):
        pa.array[...] = cubes[k].cpu().numpy()
    pageable = [np.array(pa.array) for pa in pinned]
    ingest = {}
    for name, srcs in (("pinned", [pa.array for pa in pinned]), ("pageable", pageable)):
        dst = hm.engine.DeviceBuffer(V * host_bytes)
        for _ in range(2):
            rs.ingest_forcing(srcs, None, layout="trc", out=dst)
        reps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            rs.ingest_forcing(srcs, None, layout="trc", out=dst)
        dt = (time.perf_counter() - t0) / reps
        dst.free()
        ingest[name] = {"s_per_pass": dt, "host_GBps": V * host_bytes / dt / 1e9, "cell_timesteps_per_s": n * n * T / dt}
    check(lib().hb_pack_forcing_dense_device(ptrs.data_ptr(), V, T, n * n, n, 1, n, n, out.data_ptr(), None))
    o = rs.ingest_forcing(pageable, None, layout="trc")
    same = bool(torch.equal(torch.from_numpy(o.to_host((T, n * n, V))), out.cpu()))
    o.free()
    for v in ingest.values():
        v["frac_of_h2d_probe"] = v["host_GBps"] / probe["sum_GBps"]["h2d_alone"]
    for pa in pinned:
        pa.free()
    res["ingest_from_host"] = dict(ingest, bytes_per_pass=V * host_bytes, pcie_probe=probe, identical_to_device_gather=same,
                                   note="hb_ingest_forcing: two chunk slots, copy of chunk c+1 under the gather of chunk c; wall clock of the call, output buffer reused")
    print(json.dumps({"metric": "raster ingest (cells or cell-steps)/s", "value": n * n * T / (res["pack_forcing_trc"]["ms"] * 1e-3), "unit": "cell-timesteps/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": res["pack_forcing_trc"]["ms"], "higher_is_better": True,
                      "dtype": "f64", "data": "synthetic", "config": {"workload": "raster_ingest_4096", "grid": [n, n], "timesteps": T, "variables": V,
                                                                   "l2": "every launch moves > 1 GB, far beyond the 126 MB L2"},
                      "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "kernels": res},
                      "gpu_launches": 5 * (args.steps + 3)}))
    return 0


def cpu_reference_rate_grid(T, target_seconds=10.0):
    """CPU baseline of the distributed workload: the C oracle (per-component passes + D8 Jacobi routing, OpenMP) on a
    square sample of the same synthetic grid model."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hydro_oracle_c as hoc
    cores = hoc.max_threads()

    def problem(side):
        rng = np.random.default_rng(sy.SEED)
        n = side * side
        flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128]), size=(side, side), p=[.3, .2, .3, .05, .03, .02, .05, .05])
        rr, cc = np.divmod(np.arange(n), side)
        model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
        pp = sy.parameters("exphydro", 0, n)
        pp.update(area_coef=np.full(n, 1.0), lag=rng.uniform(0.1, 2.0, n))
        init = dict(snowpack=np.zeros(n), soilwater=np.full(n, 1303.0), s_river=np.zeros(n))
        return model, sy.forcing("exphydro", 0, n, T), {"params": pp}, init
    side = 128
    model, inp, p, init = problem(side)
    hoc.run_model(model, inp, p, initstates=init)
    t0 = time.perf_counter()
    hoc.run_model(model, inp, p, initstates=init, n_threads=1)
    rate1 = side * side * T / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    hoc.run_model(model, inp, p, initstates=init)
    est = side * side * T / (time.perf_counter() - t0)
    side = int(min(1024, max(128, (est * target_seconds / 2 / T) ** 0.5))) // 32 * 32
    model, inp, p, init = problem(side)
    hoc.run_model(model, inp, p, initstates=init)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        hoc.run_model(model, inp, p, initstates=init)
        best = min(best, time.perf_counter() - t0)
    return (side * side * T / best, best, f"{side} x {side} cells x {T} steps of the same synthetic grid model, C oracle (oracle/hydro_oracle.c: "
            f"component passes + D8 Jacobi routing, OpenMP, {cores} threads; best of 2)", cores, rate1)


def bench_single_basin(args):
    """BASELINE config 1: the README example — ExpHydro, ONE basin, 365 days, through the functor with host arrays.  A single
    thread of a single warp: launch + copy latency, not bandwidth; reported next to the C oracle on one host core."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hydro_oracle_c as hoc
    hm.engine.init(0)
    T = 365
    model = hm.models.exphydro()
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    p = {"params": {k: float(np.asarray(v)[0]) for k, v in sy.parameters("exphydro", 0, 1).items()}}
    init = dict(snowpack=0.0, soilwater=1303.0)
    eng = hm.B200Model(model)
    timing = {}
    for _ in range(5):
        out = eng(inp, p, initstates=init, timing=timing)
    n = max(20, args.steps * 20)
    t0 = time.perf_counter()
    for _ in range(n):
        eng(inp, p, initstates=init, timing=timing)
    call_us = (time.perf_counter() - t0) / n * 1e6
    ref = hoc.run_model(model, inp, p, initstates=init, n_threads=1)
    t0 = time.perf_counter()
    for _ in range(200):
        hoc.run_model(model, inp, p, initstates=init, n_threads=1)
    cpu_us = (time.perf_counter() - t0) / 200 * 1e6
    err = float(np.max(np.abs(out - ref) - (1e-10 * np.abs(ref) + 1e-12)))     # the parity tests' criterion: <= 0 passes
    print(json.dumps({"metric": "basin-timesteps/sec", "value": T / (call_us * 1e-6), "unit": "basin-timesteps/s", "n_gpus": 1, "steps": n, "warmup": 5,
                      "ms_per_step": call_us * 1e-3, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "exphydro_single_basin_365d", "basins": 1, "timesteps": T, "api": "B200Model.__call__ with host arrays"},
                      "latency_us": {"functor_call": call_us, "kernel": timing.get("kernel_ms", 0.0) * 1e3, "c_oracle_one_core": cpu_us,
                                     "reference_published_per_1000_steps_ms": 0.95},
                      "parity_excess_vs_oracle": err, "parity_ok": bool(err <= 0.0), "gpu_launches": n + 5,
                      "note": "one thread of one warp: latency-bound by construction; a single basin belongs on the CPU, the GPU path exists for API parity"}))
    return 0


def bench_irf(args):
    """IRF routing of a synthetic RIVER TREE — 65 536 nodes, horizon 32, 3650 steps.  The aggregated
    kernel (one row per (destination, upstream source) pair, ~1.1 M rows) is built end to end by `aggregate_route_kernel`
    (host plan from the sparse edge list + level-synchronous composition on the device, timed), and `SparseRouteConvolution`
    then runs on it: 2 * nnz * T * H flop per pass, FP64-FMA bound.  `--workload irf_route_ragged` keeps the Pareto-count
    synthetic CSR of round 1 for comparison."""
    import torch
    import ctypes as C
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import irf
    from hydromodels_jl_b200.engine import check, lib
    hm.engine.init(0)
    n, H, T = 65536, 32, 3650
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    ragged = args.workload == "irf_route_ragged"
    build = None
    L = lib()
    L.hb_sparse_route_conv_ordered_device.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64] + [C.c_void_p] * 7
    if ragged:
        # row counts of an aggregated kernel are skewed (headwaters see themselves, outlets see hundreds of upstream nodes):
        # Pareto-distributed counts, mean ~8, capped at 256
        cnt = np.minimum(256, 1 + np.floor(3.2 * np.random.default_rng(3).pareto(1.4, n))).astype(np.int64)
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        nnz = int(ptr[-1])
        row_ptr = torch.tensor(ptr, dtype=torch.int32, device="cuda")
        dest_of = torch.repeat_interleave(torch.arange(n, device="cuda"), torch.tensor(cnt, device="cuda"))
        src = (dest_of - torch.randint(0, 256, (nnz,), generator=g, device="cuda")).clamp_(min=0).to(torch.int32).contiguous()
        order = torch.tensor(irf.balanced_destination_order(cnt, int(os.environ.get("HB_IRF_ORDER_WINDOW", "4096"))), device="cuda")
        w = torch.rand(H, nnz, generator=g, device="cuda", dtype=torch.float64) / H
        ptrs = (row_ptr.data_ptr(), src.data_ptr(), w.data_ptr(), order.data_ptr())
    else:
        down = irf.synthetic_river_tree(n, seed=3)
        up = np.nonzero(down >= 0)[0]
        rng = np.random.default_rng(3)
        route = irf.RouteIRF(["k"], lambda p, dt, h: (1 - np.exp(-dt / p[0])) * np.exp(-np.arange(h) * dt / p[0]))   # linear reservoir
        kernels = irf.build_irf_kernels(route, rng.uniform(0.5, 3.0, (n, 1)), horizon=H)
        irf.aggregate_route_kernel((down[up][:64], up[:64]), kernels[:max(int(down[up][:64].max()) + 1, 65)], horizon=H)   # warm-up: module load
        t0 = time.perf_counter()
        K = irf.aggregate_route_kernel((down[up], up), kernels, horizon=H, keep_on_device=True)
        build_s = time.perf_counter() - t0
        conv = irf.SparseRouteConvolution(K)                      # destinations longer than 64 rows are split into segments
        conv_whole = irf.SparseRouteConvolution(K, segment_rows=None)
        nnz = int(K.indices.shape[0])
        cnt = np.bincount(K.indices[:, 0] - 1, minlength=n)
        t0 = time.perf_counter()
        plan = irf.plan_route_composition((down[up], up), n, None)
        plan_s = time.perf_counter() - t0
        build = {"aggregate_route_kernel_s": build_s, "host_plan_s": plan_s, "levels": len(plan.level_row_ptr) - 1, "rows": nnz,
                 "truncated_convolutions": int(plan.contrib_row.size), "network": "synthetic_river_tree(65536, seed=3): 1-3 tributaries per node"}
    x = torch.rand(T, n, generator=g, device="cuda", dtype=torch.float64)
    out = torch.empty(T, n, device="cuda", dtype=torch.float64)

    def timed(fn, steps):
        for _ in range(min(3, steps)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    if ragged:
        call = lambda order_ptr: (lambda: check(L.hb_sparse_route_conv_ordered_device(n, n, T, H, nnz, ptrs[0], ptrs[1], ptrs[2], x.data_ptr(),
                                                                                       order_ptr, out.data_ptr(), None)))
        ms = timed(call(ptrs[3]), args.steps)
        ref = out.clone()
        extra = {"unordered_ms": timed(call(None), args.steps), "rows_per_destination": {"mean": nnz / n, "max": int(cnt.max())}}
        extra["bit_identical_to_unordered"] = bool(torch.equal(ref, out))
    else:
        ms = timed(lambda: conv.launch(x.data_ptr(), out.data_ptr(), T), args.steps)
        ref = out.clone()
        whole_ms = timed(lambda: conv_whole.launch(x.data_ptr(), out.data_ptr(), T), 2)
        rel = float(((ref - out).abs() / out.abs().clamp_min(1e-300)).max().item())
        extra = {"rows_per_destination": {"mean": nnz / n, "max": int(cnt.max())}, "segment_rows": 64, "virtual_destinations": conv.n_virtual,
                 "whole_destinations_ms": whole_ms, "max_relative_difference_vs_whole_destinations": rel, "kernel_build": build}
    flop = 2.0 * nnz * T * H
    v = C.c_double()
    L.hb_probe_fp64_tflops(C.byref(v))
    peak, peak_src = peaks()
    print(json.dumps({"metric": "node-timesteps/sec of the IRF routing convolution", "value": n * T / (ms * 1e-3), "unit": "node-timesteps/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": ("irf_route_ragged_65536x3650_h32" if ragged else "irf_route_river_tree_65536x3650_h32"), "nodes": n, "nnz": nnz,
                                 "horizon": H, "timesteps": T, **extra},
                      "roofline": {"bound": "fp64", "achieved": flop / ms / 1e9, "peak": v.value, "unit": "TFLOP/s", "frac": flop / ms / 1e9 / v.value,
                                   "peak_source": "measured DFMA probe (hb_probe_fp64_tflops)", "kernel": "hb_sparse_conv", "kernel_ms": ms,
                                   "hbm_GBps_compulsory": (2 * n * T * 8 + nnz * H * 8) / ms / 1e6, "hbm_peak": peak},
                      "gpu_launches": 3 * (args.steps + 3) + (build["levels"] + 2 if build else 0)}))
    return 0


def bench_calibration(args):
    """differential evolution on the device — ExpHydro, 65 536 members x 10 years x `steps`
    generations; every generation is trial kernel + ensemble objective kernel + selection kernel, no host round trip."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import calibration as cal, synthetic as sy
    hm.engine.init(0)
    T, N, gens = 3650, 65536, max(1, args.steps)
    model = hm.models.exphydro()
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    lb = np.array([-3.0, 0.0, 0.5, 500.0, 5.0, 0.005]); ub = np.array([0.0, 3.0, 5.0, 2500.0, 40.0, 0.1])
    truth = np.array([-2.09, 0.17, 2.674, 1709.46, 18.47, 0.0167])
    init = dict(snowpack=0.0, soilwater=1303.0)
    target = model(inp, {"params": dict(zip(model.get_param_names(), truth))}, initstates=init)[-1]
    prob = cal.OptimizationProblem(model, inp, target, metric="NSE", warm_up=366, lb_pas=lb, ub_pas=ub, default_initstates=init)
    cal.solve(prob, cal.DifferentialEvolution(popsize=N, seed=1), maxiters=3)            # warm-up (NVRTC, allocations)
    sol = cal.solve(prob, cal.DifferentialEvolution(popsize=N, seed=2), maxiters=gens)
    rate = N * T * (gens + 1) / sol.elapsed_s
    print(json.dumps({"metric": "member-timesteps/sec inside a device-resident calibration", "value": rate, "unit": "member-timesteps/s",
                      "n_gpus": 1, "steps": gens, "warmup": 3, "ms_per_step": 1e3 * sol.elapsed_s / (gens + 1), "higher_is_better": True,
                      "dtype": "f64", "data": "synthetic", "config": {"workload": "calibrate_exphydro_65536x3650", "members": N, "timesteps": T,
                                                                   "generations": gens, "algorithm": "DE/rand/1/bin F=0.7 CR=0.9"},
                      "best_objective": sol.objective, "evaluations": sol.n_evaluations, "evaluations_per_s": sol.n_evaluations / sol.elapsed_s,
                      "gpu_launches": sol.gpu_launches, "timing": "host wall clock around the whole loop incl. final D2H of the population"}))
    return 0


def _fp64_instructions_per_step(ck):
    """FP64-pipe instructions (DFMA/DMUL/DADD/DSETP) inside the time loop of a compiled stepper, counted in its SASS
    (static count of the largest backward-branch region; a few belong to rarely taken division slow paths)."""
    import re, subprocess, tempfile
    try:
        with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
            f.write(ck.cubin()); f.flush()
            sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True, timeout=120).stdout
        rows = []
        for l in sass.splitlines():
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                rows.append((int(m.group(1), 16), m.group(2).strip()))
        loops = []
        for a, ins in rows:
            if "BRA" in ins:
                t = re.findall(r"0x([0-9a-f]+)", ins)
                if t and int(t[-1], 16) < a:
                    loops.append((int(t[-1], 16), a))
        lo, hi = max(loops, key=lambda x: x[1] - x[0])
        ops = [re.sub(r"^@!?\w+\s+", "", ins).split()[0].split(".")[0] for a, ins in rows if lo <= a <= hi]
        return sum(o in ("DFMA", "DMUL", "DADD", "DSETP") for o in ops), len(ops)
    except Exception:
        return None, None


def bench_gradient(args):
    """SURVEY.md §8f row 3 (second half): d NSE / d parameters of an ExpHydro calibration ensemble, 262 144 members x 10
    years, 6 directions per launch (forward-mode tangents in the fused stepper) beside the plain objective launch."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    hm.engine.init(0)
    T, N = 3650, 262144
    model = hm.models.exphydro()
    eng = hm.B200Model(model)
    inp = sy.forcing("exphydro", 0, 1, T)[:, 0, :]
    p = {"params": sy.parameters("exphydro", 0, N)}
    init = dict(snowpack=0.0, soilwater=1303.0)
    target = 0.5 + 5.0 * np.random.default_rng(1).random(T)
    kw = dict(metric="NSE", warm_up=366, initstates=init, n_members=N)
    ms_g, ms_o = [], []
    for i in range(args.warmup + args.steps):
        tg, to = {}, {}
        loss, g = eng.gradient(inp, p, target, timing=tg, **kw)
        obj = eng.objective(inp, p, target, timing=to, **kw)
        if i >= args.warmup:
            ms_g.append(tg["kernel_ms"]); ms_o.append(to["kernel_ms"])
    info = tg["kernel"].info()
    mg, mo = float(np.mean(ms_g)), float(np.mean(ms_o))
    import ctypes as C
    pk = C.c_double()
    hm.engine.lib().hb_probe_fp64_tflops(C.byref(pk))
    roof = {}
    for tag, ck, ms in (("gradient", tg["kernel"], mg), ("objective", to["kernel"], mo)):
        fp, tot = _fp64_instructions_per_step(ck)
        if fp:
            rate = fp * N * T / (ms * 1e-3)                  # FP64-pipe thread-instructions per second
            roof[tag] = {"fp64_instructions_per_step": fp, "instructions_per_step": tot, "achieved": rate / 1e12,
                         "peak": pk.value / 2.0, "unit": "T FP64 instr/s", "frac": rate / 1e12 / (pk.value / 2.0)}
    print(json.dumps({"metric": "member-timesteps/sec of a gradient run (objective + 6 derivatives per member)", "value": N * T / (mg * 1e-3),
                      "unit": "member-timesteps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mg,
                      "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "gradient_exphydro_262144x3650", "members": N, "timesteps": T, "directions": 6, "metric": "NSE"},
                      "objective_only_ms": mo, "gradient_over_objective": mg / mo, "registers": info["registers_per_thread"],
                      "spill_bytes": info["local_bytes_per_thread"], "blocks_per_sm": info["max_blocks_per_sm"],
                      "roofline": {"bound": "fp64 issue", "kernel": "hb_stepper (HB_TANGENTS 6)", "peak_source": "measured DFMA probe (hb_probe_fp64_tflops) / 2 flop",
                                   **roof},
                      "max_abs_loss_difference_vs_objective_kernel": float(np.abs(loss - obj).max()), "gpu_launches": 2 * (args.warmup + args.steps),
                      "timing": "CUDA events around each kernel (hb_run elapsed_ms)"}))
    return 0


def bench_adjoint(args):
    """SURVEY.md §8f row 3 (reverse mode): d SSE / d (690 MLP weights, 11 physical parameters, 2 initial states) of M50 on
    50 000 basins x 3650 steps in ONE backward sweep over the stored state trajectory (templates/adjoint.cu), next to the
    forward run that stores the trajectory.  What the reference does with Zygote through its ImmutableSolver
    (/root/reference/README.md:161-187)."""
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    hm.engine.init(0)
    N, T = 50_000, 3650
    model = hm.models.m50(htypes=np.arange(1, N + 1))
    et, q = hm.models.m50_chains()
    p = {"params": sy.parameters("m50", 0, N), "nns": {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}}
    init = {k: np.full(N, v) for k, v in sy.INIT["m50"].items()}
    inp = sy.forcing("m50", 0, N, T)
    target = np.abs(0.3 * inp[0, 0] + 0.5)                 # one observed series shared by all basins
    eng = hm.B200Model(model)
    ms_a, ms_f = [], []
    for i in range(max(1, args.warmup // 2) + args.steps):
        tm = {}
        loss, g = eng.gradient(inp, p, target, metric="SSE", warm_up=366, initstates=init, timing=tm)
        if i >= max(1, args.warmup // 2):
            ms_a.append(tm["kernel_ms"]); ms_f.append(tm["forward_ms"])
    info = tm["kernel"].info()
    ma, mf = float(np.mean(ms_a)), float(np.mean(ms_f))
    gn = np.concatenate([g["nns"]["etnn"], g["nns"]["qnn"]])
    print(json.dumps({"metric": "basin-timesteps/sec of a reverse-mode gradient run (backward sweep)", "value": N * T / (ma * 1e-3),
                      "unit": "basin-timesteps/s", "n_gpus": 1, "steps": args.steps, "warmup": max(1, args.warmup // 2), "ms_per_step": ma,
                      "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "gradient_m50_adjoint_50k_basins_x_3650d", "basins": N, "timesteps": T, "metric": "SSE",
                                 "gradient_entries": {"mlp_weights": int(gn.size), "physical_parameters_per_basin": len(model.param_names),
                                                      "initial_states_per_basin": len(model.state_names)}},
                      "forward_trajectory_ms": mf, "adjoint_over_forward": ma / mf if mf else None,
                      "registers": info["registers_per_thread"], "spill_bytes": info["local_bytes_per_thread"], "blocks_per_sm": info["max_blocks_per_sm"],
                      "finite": bool(np.all(np.isfinite(gn)) and np.all(np.isfinite(loss))), "weight_gradient_norm": float(np.linalg.norm(gn)),
                      "gpu_launches": 2 * (args.steps + max(1, args.warmup // 2)), "timing": "CUDA events around each kernel"}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS) + ["raster_ingest_4096", "calibrate_exphydro", "gradient_exphydro", "gradient_m50_adjoint", "irf_route", "irf_route_ragged", "exphydro_single_basin_365d"])
    ap.add_argument("--basins", type=int, default=0, help="tuning experiments: override the workload's basin count")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"], help="opt-in Float32 mode (per-node workloads; device-resident value only)")
    ap.add_argument("--nn-tensor", type=int, default=-1, help="neural fluxes on the tensor cores: 1 / 0 / -1 = the engine's default for the precision")
    ap.add_argument("--no-bind", action="store_true", help="N > 1: do not bind each rank to its own slice of the host cores")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host-array leg of the end-to-end measurement")
    ap.add_argument("--halo-exchange", default="peer", choices=["peer", "nccl", "torch"],
                    help="sharded grid runs: how ghost states travel (library peer stores over NVLink | library NCCL | torch.distributed)")
    ap.add_argument("--no-secondary", action="store_true", help="N > 1: skip the config-5 (one 4096 x 4096 grid over all ranks) secondary line")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--block", type=int, default=0, help="tuning: threads per block of the stepper")
    ap.add_argument("--stages", type=int, default=0, help="tuning: forcing pipeline depth (time steps)")
    ap.add_argument("--maxreg", type=int, default=0, help="tuning: register budget per thread")
    ap.add_argument("--npw", type=int, default=0, help="tuning: nodes per warp of tensor-path MLP bodies (16 | 32; 0 = by node count)")
    ap.add_argument("--grid-tc", type=int, default=0, help="grid kernels: time steps per launch (0 = auto for wave: from the grid width; 2 for tile)")
    ap.add_argument("--grid-bw", type=int, default=32, help="fused grid kernel: region width (cells)")
    ap.add_argument("--grid-bh", type=int, default=32, help="fused grid kernel: region height (cells)")
    ap.add_argument("--grid-kernel", default="wave", choices=["wave", "tile", "generic"],
                    help="grid workload: wave = single-pass L2-dataflow kernel (gridwave.cu, default); tile = tile + halo "
                         "(grid.cu); generic = stepper over all T, then one routing launch per step")
    ap.add_argument("--route-passes", type=int, default=3, choices=[2, 3],
                    help="--grid-kernel generic: 3 = stepper / route on compact planes / stepper (records written once); 2 = round 1's "
                         "stepper + in-place route patches")
    ap.add_argument("--fused-grid", action="store_true", help="(round-1 flag) same as --grid-kernel tile")
    ap.add_argument("--grid-panels", type=int, default=-1,
                    help="single-GPU grid workload, wave kernel: run the grid as this many column panels one after the other "
                         "(0 = auto: one per ~1024 columns; 1 = one kernel over the whole grid; -1 = auto for grids wider than 3000 columns)")
    ap.add_argument("--grid-split", default="rows", choices=["rows", "cols"],
                    help="multi-GPU grid workload: row strips (the ranks' square grids stacked vertically) or column panels (side by side; "
                         "the wave kernel runs tall narrow sub-grids faster than short wide ones)")
    args = ap.parse_args()
    if args.fused_grid:
        args.grid_kernel = "tile"
    if args.grid_tc <= 0 and args.grid_kernel != "wave":
        args.grid_tc = 2
    if args.grid_tc < 0:
        args.grid_tc = 0          # wave: 0 = the library chooses from the grid width
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if local_world > 1 and not args.no_bind and hasattr(os, "sched_setaffinity"):
        # one contiguous slice of the host cores per rank: the library's host threads (pageable staging) and the
        # launching thread of a rank stay off the other ranks' cores
        cores = sorted(os.sched_getaffinity(0))
        k = len(cores) // local_world
        if k >= 1:
            os.sched_setaffinity(0, cores[local_rank * k:(local_rank + 1) * k])
    if args.workload == "raster_ingest_4096":
        return bench_raster_ingest(args) if rank == 0 else 0
    if args.workload == "calibrate_exphydro":
        return bench_calibration(args) if rank == 0 else 0
    if args.workload == "gradient_exphydro":
        return bench_gradient(args) if rank == 0 else 0
    if args.workload == "gradient_m50_adjoint":
        return bench_adjoint(args) if rank == 0 else 0
    if args.workload in ("irf_route", "irf_route_ragged"):
        return bench_irf(args) if rank == 0 else 0
    if args.workload == "exphydro_single_basin_365d":
        return bench_single_basin(args) if rank == 0 else 0
    model_name, N, T = WORKLOADS[args.workload]
    if args.basins:          # tuning experiments only (the config key `basins_per_gpu` reports what ran)
        N = int(args.basins)
    total_grid = "_total_" in args.workload
    if total_grid:
        N //= int(os.environ.get("WORLD_SIZE", "1"))          # cells owned per rank
    rows_of = {"exphydro": 10, "gr4j": 15, "hbv": 18, "m50": 12, "hbv_ensemble": 0, "exphydro_grid": 13}[model_name]
    v_of = {"exphydro": 3, "gr4j": 2, "hbv": 3, "m50": 3, "hbv_ensemble": 0, "exphydro_grid": 3}[model_name]
    unit = "basin-timesteps/s"
    config = {"workload": args.workload, "bucket_model": model_name, "basins_per_gpu": N, "timesteps": T,
              "rows_out": rows_of, "forcing_rows": v_of, "output": "full [rows,N,T] as the reference returns",
              "sharding": f"basins sharded over {world} rank(s), no data-path collective",
              "l2": "inputs+outputs per step (GBs) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, args.steps)
        rates = []
        rate1 = None
        for _ in range(min(steps, 3)):
            rate, dt, sample, cores, rate1 = cpu_reference_rate_grid(T, 8.0) if model_name == "exphydro_grid" else cpu_reference_rate(model_name, T, 8.0)
            rates.append(rate)
        v = float(np.median(rates))
        print(json.dumps({
            "impl": "reference", "metric": "basin-timesteps/sec", "value": v, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             "one_thread_value": rate1},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference is pure Julia (no Julia in this image): timed the CPU oracle port; the reference's own "
                    "published figure is ~1.05e6 basin-timesteps/s on one thread (docs/src/get_start_en.md:408)"}))
        return 0

    import torch
    import hydromodels_jl_b200 as hm
    from hydromodels_jl_b200 import synthetic as sy
    from hydromodels_jl_b200.engine import DeviceRun

    torch.cuda.set_device(local_rank)
    hm.engine.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n0 = rank * N
    special = model_name in ("hbv_ensemble", "exphydro_grid")
    if model_name == "hbv_ensemble":
        # one basin, N parameter sets: forcing [T][1][3] shared by all members, objective = -NSE
        args.no_e2e = True
        model = hm.models.hbv()
        p = {"params": sy.parameters("hbv", n0, n0 + N)}
        init = {k: np.full(N, v) for k, v in sy.INIT["hbv"].items()}
        forcing1 = sy.forcing("hbv", 0, 1, T)
        target = np.abs(forcing1[0, 0] * 0.3 + 0.5)
        eng = hm.B200Model(model, block_threads=args.block, stages=args.stages, max_registers=args.maxreg)
        run = DeviceRun(eng, N, T, p, initstates=init, objective="NSE", target=target, warm_up=366, ensemble=True)
        d_in = torch.from_numpy(np.ascontiguousarray(forcing1[:, 0, :].T)).cuda()
        d_out = torch.empty((N,), device="cuda", dtype=torch.float64)
    elif model_name == "exphydro_grid":
        # weak scaling: every rank owns a `side`-row strip of a (side*world) x side grid; the strips are
        # coupled through the D8 flow directions, so boundary outflows are exchanged every step
        from hydromodels_jl_b200.engine import RoutedDeviceRun
        from hydromodels_jl_b200.parallel import ShardedRoutedRun
        args.no_e2e = True
        side = int(round(N ** 0.5))
        Ng = N * world
        rng = np.random.default_rng(sy.SEED)
        panels = world > 1 and args.grid_kernel == "wave" and args.grid_split == "cols"
        g_rows, g_cols = (side, side * world) if panels else (side * world, side)
        if total_grid:
            g_rows = g_cols = int(round(Ng ** 0.5))
            side = g_cols
            config["scaling_note"] = f"one {g_rows} x {g_cols} grid cut into {world} {'column panels' if panels else 'row strips'}"
            if world > 1 and args.grid_kernel != "wave":
                raise SystemExit("the total-grid workload is sharded with the wave kernel only")
        flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128]), size=(g_rows, g_cols), p=[.3, .2, .3, .05, .03, .02, .05, .05])
        rr, cc = np.divmod(np.arange(Ng), g_cols)
        pos = np.stack([rr + 1, cc + 1], axis=1)
        model = hm.models.exphydro_grid(flw, pos)
        pp = sy.parameters("exphydro", 0, Ng)
        pp.update(area_coef=np.full(Ng, 1.0), lag=rng.uniform(0.1, 2.0, Ng))
        p = {"params": pp}
        init = dict(snowpack=np.zeros(Ng), soilwater=np.full(Ng, 1303.0), s_river=np.zeros(Ng))
        eng = hm.B200Model(model, block_threads=args.block, stages=args.stages, max_registers=args.maxreg)
        if panels:
            # column panels + 32 ghost columns, states exchanged every 32 steps
            from hydromodels_jl_b200.parallel import ShardedPanelRun
            sharded = ShardedPanelRun(eng, g_rows, g_cols, T, p, initstates=init, halo=32, stages=args.stages,
                                      max_registers=args.maxreg, block_threads=args.block, steps_per_launch=args.grid_tc)
            assert sharded.N == N
            eng.last_kernel = sharded.run.ck
            grid_run = sharded
            config["grid_split"] = "column panels, 32 ghost columns"

            class _Run:
                def launch(self, ip, op, stream=None):
                    sharded.launch(ip, op)
            run = _Run()
        elif world > 1 and args.grid_kernel == "wave":
            # row strips + ghost rows, bucket / river states exchanged every `grid_tc` steps (temporal blocking)
            from hydromodels_jl_b200.parallel import ShardedGridRun
            sharded = ShardedGridRun(eng, g_rows, g_cols, T, p, initstates=init, halo=args.grid_tc or 16, stages=args.stages,
                                     max_registers=args.maxreg, block_threads=args.block)
            assert sharded.N == N
            eng.last_kernel = sharded.run.ck
            grid_run = sharded
            n_ghost = sharded.plan.n_local - N

            class _Run:
                def launch(self, ip, op, stream=None):
                    sharded.launch(ip, op)
            run = _Run()
        elif world > 1:
            sharded = ShardedRoutedRun(eng, Ng, T, p, initstates=init)
            assert sharded.N == N

            class _Run:                      # same launch signature as the single-GPU runs
                def launch(self, ip, op, stream=None):
                    sharded.launch(ip, op)
            run = _Run()
        elif args.grid_kernel == "generic":
            run = RoutedDeviceRun(eng, N, T, p, initstates=init, legacy=(args.route_passes == 2))
            config["generic_route_passes"] = 2 if run.legacy else 3
        else:
            from hydromodels_jl_b200.engine import GridDeviceRun
            n_panels = args.grid_panels if args.grid_panels >= 0 else (0 if g_cols > 3000 else 1)
            if args.grid_kernel == "wave" and n_panels != 1:
                from hydromodels_jl_b200.parallel import PanelledGridRun
                run = PanelledGridRun(eng, g_rows, g_cols, T, p, initstates=init, panels=n_panels, steps_per_launch=args.grid_tc,
                                      block_threads=args.block, stages=args.stages, max_registers=args.maxreg)
                config["grid_panels"] = run.K
            elif args.grid_kernel == "wave":
                run = GridDeviceRun(eng, N, T, p, initstates=init, variant="wave", steps_per_chunk=args.grid_tc,
                                    block_w=args.block, stages=args.stages, max_registers=args.maxreg)
            else:
                run = GridDeviceRun(eng, N, T, p, initstates=init, steps_per_chunk=args.grid_tc, block_w=args.grid_bw,
                                    block_h=args.grid_bh)
            eng.last_kernel = run.ck
            grid_run = run
        d_in = device_forcing(torch, "exphydro", sharded.plan.n_local if (total_grid and world > 1) else N, T, sy.SEED + 17 * rank)
        if total_grid:
            pass          # one seed for the whole computed region of the rank (ghost cells included): timing only
        elif panels:
            # forcing of the ghost columns: the adjacent columns of the neighbouring ranks' panels (same generator, their seed)
            pl = sharded.plan
            parts = []
            if rank > 0:
                left = device_forcing(torch, "exphydro", N, T, sy.SEED + 17 * (rank - 1)).view(T, side, side, v_of)
                parts.append(left[:, :, side - (pl.c0 - pl.g0):, :].clone())
                del left
            parts.append(d_in.view(T, side, side, v_of))
            if rank < world - 1:
                right = device_forcing(torch, "exphydro", N, T, sy.SEED + 17 * (rank + 1)).view(T, side, side, v_of)
                parts.append(right[:, :, :pl.g1 - pl.c1, :].clone())
                del right
            d_in = torch.cat(parts, dim=2).reshape(T, pl.n_local, v_of).contiguous()
            del parts
            torch.cuda.empty_cache()
        elif world > 1 and args.grid_kernel == "wave":
            # forcing of the ghost rows: the adjacent rows of the neighbouring ranks' forcing (same generator, their seed)
            pl = sharded.plan
            parts = []
            if rank > 0:
                up = device_forcing(torch, "exphydro", N, T, sy.SEED + 17 * (rank - 1))
                parts.append(up[:, N - (pl.r0 - pl.g0) * side:, :].clone())
                del up
            parts.append(d_in)
            if rank < world - 1:
                dn = device_forcing(torch, "exphydro", N, T, sy.SEED + 17 * (rank + 1))
                parts.append(dn[:, :(pl.g1 - pl.r1) * side, :].clone())
                del dn
            d_in = torch.cat(parts, dim=1).contiguous()
            del parts
            torch.cuda.empty_cache()
        d_out = torch.empty((T, N, rows_of), device="cuda", dtype=torch.float64)
    else:
        model, p, init = build_problem(hm, sy, model_name, n0, N, T)
        eng = hm.B200Model(model, block_threads=args.block, stages=args.stages, max_registers=args.maxreg, precision=args.precision,
                           nn_tensor=None if args.nn_tensor < 0 else bool(args.nn_tensor))
        eng.nodes_per_warp = args.npw
        run = DeviceRun(eng, N, T, p, initstates=init)
        d_in = device_forcing(torch, model_name, N, T, sy.SEED + 17 * rank)
        d_out = torch.empty((T, N, rows_of), device="cuda", dtype=torch.float64)
        if args.precision == "f32":
            args.no_e2e = True
            d_in, d_out = d_in.float().contiguous(), d_out.float()
            config.update(precision="f32 (opt-in mode: forcing, arithmetic and records in Float32)",
                          nn_path=("tcgen05.mma kind::tf32, 3xTF32 split" if run.ck.spec.nn_tc_bytes else "per-thread FFMA") if model_name == "m50" else None)
    stream = torch.cuda.current_stream()
    info = eng.last_kernel.info()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        run.launch(d_in.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 10.0:     # nvidia-smi needs seconds to enumerate an 8-GPU box
        time.sleep(0.05)
    time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    e0.record(stream)
    for a, b in evs:
        a.record(stream)
        run.launch(d_in.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        b.record(stream)
    e1.record(stream)
    barrier()
    wall1 = time.time()
    total_ms = e0.elapsed_time(e1)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    clocks = sampler.summary(wall0, wall1)
    t_total = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_total.item())
    value = world * N * T * args.steps / (total_ms_max * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean launch time
    peak, peak_src = peaks()
    bytes_per_unit = (4 if args.precision == "f32" else 8) * (v_of + rows_of)
    launches_per_step = 1
    if model_name == "exphydro_grid":
        generic = args.grid_kernel == "generic" or (world > 1 and args.grid_kernel != "wave")
        if not generic:
            args.grid_tc = grid_run.check()          # raises if a warp of the wave kernel gave up waiting
            config["grid_kernel"], config["steps_per_launch"] = args.grid_kernel, args.grid_tc
        launches_per_step = (T + 2 + (0 if (world > 1 or args.route_passes == 2) else 1)) if generic else -(-T // args.grid_tc)
    k_ms = float(np.mean(kernel_ms))
    achieved = bytes_per_unit * N * T / (k_ms * 1e-3) / 1e9
    kname = {"exphydro_grid": ("hb_stepper + T x hb_route_step (in-place patches)" if (world > 1 or args.route_passes == 2) else
                               "hb_stepper (route inputs) + T x hb_route_step (planes) + hb_stepper (records)") if (args.grid_kernel == "generic" or (world > 1 and args.grid_kernel != "wave")) else
             ("hb_grid_wave" if args.grid_kernel == "wave" else "hb_grid_chunk"),
             "hbv_ensemble": "hb_stepper (objective mode)"}.get(model_name, "hb_stepper")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_basin_timestep": bytes_per_unit, "kernel_ms": k_ms,
                "registers": info["registers_per_thread"], "blocks_per_sm": info["max_blocks_per_sm"],
                "dyn_smem": info["dynamic_smem_bytes"], "block": info["block_threads"], "spill_bytes": info["local_bytes_per_thread"]}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            ent = json.load(open(tr)).get(args.workload)
            if isinstance(ent, dict):
                roofline["traffic"] = ent.get("bytes")
                roofline["traffic_source"] = {"kind": "static_profile (ncu --set full capture, not measured in this run)",
                                              **{k: v for k, v in ent.items() if k != "bytes"}}
        except Exception:
            pass

    # ---- end to end through the public functor, HOST buffers (the headline against the CPU arm) ------------------
    # Same N and T at every world size.  Each rank pins its own buffers through the C-ABI (hb_malloc_host, after hb_init on
    # its device); a second leg passes ordinary pageable numpy arrays (what a Julia `Array` is) through the same call.
    e2e = None
    launches = args.steps * launches_per_step
    if not args.no_e2e:
        try:
            e2e = e2e_leg(torch, dist, hm, eng, p, init, d_in, T, N, v_of, rows_of, world, rank, local_rank, args, unit, barrier)
            launches += e2e.pop("_launches", 0)
        except Exception as ex:  # e.g. not enough pinnable host memory on the box
            e2e = {"value": None, "unit": unit, "error": f"{type(ex).__name__}: {ex}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and (not special or model_name == "exphydro_grid"):
        rate, dt, sample, cores, rate1 = cpu_reference_rate_grid(T) if model_name == "exphydro_grid" else cpu_reference_rate(model_name, T)
        cpu = {"value": rate, "unit": unit, "cores": cores, "kind": "port", "sample": sample, "seconds": dt,
               "one_thread_value": rate1,
               "reference_published_one_thread": 1.05e6}    # docs/src/get_start_en.md:408 (hardware not stated)

    secondary = None
    if args.workload == DEFAULT_WORKLOAD and not args.no_secondary and args.precision == "f64":
        # config 5 (the routed grid over all ranks) next to the headline, so that the driver's 1/2/4/8 runs carry it
        try:
            del d_in, d_out
        except NameError:
            pass
        torch.cuda.empty_cache()
        try:
            secondary = bench_grid_total(torch, dist, hm, sy, args, rank, world, max(2, min(args.steps, 5)), barrier, args.halo_exchange)
        except Exception as ex:
            secondary = {"workload": "exphydro_grid_4096x4096_total_x_60d", "value": None, "error": f"{type(ex).__name__}: {ex}"[:300]}
        launches += (secondary or {}).get("gpu_launches", 0) or 0

    sampler.stop()
    if rank == 0:
        print(json.dumps({
            "metric": "basin-timesteps/sec", "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if total_grid else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "secondary": secondary, "gpu_launches": launches, "clocks": clocks,
            "kernel_ms_each": [round(x, 4) for x in kernel_ms]}))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
