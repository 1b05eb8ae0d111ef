"""Experiment: how does the wave kernel's pass time depend on the grid's aspect ratio (= the dispatch-order distance
between vertically adjacent blocks) at a fixed cell count?"""
import sys, os, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import GridDeviceRun
import bench

hm.engine.init(0)
T = 60
for rows, cols, tc in [(544, 4096, 16), (544, 4096, 8), (4096, 544, 16), (4096, 576, 16), (4096, 576, 32), (2048, 1088, 16)]:
    rng = np.random.default_rng(1)
    n = rows * cols
    flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128]), size=(rows, cols), p=[.3, .2, .3, .05, .03, .02, .05, .05])
    rr, cc = np.divmod(np.arange(n), cols)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    pp = sy.parameters("exphydro", 0, n)
    pp.update(area_coef=np.full(n, 1.0), lag=rng.uniform(0.1, 2.0, n))
    init = dict(snowpack=np.zeros(n), soilwater=np.full(n, 1303.0), s_river=np.zeros(n))
    eng = hm.B200Model(model)
    run = GridDeviceRun(eng, n, T, {"params": pp}, initstates=init, variant="wave", steps_per_chunk=tc)
    d_in = bench.device_forcing(torch, "exphydro", n, T, 7)
    d_out = torch.empty((T, n, 13), device="cuda", dtype=torch.float64)
    for _ in range(3):
        run.launch(d_in.data_ptr(), d_out.data_ptr(), None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        run.launch(d_in.data_ptr(), d_out.data_ptr(), None)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{rows} x {cols}: {ms:.2f} ms per pass, steps per launch {run.check()}", flush=True)
    del d_in, d_out, run, eng, model
    torch.cuda.empty_cache()
