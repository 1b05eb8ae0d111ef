"""Experiment: do independent strips of the grid, run as CONCURRENT wave kernels on separate streams, fill each other's
lock-step wait phases?  K independent (side/K x side) problems on K streams vs one side x side problem."""
import sys, os, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import GridDeviceRun
import bench

hm.engine.init(0)
side, T = 2048, 60


def make(rows, cols, tc, seed):
    rng = np.random.default_rng(seed)
    n = rows * cols
    flw = rng.choice(np.array([1, 2, 4, 8, 16, 32, 64, 128]), size=(rows, cols), p=[.3, .2, .3, .05, .03, .02, .05, .05])
    rr, cc = np.divmod(np.arange(n), cols)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    pp = sy.parameters("exphydro", 0, n)
    pp.update(area_coef=np.full(n, 1.0), lag=rng.uniform(0.1, 2.0, n))
    init = dict(snowpack=np.zeros(n), soilwater=np.full(n, 1303.0), s_river=np.zeros(n))
    eng = hm.B200Model(model)
    run = GridDeviceRun(eng, n, T, {"params": pp}, initstates=init, variant="wave", steps_per_chunk=tc)
    d_in = bench.device_forcing(torch, "exphydro", n, T, seed)
    d_out = torch.empty((T, n, 13), device="cuda", dtype=torch.float64)
    return eng, run, d_in, d_out


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for K, tc in [(1, 16), (2, 8), (2, 4), (4, 4), (4, 2), (8, 2)]:
    probs = [make(side // K, side, tc, 100 + k) for k in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]

    def conc():
        for (eng, run, di, do), st in zip(probs, streams):
            run.launch(di.data_ptr(), do.data_ptr(), st.cuda_stream)

    def seq():
        for eng, run, di, do in probs:
            run.launch(di.data_ptr(), do.data_ptr(), None)
    ms_c, ms_s = timeit(conc), timeit(seq)
    for eng, run, di, do in probs:
        run.check()
    print(f"K {K} tc {tc}: concurrent {ms_c:.2f} ms  sequential {ms_s:.2f} ms  (cells {side * side}, T {T})", flush=True)
    del probs
    torch.cuda.empty_cache()
