"""How many CTAs of the tcgen05 M50 kernel share an SM?  Time N = 148 * 128 * k basins (k CTAs per SM if co-resident)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceRun
hm.engine.init(0)
T = 1000
res = {}
for nt in (1, 0):
    for k in (1, 2, 3, 4, 6):
        N = 148 * 128 * k
        model = hm.models.m50(htypes=np.arange(1, N + 1))
        et, q = hm.models.m50_chains()
        p = {"params": sy.parameters("m50", 0, N), "nns": {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}}
        init = {kk: np.full(N, v) for kk, v in sy.INIT["m50"].items()}
        eng = hm.B200Model(model, precision="f32", nn_tensor=bool(nt))
        run = DeviceRun(eng, N, T, p, initstates=init)
        d_in = torch.rand((T, N, 3), device="cuda", dtype=torch.float32) * 10
        d_out = torch.empty((T, N, 12), device="cuda", dtype=torch.float32)
        for _ in range(2):
            run.launch(d_in.data_ptr(), d_out.data_ptr(), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run.launch(d_in.data_ptr(), d_out.data_ptr(), None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[f"nn_tensor={nt} ctas_per_sm={k}"] = {"ms": ms, "occupancy_api": eng.last_kernel.info()["max_blocks_per_sm"], "regs": eng.last_kernel.info()["registers_per_thread"]}
        print(nt, k, ms, flush=True)
        del d_in, d_out, run
print(json.dumps(res))
