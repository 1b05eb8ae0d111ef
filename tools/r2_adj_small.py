"""One small adjoint run (M50, 18 944 basins x 300 steps) for an ncu capture of hb_adjoint."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
hm.engine.init(0)
N, T = 148 * 128, 300
model = hm.models.m50(htypes=np.arange(1, N + 1))
et, q = hm.models.m50_chains()
p = {"params": sy.parameters("m50", 0, N), "nns": {"etnn": et.init_params(1) * 0.5, "qnn": q.init_params(2) * 0.5}}
init = {k: np.full(N, v) for k, v in sy.INIT["m50"].items()}
inp = sy.forcing("m50", 0, N, T)
tm = {}
loss, g = hm.B200Model(model).gradient(inp, p, np.abs(0.3 * inp[0, 0] + 0.5), metric="SSE", warm_up=30, initstates=init, timing=tm)
print("adjoint ms", tm["kernel_ms"], "forward ms", tm["forward_ms"], float(np.linalg.norm(g["nns"]["qnn"])))
