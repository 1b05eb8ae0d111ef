"""Random-shape fuzz of the wave grid kernel (whole grid, strips, panels, panelled single-GPU run) against the generic
two-pass path: every value must be bit-identical.  usage: python tools/fuzz_wave.py [n_cases] [seed]"""
import sys, os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np, torch
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy, parallel as par
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    R = int(rng.choice([1, 2, 3, 7, 16, 33, 64, 100, 257]))
    Cc = int(rng.choice([1, 2, 5, 31, 32, 33, 64, 96, 130, 256, 515]))
    T = int(rng.choice([1, 2, 5, 17, 33, 48]))
    tc = int(rng.choice([1, 3, 16]))
    N = R * Cc
    flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0, 9], size=(R, Cc))
    rr, cc = np.divmod(np.arange(N), Cc)
    model = hm.models.exphydro_grid(flw, np.stack([rr + 1, cc + 1], axis=1))
    p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))}
    init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
    inp = sy.forcing("exphydro", 0, N, T)
    g = hm.B200Model(model); g.grid_kernel = "generic"
    ref = g(inp, p, initstates=init)
    w = hm.B200Model(model); w.grid_steps_per_launch = tc
    got = w(inp, p, initstates=init)
    ok = np.array_equal(got, ref)
    msg = f"case {case}: {R} x {Cc}, T {T}, tc {tc}: wave {'ok' if ok else 'MISMATCH'}"
    K = int(rng.choice([2, 3]))
    halo = int(rng.choice([1, 4, 32]))
    if Cc // K >= halo and T > 1:
        run = par.PanelledGridRun(hm.B200Model(model), R, Cc, T, p, initstates=init, panels=K, halo=halo, steps_per_launch=tc)
        d_in = torch.from_numpy(np.ascontiguousarray(inp.transpose(2, 1, 0))).cuda()
        rec = torch.zeros((T, N, 13), dtype=torch.float64, device="cuda")
        run.launch(d_in.data_ptr(), rec.data_ptr()); run.check()
        ok2 = np.array_equal(rec.cpu().numpy().transpose(2, 1, 0), ref)
        msg += f"; {run.K} panels halo {halo}: {'ok' if ok2 else 'MISMATCH'}"
        ok = ok and ok2
    bad += 0 if ok else 1
    print(msg, flush=True)
print("FUZZ", "FAILED" if bad else "PASSED", f"({n_cases} cases, {bad} bad)")
