import sys, os, numpy as np, torch
sys.path.insert(0, "oracle"); sys.path.insert(0, ".")
import hydromodels_jl_b200 as hm, bench
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceRun
import hydro_oracle as ho
hm.engine.init(0)
N, T = 100_000, 3650
model, p, init = bench.build_problem(hm, sy, "gr4j", 0, N, T)
d_in = bench.device_forcing(torch, "gr4j", N, T, 1234)
d_out = torch.empty((T, N, 15), device="cuda", dtype=torch.float64)
DeviceRun(hm.B200Model(model), N, T, p, initstates=init).launch(d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
bad = (~torch.isfinite(d_out)).any(dim=2).any(dim=0)
idx = torch.nonzero(bad).flatten().cpu().numpy()
print("bad nodes", len(idx), idx[:10])
i = idx[:4]
ti = torch.from_numpy(i).cuda()
got = d_out.index_select(1, ti).cpu().numpy().transpose(2, 1, 0)
inp = np.asfortranarray(d_in.index_select(1, ti).cpu().numpy().transpose(2, 1, 0))
sub = hm.models.gr4j(htypes=np.arange(1, len(i) + 1))
ps = {"params": {k: np.asarray(v)[i] for k, v in p["params"].items()}}
print(ps)
ref = ho.run_model(sub, inp, ps, initstates={k: v[i] for k, v in init.items()})
for k in range(len(i)):
    tb_g = np.argmax(~np.isfinite(got[:, k, :]).all(axis=0)); tb_r = np.argmax(~np.isfinite(ref[:, k, :]).all(axis=0))
    print("node", i[k], "first bad t gpu", tb_g, "oracle", tb_r, "oracle finite:", np.isfinite(ref[:, k, :]).all())
    t0 = max(0, tb_g - 2)
    print(" gpu", got[:3, k, t0:tb_g+1]); print(" ref", ref[:3, k, t0:tb_g+1])
