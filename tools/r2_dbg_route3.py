import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hydromodels_jl_b200 as hm
from hydromodels_jl_b200 import synthetic as sy
from hydromodels_jl_b200.engine import DeviceBuffer, DeviceRun, RoutedDeviceRun
hm.engine.init(0)
rng = np.random.default_rng(17)
R, Cc, T = 23, 31, 45
flw = rng.choice([1, 2, 4, 8, 16, 32, 64, 128, 0, 3], size=(R, Cc))
pos = [(r + 1, c + 1) for r in range(R) for c in range(Cc) if rng.random() < 0.85]
N = len(pos)
model = hm.models.exphydro_grid(flw, pos)
inp = sy.forcing("exphydro", 0, N, T)
p = {"params": dict(sy.parameters("exphydro", 0, N), area_coef=rng.uniform(0.5, 2.0, N), lag=rng.uniform(0.1, 3.0, N))}
init = dict(snowpack=np.zeros(N), soilwater=np.full(N, 1303.0), s_river=rng.uniform(0, 2, N))
e3, e2 = hm.B200Model(model), hm.B200Model(model)
e3.grid_kernel = e2.grid_kernel = "generic"
e2.generic_route_passes = 2
g3, g2 = e3(inp, p, initstates=init), e2(inp, p, initstates=init)
names = model.var_names
for r, nm in enumerate(names):
    d = g3[r] != g2[r]
    if d.any():
        idx = np.argwhere(d)
        print(nm, int(d.sum()), "first", idx[0], g3[r][tuple(idx[0])].hex(), g2[r][tuple(idx[0])].hex(), "steps", np.unique(idx[:, 1])[:8])
# q plane of pass 1 vs q row
run = RoutedDeviceRun(e3, N, T, p, initstates=init)
din = DeviceBuffer.from_host(np.asfortranarray(inp).ravel(order="K")); rec = DeviceBuffer(T * N * 13 * 8)
run.launch(din.ptr, rec.ptr, None); hm.engine.check(hm.engine.lib().hb_stream_synchronize(None))
q1 = run.xin.to_host((T, N)); full = rec.to_host((13, N, T), order="F")
print("q plane == q row:", np.array_equal(q1.T, full[names.index("q")]), "q row 3p == 2p:", np.array_equal(full[names.index("q")], g2[names.index("q")]))
o2 = run.out2.to_host((T, N, 2))
print("out2 vs rows:", np.array_equal(o2[:, :, 0].T, full[names.index("s_river")]), np.array_equal(o2[:, :, 1].T, full[names.index("q_routed")]))
