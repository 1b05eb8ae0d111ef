"""TEST-ONLY: run a generated stepper *body* on the CPU to validate the lowering without a GPU.

The CUDA C body that `lowering.lower_stepper` emits (sections PROLOGUE / STEP / HELPERS) is plain
C arithmetic.  This harness splices the same text into a trivial serial C++ loop (one node at a
time, per-thread I/O replaced by array indexing), compiles it with g++ and runs it through
ctypes.  It exercises the lowering (CSE, hoisting, carried state-only values, UH rings, MLP
helpers, row order) and `B200Model.prepare` marshalling in the `-m "not gpu"` suite; the kernel
template, TMA staging and launch logic are only exercised by the `-m gpu` tests.  Nothing under
`hydromodels.jl_b200/` imports this file.
"""
import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_SRC = r'''
#include <cmath>
#include <cstdint>
typedef long long i64;
#define HB_HOST_TEST 1
#define HB_TANGENTS @TANGENTS@
#include "@MATH_HEADER@"
#if HB_TANGENTS > 0
typedef hb_dual treal;
#else
typedef double treal;
#endif
#define __device__
#define __forceinline__ inline
#define __restrict__
#define HB_ARR(n) ((n) > 0 ? (n) : 1)
static const double* hb_nn_host = 0;      /* stands in for the kernel's __constant__ hb_nn_const[] */
#define HB_NNW(i) hb_nn_host[i]
@DEFINES@
@HELPERS@
extern "C" int run(i64 N, i64 T, int shared, const double* in, const double* params, const int* p_off,
                   const int* p_mode, const int* htypes, const double* init, const double* uhw,
                   const double* nn, double minv, double* out, double* sim, double* final_states, double* dsim) {
    const double* sNN = nn; (void)sNN; hb_nn_host = nn;
    for (i64 n = 0; n < N; ++n) {
        treal hb_p[HB_ARR(HB_NP)], hb_u[HB_ARR(HB_S)], hb_uhw[HB_ARR(HB_UHW_TOTAL)], hb_uhh[HB_ARR(HB_UHH_TOTAL)], hb_c[HB_ARR(HB_NC)];
        for (int j = 0; j < HB_NP; ++j) {
            i64 idx = p_off[j];
            if (p_mode[j] == 1) idx += n; else if (p_mode[j] >= 2) idx += htypes[(i64)(p_mode[j] - 2) * N + n];
            hb_p[j] = params[idx];
        }
#if HB_TANGENTS > 0
        { const int tdir[HB_ARR(HB_NP)] = HB_TDIR;
          for (int j = 0; j < HB_NP; ++j) for (int k = 0; k < HB_TANGENTS; ++k) hb_p[j].d[k] = tdir[j] == k ? 1.0 : 0.0; }
#endif
        for (int s = 0; s < HB_S; ++s) hb_u[s] = init ? init[(i64)s * N + n] : 0.0;
#if HB_TANGENTS > 0
        { const int sdir[HB_ARR(HB_S)] = HB_TSDIR;
          for (int s = 0; s < HB_S; ++s) for (int k = 0; k < HB_TANGENTS; ++k) hb_u[s].d[k] = sdir[s] == k ? 1.0 : 0.0; }
#endif
        for (int k = 0; k < HB_UHW_TOTAL; ++k) hb_uhw[k] = uhw[(i64)k * N + n];
        for (int k = 0; k < HB_UHH_TOTAL; ++k) hb_uhh[k] = 0.0;
        const double hb_minv = minv;
        (void)hb_p; (void)hb_u; (void)hb_uhw; (void)hb_uhh; (void)hb_c; (void)hb_minv;
@PROLOGUE@
        for (i64 t = 0; t < T; ++t) {
            double hb_x[HB_ARR(HB_V)], hb_out[HB_ARR(HB_R)];
            treal hb_sim = 0.0;
            for (int v = 0; v < HB_V; ++v) hb_x[v] = shared ? in[t * HB_V + v] : in[(t * N + n) * HB_V + v];
            (void)hb_out; (void)hb_sim;
@STEP@
            for (int r = 0; r < HB_R; ++r) out[(t * N + n) * HB_R + r] = hb_out[r];
            if (sim) sim[t * N + n] = hb_val(hb_sim);
#if HB_TANGENTS > 0
            if (dsim) for (int k = 0; k < HB_TANGENTS; ++k) dsim[((i64)k * T + t) * N + n] = hb_sim.d[k];
#endif
        }
        if (final_states) for (int s = 0; s < HB_S; ++s) final_states[(i64)s * N + n] = hb_val(hb_u[s]);
    }
    return 0;
}
'''


def _sections(body: str):
    secs, cur = {}, None
    for line in body.splitlines():
        if line.startswith("//@@HB:"):
            cur = line[7:].strip()
            secs.setdefault(cur, [])
        elif cur is not None:
            secs[cur].append(line)
    return {k: "\n".join(v) for k, v in secs.items()}


def run_body_on_cpu(eng, input, params, config=None, initstates=None, rows=None, objective=False,
                    ensemble_members=None, return_final_states=False, uh_on_device=False, uh_kmax=None, tangents=None):
    """Mirror of `B200Model.__call__` that executes the generated body with g++ instead of NVRTC."""
    from hydromodels_jl_b200.components import normalize_config
    from hydromodels_jl_b200.lowering import lower_stepper

    cmin = None
    if isinstance(config, (tuple, list)):                 # per-component configs: floors baked in as literals
        cfgs = [normalize_config(c) for c in config]
        cfg, cmin = cfgs[0], [c.min_value for c in cfgs]
    else:
        cfg = normalize_config(config)
    input = np.asarray(input, dtype=np.float64)
    ensemble = ensemble_members is not None
    N, T = eng._dims(input, ensemble, ensemble_members)
    inp = np.asfortranarray(input)
    probe = lower_stepper(eng.component, rows=rows, objective=objective, fast=eng.fast)
    kmax = uh_kmax or eng.prepare(probe, params, N, ensemble=ensemble)[5]
    prog = lower_stepper(eng.component, rows=rows, objective=objective, uh_kmax=kmax or None, fast=eng.fast, uh_on_device=uh_on_device,
                         component_min_values=cmin, tangents=tangents)
    flat, p_off, p_mode, htypes, uhw, _, nn = eng.prepare(prog, params, N, ensemble=ensemble)
    if uh_on_device and uhw:
        uhw = [np.full((k, N), np.nan) for k in kmax]       # the prologue must overwrite every ordinate
    uhw_flat = np.ascontiguousarray(np.concatenate(uhw, axis=0)) if uhw else np.zeros(1)
    init = eng._initstates(prog, initstates, N)
    R = 0 if objective else len(prog.row_names)
    secs = _sections(prog.body)
    defines = "\n".join([
        f"#define HB_V {len(prog.input_names)}", f"#define HB_R {R}", f"#define HB_S {len(prog.state_names)}",
        f"#define HB_NP {len(prog.param_slots)}", f"#define HB_UHW_TOTAL {sum(prog.uh_kmax)}",
        f"#define HB_UHH_TOTAL {sum(k - 1 for k in prog.uh_kmax)}", secs.get("DEFINES", "")])
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hydromodels.jl_b200", "csrc",
                       "templates", "hb_math.cuh")
    D = len(tangents) if tangents else 0
    src = (_SRC.replace("@DEFINES@", defines).replace("@TANGENTS@", str(D)).replace("@MATH_HEADER@", hdr).replace("@HELPERS@", secs.get("HELPERS", ""))
           .replace("@PROLOGUE@", secs.get("PROLOGUE", "")).replace("@STEP@", secs.get("STEP", "")))
    tag = hashlib.sha1((src + open(hdr).read()).encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "hb_cpu_harness")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"h{tag}.so")
    if not os.path.exists(so):
        cpp = os.path.join(d, f"h{tag}.cpp")
        open(cpp, "w").write(src)
        # -ffp-contract=off: no FMA contraction beyond the explicit fma() calls; plain IEEE
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", cpp, "-o", so], check=True)
    L = C.CDLL(so)
    out = np.zeros((R, N, T) if eng.multinode or ensemble or (getattr(eng, "any_rank", False) and input.ndim == 3) else (R, T), order="F")
    sim = np.zeros((T, N)) if objective else None
    S = len(prog.state_names)
    final = np.zeros((S, N)) if return_final_states else None
    P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    dsim = np.zeros((D, T, N)) if D else None
    L.run.argtypes = [C.c_int64, C.c_int64, C.c_int] + [C.c_void_p] * 8 + [C.c_double] + [C.c_void_p] * 4
    L.run(N, T, 1 if ensemble else 0, P(inp), P(flat), P(p_off), P(p_mode), P(htypes), P(init), P(uhw_flat),
          P(nn if nn is not None else np.zeros(1)), cfg.min_value, P(out), P(sim), P(final), P(dsim))
    if D:
        return sim.T.copy(), np.ascontiguousarray(dsim.transpose(0, 2, 1))   # [N, T], [D, N, T]
    if objective:
        return sim.T.copy()          # [N, T] simulated last row
    if return_final_states:
        return out, final
    return out
