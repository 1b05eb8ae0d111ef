/*
 * CPU ORACLE (C restatement) — test infrastructure only, never the product path.
 *
 * Plain-C restatement of the reference's hot loops for the forward run of HydroModels.jl
 * (v0.6.2), structured exactly like the Julia code it follows:
 *
 *   ho_run_bucket : HydroBucket functor  (/root/reference/src/bucket.jl:169-256)
 *       pass 1  hydrosolve(MutableSolver) (/root/reference/src/solve.jl:31-55):
 *               for t: du = ode_func(x_t, u, p); u = max(min_value, u + du); states[:, t] = u
 *       pass 2  flux_func over the whole series from the post-update states, then
 *               vcat(states, fluxes)      (/root/reference/src/bucket.jl:190-192,242-243)
 *   ho_uh_conv    : hydro_conv = normalise + causal convolution (/root/reference/src/uh.jl:187-202)
 *   ho_d8_push    : build_aggr_func(flwdir, positions): scatter, eight masked pad_zeros shifts,
 *                   sum, clip, gather    (/root/reference/src/route.jl:342-365)
 *   ho_run_route  : HydroRoute functor, Jacobi routing step (/root/reference/src/route.jl:367-404)
 *
 * The flux arithmetic (HydroModelCore's runtime-generated functions, not vendored) is evaluated
 * by a small vector VM: one instruction = one broadcast operation over a block of nodes, i.e.
 * the same "one temporary per expression, broadcast over nodes" structure as the generated Julia
 * code, with libm functions and NO floating-point contraction (compile with -ffp-contract=off).
 * Blocks of nodes are independent (src/bucket.jl:229-233 broadcasts over nodes with no
 * cross-node term), so `#pragma omp parallel for` over node blocks is the multithreaded CPU
 * baseline bench.py reports ("kind": "port").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HO_BLK 256

enum { HO_ADD, HO_SUB, HO_MUL, HO_DIV, HO_NEG, HO_POW, HO_EXP, HO_LOG, HO_SQRT, HO_ABS, HO_TANH, HO_STEP, HO_MIN, HO_MAX, HO_NN };

typedef struct { int32_t op, dst, a, b; } ho_instr;

/* one dense layer chain: act(W x + b); activation codes 0 identity 1 tanh 2 leakyrelu 3 relu 4 sigmoid */
typedef struct {
    int32_t n_layers;
    int32_t n_in[8], n_out[8], act[8];
    int32_t offset;          /* into the flat nn parameter vector */
    int32_t in_regs[16];
    int32_t out_regs[16];
} ho_nn;

typedef struct {
    int32_t n_instr, n_reg;
    const ho_instr *instr;
    int32_t n_const;
    const int32_t *const_regs;
    const double *const_vals;
    int32_t n_nn;
    const ho_nn *nns;
    const double *nn_params;
} ho_prog;

int ho_version(void) { return 1; }
int ho_max_threads(void) {   /* host cores available to the process */
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

static double ho_act(int a, double z) {
    switch (a) {
        case 1: return tanh(z);
        case 2: return fmax(0.01 * z, z);   /* NNlib.leakyrelu, slope 0.01 */
        case 3: return fmax(0.0, z);
        case 4: return 1.0 / (1.0 + exp(-z));
        default: return z;
    }
}

/* LuxCore.apply(chain, x, ps): per layer weight[out,in] column-major then bias[out] (src/nn.jl:56-58) */
static void ho_eval_nn(const ho_nn *nn, const double *params, double *R, int n) {
    double a[64], z[64];
    for (int i = 0; i < n; ++i) {
        int nin = nn->n_in[0];
        for (int k = 0; k < nin; ++k) a[k] = R[(size_t)nn->in_regs[k] * HO_BLK + i];
        const double *w = params + nn->offset;
        for (int l = 0; l < nn->n_layers; ++l) {
            int ni = nn->n_in[l], no = nn->n_out[l];
            const double *bias = w + (size_t)ni * no;
            for (int o = 0; o < no; ++o) {
                double s = 0.0;
                for (int k = 0; k < ni; ++k) s += w[(size_t)k * no + o] * a[k];
                z[o] = ho_act(nn->act[l], s + bias[o]);
            }
            for (int o = 0; o < no; ++o) a[o] = z[o];
            w += (size_t)ni * no + no;
        }
        int no = nn->n_out[nn->n_layers - 1];
        for (int o = 0; o < no; ++o) R[(size_t)nn->out_regs[o] * HO_BLK + i] = a[o];
    }
}

static void ho_eval(const ho_prog *p, double *R, int n) {
    for (int q = 0; q < p->n_instr; ++q) {
        const ho_instr in = p->instr[q];
        if (in.op == HO_NN) { ho_eval_nn(&p->nns[in.a], p->nn_params, R, n); continue; }
        double *d = R + (size_t)in.dst * HO_BLK;
        const double *a = R + (size_t)in.a * HO_BLK;
        const double *b = R + (size_t)(in.b < 0 ? 0 : in.b) * HO_BLK;
        switch (in.op) {
            case HO_ADD: for (int i = 0; i < n; ++i) d[i] = a[i] + b[i]; break;
            case HO_SUB: for (int i = 0; i < n; ++i) d[i] = a[i] - b[i]; break;
            case HO_MUL: for (int i = 0; i < n; ++i) d[i] = a[i] * b[i]; break;
            case HO_DIV: for (int i = 0; i < n; ++i) d[i] = a[i] / b[i]; break;
            case HO_NEG: for (int i = 0; i < n; ++i) d[i] = -a[i]; break;
            case HO_POW: for (int i = 0; i < n; ++i) d[i] = pow(a[i], b[i]); break;
            case HO_EXP: for (int i = 0; i < n; ++i) d[i] = exp(a[i]); break;
            case HO_LOG: for (int i = 0; i < n; ++i) d[i] = log(a[i]); break;
            case HO_SQRT: for (int i = 0; i < n; ++i) d[i] = sqrt(a[i]); break;
            case HO_ABS: for (int i = 0; i < n; ++i) d[i] = fabs(a[i]); break;
            case HO_TANH: for (int i = 0; i < n; ++i) d[i] = tanh(a[i]); break;
            /* step_func(x) = (tanh(5.0 * x) + 1.0) * 0.5   (src/HydroModels.jl:204) */
            case HO_STEP: for (int i = 0; i < n; ++i) d[i] = (tanh(5.0 * a[i]) + 1.0) * 0.5; break;
            case HO_MIN: for (int i = 0; i < n; ++i) d[i] = a[i] < b[i] ? a[i] : b[i]; break;
            case HO_MAX: for (int i = 0; i < n; ++i) d[i] = a[i] > b[i] ? a[i] : b[i]; break;
            default: break;
        }
    }
}

/*
 * Bucket functor.  Registers: [0,V) inputs, [V,V+S) states, [V+S,V+S+P) params, the rest constants
 * and temporaries.  input [T][N][V], params [P][N] (already p[htypes], src/utils.jl:174-201),
 * init [S][N] (NULL = zeros), out [T][N][S+O].
 */
int ho_run_bucket(int64_t N, int64_t T, int V, int S, int P, int O, const double *input, const double *params,
                  const double *init, double min_value, const ho_prog *prog, const int32_t *flux_regs,
                  const int32_t *dstate_regs, double *out, int n_threads) {
    const int rows = S + O;
    const int64_t nblk = (N + HO_BLK - 1) / HO_BLK;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    int fail = 0;
#pragma omp parallel
    {
        double *R = (double *)malloc(sizeof(double) * (size_t)prog->n_reg * HO_BLK);
        if (!R) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp for schedule(dynamic, 1)
        for (int64_t blk = 0; blk < nblk; ++blk) {
            if (!R) continue;
            const int64_t n0 = blk * HO_BLK;
            const int n = (int)((N - n0) < HO_BLK ? (N - n0) : HO_BLK);
            for (int j = 0; j < P; ++j) memcpy(R + (size_t)(V + S + j) * HO_BLK, params + (size_t)j * N + n0, sizeof(double) * n);
            for (int c = 0; c < prog->n_const; ++c)
                for (int i = 0; i < n; ++i) R[(size_t)prog->const_regs[c] * HO_BLK + i] = prog->const_vals[c];
            /* ---- pass 1: hydrosolve(MutableSolver) ---- */
            for (int s = 0; s < S; ++s)
                for (int i = 0; i < n; ++i) R[(size_t)(V + s) * HO_BLK + i] = init ? init[(size_t)s * N + n0 + i] : 0.0;
            if (S > 0) {
                for (int64_t t = 0; t < T; ++t) {
                    for (int i = 0; i < n; ++i)
                        for (int v = 0; v < V; ++v) R[(size_t)v * HO_BLK + i] = input[((size_t)t * N + n0 + i) * V + v];
                    ho_eval(prog, R, n);
                    for (int s = 0; s < S; ++s) {
                        double *u = R + (size_t)(V + s) * HO_BLK;
                        const double *du = R + (size_t)dstate_regs[s] * HO_BLK;
                        for (int i = 0; i < n; ++i) {
                            const double x = u[i] + du[i];
                            u[i] = x > min_value ? x : min_value;             /* max.(min_val, u .+ du) */
                            out[((size_t)t * N + n0 + i) * rows + s] = u[i];  /* saved after the update */
                        }
                    }
                }
            }
            /* ---- pass 2: all fluxes from the post-update states ---- */
            for (int64_t t = 0; t < T; ++t) {
                for (int i = 0; i < n; ++i) {
                    for (int v = 0; v < V; ++v) R[(size_t)v * HO_BLK + i] = input[((size_t)t * N + n0 + i) * V + v];
                    for (int s = 0; s < S; ++s) R[(size_t)(V + s) * HO_BLK + i] = out[((size_t)t * N + n0 + i) * rows + s];
                }
                ho_eval(prog, R, n);
                for (int o = 0; o < O; ++o) {
                    const double *f = R + (size_t)flux_regs[o] * HO_BLK;
                    for (int i = 0; i < n; ++i) out[((size_t)t * N + n0 + i) * rows + S + o] = f[i];
                }
            }
        }
        free(R);
    }
    return fail;
}

/* hydro_conv for every node: x [T][N], w [K][N] raw S-curve differences (normalised here), K_n per node. */
int ho_uh_conv(int64_t N, int64_t T, int Kmax, const double *x, const double *w, const int32_t *K, double *y, int n_threads) {
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        const int k = K[n];
        if (k <= 0) {                                /* length(uh_weight) == 0 → return input */
            for (int64_t t = 0; t < T; ++t) y[t * N + n] = x[t * N + n];
            continue;
        }
        double sum = 0.0;
        for (int j = 0; j < k; ++j) sum += w[(size_t)j * N + n];
        for (int64_t t = 0; t < T; ++t) {
            double acc = 0.0;
            const int lim = (t + 1) < k ? (int)(t + 1) : k;
            for (int j = 0; j < lim; ++j) acc += (w[(size_t)j * N + n] / sum) * x[(t - j) * N + n];
            y[t * N + n] = acc;
        }
    }
    (void)Kmax;
    return 0;
}

/* grid_routing(input): scatter to [R][C], for each D8 code add the masked array shifted by the
 * code's pad offsets into a padded [R+2][C+2] array, clip [1..R][1..C], gather (src/route.jl:349-364). */
int ho_d8_push(int R, int C, const int32_t *flwdir, int64_t N, const int32_t *rows0, const int32_t *cols0, const double *q,
               double *inflow, double *work /* (R*C + (R+2)*(C+2)) doubles */) {
    static const int codes[8] = {1, 2, 4, 8, 16, 32, 64, 128};
    static const int pad_r[8] = {1, 2, 2, 2, 1, 0, 0, 0};   /* rows padded before */
    static const int pad_c[8] = {2, 2, 1, 0, 0, 0, 1, 2};   /* cols padded before */
    double *arr = work, *routed = work + (size_t)R * C;
    memset(arr, 0, sizeof(double) * (size_t)R * C);
    memset(routed, 0, sizeof(double) * (size_t)(R + 2) * (C + 2));
    for (int64_t n = 0; n < N; ++n) arr[(size_t)rows0[n] * C + cols0[n]] = q[n];
    for (int k = 0; k < 8; ++k)
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c)
                if (flwdir[(size_t)r * C + c] == codes[k])
                    routed[(size_t)(r + pad_r[k]) * (C + 2) + (c + pad_c[k])] += arr[(size_t)r * C + c];
    for (int64_t n = 0; n < N; ++n) inflow[n] = routed[(size_t)(rows0[n] + 1) * (C + 2) + (cols0[n] + 1)];
    return 0;
}

/* HydroRoute functor, one state / one outflow (registers as in ho_run_bucket with S = 1).
 * down[n] = 0-based downstream node or -1 (graph variant: adjacency' * q); when flwdir != NULL the
 * literal D8 push above is used instead. */
int ho_run_route(int64_t N, int64_t T, int V, int P, const double *input, const double *params, const double *init,
                 double min_value, const ho_prog *prog, int outflow_reg, int dstate_reg, int R, int C,
                 const int32_t *flwdir, const int32_t *rows0, const int32_t *cols0, const int32_t *down, double *out) {
    if (N > HO_BLK * 4096) return 2;
    const int64_t nblk = (N + HO_BLK - 1) / HO_BLK;
    double *Rg = (double *)malloc(sizeof(double) * (size_t)prog->n_reg * HO_BLK);
    double *u = (double *)calloc((size_t)N, sizeof(double));
    double *q = (double *)malloc(sizeof(double) * (size_t)N);
    double *ds = (double *)malloc(sizeof(double) * (size_t)N);
    double *inflow = (double *)malloc(sizeof(double) * (size_t)N);
    double *work = flwdir ? (double *)malloc(sizeof(double) * ((size_t)R * C + (size_t)(R + 2) * (C + 2))) : NULL;
    if (!Rg || !u || !q || !ds || !inflow || (flwdir && !work)) return 1;
    if (init) memcpy(u, init, sizeof(double) * (size_t)N);
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t t = 0; t < T; ++t) {
            for (int64_t blk = 0; blk < nblk; ++blk) {
                const int64_t n0 = blk * HO_BLK;
                const int n = (int)((N - n0) < HO_BLK ? (N - n0) : HO_BLK);
                for (int j = 0; j < P; ++j) memcpy(Rg + (size_t)(V + 1 + j) * HO_BLK, params + (size_t)j * N + n0, sizeof(double) * n);
                for (int c = 0; c < prog->n_const; ++c)
                    for (int i = 0; i < n; ++i) Rg[(size_t)prog->const_regs[c] * HO_BLK + i] = prog->const_vals[c];
                for (int i = 0; i < n; ++i) {
                    for (int v = 0; v < V; ++v) Rg[(size_t)v * HO_BLK + i] = input[((size_t)t * N + n0 + i) * V + v];
                    Rg[(size_t)V * HO_BLK + i] = pass == 0 ? u[n0 + i] : out[((size_t)t * N + n0 + i) * 2];
                }
                ho_eval(prog, Rg, n);
                for (int i = 0; i < n; ++i) {
                    if (pass == 0) {
                        q[n0 + i] = Rg[(size_t)outflow_reg * HO_BLK + i];
                        ds[n0 + i] = Rg[(size_t)dstate_reg * HO_BLK + i];
                    } else {
                        out[((size_t)t * N + n0 + i) * 2 + 1] = Rg[(size_t)outflow_reg * HO_BLK + i];
                    }
                }
            }
            if (pass == 0) {
                if (flwdir) ho_d8_push(R, C, flwdir, N, rows0, cols0, q, inflow, work);
                else {
                    memset(inflow, 0, sizeof(double) * (size_t)N);
                    for (int64_t n = 0; n < N; ++n) if (down[n] >= 0) inflow[down[n]] += q[n];
                }
                for (int64_t n = 0; n < N; ++n) {
                    const double x = u[n] + (ds[n] + inflow[n]);      /* tmp_states .+ tmp_inflow_arr */
                    u[n] = x > min_value ? x : min_value;
                    out[((size_t)t * N + n) * 2] = u[n];
                }
            }
        }
    }
    free(Rg); free(u); free(q); free(ds); free(inflow); free(work);
    return 0;
}

/* _collect_rows / vcat of 1-row views (/root/reference/src/model.jl:154-192): dst[i][j] =
 * src_j[i][row_j] for i over all (node, time) pairs; src_j has rows_j variables per pair. */
int ho_collect_rows(int64_t NT, int k, const double *const *src, const int32_t *src_rows, const int32_t *row, double *dst,
                    int n_threads) {
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < NT; ++i)
        for (int j = 0; j < k; ++j) dst[i * k + j] = src[j][i * src_rows[j] + row[j]];
    return 0;
}
