// =============================================================================================
// hb_adjoint — reverse-mode gradient of the objective of a per-node bucket model (with neural fluxes) on B200.
//
// What the reference obtains with Zygote through its ImmutableSolver (/root/reference/README.md:161-187,
// src/solve.jl:62-79; training loop docs/src/tutorials/neuralnetwork_embeding.md): d loss / d (parameters, initial states,
// neural-network weights) for   loss_n = scale_n * sum_{t >= warm_up} (sim_n(t) - obs_n(t))^2   (SSE, MSE, -NSE + 1:
// every metric that is linear in the squared errors).  The forward recurrence is the explicit stepper of solve.jl:47-52,
//     u_t = max(min_value, u_{t-1} + du(u_{t-1}, x_t, p)),      sim_t = g(u_t, x_t, p),
// whose states the forward kernel stored ([T][N][S]).  One thread per node walks t = T .. 1 with the adjoint lambda of its
// S states in registers:
//     recompute the step from u_{t-1} (stored) and x_t      -> every intermediate value, sim_t
//     seed   d loss / d sim_t = 2 scale (sim_t - obs_t)      (0 before warm_up)
//     sweep  the step's vector-Jacobian products in reverse  -> lambda_{t-1}, parameter gradients, and through
//            hb_mlp_bwd_k the gradients of the neural fluxes' inputs and WEIGHTS
// No tape: the step is cheap to recompute and the only stored quantity is the state trajectory the forward run writes anyway.
//
// Weight gradients are shared by all nodes: inside a warp the 32 nodes' products delta_o * a_i of 32 consecutive weights go
// through a butterfly of shuffles that leaves lane l with the warp total of weight 32 g + l (hb_transpose_reduce32), which
// the lane adds to its accumulator register hb_gw[g] for the whole run; at the end every warp adds its accumulators to the
// global gradient with one atomicAdd per weight.  Lanes past the last node shadow it with zero seeds (they must take part
// in the shuffles) and store nothing.
//
// TEMPLATE: //@@HB:<SECTION> markers are replaced by host-generated text (lowering._emit_adjoint).
// =============================================================================================

typedef long long i64;
typedef unsigned int u32;

//@@HB:DEFINES
// (generated) HB_V HB_S HB_NP HB_NN_WORDS HB_BLOCK HB_NG HB_GOFF_k

#define HB_ARR(n) ((n) > 0 ? (n) : 1)
typedef double real;

struct HbAdjArgs {
    const double *in;         // [T][N][V] forcing
    const double *traj;       // [T][N][S] post-update states of every step (forward run with rows = states)
    const double *params;
    const int *htypes;
    const double *init;       // [S][N] initial states or NULL (zeros)
    const double *target;     // [T] or [T][N]
    const double *scale;      // [N] weight of the squared errors of node n (1: SSE, 1/n: MSE, 1/sum (obs - mean obs)^2: NSE)
    double *grad_params;      // [NP][N]   d loss_n / d p_j(n)
    double *grad_init;        // [S][N]    d loss_n / d u_s(0)
    double *grad_nn;          // [NN_WORDS] d sum_n loss_n / d weights (atomicAdd: zero it before the launch)
    double *loss;             // [N]       scale_n * sum of squared errors
    i64 N;
    i64 T;
    double min_value;
    int warm_up;              // 1-based first step that counts
    int target_per_node;
    int p_off[HB_ARR(HB_NP)];
    int p_mode[HB_ARR(HB_NP)];
};

//@@HB:MATH

#if HB_NN_WORDS > 0
__constant__ double hb_nn_const[HB_NN_WORDS];
#define HB_NNW(i) (hb_nn_const[i])
#endif

// 32 values per lane, 32 lanes: returns in lane l the sum over all lanes of p[l] (31 shuffle-add steps instead of 32 x 5)
__device__ __forceinline__ double hb_transpose_reduce32(double (&p)[32], const int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const double send = up ? p[j] : p[j + s];
            const double keep = up ? p[j + s] : p[j];
            p[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return p[0];
}

//@@HB:HELPERS

extern "C" __global__ void __launch_bounds__(HB_BLOCK) hb_adjoint(const __grid_constant__ HbAdjArgs a) {
    const int lane = threadIdx.x & 31;
    const i64 n = (i64)blockIdx.x * HB_BLOCK + threadIdx.x;
    hb_math_init();   // exp table -> shared memory (block barrier inside)
    const i64 n0w = n - lane;
    if (n0w >= a.N) return;                       // whole warp past the last node
    const bool valid = n < a.N;
    const i64 nc = valid ? n : a.N - 1;
    const double *sNN = nullptr;
    (void)sNN; (void)lane;

    double hb_p[HB_ARR(HB_NP)], hb_gp[HB_ARR(HB_NP)];
#pragma unroll
    for (int j = 0; j < HB_NP; ++j) {
        const int mode = a.p_mode[j];
        i64 idx = a.p_off[j];
        if (mode == 1) idx += nc;
        else if (mode >= 2) idx += a.htypes[(i64)(mode - 2) * a.N + nc];
        hb_p[j] = a.params[idx];
        hb_gp[j] = 0.0;
    }
    double hb_lam[HB_ARR(HB_S)];
#pragma unroll
    for (int s = 0; s < HB_S; ++s) hb_lam[s] = 0.0;
    double hb_gw[HB_ARR(HB_NG)];
#pragma unroll
    for (int g = 0; g < HB_NG; ++g) hb_gw[g] = 0.0;
    const double hb_minv = a.min_value;
    const double scale = valid ? a.scale[nc] : 0.0;
    double loss = 0.0;
    (void)hb_p; (void)hb_gp; (void)hb_lam; (void)hb_gw; (void)hb_minv;

    //@@HB:PROLOGUE

#pragma unroll 1
    for (i64 t = a.T - 1; t >= 0; --t) {
        double hb_x[HB_ARR(HB_V)], hb_u[HB_ARR(HB_S)];
#pragma unroll
        for (int v = 0; v < HB_V; ++v) hb_x[v] = a.in[(t * a.N + nc) * HB_V + v];
        // pre-update states of step t: what step t-1 stored, or the (unclamped) initial states (solve.jl:47-49)
#pragma unroll
        for (int s = 0; s < HB_S; ++s)
            hb_u[s] = t > 0 ? a.traj[((t - 1) * a.N + nc) * HB_S + s] : (a.init ? a.init[(i64)s * a.N + nc] : 0.0);
        const double obs = a.target_per_node ? a.target[t * a.N + nc] : a.target[t];
        const bool counts = (t + 1 >= a.warm_up);
        double hb_sim = 0.0, hb_seed = 0.0;
        (void)hb_x; (void)hb_u;

#define HB_SEED_STMT { const double e = hb_sim - obs; if (counts) { hb_seed = 2.0 * scale * e; loss = fma(scale * e, e, loss); } }
        //@@HB:STEP
#undef HB_SEED_STMT
    }

    //@@HB:EPILOGUE

    // ---- results ------------------------------------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int j = 0; j < HB_NP; ++j) a.grad_params[(i64)j * a.N + n] = hb_gp[j];
#pragma unroll
        for (int s = 0; s < HB_S; ++s) a.grad_init[(i64)s * a.N + n] = hb_lam[s];
        a.loss[n] = loss;
    }
#if HB_NN_WORDS > 0
#pragma unroll
    for (int g = 0; g < HB_NG; ++g) {
        const int w = HB_GROUP_WEIGHT(g, lane);       // flat index of the weight this lane holds for group g, or -1
        if (w >= 0) atomicAdd(a.grad_nn + w, hb_gw[g]);
    }
#endif
}
