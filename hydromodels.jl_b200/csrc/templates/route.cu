// =============================================================================================
// hb_route — HydroRoute (Jacobi routing) on B200: one launch per time step, one thread per node.
//
// Reference semantics (/root/reference/src/route.jl:367-404): at every step each
// node's outflow is computed from its PREVIOUS state, pushed one hop downstream by `aggr_func`
// (D8 shift-and-add, src/route.jl:342-365, or adjacency' * q, :336-340) and added to the
// downstream node's increment in the SAME step:
//     q_out = rflux(x_t, u_prev, p);  dS = g(x_t, u_prev, q_out, p);  inflow_n = sum_{m -> n} q_out_m
//     u <- max(min_value, u_prev + (dS + inflow));   row(t) = rflux(x_t, u, p)   (post-update)
// There is no intra-step dependency between nodes, only a barrier between steps, so step t is one
// kernel:  gather the upstream outflows of step t (written by the previous launch) through a CSR
// upstream list, update the state, write the two result rows INTO the model's [T][N][R] records
// (the per-node stepper left those two rows blank), and already compute this node's outflow for
// step t+1 (it depends only on the node's own new state and next input) into the other half of a
// ping-pong buffer.  The push of the reference (scatter to 8 padded arrays) becomes a gather: no
// atomics, deterministic summation order (upstream nodes sorted by D8 code, the order in which the
// reference adds its eight shifted arrays).
//
// TEMPLATE: //@@HB:<SECTION> markers are replaced by host-generated text (lowering.lower_route).
// =============================================================================================

typedef long long i64;
typedef unsigned int u32;

//@@HB:DEFINES
// (generated) HB_RV (route inputs) HB_NP HB_BLOCK

#define HB_ARR(n) ((n) > 0 ? (n) : 1)

struct HbRouteArgs {
    double *rec;            // [T][N][R] model records (inputs are read from, results written into)
    const double *params;
    const int *htypes;
    const double *s_prev;   // [N] state before this step
    double *s_next;         // [N] state after this step
    const double *q_cur;    // [N] outflows of this step (from the previous launch)
    double *q_next;         // [N] outflows of the next step
    const int *up_ptr;      // [N+1] CSR of upstream nodes
    const int *up_idx;
    const double *xin;      // planes mode (non-NULL): route inputs as a compact [T][N][HB_RV] array instead of record rows
    double *out2;           // planes mode (non-NULL): [T][N][2] = {state, outflow} instead of the two record rows — both fully
                            //   coalesced; the records are then written ONCE by a second stepper pass that takes out2 as its
                            //   auxiliary input (no 8-byte patches into 104-byte records: ncu r01 saw 312 B per cell-step there)
    i64 N;
    i64 t;                  // 0-based step; t == -1: prologue launch (only q_next for step 0)
    i64 T;
    double min_value;
    int R;                  // rows per record
    int s_row;              // record row of the route state
    int o_row;              // record row of the route output
    int in_row[HB_ARR(HB_RV)];
    int p_off[HB_ARR(HB_NP)];
    int p_mode[HB_ARR(HB_NP)];
};

//@@HB:MATH

//@@HB:HELPERS
// (generated)  __device__ double hb_rflux(const double* hb_x, double hb_s, const double* hb_p);
//              __device__ double hb_rds  (const double* hb_x, double hb_s, double hb_q, const double* hb_p);

extern "C" __global__ void __launch_bounds__(HB_BLOCK) hb_route_step(const __grid_constant__ HbRouteArgs a) {
    const i64 n = (i64)blockIdx.x * HB_BLOCK + threadIdx.x;
    hb_math_init();   // exp table -> shared memory (block barrier inside)
    if (n >= a.N) return;
    double hb_p[HB_ARR(HB_NP)];
#pragma unroll
    for (int j = 0; j < HB_NP; ++j) {
        const int mode = a.p_mode[j];
        i64 idx = a.p_off[j];
        if (mode == 1) idx += n;
        else if (mode >= 2) idx += a.htypes[(i64)(mode - 2) * a.N + n];
        hb_p[j] = a.params[idx];
    }
    double hb_x[HB_ARR(HB_RV)];
    double s_new;
    if (a.t >= 0) {
        double *rec = a.rec + (a.t * a.N + n) * a.R;
        if (a.xin) {
#pragma unroll
            for (int v = 0; v < HB_RV; ++v) hb_x[v] = a.xin[(a.t * a.N + n) * HB_RV + v];
        } else {
#pragma unroll
            for (int v = 0; v < HB_RV; ++v) hb_x[v] = rec[a.in_row[v]];
        }
        const double s_prev = a.s_prev[n];
        const double q_own = a.q_cur[n];
        const double ds = hb_rds(hb_x, s_prev, q_own, hb_p);
        double inflow = 0.0;
        for (int e = a.up_ptr[n]; e < a.up_ptr[n + 1]; ++e) inflow += a.q_cur[a.up_idx[e]];
        s_new = hb_max(a.min_value, s_prev + (ds + inflow));      // tmp_states .+ tmp_inflow_arr, then the clamp
        const double o_new = hb_rflux(hb_x, s_new, hb_p);          // flux_func over post-update states
        if (a.out2) {
            *reinterpret_cast<double2 *>(a.out2 + (a.t * a.N + n) * 2) = make_double2(s_new, o_new);
        } else {
            rec[a.s_row] = s_new;                                  // saved after the update (solve.jl:51)
            rec[a.o_row] = o_new;
        }
        a.s_next[n] = s_new;
    } else {
        s_new = a.s_prev[n];                                       // initial state, used unclamped (solve.jl:47-49)
    }
    if (a.t + 1 < a.T) {
        const double *recn = a.rec + ((a.t + 1) * a.N + n) * a.R;
        if (a.xin) {
#pragma unroll
            for (int v = 0; v < HB_RV; ++v) hb_x[v] = a.xin[((a.t + 1) * a.N + n) * HB_RV + v];
        } else {
#pragma unroll
            for (int v = 0; v < HB_RV; ++v) hb_x[v] = recn[a.in_row[v]];
        }
        a.q_next[n] = hb_rflux(hb_x, s_new, hb_p);
    }
}
