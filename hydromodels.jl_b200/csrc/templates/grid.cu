// =============================================================================================
// hb_grid_chunk — fused distributed model on a dense D8 grid: per-cell buckets + HydroRoute in ONE pass.
//
// The generic path for models that end in a HydroRoute (stepper.cu over all T, then route.cu once per
// step) writes the [T][N][R] records twice: the stepper leaves the two route rows blank and the
// route kernel patches 8-byte values into 104-byte records, which HBM turns into 64-byte
// read-modify-writes (ncu r01: 312 B per cell-step for the route alone).  On a dense grid
// (positions = every cell, row-major) this kernel removes the second pass: a thread block owns a
// (HB_BW-2H) x (HB_BH-2H) tile of cells plus a halo of H = HB_TC cells, keeps bucket and river states in
// registers for HB_TC consecutive time steps, exchanges the Jacobi outflows of every step through
// shared memory, and writes each finished record exactly once.  Halo cells are recomputed redundantly
// (buckets included): their river state goes stale from the region border inwards by one cell per step,
// which is exactly the halo depth, so every tile cell stays exact for the whole chunk.  States travel
// between chunk launches through ping-pong arrays in global memory.
//
// Semantics per step = the reference's (src/model.jl:273-309 sequencing; src/solve.jl:47-52 update;
// src/route.jl:387-403 route step; src/route.jl:342-365 D8 push as an 8-neighbour gather in the
// reference's code order 1,2,4,...,128).
// TEMPLATE: //@@HB:<SECTION> markers are replaced by host-generated text (lowering.lower_grid).
// =============================================================================================

typedef long long i64;
typedef unsigned int u32;

//@@HB:DEFINES
// (generated) HB_V HB_R HB_S HB_NP HB_NPS (params of the per-cell part; the route's follow) HB_NC HB_RV
//             HB_SROW HB_OROW HB_TC HB_BW HB_BH   and   HB_ROUTE_IN(v) -> hb_out[row of route input v]

#define HB_H HB_TC
#define HB_BLOCK (HB_BW * HB_BH)
#define HB_ARR(n) ((n) > 0 ? (n) : 1)
typedef double real;

struct HbGridArgs {
    const double *in;        // [T][N][V] forcing, node n = r * C + c
    double *out;             // [T][N][R] records
    const double *params;
    const int *htypes;
    const double *su_in;     // [S][N] bucket states at the start of the chunk (NULL: zeros)
    double *su_out;          // [S][N] ... at the end
    const double *rs_in;     // [N] river state at the start of the chunk (NULL: zeros)
    double *rs_out;
    const unsigned char *flw; // [R][C] D8 codes (anything else: sink)
    i64 N;
    i64 t0;                  // first step of this chunk
    int nsteps;              // steps in this chunk (<= HB_TC)
    int rows, cols;          // grid shape
    double min_value;
    int p_off[HB_ARR(HB_NP)];
    int p_mode[HB_ARR(HB_NP)];
};

//@@HB:MATH

//@@HB:HELPERS

extern "C" __global__ void __launch_bounds__(HB_BLOCK, 1) hb_grid_chunk(const __grid_constant__ HbGridArgs a) {
    extern __shared__ __align__(16) unsigned char hb_smem[];
    double *qs = reinterpret_cast<double *>(hb_smem);                          // [BH][BW] outflows of this step
    double *stage = qs + HB_BLOCK;                                             // [warps][32 * R] record staging
    unsigned char *fs = reinterpret_cast<unsigned char *>(stage + (HB_BLOCK / 32) * 32 * HB_R);   // [BH][BW] D8 codes
    const int tid = threadIdx.x;
    const int lx = tid % HB_BW, ly = tid / HB_BW;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx0 = (int)blockIdx.x * (HB_BW - 2 * HB_H) - HB_H;
    const int ty0 = (int)blockIdx.y * (HB_BH - 2 * HB_H) - HB_H;
    const int r = ty0 + ly, c = tx0 + lx;
    const bool in_grid = r >= 0 && r < a.rows && c >= 0 && c < a.cols;
    const bool core = in_grid && lx >= HB_H && lx < HB_BW - HB_H && ly >= HB_H && ly < HB_BH - HB_H;
    const i64 n = in_grid ? (i64)r * a.cols + c : 0;
    hb_math_init();

    fs[tid] = in_grid ? a.flw[n] : (unsigned char)0;
    real hb_p[HB_ARR(HB_NP)];
#pragma unroll
    for (int j = 0; j < HB_NP; ++j) {
        const int mode = a.p_mode[j];
        i64 idx = a.p_off[j];
        if (mode == 1) idx += n;
        else if (mode >= 2) idx += a.htypes[(i64)(mode - 2) * a.N + n];
        hb_p[j] = a.params[idx];
    }
    real hb_u[HB_ARR(HB_S)];
#pragma unroll
    for (int s = 0; s < HB_S; ++s) hb_u[s] = a.su_in ? a.su_in[(i64)s * a.N + n] : 0.0;
    double rs = a.rs_in ? a.rs_in[n] : 0.0;
    real hb_c[HB_ARR(HB_NC)];
    real hb_uhw[1], hb_uhh[1];
    const real hb_minv = a.min_value;
    const double *sNN = nullptr;
    (void)hb_p; (void)hb_u; (void)hb_c; (void)hb_uhw; (void)hb_uhh; (void)hb_minv; (void)sNN; (void)lane;

    //@@HB:PROLOGUE

    // D8: the neighbour at (ly - dr, lx - dc) drains into this cell when its code is `code`
    const int kCode[8] = {1, 2, 4, 8, 16, 32, 64, 128};
    const int kDr[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const int kDc[8] = {1, 1, 0, -1, -1, -1, 0, 1};

    for (int tt = 0; tt < a.nsteps; ++tt) {
        const i64 t = a.t0 + tt;
        real hb_x[HB_ARR(HB_V)];
        real hb_out[HB_ARR(HB_R)];
#pragma unroll
        for (int rr = 0; rr < HB_R; ++rr) hb_out[rr] = 0.0;
        double q_out = 0.0;
        real xr[HB_ARR(HB_RV)];
        if (in_grid) {
            const double *gx = a.in + (t * a.N + n) * HB_V;
#pragma unroll
            for (int v = 0; v < HB_V; ++v) hb_x[v] = gx[v];

            //@@HB:STEP

#pragma unroll
            for (int v = 0; v < HB_RV; ++v) xr[v] = HB_ROUTE_IN(v);
            q_out = hb_rflux(xr, rs, hb_p + HB_NPS);                       // outflow from the PREVIOUS river state
        }
        qs[tid] = q_out;
        __syncthreads();
        if (in_grid) {
            double inflow = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {                                   // the reference's summation order
                const int ny = ly - kDr[k], nx = lx - kDc[k];
                if (ny >= 0 && ny < HB_BH && nx >= 0 && nx < HB_BW && fs[ny * HB_BW + nx] == kCode[k])
                    inflow += qs[ny * HB_BW + nx];
            }
            const double ds = hb_rds(xr, rs, q_out, hb_p + HB_NPS);
            rs = hb_max(a.min_value, rs + (ds + inflow));
            hb_out[HB_SROW] = rs;
            hb_out[HB_OROW] = hb_rflux(xr, rs, hb_p + HB_NPS);
        }
        // ---- records of the tile cells of this region row (= this warp when HB_BW == 32) ------------
        {
            double *st = stage + warp * 32 * HB_R;
#pragma unroll
            for (int rr = 0; rr < HB_R; ++rr) st[lane * HB_R + rr] = hb_out[rr];
            __syncwarp();
            const int wy = (warp * 32) / HB_BW;                             // region row of this warp
            const int wx0 = (warp * 32) % HB_BW;                            // first region column of this warp
            const int gr = ty0 + wy;
            if (wy >= HB_H && wy < HB_BH - HB_H && gr >= 0 && gr < a.rows) {
                int lo = HB_H - wx0, hi = HB_BW - HB_H - wx0;               // tile columns, as lanes of this warp
                if (lo < 0) lo = 0;
                if (hi > 32) hi = 32;
                if (tx0 + wx0 + lo < 0) lo = -(tx0 + wx0);
                if (tx0 + wx0 + hi > a.cols) hi = a.cols - (tx0 + wx0);
                if (hi > lo) {
                    double *dst = a.out + (t * a.N + (i64)gr * a.cols + (tx0 + wx0 + lo)) * HB_R;
                    const double *src = st + lo * HB_R;
                    const int cnt = (hi - lo) * HB_R;
                    for (int i = lane; i < cnt; i += 32) dst[i] = src[i];
                }
            }
        }
        __syncthreads();                                                    // qs / stage are reused next step
    }
    if (core) {
#pragma unroll
        for (int s = 0; s < HB_S; ++s) a.su_out[(i64)s * a.N + n] = hb_u[s];
        a.rs_out[n] = rs;
    }
}
