// =============================================================================================
// hb_grid_wave — distributed model on a dense D8 grid in ONE pass: per-cell buckets + HydroRoute,
// every record written exactly once, no halo recomputation and no block-wide barrier in the time loop.
//
// Reference semantics (src/model.jl:273-309 sequencing; src/solve.jl:47-52 update; src/route.jl:387-403
// Jacobi route step; src/route.jl:342-365 D8 push, here an 8-neighbour gather in the reference's code
// order 1,2,4,...,128):
//     q_out(t,n) = rflux(x_route(t,n), s(t-1,n))                 x_route = rows the cell's buckets produced at t
//     s(t,n)     = max(min_value, s(t-1,n) + (ds(x_route, s(t-1,n), q_out) + sum_{m drains into n} q_out(t,m)))
//     rows(t,n)  = [bucket rows..., s(t,n), rflux(x_route, s(t,n))]
// The only coupling between cells is one hop per step.  The generic path (stepper.cu for all T, then
// route.cu once per step) pays for it with a second pass that patches 8-byte values into 104-byte
// records (ncu r01: 312 B per cell-step for the route alone); the tile+halo kernel (grid.cu) pays with
// redundant buckets and two block barriers per step.  Here the coupling is a dataflow between WARPS:
//
//   * thread per cell, warp = 32 consecutive cells of the row-major grid, bucket and river state in
//     registers for the T_c steps of one launch (states cross launches through [S][N] arrays);
//   * each step a cell publishes its outflow as ONE 16-byte word {q, step tag} into a two-slot (step parity)
//     array in global memory (L2-resident); to finish the step it polls the words of the cells that drain
//     into it — the poll IS the gather — and of the cell it drains into (back-pressure: that neighbour's
//     tag t+1 proves it has consumed this cell's step t-1 value, so the slot may be overwritten).  Value and
//     tag travel in one aligned 128-bit store / load (single L2 transaction, never torn — the same property
//     CUB's decoupled look-back relies on for 64-bit payloads), so no fence is needed on either side: the
//     first version (separate release/acquire step counters) spent 60 % of its warp time in
//     MEMBAR.GPU/ERRBAR and CCTL.IVALL (profiles/r02_wave_flags_ncu.txt);
//   * software pipelining: the bucket part of step t+1 (independent of the river) is computed between
//     publishing step t and polling for it, so the L2 round trip is hidden behind FP64 work;
//   * forcing arrives through the same per-warp TMA ring as stepper.cu, finished records leave through
//     per-warp TMA bulk stores (3 tiles: one being filled by the next step's buckets, one waiting for its
//     river rows, one draining).
//
// Forward progress: blocks take their logical index from an atomic ticket, so logical order = dispatch
// order; a cell only ever waits on cells of the rows above / below, i.e. on logical blocks at most
// (cols/HB_BLOCK + 1) behind or ahead, at most T_c steps deep.  The host clamps T_c so that this dependency
// cone fits well inside the number of co-resident blocks.  Every poll loop is bounded: on expiry the lane
// raises the error word and stops waiting (the host turns it into an error; results are then invalid).
//
// TEMPLATE: //@@HB:<SECTION> markers are replaced by host-generated text (lowering.lower_grid).
// =============================================================================================

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

//@@HB:DEFINES
// (generated) HB_V HB_R HB_S HB_NP HB_NPS HB_NC HB_RV HB_SROW HB_OROW HB_BLOCK HB_STAGES HB_MINBLOCKS
//             and   HB_ROUTE_IN(v) -> hb_out[row of route input v]

#ifndef HB_OBUF
#define HB_OBUF 3                   // record tiles per warp (2: the store of step i-2 must have drained before bucket i)
#endif
#define HB_WARPS (HB_BLOCK / 32)
#define HB_ARR(n) ((n) > 0 ? (n) : 1)
typedef double real;

struct HbWaveArgs {
    const double *in;         // [T][N][V] forcing, node n = r * cols + c
    double *out;              // [T][N][R] records
    const double *params;
    const int *htypes;
    const double *su_in;      // [S][N] bucket states at the start of the chunk (NULL: zeros)
    double *su_out;           // [S][N] ... at the end
    const double *rs_in;      // [N] river state at the start of the chunk (NULL: zeros)
    double *rs_out;
    const unsigned char *flw; // [rows][cols] D8 codes (anything else: sink)
    double *qbuf;             // [HB_QSLOTS][N] words {q, tag}: outflow of step t in slot t % HB_QSLOTS, tag = t + 1
    int *unused;
    u32 *ticket;              // [0] block ticket (monotone over the run), [1] error flag
    i64 N;
    i64 t0;                   // first step of this chunk
    int out_r0, out_nr;       // records are written for the cells of rows [out_r0, out_r0 + out_nr) x columns
    int out_c0, out_nc;       //   [out_c0, out_c0 + out_nc) only, as out[T][out_nr][out_nc][R]: a strip / panel of a multi-GPU run
                              //   computes ghost rows / columns whose records belong to the neighbouring rank
    u32 ticket_base;          // value of the ticket when this launch started
    int nsteps;               // steps in this chunk
    int rows, cols;
    int tma_in_ok, tma_out_ok;
    double min_value;
    // pitched addressing: a sub-grid (column panel) run in place on the arrays of a wider grid.  Forcing of local cell (r, c) at
    // step t: in[(t * in_tstride + r * in_pitch + in_col0 + c) * V]; record of an owned cell:
    // out[(t * out_tstride + (r - out_r0) * out_pitch + out_base + (c - out_c0)) * R].  Compact sub-grids: in_pitch = cols,
    // in_col0 = 0, in_tstride = N; out_pitch = out_nc, out_base = 0, out_tstride = out_nr * out_nc.
    i64 in_tstride, out_tstride;
    int in_pitch, in_col0, out_pitch, out_base;
    int p_off[HB_ARR(HB_NP)];
    int p_mode[HB_ARR(HB_NP)];
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ u32 hb_smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hb_mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void hb_mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hb_mbar_wait(u32 bar, u32 parity) {
    u32 ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void hb_bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void hb_bulk_s2g(void *dst, u32 src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void hb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void hb_bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void hb_bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void hb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// inter-cell dataflow: {value, tag} as one 128-bit word; relaxed GPU-scope accesses are served by L2, the
// point of coherence between SMs
__device__ __forceinline__ void hb_publish(double *slot, double v, i64 tag) {
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(slot), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ i64 hb_peek(const double *slot, double &v) {
    i64 bits, tag;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(bits), "=l"(tag) : "l"(slot) : "memory");
    v = __longlong_as_double(bits);
    return tag;
}

//@@HB:MATH

//@@HB:HELPERS

#ifndef HB_QSLOTS
#define HB_QSLOTS 2               // outflow words per cell (ring over the step index); > 2 relaxes the back-pressure on the downstream cell
#endif
#ifndef HB_SPIN_LIMIT
#define HB_SPIN_LIMIT (1 << 22)   // polls (with back-off ~ 0.1-0.2 us each) before a warp gives up
#endif

extern "C" __global__ void __launch_bounds__(HB_BLOCK, HB_MINBLOCKS) hb_grid_wave(const __grid_constant__ HbWaveArgs a) {
    extern __shared__ __align__(128) unsigned char hb_smem[];
    __shared__ u32 s_ticket;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u) - a.ticket_base;
    hb_math_init();   // exp table -> shared memory; its block barrier also publishes s_ticket
    const i64 n = (i64)s_ticket * HB_BLOCK + threadIdx.x;
    const i64 n0w = n - lane;
    if (n0w >= a.N) return;                     // whole warp past the grid (warp-uniform)
    const bool valid = n < a.N;
    const bool full_warp = (n0w + 32) <= a.N;
    const i64 nc = valid ? n : a.N - 1;         // lanes past the last cell shadow it and never store

    constexpr int kInStage = 32 * HB_V * 8;
    constexpr int kInBytes = ((HB_STAGES * kInStage + 127) / 128) * 128;
    constexpr int kOutTile = 32 * HB_R * 8;
    constexpr int kOutBytes = ((HB_OBUF * kOutTile + 127) / 128) * 128;
    unsigned char *warp_base = hb_smem + (size_t)warp * (kInBytes + kOutBytes);
    real *sIn = reinterpret_cast<real *>(warp_base);
    real *sOut = reinterpret_cast<real *>(warp_base + kInBytes);
    u64 *bars = reinterpret_cast<u64 *>(hb_smem + (size_t)HB_WARPS * (kInBytes + kOutBytes)) + warp * HB_STAGES;

    const i64 in_cell = (nc / a.cols) * a.in_pitch + a.in_col0 + (nc % a.cols);        // this cell / the warp's first cell in the forcing array
    const i64 in_cell_w = (n0w / a.cols) * a.in_pitch + a.in_col0 + (n0w % a.cols);
    const bool tma_in = a.tma_in_ok && full_warp && (n0w % a.cols) + 32 <= a.cols && ((in_cell_w * HB_V * 8) & 15) == 0;
    // position of this cell / of the warp's first cell in the output window
    const int o_r = (int)(nc / a.cols) - a.out_r0, o_c = (int)(nc % a.cols) - a.out_c0;
    const int ow_r = (int)(n0w / a.cols) - a.out_r0, ow_c = (int)(n0w % a.cols) - a.out_c0;
    const bool out_valid = valid && o_r >= 0 && o_r < a.out_nr && o_c >= 0 && o_c < a.out_nc;
    const i64 o_rel = (i64)o_r * a.out_pitch + a.out_base + o_c, ow_rel = (i64)ow_r * a.out_pitch + a.out_base + ow_c;
    // the bulk store needs the warp's 32 cells in one row, inside the window, at a 16-byte aligned record offset
    const bool tma_out = a.tma_out_ok && full_warp && (n0w % a.cols) + 32 <= a.cols && ow_r >= 0 && ow_r < a.out_nr && ow_c >= 0 &&
                         ow_c + 32 <= a.out_nc && ((ow_rel * HB_R * 8) & 15) == 0;

    // ---- per-cell constants ---------------------------------------------------------------------
    real hb_p[HB_ARR(HB_NP)];
#pragma unroll
    for (int j = 0; j < HB_NP; ++j) {
        const int mode = a.p_mode[j];
        i64 idx = a.p_off[j];
        if (mode == 1) idx += nc;
        else if (mode >= 2) idx += a.htypes[(i64)(mode - 2) * a.N + nc];
        hb_p[j] = a.params[idx];
    }
    real hb_u[HB_ARR(HB_S)];
#pragma unroll
    for (int s = 0; s < HB_S; ++s) hb_u[s] = a.su_in ? a.su_in[(i64)s * a.N + nc] : 0.0;
    double rs = a.rs_in ? a.rs_in[nc] : 0.0;
    real hb_c[HB_ARR(HB_NC)];
    real hb_uhw[1], hb_uhh[1];
    const real hb_minv = a.min_value;
    const double *sNN = nullptr;
    (void)hb_p; (void)hb_u; (void)hb_c; (void)hb_uhw; (void)hb_uhh; (void)hb_minv; (void)sNN;

    // D8: the neighbour at (r - dr, c - dc) drains into this cell when its code is kCode[k]
    const int kCode[8] = {1, 2, 4, 8, 16, 32, 64, 128};
    const int kDr[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const int kDc[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int cr = (int)(nc / a.cols), cc = (int)(nc % a.cols);
    u32 inmask = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ny = cr - kDr[k], nx = cc - kDc[k];
        if (ny >= 0 && ny < a.rows && nx >= 0 && nx < a.cols && a.flw[(i64)ny * a.cols + nx] == kCode[k]) inmask |= 1u << k;
    }
    // bit 8: the cell this one drains into (polled for back-pressure only)
    i64 down_off = 0;
    {
        const int code = a.flw[nc];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ny = cr + kDr[k], nx = cc + kDc[k];
            if (code == kCode[k] && ny >= 0 && ny < a.rows && nx >= 0 && nx < a.cols) {
                inmask |= 1u << 8;
                down_off = (i64)kDr[k] * a.cols + kDc[k];
            }
        }
    }
    if (!valid) inmask = 0;
    bool gave_up = false;
#ifndef HB_PREFETCH_POLL
#define HB_PREFETCH_POLL 0       // 1: request the first two upstream words and the downstream word of step i-1 BEFORE the bucket
#endif                           //    part of step i and consume them after it (the L2 round trip of the common first poll round
                                 //    overlaps the bucket arithmetic; cells with more inflows finish in the normal poll loop)
#if HB_PREFETCH_POLL
    int pk0 = -1, pk1 = -1, poff0 = 0, poff1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if ((inmask >> k) & 1u) {
            const int off = 2 * (kDr[k] * a.cols + kDc[k]);
            if (pk0 < 0) { pk0 = k; poff0 = off; }
            else if (pk1 < 0) { pk1 = k; poff1 = off; }
        }
    }
#endif

    //@@HB:PROLOGUE

    const int ns = a.nsteps;
    const i64 in_step = a.in_tstride * HB_V;
    const i64 out_step = a.out_tstride * HB_R;
    const real *in_warp = a.in + (a.t0 * a.in_tstride + in_cell_w) * HB_V;
    const real *gx = a.in + (a.t0 * a.in_tstride + in_cell) * HB_V;
    real *ow_ptr = a.out + a.t0 * out_step + ow_rel * HB_R;
    real *go = a.out + a.t0 * out_step + o_rel * HB_R;

    if (tma_in) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < HB_STAGES; ++s) hb_mbar_init(hb_smem_u32(&bars[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            hb_fence_proxy_async();
#pragma unroll
            for (int s = 0; s < HB_STAGES; ++s) {
                if (s < ns) {
                    hb_mbar_expect_tx(hb_smem_u32(&bars[s]), kInStage);
                    hb_bulk_g2s(hb_smem_u32(sIn + s * 32 * HB_V), in_warp + (i64)s * in_step, kInStage, hb_smem_u32(&bars[s]));
                }
            }
        }
        __syncwarp();
    }
    const real *pf_ptr = in_warp + (i64)HB_STAGES * in_step;
    int slot = 0;
    u32 parity = 0;

    real xr[HB_ARR(HB_RV)];      // route inputs of the step whose river part is pending
    double q_cur = 0.0;          // this cell's outflow of that step
    double row_s = 0.0, row_o = 0.0;

    // iteration i: buckets of step i (i < ns)  |  river part + record of step i-1 (i > 0)  |  publish step i
#pragma unroll 1
    for (int i = 0; i <= ns; ++i) {
        real xr_next[HB_ARR(HB_RV)];
#if HB_PREFETCH_POLL
        double pfv0 = 0.0, pfv1 = 0.0, pfvd = 0.0;
        i64 pft0 = 0, pft1 = 0, pftd = 0;      // tag 0 never matches (tags are step index + 1 >= 1)
        if (i > 0 && !gave_up) {
            const i64 t_abs = a.t0 + (i - 1);
            const double *qb = a.qbuf + 2 * ((t_abs % HB_QSLOTS) * a.N + nc);
            const i64 bp_step = t_abs + 2 - HB_QSLOTS;
            if (pk0 >= 0) pft0 = hb_peek(qb - poff0, pfv0);
            if (pk1 >= 0) pft1 = hb_peek(qb - poff1, pfv1);
            if ((inmask >> 8) & 1u)
                pftd = hb_peek(a.qbuf + 2 * ((((bp_step % HB_QSLOTS) + HB_QSLOTS) % HB_QSLOTS) * a.N + nc + down_off), pfvd);
        }
#endif
        if (i < ns) {
            real hb_x[HB_ARR(HB_V)];
            if (tma_in) {
                hb_mbar_wait(hb_smem_u32(&bars[slot]), parity);
                const real *sx = sIn + slot * 32 * HB_V + lane * HB_V;
#pragma unroll
                for (int v = 0; v < HB_V; ++v) hb_x[v] = sx[v];
                // generic-proxy reads of the slot -> ordered before the async-proxy refill of the same slot.  Without this
                // fence the TMA write can overtake the loads above (seen as wrong forcing from the third step of a launch
                // on in ~40 % of cold first launches; __syncwarp alone orders the generic proxy only)
                hb_fence_proxy_async();
                __syncwarp();
                if (lane == 0 && i + HB_STAGES < ns) {
                    hb_mbar_expect_tx(hb_smem_u32(&bars[slot]), kInStage);
                    hb_bulk_g2s(hb_smem_u32(sIn + slot * 32 * HB_V), pf_ptr, kInStage, hb_smem_u32(&bars[slot]));
                }
                pf_ptr += in_step;
                if (++slot == HB_STAGES) { slot = 0; parity ^= 1; }
            } else {
#pragma unroll
                for (int v = 0; v < HB_V; ++v) hb_x[v] = gx[v];
                gx += in_step;
            }
            real hb_out[HB_ARR(HB_R)];
            (void)hb_x;

            //@@HB:STEP

#pragma unroll
            for (int v = 0; v < HB_RV; ++v) xr_next[v] = HB_ROUTE_IN(v);
            // bucket rows -> tile (i % 3); the bulk store that last read this tile (step i-3) must have drained it
            real *so = sOut + (i % HB_OBUF) * 32 * HB_R;
            if (tma_out && i >= HB_OBUF) {
                if (lane == 0) hb_bulk_wait_read<HB_OBUF - 2>();
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < HB_R; ++r) so[lane * HB_R + r] = hb_out[r];
        }
        if (i > 0) {
            // ---- river part of step i-1: poll the upstream words (= gather) and the downstream word ----
            const i64 t_abs = a.t0 + (i - 1);
            const i64 need = t_abs + 1;
            const double *qb = a.qbuf + 2 * ((t_abs % HB_QSLOTS) * a.N + nc);
            // back-pressure: publishing step t+1 overwrites step t+1-HB_QSLOTS, which the downstream cell read in ITS river part
            // of that step, i.e. before it published step t+2-HB_QSLOTS (tag t+3-HB_QSLOTS)
            const i64 bp_step = t_abs + 2 - HB_QSLOTS;
            const double *qdown = a.qbuf + 2 * ((((bp_step % HB_QSLOTS) + HB_QSLOTS) % HB_QSLOTS) * a.N + nc + down_off);
            double qin[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) qin[k] = 0.0;
            u32 pending = gave_up ? 0u : inmask;
#if HB_PREFETCH_POLL
            if (pending) {
                const bool ok0 = pk0 >= 0 && pft0 == need, ok1 = pk1 >= 0 && pft1 == need;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (ok0 && k == pk0) { qin[k] = pfv0; pending &= ~(1u << k); }
                    if (ok1 && k == pk1) { qin[k] = pfv1; pending &= ~(1u << k); }
                }
                if (((pending >> 8) & 1u) && pftd >= bp_step + 1) pending &= ~(1u << 8);
            }
#endif
            int spins = 0;
            while (pending) {
                // all outstanding words are requested back to back (one L2 round trip per poll round)
                double pv[9];
                i64 ptag[9];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    ptag[k] = 0;
                    if ((pending >> k) & 1u) ptag[k] = hb_peek(qb - 2 * ((i64)kDr[k] * a.cols + kDc[k]), pv[k]);
                }
                ptag[8] = 0;
                if ((pending >> 8) & 1u) ptag[8] = hb_peek(qdown, pv[8]);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (((pending >> k) & 1u) && ptag[k] == need) {
                        qin[k] = pv[k];
                        pending &= ~(1u << k);
                    }
                }
                if (((pending >> 8) & 1u) && ptag[8] >= bp_step + 1) pending &= ~(1u << 8);
                if (pending) {
                    if (++spins > HB_SPIN_LIMIT) {
                        gave_up = true;
                        atomicExch(a.ticket + 1, 1u);
                        break;
                    }
                    if (spins > 2) __nanosleep(spins > 64 ? 200 : 20);
                }
            }
            double inflow = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) inflow += qin[k];                  // the reference's summation order
            const double ds = hb_rds(xr, rs, q_cur, hb_p + HB_NPS);
            rs = hb_max(a.min_value, rs + (ds + inflow));
            row_s = rs;
            row_o = hb_rflux(xr, rs, hb_p + HB_NPS);
        }
        if (i < ns) {
            // ---- outflow of step i from the river state after step i-1; publish (before the record store: the
            //      neighbours' critical path runs through this word) ----------------------------------------
#pragma unroll
            for (int v = 0; v < HB_RV; ++v) xr[v] = xr_next[v];
            q_cur = hb_rflux(xr, rs, hb_p + HB_NPS);
            const i64 t_abs = a.t0 + i;
            if (valid) hb_publish(a.qbuf + 2 * ((t_abs % HB_QSLOTS) * a.N + nc), q_cur, t_abs + 1);
        }
        if (i > 0) {
            // ---- record of step i-1 leaves ------------------------------------------------------------------
            real *so = sOut + ((i - 1) % HB_OBUF) * 32 * HB_R;
            so[lane * HB_R + HB_SROW] = row_s;
            so[lane * HB_R + HB_OROW] = row_o;
            if (tma_out) {
                hb_fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    hb_bulk_s2g(ow_ptr, hb_smem_u32(so), kOutTile);
                    hb_bulk_commit();
                }
            } else if (out_valid) {
#pragma unroll
                for (int r = 0; r < HB_R; ++r) go[r] = so[lane * HB_R + r];
            }
            ow_ptr += out_step;
            go += out_step;
        }
    }
    if (tma_out && lane == 0) hb_bulk_wait<0>();   // shared tiles must outlive the last bulk stores

    if (valid) {
#pragma unroll
        for (int s = 0; s < HB_S; ++s) a.su_out[(i64)s * a.N + n] = hb_u[s];
        a.rs_out[n] = rs;
    }
}
