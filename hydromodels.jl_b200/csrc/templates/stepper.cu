// =============================================================================================
// hb_stepper — fused explicit time-stepper for per-node bucket models on B200 (sm_100a).
//
// Replaces, in ONE streaming pass over t, what the reference does per component in two passes
// over the whole series (hydrosolve MutableSolver: /root/reference/src/solve.jl:31-55; bucket
// functors: src/bucket.jl:169-256; model sequencing: src/model.jl:241-309; UH: src/uh.jl:187-237):
//
//   for t in 1..T:   x  <- input[:, n, t]                      (ConstantInterpolation, interpolate.jl:49)
//     for each component k (declaration order):
//        fluxes(x, u_pre)  -> du ;  u <- max(min_value, u + du) ;  fluxes(x, u_post) -> rows
//        (stateless flux: one evaluation;  unit hydrograph: causal ring-buffer dot product)
//     out[:, n, t] <- [states..., outputs...]
//
// This file is a TEMPLATE: the host lowers the model's expression trees to CUDA C and the
// library splices the text in at the //@@HB:<SECTION> markers, then NVRTC-compiles for sm_100a.
//
// Execution model: one thread per node (basin / grid cell / ensemble member), all state in
// registers for the whole run.  Warps are autonomous — no block-wide barrier in the time loop.
// Every warp owns 32 consecutive nodes, so its forcing for one step is ONE contiguous
// 32*V*8-byte run of the time-major [T][N][V] array and its results for one step ONE contiguous
// 32*R*8-byte run of [T][N][R]:
//   * forcing:  lane 0 issues a TMA bulk copy (cp.async.bulk global->shared, mbarrier
//     complete_tx) HB_STAGES time steps ahead into a per-warp ring; lanes read their V values
//     from shared memory.  Every HBM read is a full-line, coalesced bulk transfer.
//   * results:  lanes write their R values into a per-warp shared tile, lane 0 issues a TMA
//     bulk store (cp.async.bulk shared->global, bulk_group) — double-buffered, fire and forget.
// Shapes that cannot meet the 16-byte TMA alignment (odd N*V or N*R, the ragged last warp,
// N == 1) take the per-thread ld.global / st.global path of the same kernel (warp-uniform
// branch); results are identical.
// =============================================================================================

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

//@@HB:DEFINES
// (generated) HB_V HB_R HB_S HB_NP HB_NC HB_UHW_TOTAL HB_UHH_TOTAL HB_NN_WORDS HB_MODE HB_METRIC
//             HB_SHARED_FORCING HB_BLOCK HB_STAGES HB_MINBLOCKS

#ifndef HB_NN_MMA
#define HB_NN_MMA 0
#endif
#ifndef HB_V2
#define HB_V2 0       // auxiliary per-node inputs (a second [T][N][V2] stream riding the same TMA ring): rows another kernel
#endif                //   produced for this model — the route state and outflow the generic routed path computes between two
                      //   stepper passes (hb_route_step in planes mode) — which the body copies into the records (hb_aux[k])
#ifndef HB_NN_TC
#define HB_NN_TC 0    // 1: Float32 mode, tensor-shaped dense layers through tcgen05.mma (templates/hb_tc.cuh): the CTA's 128
#endif                //    threads / basins are the M dimension, so the CTA meets at a barrier in every neural-flux evaluation
#ifndef HB_F32
#define HB_F32 0
#endif
// `real` = the arithmetic type of the run: double (the reference's promote_type, src/solve.jl:42) or,
// opt-in, float (forcing and result streams are then float32 too; parameters, initial states, UH
// ordinates, MLP parameters, targets and objectives stay double in memory)
#if HB_F32
typedef float real;
#else
typedef double real;
#endif
#ifndef HB_TANGENTS
#define HB_TANGENTS 0 // D > 0: gradient run — parameters carry D unit tangents (HB_TDIR: direction of each parameter slot or -1),
#endif                //   states / fluxes / the objective propagate them (hb_dual, hb_math.cuh); objective mode only, Float64 only
#ifndef HB_NPW
#define HB_NPW 32     // nodes per warp.  16: lanes l and l+16 carry the SAME node — the scalar code runs identically in both halves
#endif                //   (free under SIMT) while the warp-cooperative tensor-core MLPs spread half as many basins over the 32 lanes:
                      //   twice the warps for a small, MLP-heavy (latency-bound) run such as M50 on 50 k basins
#define HB_OBUF 2
#define HB_WARPS (HB_BLOCK / 32)
#define HB_ARR(n) ((n) > 0 ? (n) : 1)

struct HbArgs {
    const real *in;
    const double *params;
    const int *htypes;
    const double *init;
    const double *uhw;
    const double *uhh_in;
    const double *nn;
    const double *target;
    real *out;
    double *obj;
    double *final_states;
    double *uhh_out;
    const real *in2;          // [T][N][HB_V2] auxiliary inputs (HB_V2 > 0)
    i64 N;
    i64 T;
    double min_value;
    int warm_up;
    int target_per_node;
    int tma_in_ok;
    int tma_out_ok;
    int compact_n;            // > 0: store only rows compact_rows[0..compact_n) of every step, as out[T][N][compact_n] — the SAME
    int compact_rows[4];      //   cubin (hence bit-identical values) then serves as the first pass of the generic routed path
    int p_off[HB_ARR(HB_NP)];
    int p_mode[HB_ARR(HB_NP)];
};

// ---- PTX wrappers (mbarrier + TMA bulk copies) ------------------------------------------------
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
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void hb_bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared -> global, tracked by bulk async-groups
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

// ---- math helpers shared by all generated bodies: templates/hb_math.cuh -----------------------
//@@HB:MATH

// `treal`: what parameters, states and everything derived from them are made of
#if HB_TANGENTS > 0
typedef hb_dual treal;
typedef hb_dual tacc;    // objective accumulators
#else
typedef real treal;
typedef double tacc;
#endif

// ---- MLP parameters: module-scope constant bank (filled by the host before each launch) so that
//      every weight is a c[3][imm] operand of its DFMA; shared across all nodes (src/nn.jl:140-165)
#if HB_NN_WORDS > 0
// in the arithmetic type of the run: a Float32 body must not pay a double -> float conversion per weight use (the host
// converts the parameters once per launch, hb_api.cpp launch_stepper)
__constant__ real hb_nn_const[HB_NN_WORDS];
#define HB_NNW(i) (hb_nn_const[i])
#endif
// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor pipe (SASS DMMA).  Fragments: A lane l -> A[l/4][l%4],
// B lane l -> B[l%4][l/4], C/D lane l -> C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void hb_dmma(double &d0, double &d1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// the first product of an accumulation reads its addend (the bias pair, shared by all M tiles) from other registers: no copies
__device__ __forceinline__ void hb_dmma_c(double &d0, double &d1, const double a, const double b, const double c0, const double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};" : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

#if HB_NN_TC
//@@HB:TC
#endif

//@@HB:HELPERS

// ---- the kernel -------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(HB_BLOCK, HB_MINBLOCKS) hb_stepper(const __grid_constant__ HbArgs a) {
    extern __shared__ __align__(128) unsigned char hb_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int ln = lane % HB_NPW;                       // this lane's node within the warp (HB_NPW == 16: lane and lane+16 share it)
    const i64 n0w = ((i64)blockIdx.x * HB_WARPS + warp) * HB_NPW;
    const i64 n = n0w + ln;
    const bool valid = n < a.N;
    const bool full_warp = (n0w + HB_NPW) <= a.N;
    hb_math_init();   // exp table -> shared memory (block barrier inside)

    // shared-memory carve-up: [MLP params + per-warp MMA scratch][per-warp forcing ring][per-warp result tiles][mbarriers]
#if HB_NN_MMA
    constexpr int kNNBytes = ((HB_NN_WORDS * 8 + 127) / 128) * 128 + HB_WARPS * 1024;
#elif HB_NN_TC
    constexpr int kNNBytes = (((2 * HB_TC_A_FLOATS + HB_TC_B_FLOATS) * 4 + 127) / 128) * 128;   // A tiles (hi, lo) + B tiles of every tensor layer
#else
    constexpr int kNNBytes = 0;   // per-thread MLP path: parameters live in the constant bank (hb_nn_const)
#endif
    constexpr int kInStage = HB_NPW * HB_V * (int)sizeof(real);                      // bytes of one warp-step of forcing
    constexpr int kIn2Stage = HB_NPW * HB_V2 * (int)sizeof(real);                    // ... of auxiliary inputs
    constexpr int kIn1Bytes = ((HB_STAGES * kInStage + 127) / 128) * 128;
    constexpr int kInBytes = kIn1Bytes + ((HB_STAGES * kIn2Stage + 127) / 128) * 128;
    constexpr int kOutTile = HB_NPW * HB_R * (int)sizeof(real);                      // bytes of one warp-step of results
    constexpr int kOutBytes = ((HB_OBUF * kOutTile + 127) / 128) * 128;
    real *sNN = reinterpret_cast<real *>(hb_smem);   // tensor path only (Float64): staged MLP parameters
    double *hb_scr = reinterpret_cast<double *>(hb_smem + ((HB_NN_WORDS * 8 + 127) / 128) * 128) + warp * 128;   // 1 KB per warp
    unsigned char *warp_base = hb_smem + kNNBytes + (size_t)warp * (kInBytes + kOutBytes);
    real *sIn = reinterpret_cast<real *>(warp_base);
    real *sIn2 = reinterpret_cast<real *>(warp_base + kIn1Bytes);
    real *sOut = reinterpret_cast<real *>(warp_base + kInBytes);
    (void)sIn2;
    u64 *bars = reinterpret_cast<u64 *>(hb_smem + kNNBytes + (size_t)HB_WARPS * (kInBytes + kOutBytes)) + warp * HB_STAGES;
    (void)sNN; (void)hb_scr; (void)sIn; (void)sOut; (void)bars;
#if HB_NN_MMA
    // tensor-core MLP path: dense-layer weights are read as MMA fragments (lane-dependent addresses)
    for (int i = threadIdx.x; i < HB_NN_WORDS; i += HB_BLOCK) sNN[i] = a.nn[i];
    __syncthreads();
#endif
#if HB_NN_TC
    // tcgen05 path: weights -> canonical K-major B tiles (hi / lo halves of the 3xTF32 split), TMEM columns for the
    // accumulator, one mbarrier for the MMA completions
    __shared__ u64 hb_tc_bar;
    __shared__ u32 hb_tc_tmem_base;
    HbTc hb_tc;
    {
        float *sTcA = reinterpret_cast<float *>(hb_smem);
        float *sTcB = sTcA + 2 * HB_TC_A_FLOATS;
        HB_TC_STAGE(sTcB, a.nn, (int)threadIdx.x, HB_BLOCK)
        if (warp == 0) hb_tc_alloc(hb_smem_u32(&hb_tc_tmem_base), HB_TC_COLS);
        if (threadIdx.x == 0) {
            hb_mbar_init(hb_smem_u32(&hb_tc_bar), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        hb_fence_proxy_async();
        hb_tc_fence_before();
        __syncthreads();
        hb_tc_fence_after();
        hb_tc.sA = sTcA;
        hb_tc.sA_u32 = hb_smem_u32(sTcA);
        hb_tc.sB_u32 = hb_smem_u32(sTcB);
        hb_tc.bar_u32 = hb_smem_u32(&hb_tc_bar);
        hb_tc.tmem = hb_tc_tmem_base;
        hb_tc.phase = 0;
        hb_tc.a_lo_floats = HB_TC_A_FLOATS;
    }
#endif

    // warp-uniform choice of the I/O path
    const bool tma_in = (HB_SHARED_FORCING == 0) && (HB_V > 0) && a.tma_in_ok && full_warp;
    const bool tma_out = (HB_MODE == 0) && (HB_R > 0) && a.tma_out_ok && full_warp;
    // Tensor-path bodies contain warp-collective MMAs / shuffles: lanes past the last node of a ragged warp
    // stay alive, shadow the last node (`nc`) and never store.  Every other body lets them exit here (the
    // ragged warp is on the per-thread path, which has no warp syncs) — measurably faster code
    // (GR4J 100k x 3650: 8.4 vs 9.4 ms).
    const i64 nc = valid ? n : a.N - 1;
#if !HB_NN_MMA && !HB_NN_TC
    if (!valid) return;
#endif

    // ---- per-node constants: parameters (p[htypes], /root/reference/src/utils.jl:174-201) ------
    treal hb_p[HB_ARR(HB_NP)];
#pragma unroll
    for (int j = 0; j < HB_NP; ++j) {
        const int mode = a.p_mode[j];
        i64 idx = a.p_off[j];
        if (mode == 1) idx += nc;
        else if (mode >= 2) idx += a.htypes[(i64)(mode - 2) * a.N + nc];
        hb_p[j] = (real)a.params[idx];
    }
#if HB_TANGENTS > 0
    {
        constexpr int tdir[HB_ARR(HB_NP)] = HB_TDIR;   // d p_j / d theta_k = [tdir[j] == k]
#pragma unroll
        for (int j = 0; j < HB_NP; ++j)
#pragma unroll
            for (int k = 0; k < HB_TANGENTS; ++k) hb_p[j].d[k] = tdir[j] == k ? 1.0 : 0.0;
    }
#endif
    // ---- states: initial value is used UNCLAMPED for the first increment (solve.jl:47-49) ------
    treal hb_u[HB_ARR(HB_S)];
#pragma unroll
    for (int s = 0; s < HB_S; ++s) hb_u[s] = a.init ? (real)a.init[(i64)s * a.N + nc] : (real)0.0;
#if HB_TANGENTS > 0
    {
        constexpr int sdir[HB_ARR(HB_S)] = HB_TSDIR;   // d u_s(0) / d theta_k = [sdir[s] == k]: derivatives w.r.t. initial states
#pragma unroll
        for (int s = 0; s < HB_S; ++s)
#pragma unroll
            for (int k = 0; k < HB_TANGENTS; ++k) hb_u[s].d[k] = sdir[s] == k ? 1.0 : 0.0;
    }
#endif
    // ---- unit-hydrograph ordinates and history -----------------------------------------------
    treal hb_uhw[HB_ARR(HB_UHW_TOTAL)];
    treal hb_uhh[HB_ARR(HB_UHH_TOTAL)];
#pragma unroll
    for (int k = 0; k < HB_UHW_TOTAL; ++k) hb_uhw[k] = (real)a.uhw[(i64)k * a.N + nc];
#pragma unroll
    for (int k = 0; k < HB_UHH_TOTAL; ++k) hb_uhh[k] = a.uhh_in ? (real)a.uhh_in[(i64)k * a.N + nc] : (real)0.0;
    treal hb_c[HB_ARR(HB_NC)];   // state-only subexpressions carried from one step to the next
    const real hb_minv = (real)a.min_value;
    (void)hb_p; (void)hb_u; (void)hb_uhw; (void)hb_uhh; (void)hb_c; (void)hb_minv;

#if HB_MODE == 1
    // objective accumulators (shifted sums; metrics of …OptimizationExt.jl:34-68)
    double q_n = 0.0, q_so = 0.0, q_soo = 0.0, q_shift = 0.0, q_shift_s = 0.0;
    tacc q_ss = 0.0, q_sss = 0.0, q_sos = 0.0, q_sse = 0.0;   // the sums the simulated series enters
#endif

    //@@HB:PROLOGUE

    const i64 T = a.T;
    const i64 in_step = (HB_SHARED_FORCING ? 1 : a.N) * (i64)HB_V;   // doubles per time step
    const real *in_warp = a.in + (HB_SHARED_FORCING ? 0 : n0w * HB_V);
    const real *in_lane = a.in + (HB_SHARED_FORCING ? 0 : nc * HB_V);
    real *out_warp = a.out + n0w * HB_R;
    real *out_lane = a.out + n * HB_R;
    (void)in_warp; (void)out_warp; (void)out_lane; (void)in_lane;
#if HB_V2 > 0
    // the auxiliary stream is optional at run time (NULL: hb_aux reads as 0 — the first pass of the generic routed path runs
    // this same kernel before the route rows exist)
    const int in2_bytes = a.in2 ? kIn2Stage : 0;
    const i64 in2_step = a.N * (i64)HB_V2;
    const real *in2_warp = a.in2 + n0w * HB_V2;
    const real *gx2 = a.in2 + nc * HB_V2;
    const real *pf2_ptr = in2_warp + (i64)HB_STAGES * in2_step;
#else
    constexpr int in2_bytes = 0;
#endif

    if (tma_in) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < HB_STAGES; ++s) hb_mbar_init(hb_smem_u32(&bars[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            hb_fence_proxy_async();
#pragma unroll
            for (int s = 0; s < HB_STAGES; ++s) {
                if (s < T) {
                    hb_mbar_expect_tx(hb_smem_u32(&bars[s]), kInStage + in2_bytes);
                    hb_bulk_g2s(hb_smem_u32(sIn + s * HB_NPW * HB_V), in_warp + (i64)s * in_step, kInStage, hb_smem_u32(&bars[s]));
#if HB_V2 > 0
                    if (in2_bytes) hb_bulk_g2s(hb_smem_u32(sIn2 + s * HB_NPW * HB_V2), in2_warp + (i64)s * in2_step, kIn2Stage, hb_smem_u32(&bars[s]));
#endif
                }
            }
        }
        __syncwarp();
    }

    int slot = 0;
    u32 parity = 0;
    int ob = 0;
    // running pointers (strength-reduced: no 64-bit multiplies in the time loop)
    const i64 out_step = a.N * (i64)HB_R;
    const real *pf_ptr = in_warp + (i64)HB_STAGES * in_step;   // forcing prefetched HB_STAGES steps ahead
    const real *gx = in_lane;
    real *ow_ptr = out_warp;
    real *go = out_lane;
    (void)pf_ptr; (void)gx; (void)ow_ptr; (void)go; (void)out_step;
#pragma unroll 1
    for (i64 t = 0; t < T; ++t) {
        // ---- forcing of this step -------------------------------------------------------------
        real hb_x[HB_ARR(HB_V)];
        real hb_aux[HB_ARR(HB_V2)];
        (void)hb_aux;
        if (tma_in) {
            hb_mbar_wait(hb_smem_u32(&bars[slot]), parity);
            const real *sx = sIn + slot * HB_NPW * HB_V + ln * HB_V;
#pragma unroll
            for (int v = 0; v < HB_V; ++v) hb_x[v] = sx[v];
#if HB_V2 > 0
#pragma unroll
            for (int v = 0; v < HB_V2; ++v) hb_aux[v] = in2_bytes ? sIn2[slot * HB_NPW * HB_V2 + ln * HB_V2 + v] : (real)0.0;
#endif
            // cross-proxy write-after-read: the lanes' generic-proxy loads of the slot must be ordered before the
            // async-proxy (TMA) refill of the same slot; __syncwarp orders the generic proxy only (the wave grid
            // kernel, which shares this ring, showed overtaken loads on cold launches without the fence)
            hb_fence_proxy_async();
            __syncwarp();   // every lane has consumed the slot -> refill it HB_STAGES steps ahead
            if (lane == 0 && t + HB_STAGES < T) {
                hb_mbar_expect_tx(hb_smem_u32(&bars[slot]), kInStage + in2_bytes);
                hb_bulk_g2s(hb_smem_u32(sIn + slot * HB_NPW * HB_V), pf_ptr, kInStage, hb_smem_u32(&bars[slot]));
#if HB_V2 > 0
                if (in2_bytes) hb_bulk_g2s(hb_smem_u32(sIn2 + slot * HB_NPW * HB_V2), pf2_ptr, kIn2Stage, hb_smem_u32(&bars[slot]));
#endif
            }
            pf_ptr += in_step;
#if HB_V2 > 0
            pf2_ptr += in2_step;
#endif
            if (++slot == HB_STAGES) { slot = 0; parity ^= 1; }
        } else {
#pragma unroll
            for (int v = 0; v < HB_V; ++v) hb_x[v] = gx[v];
            gx += in_step;
#if HB_V2 > 0
#pragma unroll
            for (int v = 0; v < HB_V2; ++v) hb_aux[v] = in2_bytes ? gx2[v] : (real)0.0;
            gx2 += in2_step;
#endif
        }
        real hb_out[HB_ARR(HB_R)];
        (void)hb_x; (void)hb_out;
#if HB_MODE == 1
        treal hb_sim = (real)0.0;
#endif

        //@@HB:STEP

        // ---- results of this step -------------------------------------------------------------
#if HB_MODE == 0 && HB_R > 0
        if (a.compact_n > 0) {
            // selected rows only, as out[T][N][compact_n]: through the warp's shared tile (a run-time row index into the
            // register array hb_out would spill it), then one coalesced store per selected row
            real *so = sOut;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < HB_R; ++r) so[ln * HB_R + r] = hb_out[r];
            __syncwarp();
            if (valid) {
                for (int k = 0; k < a.compact_n; ++k) a.out[(t * a.N + n) * a.compact_n + k] = so[ln * HB_R + a.compact_rows[k]];
            }
        } else if (tma_out) {
            real *so = sOut + ob * HB_NPW * HB_R;
            if (t >= HB_OBUF) {   // the bulk store that last read this tile must have drained it
                if (lane == 0) hb_bulk_wait_read<HB_OBUF - 1>();
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < HB_R; ++r) so[ln * HB_R + r] = hb_out[r];   // HB_NPW == 16: both copies of a node store the same values
            hb_fence_proxy_async();   // generic-proxy writes -> visible to the async proxy
            __syncwarp();
            if (lane == 0) {
                hb_bulk_s2g(ow_ptr, hb_smem_u32(so), kOutTile);
                hb_bulk_commit();
            }
            ow_ptr += out_step;
            ob ^= 1;
        } else {
            if (valid) {
#pragma unroll
                for (int r = 0; r < HB_R; ++r) go[r] = hb_out[r];
            }
            go += out_step;
        }
#endif
#if HB_MODE == 1
        if (t + 1 >= a.warm_up) {
            double o = a.target_per_node ? a.target[t * a.N + nc] : a.target[t];
            tacc s = (tacc)hb_sim;
#if HB_METRIC == 3
            o = log(o + 1e-6);
            s = log(s + 1e-6);
#endif
            // each series is shifted by its OWN first value: a simulated series many orders of magnitude below the observed
            // one (degenerate parameter sets are routine inside a calibration) keeps its variance and covariance
            if (q_n == 0.0) { q_shift = o; q_shift_s = hb_val(s); }   // the shifts are constants of the run (no tangent)
            const double od = o - q_shift;
            const auto sd = s - q_shift_s, e = o - s;
            q_n += 1.0; q_so += od; q_ss += sd;
            q_soo = fma(od, od, q_soo); q_sss = fma(sd, sd, q_sss); q_sos = fma(od, sd, q_sos);
            q_sse = fma(e, e, q_sse);
        }
#endif
    }
    if (tma_out && lane == 0) hb_bulk_wait<0>();   // shared tiles must outlive the last bulk stores

    // ---- epilogue: final states / UH history for chunked-T runs, objective ---------------------
#if HB_NN_TC
    // the TMEM columns go back before the CTA ends; every thread (also the shadows past the last node) reaches this point
    hb_tc_fence_before();
    __syncthreads();
    if (warp == 0) hb_tc_dealloc(hb_tc.tmem, HB_TC_COLS);
#endif
    if (!valid) return;
    if (a.final_states) {
#pragma unroll
        for (int s = 0; s < HB_S; ++s) a.final_states[(i64)s * a.N + n] = hb_val(hb_u[s]);
    }
    if (a.uhh_out) {
#pragma unroll
        for (int k = 0; k < HB_UHH_TOTAL; ++k) a.uhh_out[(i64)k * a.N + n] = hb_val(hb_uhh[k]);
    }
#if HB_MODE == 1
    {
        const double var_o = q_soo - q_so * q_so / q_n;   // = sum (o - mean o)^2
        tacc val;
#if HB_METRIC == 0
        val = q_sse / q_n;                                 // MSE
#elif HB_METRIC == 1
        val = -(1.0 - q_sse / var_o);                      // -NSE
#elif HB_METRIC == 4
        val = q_sse;                                       // sum of squared errors (the loss of README.md:161-187)
#else
        const auto var_s = q_sss - q_ss * q_ss / q_n;
        const auto cov = q_sos - q_so * q_ss / q_n;
        const auto r = cov / sqrt(var_o * var_s);
        const auto alpha = sqrt(var_s / var_o);            // std(sim)/std(obs), (n-1) cancels
        const auto beta = (q_ss + q_n * q_shift_s) / (q_so + q_n * q_shift);
        val = -(1.0 - sqrt((r - 1.0) * (r - 1.0) + (alpha - 1.0) * (alpha - 1.0) + (beta - 1.0) * (beta - 1.0)));
#endif
        a.obj[n] = hb_val(val);
#if HB_TANGENTS > 0
        // gradient rows [D][N] (the `out` slot of the argument block carries the buffer in a gradient run)
        double *grad = reinterpret_cast<double *>(a.out);
#pragma unroll
        for (int k = 0; k < HB_TANGENTS; ++k) grad[(i64)k * a.N + n] = val.d[k];
#endif
    }
#endif
    //@@HB:EPILOGUE
}
