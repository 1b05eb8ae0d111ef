// =============================================================================================
// hb_tc.cuh — 5th-generation tensor cores (tcgen05 + TMEM) for the dense layers of neural fluxes in the opt-in
// Float32 mode (north star: "a tcgen05 tensor-core path where a basin batch makes the MLP a true dense contraction";
// reference op: LuxCore.apply of a Dense layer, /root/reference/src/nn.jl:56-58,148,161-163).
//
// A CTA of 128 threads carries 128 basins = the M dimension of  D[basin][n] = sum_k A[basin][k] * W[n][k]:
//   * every thread writes its K activations into the A tile in shared memory (canonical K-major, no swizzle: core
//     matrices of 8 rows x 16 bytes, LBO = 128 B between the two K-chunks of an MMA, SBO = 256 B between 8-row groups),
//   * ONE elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N x K8 per instruction), accumulator in TMEM,
//   * tcgen05.commit -> mbarrier; every thread then reads its own row of D with tcgen05.ld.32x32b (thread t <-> TMEM lane t).
// Plain TF32 keeps 10 mantissa bits — not enough for the stated Float32-mode tolerance — so operands are split
// x = hi + lo (hi = the 19 bits the tensor core reads, lo = x - hi exactly) and D = lo*hi + hi*lo + hi*hi ("3xTF32",
// error ~2^-21 relative, small terms first).  Weights are split once per launch, activations per evaluation.
// Spliced into templates/stepper.cu (HB_NN_TC) and #included by aot_kernels.cu for the stand-alone probe.
// =============================================================================================
#ifndef HB_TC_CUH
#define HB_TC_CUH

struct HbTc {
    float *sA;              // A tiles: hi at sA, lo at sA + HB_TC_A_FLOATS (per K-step 128 x 8 floats = 4096 B)
    unsigned sA_u32;        // shared-window address of sA
    unsigned sB_u32;        // shared-window address of the B tiles (all tensor layers, hi then lo per layer)
    unsigned bar_u32;       // mbarrier the MMA commits arrive on
    unsigned tmem;          // TMEM base address (lane 0, first allocated column)
    unsigned phase;         // parity of the next completion
    int a_lo_floats;        // offset of the lo tiles inside sA, in floats
};

__device__ __forceinline__ void hb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole warp (sync.aligned): allocate `ncols` (power of two >= 32) TMEM columns, base address -> shared word
__device__ __forceinline__ void hb_tc_alloc(unsigned smem_dst_u32, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst_u32), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hb_tc_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle, LBO 128 B, SBO 256 B, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long hb_tc_desc(unsigned smem_addr_u32) {
    return (unsigned long long)((smem_addr_u32 >> 4) & 0x3fffu) | ((unsigned long long)(128u >> 4) << 16) |
           ((unsigned long long)(256u >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D f32, A/B tf32, both K-major, M = 128, N = n
__device__ __forceinline__ constexpr unsigned hb_tc_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__device__ __forceinline__ void hb_tc_mma_tf32(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                               unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void hb_tc_commit(unsigned bar_u32) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_u32) : "memory");
}
__device__ __forceinline__ void hb_tc_mbar_wait(unsigned bar_u32, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_u32), "r"(parity)
            : "memory");
    } while (!ok);
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void hb_tc_ld16(unsigned taddr, float *r) {
    unsigned u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ float hb_tc_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// B tiles of one layer (weights W[n + N * k], Lux layout: weight[out, in] column-major) -> canonical K-major tiles, hi then lo
template <int K, int N>
__device__ __forceinline__ void hb_tc_stage_b(float *sB, const double *W, int tid, int nthreads) {
    for (int idx = tid; idx < N * K; idx += nthreads) {
        const int n = idx % N, k = idx / N;
        const float w = (float)W[idx];
        const float hi = hb_tc_hi(w);
        const int off = (k / 8) * (N * 8) + (n / 8) * 64 + ((k % 8) / 4) * 32 + (n % 8) * 4 + (k % 4);
        sB[off] = hi;
        sB[N * K + off] = w - hi;
    }
}

// One dense layer for the CTA's 128 basins: z[n] = sum_k x[k] * W[n][k] (no bias, no activation).  Called by ALL 128
// threads at the same point; `b_off_floats` = offset of this layer's B tiles inside the B region.
template <int K, int N>
__device__ __forceinline__ void hb_tc_layer(HbTc &tc, const float *x, int b_off_floats, float *z) {
    static_assert(K % 8 == 0 && N % 16 == 0 && N <= 64, "tensor layer shape");
    const int r = threadIdx.x;
    float *row_hi = tc.sA + (r / 8) * 64 + (r % 8) * 4;
#pragma unroll
    for (int kc = 0; kc < K / 4; ++kc) {
        float4 h, l;
        h.x = hb_tc_hi(x[4 * kc + 0]); l.x = x[4 * kc + 0] - h.x;
        h.y = hb_tc_hi(x[4 * kc + 1]); l.y = x[4 * kc + 1] - h.y;
        h.z = hb_tc_hi(x[4 * kc + 2]); l.z = x[4 * kc + 2] - h.z;
        h.w = hb_tc_hi(x[4 * kc + 3]); l.w = x[4 * kc + 3] - h.w;
        float *p = row_hi + (kc / 2) * 1024 + (kc % 2) * 32;      // K-step stride 4096 B, K-chunk stride 128 B
        *reinterpret_cast<float4 *>(p) = h;
        *reinterpret_cast<float4 *>(p + tc.a_lo_floats) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile writes -> visible to the tensor core
    hb_tc_fence_before();                                          // my tcgen05.ld of the previous result precedes the barrier
    __syncthreads();
    if (threadIdx.x == 0) {
        hb_tc_fence_after();
        const unsigned idesc = hb_tc_idesc(N);
        const unsigned a_hi = tc.sA_u32, a_lo = tc.sA_u32 + (unsigned)tc.a_lo_floats * 4u;
        const unsigned b_hi = tc.sB_u32 + (unsigned)b_off_floats * 4u, b_lo = b_hi + (unsigned)(N * K) * 4u;
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
            const unsigned ao = (unsigned)ks * 4096u, bo = (unsigned)ks * (unsigned)(N * 32);
            hb_tc_mma_tf32(tc.tmem, hb_tc_desc(a_lo + ao), hb_tc_desc(b_hi + bo), idesc, ks > 0 ? 1u : 0u);
            hb_tc_mma_tf32(tc.tmem, hb_tc_desc(a_hi + ao), hb_tc_desc(b_lo + bo), idesc, 1u);
            hb_tc_mma_tf32(tc.tmem, hb_tc_desc(a_hi + ao), hb_tc_desc(b_hi + bo), idesc, 1u);
        }
        hb_tc_commit(tc.bar_u32);
    }
    hb_tc_mbar_wait(tc.bar_u32, tc.phase);
    tc.phase ^= 1u;
    hb_tc_fence_after();
    const unsigned lane_base = ((unsigned)(threadIdx.x & ~31) & 127u) << 16;
#pragma unroll
    for (int c = 0; c < N / 16; ++c) hb_tc_ld16(tc.tmem + lane_base + 16u * c, z + 16 * c);
}

#endif  // HB_TC_CUH
