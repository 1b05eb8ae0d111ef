// Ahead-of-time (nvcc, sm_100a) kernels that do not depend on a model: measurement probes used by
// bench.py (FP64 FMA peak for the compute-bound roofline, L2 flush) and exported through the C-ABI.
#include <cuda_runtime.h>
#include <mutex>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/hydrob200.h"
#include "templates/hb_tc.cuh"

namespace {

// Stand-alone check of the tcgen05 dense-layer path (templates/hb_tc.cuh): one CTA of 128 rows, D = A * W^T, `reps` rounds
// through the same TMEM columns and mbarrier (phase flips), the last round's result is stored.
template <int K, int N>
__global__ void __launch_bounds__(128) hb_tc_probe_kernel(const float *__restrict__ A, const double *__restrict__ W, float *__restrict__ D, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    constexpr int kAFloats = (K / 8) * 1024;
    float *sA = reinterpret_cast<float *>(smem);
    float *sB = sA + 2 * kAFloats;
    hb_tc_stage_b<K, N>(sB, W, threadIdx.x, 128);
    if (threadIdx.x < 32) hb_tc_alloc((unsigned)__cvta_generic_to_shared(&tmem_base), N < 32 ? 32 : N);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    hb_tc_fence_before();
    __syncthreads();
    hb_tc_fence_after();
    HbTc tc;
    tc.sA = sA;
    tc.sA_u32 = (unsigned)__cvta_generic_to_shared(sA);
    tc.sB_u32 = (unsigned)__cvta_generic_to_shared(sB);
    tc.bar_u32 = (unsigned)__cvta_generic_to_shared(&bar);
    tc.tmem = tmem_base;
    tc.phase = 0;
    tc.a_lo_floats = kAFloats;
    float x[K], z[N];
    for (int k = 0; k < K; ++k) x[k] = A[threadIdx.x * K + k];
    for (int rep = 0; rep < reps; ++rep) {
        float xs[K];
        const float s = (rep == reps - 1) ? 1.0f : 0.5f + rep;          // earlier rounds leave different values behind
        for (int k = 0; k < K; ++k) xs[k] = x[k] * s;
        hb_tc_layer<K, N>(tc, xs, 0, z);
    }
    for (int n = 0; n < N; ++n) D[threadIdx.x * N + n] = z[n];
    hb_tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) hb_tc_dealloc(tc.tmem, N < 32 ? 32 : N);
}

// Dependent-chain FP64 FMA probe: 8 independent accumulators per thread, ITERS x 8 DFMA each.
__global__ void __launch_bounds__(256) hb_fp64_probe(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;   // never true; keeps the chain alive
}

// FP64 tensor-core probe: mma.sync m8n8k4 f64 (SASS DMMA), 4 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256) hb_dmma_probe(double *out, int iters, double a, double b) {
    double c0[2] = {0.0, 0.0}, c1[2] = {1.0, 1.0}, c2[2] = {2.0, 2.0}, c3[2] = {3.0, 3.0};
    const double av = a + threadIdx.x * 1e-9, bv = b;
    for (int i = 0; i < iters; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(av), "d"(bv));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(av), "d"(bv));
    }
    const double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
    if (s == 123.456) out[0] = s;
}

// Co-tenant stand-in for tests: every block holds `smem` bytes of shared memory (i.e. a share of an SM) and spins until
// `ns` nanoseconds have passed on the global timer.
__global__ void hb_occupy(unsigned long long ns) {
    extern __shared__ unsigned char hog[];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) hog[0] = 1;
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}

__global__ void hb_cvt_f64_f32(const double *__restrict__ src, float *__restrict__ dst, int64_t n) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}

__global__ void hb_fill(double *p, int64_t n, double v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

extern "C" {

// Measured FP64 FMA throughput of this GPU in TFLOP/s (2 flop per FMA); the denominator for the
// compute-bound configs (MEASURED_PEAKS.json has no FP64 entry).
int hb_probe_fp64_tflops(double *tflops) {
    if (!tflops) return HB_ERR_INVALID;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HB_ERR_NO_DEVICE; }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return HB_ERR_NOMEM;
    const int iters = 1 << 14, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    hb_fp64_probe<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        hb_fp64_probe<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return HB_ERR_CUDA;
    const double flop = 2.0 * 8.0 * iters * (double)blocks * threads;
    *tflops = flop / (best * 1e-3) / 1e12;
    return HB_OK;
}

// Measured FP64 tensor-core (DMMA m8n8k4) throughput in TFLOP/s: decides whether the MLP-fused
// stepper should contract its dense layers on the tensor pipe or on the FP64 FMA pipe.
int hb_probe_dmma_tflops(double *tflops) {
    if (!tflops) return HB_ERR_INVALID;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return HB_ERR_NO_DEVICE; }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess) return HB_ERR_NOMEM;
    const int iters = 1 << 12, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    hb_dmma_probe<<<blocks, threads>>>(d, iters, 0.5, 1e-3);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        hb_dmma_probe<<<blocks, threads>>>(d, iters, 0.5, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return HB_ERR_CUDA;
    const double flop = 2.0 * 8 * 8 * 4 * 4.0 * iters * (double)blocks * (threads / 32);   // 4 MMAs of 8x8x4 per warp-iteration
    *tflops = flop / (best * 1e-3) / 1e12;
    return HB_OK;
}

// Hold `n_blocks` blocks of `smem_bytes` shared memory each busy for `ms` milliseconds on `stream` (asynchronous): the
// co-tenant the wave grid kernel's forward-progress tests run against.
int hb_probe_occupy_sms(int n_blocks, int smem_bytes, double ms, void *stream) {
    if (n_blocks <= 0 || smem_bytes < 0 || smem_bytes > 227 * 1024 || !(ms >= 0)) return HB_ERR_INVALID;
    if (smem_bytes > 48 * 1024 &&
        cudaFuncSetAttribute(hb_occupy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return HB_ERR_CUDA;
    hb_occupy<<<n_blocks, 32, smem_bytes, (cudaStream_t)stream>>>((unsigned long long)(ms * 1e6));
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// D[128][N] = A[128][K] * W^T (W[n + N*k] doubles, rounded to float) through the tcgen05 path of templates/hb_tc.cuh;
// (K, N) in {(16, 16), (24, 32), (64, 64)}.  DEVICE pointers; test / measurement aid.
int hb_probe_tc_layer(int k, int n, const float *A, const double *W, float *D, int reps, void *stream) {
    if (!A || !W || !D || reps < 1) return HB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = (2 * (k / 8) * 1024 + 2 * k * n) * 4;
    if (k == 16 && n == 16) hb_tc_probe_kernel<16, 16><<<1, 128, smem, st>>>(A, W, D, reps);
    else if (k == 24 && n == 32) hb_tc_probe_kernel<24, 32><<<1, 128, smem, st>>>(A, W, D, reps);
    else if (k == 64 && n == 64) {
        if (cudaFuncSetAttribute(hb_tc_probe_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return HB_ERR_CUDA;
        hb_tc_probe_kernel<64, 64><<<1, 128, smem, st>>>(A, W, D, reps);
    } else return HB_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// dst[i] = (float)src[i] (MLP parameters of a Float32 run -> the kernel's float constant bank)
int hb_convert_f64_f32_device(const double *src, float *dst, int64_t n, void *stream) {
    if (!src || !dst || n < 0) return HB_ERR_INVALID;
    if (n == 0) return HB_OK;
    hb_cvt_f64_f32<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// Overwrite `n` doubles (use a buffer larger than the 126 MB L2 to flush it between timed runs).
int hb_fill_device(double *p, int64_t n, double v, void *stream) {
    if (!p || n < 0) return HB_ERR_INVALID;
    hb_fill<<<1184, 256, 0, (cudaStream_t)stream>>>(p, n, v);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

}  // extern "C"

// =============================================================================================
// SparseRouteConvolution (/root/reference/src/route.jl:593-610): out[dest, :] += causal_conv(in[source, :],
// weights[row, :]) over the nnz (dest, source) rows of a SparseRouteKernel — IRF river routing.
// Rows are grouped by destination (CSR, original row order kept inside a group, so the summation order is
// the reference's); one thread per (destination, time step): out[t][d] = sum_rows sum_{lag<=min(t,H)}
// w[row][lag] * in[t-lag+1][src(row)].  Layouts are the Julia column-major arrays: in [T][n_in],
// out [T][n_out], weights [H][nnz] (CSR-ordered rows).
// =============================================================================================
namespace {
__global__ void __launch_bounds__(256) hb_sparse_conv_kernel(int n_out, int n_in, long long T, int H, long long nnz,
                                                             const int *__restrict__ row_ptr, const int *__restrict__ src,
                                                             const double *__restrict__ w, const double *__restrict__ in,
                                                             double *__restrict__ out) {
    const int d = blockIdx.x * 32 + (threadIdx.x & 31);
    const long long t = (long long)blockIdx.y * 8 + (threadIdx.x >> 5);
    if (d >= n_out || t >= T) return;
    double total = 0.0;
    const int upper = (int)((t + 1) < H ? (t + 1) : H);
    for (int row = row_ptr[d]; row < row_ptr[d + 1]; ++row) {
        const double *x = in + src[row];
        const double *wr = w + row;
        double acc = 0.0;
        for (int lag = 0; lag < upper; ++lag) acc += wr[(long long)lag * nnz] * x[(t - lag) * n_in];   // _causal_conv_add!
        total += acc;
    }
    out[t * n_out + d] = total;
}
}  // namespace

namespace {
// Tiled variant (horizon <= HP): a block owns 32 destinations x 64 time steps (8 warps x 8 steps), a thread one destination
// and TT = 8 consecutive steps.  Per kernel row the block stages, with cp.async straight into shared memory and one row
// AHEAD of the arithmetic (double buffer), the row's H weights and the 64 + H - 1 input values its 8 time tiles need; a
// thread then slides a register ring of 8 inputs over the lags (one shared-memory load per lag), so a row costs each
// thread H + 7 LDS for 8 * H DFMAs and no global load sits in front of the multiply-adds.  Operation order = the kernel
// above (lags ascending inside a row, rows in CSR order): results are bit-identical.
__device__ __forceinline__ void hb_cp_async8(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 8 : 0;                      // src-size 0: nothing is read, the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

#ifndef HB_IRF_TT
#define HB_IRF_TT 32                                   // time steps per thread (default; 8 | 16 | 32 through the environment variable HB_IRF_TT)
#endif
template <int HP, int TT>
__global__ void __launch_bounds__(256) hb_sparse_conv_tiled(int n_out, int n_in, long long T, int H, long long nnz,
                                                            const int *__restrict__ row_ptr, const int *__restrict__ src,
                                                            const double *__restrict__ w, const double *__restrict__ in,
                                                            const double *__restrict__ xt, long long Tp,
                                                            const int *__restrict__ order, double *__restrict__ out) {
    constexpr int BT = 8 * TT;                         // time steps per block (8 warps)
    constexpr int WP = BT + HP;                        // staged input window: steps [t_blk - HP, t_blk + BT)
    extern __shared__ __align__(16) double hb_irf_smem[];
    // pitch 33: the time-major staging below writes a column (32 consecutive steps of ONE source) per warp instruction and
    // the arithmetic reads a row (32 sources at one step) — both conflict-free with an odd pitch
    double(*sx)[WP][33] = reinterpret_cast<double(*)[WP][33]>(hb_irf_smem);                       // [2][WP][33]
    double(*sw)[HP][32] = reinterpret_cast<double(*)[HP][32]>(hb_irf_smem + 2 * WP * 33);        // [2][HP][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // `order`: the destinations in the sequence the blocks take them (e.g. sorted by row count, so that the 32 of a block
    // finish together — a block runs for as many rows as its longest destination); NULL = natural order
    const int slot = blockIdx.x * 32 + lane;
    const int d = slot < n_out ? (order ? order[slot] : slot) : n_out;
    const long long t_blk = (long long)blockIdx.y * BT;
    const long long t0 = t_blk + warp * TT;
    const int r_begin = d < n_out ? row_ptr[d] : 0;
    const int nrows = d < n_out ? row_ptr[d + 1] - r_begin : 0;
    int maxrows = nrows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(0xffffffffu, maxrows, o);
        maxrows = other > maxrows ? other : maxrows;
    }
    // the source index of a row is fetched ONE ROW BEFORE its window is staged (the staging addresses depend on it: an
    // L2 round trip in front of the copies otherwise stalls the warp — 15 % of the stall samples before this)
    auto src_of = [&](int r) { return r < nrows ? src[(long long)r_begin + r] : -1; };
    int src_next = src_of(0);
    auto stage = [&](int r, int buf) {                 // this thread's share of row r: window rows / lags warp, warp+8, ...
        const bool has = r < nrows;
        const long long row = (long long)r_begin + r;
        const int my_src = src_next;
        src_next = src_of(r + 1);
        if (xt) {
            // time-contiguous copy of the input (hb_irf_transpose: xt[n][HP + t], zero padded): a warp fetches 32 consecutive
            // steps of one source per instruction — full 32-byte sectors instead of one 8-byte value per sector
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = warp + 8 * i;            // destination (lane index) whose source this warp stages
                const int sj = __shfl_sync(0xffffffffu, my_src, j);
                const double *xrow = xt + (long long)(sj < 0 ? 0 : sj) * Tp + t_blk;   // padded index of step t_blk - HP
                for (int m = lane; m < WP; m += 32) hb_cp_async8(&sx[buf][m][j], xrow + m, sj >= 0);
            }
        } else {
            const double *x = in + (has ? my_src : 0);
            for (int m = warp; m < WP; m += 8) {
                const long long t = t_blk - HP + m;
                const bool ok = has && t >= 0 && t < T;
                hb_cp_async8(&sx[buf][m][lane], ok ? x + t * n_in : in, ok);
            }
        }
        for (int l = warp; l < HP; l += 8) {
            const bool ok = has && l < H;
            hb_cp_async8(&sw[buf][l][lane], ok ? w + (long long)l * nnz + row : w, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double total[TT];
#pragma unroll
    for (int k = 0; k < TT; ++k) total[k] = 0.0;
    if (maxrows > 0) stage(0, 0);
    for (int r = 0; r < maxrows; ++r) {
        const int buf = r & 1;
        if (r + 1 < maxrows) {
            stage(r + 1, buf ^ 1);                     // buffer buf^1 was released by the barrier that ended row r-1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                               // row r has landed for every thread of the block
        if (r < nrows && t0 < T) {
            const int base = HP + warp * TT;           // window index of step t0
            double ring[TT], acc[TT];
#pragma unroll
            for (int k = 0; k < TT; ++k) {
                ring[k] = sx[buf][base + k][lane];
                acc[k] = 0.0;
            }
#pragma unroll
            for (int c = 0; c < HP / TT; ++c) {
#pragma unroll
                for (int j = 0; j < TT; ++j) {
                    const int l = c * TT + j;
                    if (l > 0) ring[(TT - j) % TT] = sx[buf][base - l][lane];   // lag l needs x[t0 - l] in place of x[t0 + TT - l]
                    const double wl = sw[buf][l][lane];
#pragma unroll
                    for (int k = 0; k < TT; ++k) acc[k] = fma(wl, ring[(k - j + TT) % TT], acc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < TT; ++k) total[k] += acc[k];
        }
        __syncthreads();                               // row r's buffer may be overwritten (by the stage of row r+2)
    }
    if (d < n_out) {
#pragma unroll
        for (int k = 0; k < TT; ++k)
            if (t0 + k < T) out[(t0 + k) * n_out + d] = total[k];
    }
}

// xt[n][lead + t] = in[t][n] for 0 <= t < T, zero elsewhere in [0, Tp): 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) hb_irf_transpose(const double *__restrict__ in, int n_in, long long T, int lead, long long Tp,
                                                        double *__restrict__ xt) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long tp0 = (long long)blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long t = tp0 + ty + 8 * k - lead;
        const int n = n0 + tx;
        tile[ty + 8 * k][tx] = (t >= 0 && t < T && n < n_in) ? in[t * n_in + n] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int n = n0 + ty + 8 * k;
        const long long tp = tp0 + tx;
        if (n < n_in && tp < Tp) xt[(long long)n * Tp + tp] = tile[tx][ty + 8 * k];
    }
}

// workspace of the transposed input (grown on demand, kept for the life of the process; cudaFree waits for kernels using it)
double *g_irf_ws = nullptr;
size_t g_irf_cap = 0;
std::mutex g_irf_mu;      // growth + the launches that use the workspace are one critical section (calls are serialised per process)

template <int HP, int TT>
int hb_launch_conv_tiled(cudaStream_t st, int n_out, int n_in, long long T, int H, long long nnz, const int *row_ptr,
                         const int *src, const double *w, const double *in, const int *order, double *out) {
    constexpr int BT = 8 * TT;
    const dim3 grid((unsigned)((n_out + 31) / 32), (unsigned)((T + BT - 1) / BT));
    const int smem = (2 * (BT + HP) * 33 + 2 * HP * 32) * 8;
    // per device and per context: set on every call (a host-side attribute write, negligible next to the launch)
    if (cudaFuncSetAttribute(hb_sparse_conv_tiled<HP, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return HB_ERR_CUDA;
    const double *xt = nullptr;
    std::lock_guard<std::mutex> lk(g_irf_mu);
    const long long Tp = (long long)grid.y * BT + HP;              // every block's window [t_blk, t_blk + BT + HP) in padded steps
    if (!getenv("HB_IRF_NO_TRANSPOSE") && (n_in + 31) / 32 <= 65535) {
        const size_t need = (size_t)n_in * (size_t)Tp * 8;
        if (need > g_irf_cap) {
            if (g_irf_ws) cudaFree(g_irf_ws);
            g_irf_ws = nullptr;
            g_irf_cap = 0;
            if (cudaMalloc(&g_irf_ws, need) == cudaSuccess) g_irf_cap = need;
            else cudaGetLastError();                               // no room for the copy: gather from the [T][n_in] input instead
        }
        if (g_irf_cap >= need) {
            dim3 tg((unsigned)(Tp / 32), (unsigned)((n_in + 31) / 32));
            hb_irf_transpose<<<tg, 256, 0, st>>>(in, n_in, T, HP, Tp, g_irf_ws);
            xt = g_irf_ws;
        }
    }
    hb_sparse_conv_tiled<HP, TT><<<grid, 256, smem, st>>>(n_out, n_in, T, H, nnz, row_ptr, src, w, in, xt, Tp, order, out);
    return HB_OK;
}
}  // namespace

extern "C" int hb_sparse_route_conv_device(int n_out, int n_in, int64_t T, int H, int64_t nnz, const int32_t *row_ptr,
                                           const int32_t *src, const double *weights, const double *input, double *out,
                                           void *stream) {
    return hb_sparse_route_conv_ordered_device(n_out, n_in, T, H, nnz, row_ptr, src, weights, input, nullptr, out, stream);
}

extern "C" int hb_sparse_route_conv_ordered_device(int n_out, int n_in, int64_t T, int H, int64_t nnz, const int32_t *row_ptr,
                                                   const int32_t *src, const double *weights, const double *input,
                                                   const int32_t *dest_order, double *out, void *stream) {
    if (n_out <= 0 || n_in <= 0 || T < 0 || H <= 0 || nnz < 0 || !row_ptr || !out) return HB_ERR_INVALID;
    if (T == 0) return HB_OK;
    const bool naive = H > 64 || getenv("HB_IRF_NAIVE") != nullptr;
    if (naive) {
        dim3 grid((unsigned)((n_out + 31) / 32), (unsigned)((T + 7) / 8));
        hb_sparse_conv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_out, n_in, T, H, nnz, row_ptr, src, weights, input, out);
    } else {
        const int32_t *order = getenv("HB_IRF_NO_ORDER") ? nullptr : dest_order;
        // time steps per thread: 16 (128-step tiles, 100 / 134 KB of shared memory: 2 / 1 blocks per SM) or 8 (64-step tiles, 67 / 100 KB:
        // 3 / 2 blocks per SM) — HB_IRF_TT overrides the default
        const int tt = getenv("HB_IRF_TT") ? atoi(getenv("HB_IRF_TT")) : HB_IRF_TT;
        cudaStream_t st = (cudaStream_t)stream;
        int rc;
        if (H <= 32) rc = tt == 8 ? hb_launch_conv_tiled<32, 8>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out)
                        : tt == 32 ? hb_launch_conv_tiled<32, 32>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out)
                                   : hb_launch_conv_tiled<32, 16>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out);
        else rc = tt == 8 ? hb_launch_conv_tiled<64, 8>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out)
                : tt == 32 ? hb_launch_conv_tiled<64, 32>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out)
                           : hb_launch_conv_tiled<64, 16>(st, n_out, n_in, T, H, nnz, row_ptr, src, weights, input, order, out);
        if (rc) return rc;
    }
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// =============================================================================================
// IRF kernel COMPOSITION on the device: aggregate_route_kernel (/root/reference/src/route.jl:523-577).
// The reference walks the nodes in topological order and, for every upstream neighbour u of a node and every source s
// already known at u, adds  edge_weight * truncated_conv(self_kernel(node), K[u][s])  to K[node][s]
// (_truncated_route_convolution, :506-521: out[ia+ib-1] += a[ia] * b[ib], ia ascending, zero a[ia] skipped, plain multiply
// then add).  Here the host derives the row structure (Kahn levels over the sparse upstream lists, one output row per
// (node, source), its contributions listed by ascending upstream id = the reference's accumulation order) and the
// arithmetic runs level by level: one warp per output row, lane k owns lag k.  Same operations in the same order with
// explicitly rounded multiplies and adds, so the weights equal the reference's bit for bit.
// =============================================================================================
namespace {

__global__ void __launch_bounds__(256) hb_irf_compose_kernel(long long row0, long long n_rows, const int *__restrict__ row_dest,
                                                             const long long *__restrict__ contrib_ptr, const long long *__restrict__ contrib_row,
                                                             const double *__restrict__ contrib_ew, const double *__restrict__ node_kernels,
                                                             double *weights, int H) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = row0 + (long long)blockIdx.x * 8 + warp;
    if (r >= row0 + n_rows) return;
    double *a = sm + (size_t)warp * 3 * H, *b = a + H, *acc = b + H;
    const double *selfk = node_kernels + (long long)row_dest[r] * H;
    for (int k = lane; k < H; k += 32) a[k] = selfk[k];
    const long long c0 = contrib_ptr[r], c1 = contrib_ptr[r + 1];
    if (c0 == c1) {                                             // the node's own entry: pairwise[node][node] = self kernel
        __syncwarp();
        for (int k = lane; k < H; k += 32) weights[r * H + k] = a[k];
        return;
    }
    for (long long c = c0; c < c1; ++c) {
        const double *up = weights + contrib_row[c] * H;        // finished in an earlier level (earlier launch)
        __syncwarp();
        for (int k = lane; k < H; k += 32) b[k] = up[k];
        __syncwarp();
        const double ew = contrib_ew[c];
        for (int k = lane; k < H; k += 32) {
            double o = 0.0;
            for (int ia = 0; ia <= k; ++ia) {
                const double av = a[ia];
                if (av != 0.0) o = __dadd_rn(o, __dmul_rn(av, b[k - ia]));
            }
            if (ew != 1.0) o = __dmul_rn(o, ew);
            acc[k] = c == c0 ? o : __dadd_rn(acc[k], o);
        }
    }
    __syncwarp();
    for (int k = lane; k < H; k += 32) weights[r * H + k] = acc[k];
}

// out row i = w row perm[i]; transpose != 0 writes out[h][i] (the layout hb_sparse_route_conv_device takes), else out[i][h]
__global__ void __launch_bounds__(256) hb_irf_gather_rows_kernel(const double *__restrict__ w, const long long *__restrict__ perm, long long nnz,
                                                                 int H, int transpose, double *__restrict__ out) {
    __shared__ double tile[32][33];
    const long long i0 = (long long)blockIdx.x * 32;
    const int h0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const long long i = i0 + j;
        if (i < nnz && h0 + tx < H) tile[j][tx] = w[(perm ? perm[i] : i) * H + h0 + tx];
    }
    __syncthreads();
    if (transpose) {
        for (int j = ty; j < 32; j += 8) {
            const long long i = i0 + tx;
            if (i < nnz && h0 + j < H) out[(long long)(h0 + j) * nnz + i] = tile[tx][j];
        }
    } else {
        for (int j = ty; j < 32; j += 8) {
            const long long i = i0 + j;
            if (i < nnz && h0 + tx < H) out[i * H + h0 + tx] = tile[j][tx];
        }
    }
}

}  // namespace

namespace {
// out[t][d] = part[t][seg_ptr[d]] + part[t][seg_ptr[d] + 1] + ... in that order (a destination whose kernel rows were split
// into segments, each convolved as a "virtual destination" by a different block)
__global__ void __launch_bounds__(256) hb_irf_reduce_segments_kernel(int n_out, long long T, int n_virtual, const int *__restrict__ seg_ptr,
                                                                     const double *__restrict__ part, double *__restrict__ out) {
    const int d = blockIdx.x * 32 + (threadIdx.x & 31);
    const long long t = (long long)blockIdx.y * 8 + (threadIdx.x >> 5);
    if (d >= n_out || t >= T) return;
    const double *p = part + t * n_virtual;
    const int v0 = seg_ptr[d], v1 = seg_ptr[d + 1];
    double s = v0 < v1 ? p[v0] : 0.0;
    for (int v = v0 + 1; v < v1; ++v) s += p[v];
    out[t * n_out + d] = s;
}
}  // namespace

extern "C" int hb_irf_reduce_segments_device(int32_t n_out, int64_t n_steps, int32_t n_virtual, const int32_t *seg_ptr, const double *part,
                                             double *out, void *stream) {
    if (n_out <= 0 || n_steps < 0 || n_virtual < n_out || !seg_ptr || !part || !out) return HB_ERR_INVALID;
    if (n_steps == 0) return HB_OK;
    dim3 grid((unsigned)((n_out + 31) / 32), (unsigned)((n_steps + 7) / 8));
    hb_irf_reduce_segments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_out, n_steps, n_virtual, seg_ptr, part, out);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_irf_compose_device(int32_t n_levels, const int64_t *level_row_ptr, const int32_t *row_dest, const int64_t *contrib_ptr,
                                     const int64_t *contrib_row, const double *contrib_ew, const double *node_kernels, double *weights,
                                     int32_t horizon, void *stream) {
    if (n_levels < 0 || !level_row_ptr || !row_dest || !contrib_ptr || !node_kernels || !weights || horizon < 1 || horizon > 1024) return HB_ERR_INVALID;
    const int smem = 8 * 3 * horizon * 8;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(hb_irf_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return HB_ERR_CUDA;
    for (int l = 0; l < n_levels; ++l) {
        const long long r0 = level_row_ptr[l], n = level_row_ptr[l + 1] - r0;
        if (n <= 0) continue;
        hb_irf_compose_kernel<<<(unsigned)((n + 7) / 8), 256, smem, (cudaStream_t)stream>>>(
            r0, n, row_dest, (const long long *)contrib_ptr, (const long long *)contrib_row, contrib_ew, node_kernels, weights, horizon);
    }
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_irf_gather_rows_device(const double *w, const int64_t *perm, int64_t nnz, int32_t horizon, int32_t transpose, double *out,
                                         void *stream) {
    if (!w || !out || nnz < 0 || horizon < 1) return HB_ERR_INVALID;
    if (nnz == 0) return HB_OK;
    dim3 grid((unsigned)((nnz + 31) / 32), (unsigned)((horizon + 31) / 32));
    hb_irf_gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, (const long long *)perm, nnz, horizon, transpose, out);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// =============================================================================================
// Raster ingest (/root/reference/ext/HydroModelsRastersExt):
//
// hb_d8_flow_direction — compute_d8_flow_direction (HydroModelsRastersExt.jl:32-109): every cell drains to the
//   neighbour with the steepest downward slope (center - neighbour) / distance; neighbours visited in the order
//   E, SE, S, SW, W, NW, N, NE, a later neighbour wins only on a strictly greater slope (max starts at -Inf, so a
//   pit still points at its least-uphill neighbour); NaN cells get 0 and are skipped as neighbours; diagonal
//   distance sqrt(cx^2+cy^2), orthogonal distance max(cx, cy).  One thread per cell, a (32+2) x (8+2) tile of the
//   DEM staged in shared memory so every DEM value is read from HBM once (8 B in, 1 B out per cell).
//
// hb_pack_forcing — load_netcdf_data + the assembly of the model input (netcdf_loader.jl:29-62, :119-136): gather
//   V rasters [T][.][.] at the node positions into the stepper's forcing layout [T][N][V] (= Julia input[V,N,T]).
//   One thread per (node, step); the V values of a node are written as one contiguous run.
// =============================================================================================
namespace {

constexpr int kD8TileW = 32, kD8TileH = 8;

__global__ void __launch_bounds__(kD8TileW *kD8TileH) hb_d8_kernel(const double *__restrict__ dem, int rows, int cols, long long row_stride,
                                                                   long long col_stride, double dist_orth, double dist_diag,
                                                                   unsigned char *__restrict__ out) {
    __shared__ double tile[kD8TileH + 2][kD8TileW + 2];
    const int c0 = blockIdx.x * kD8TileW, r0 = blockIdx.y * kD8TileH;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = threadIdx.x; i < (kD8TileH + 2) * (kD8TileW + 2); i += kD8TileW * kD8TileH) {
        const int lr = i / (kD8TileW + 2), lc = i % (kD8TileW + 2);
        const int r = r0 + lr - 1, c = c0 + lc - 1;
        tile[lr][lc] = (r >= 0 && r < rows && c >= 0 && c < cols) ? dem[r * row_stride + c * col_stride] : nan;   // off-grid == NoData
    }
    __syncthreads();
    const int lc = threadIdx.x % kD8TileW, lr = threadIdx.x / kD8TileW;
    const int r = r0 + lr, c = c0 + lc;
    if (r >= rows || c >= cols) return;
    const double center = tile[lr + 1][lc + 1];
    int code = 0;
    if (center == center) {
        const int dr[8] = {0, 1, 1, 1, 0, -1, -1, -1};
        const int dc[8] = {1, 1, 0, -1, -1, -1, 0, 1};
        double best = __longlong_as_double(0xfff0000000000000ll);   // -Inf
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double nb = tile[lr + 1 + dr[k]][lc + 1 + dc[k]];
            if (nb == nb) {
                const double slope = (center - nb) / ((k & 1) ? dist_diag : dist_orth);
                if (slope > best) {
                    best = slope;
                    code = 1 << k;
                }
            }
        }
    }
    out[(long long)r * cols + c] = (unsigned char)code;
}

__global__ void __launch_bounds__(256) hb_pack_kernel(const double *const *__restrict__ vars, int n_vars, long long T, long long t_stride,
                                                      long long row_stride, long long col_stride, const int *__restrict__ node_row,
                                                      const int *__restrict__ node_col, long long N, double *__restrict__ out) {
    const long long n = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long t = blockIdx.y;
    if (n >= N) return;
    const long long off = t * t_stride + node_row[n] * row_stride + node_col[n] * col_stride;
    double *o = out + (t * N + n) * n_vars;
    for (int v = 0; v < n_vars; ++v) o[v] = vars[v][off];
}

// Dense variant: the nodes are ALL cells of the raster in row-major order (n = r * cols + c) — no position arrays to re-read
// every step, and a 32 x 32 tile per variable goes through shared memory so that BOTH sides are coalesced whatever the
// raster's memory order: reads run along its fastest axis, writes are runs of 32 * V consecutive doubles of out[t][n][v].
__global__ void __launch_bounds__(256) hb_pack_dense_kernel(const double *const *__restrict__ vars, int V, long long t_stride,
                                                            long long row_stride, long long col_stride, int rows, int cols,
                                                            double *__restrict__ out) {
    extern __shared__ double hb_pack_smem[];
    constexpr int VP = 32 * 33 + 1;                    // doubles per variable tile (odd pitches: conflict-free both ways)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const long long t = blockIdx.z;
    const bool rows_fastest = row_stride < col_stride;
    for (int v = 0; v < V; ++v) {
        const double *src = vars[v] + t * t_stride;
        double *tile = hb_pack_smem + v * VP;          // tile[r_local * 33 + c_local]
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int a = ty + 8 * k;
            const int rl = rows_fastest ? tx : a, cl = rows_fastest ? a : tx;
            const int r = r0 + rl, c = c0 + cl;
            if (r < rows && c < cols) tile[rl * 33 + cl] = src[r * row_stride + c * col_stride];
        }
    }
    __syncthreads();
    const long long N = (long long)rows * cols;
    const int ncol = cols - c0 < 32 ? cols - c0 : 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rl = ty + 8 * k, r = r0 + rl;
        if (r < rows) {
            double *o = out + (t * N + (long long)r * cols + c0) * V;
            for (int j = tx; j < ncol * V; j += 32) {
                const int cl = j / V, v = j - cl * V;
                o[j] = hb_pack_smem[v * VP + rl * 33 + cl];
            }
        }
    }
}

}  // namespace

extern "C" int hb_d8_flow_direction_device(const double *dem, int rows, int cols, int64_t row_stride, int64_t col_stride,
                                           double cellsize_x, double cellsize_y, uint8_t *flow_dir, void *stream) {
    if (!dem || !flow_dir || rows <= 0 || cols <= 0 || !(cellsize_x > 0) || !(cellsize_y > 0)) return HB_ERR_INVALID;
    const double orth = cellsize_x > cellsize_y ? cellsize_x : cellsize_y;          // max(cellsize_x, cellsize_y)
    const double diag = sqrt(cellsize_x * cellsize_x + cellsize_y * cellsize_y);
    dim3 grid((unsigned)((cols + kD8TileW - 1) / kD8TileW), (unsigned)((rows + kD8TileH - 1) / kD8TileH));
    hb_d8_kernel<<<grid, kD8TileW * kD8TileH, 0, (cudaStream_t)stream>>>(dem, rows, cols, row_stride, col_stride, orth, diag, flow_dir);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_pack_forcing_device(const double *const *vars_device_array, int n_vars, int64_t n_steps, int64_t t_stride,
                                      int64_t row_stride, int64_t col_stride, const int32_t *node_row, const int32_t *node_col,
                                      int64_t n_nodes, double *out, void *stream) {
    if (!vars_device_array || n_vars <= 0 || n_steps < 0 || n_nodes <= 0 || !node_row || !node_col || !out) return HB_ERR_INVALID;
    if (n_steps == 0) return HB_OK;
    if (n_steps > 65535) return HB_ERR_INVALID;   // one grid row per step; callers chunk the time axis
    dim3 grid((unsigned)((n_nodes + 255) / 256), (unsigned)n_steps);
    hb_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(vars_device_array, n_vars, n_steps, t_stride, row_stride, col_stride, node_row,
                                                          node_col, n_nodes, out);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_pack_forcing_dense_device(const double *const *vars_device_array, int n_vars, int64_t n_steps, int64_t t_stride,
                                            int64_t row_stride, int64_t col_stride, int rows, int cols, double *out, void *stream) {
    if (!vars_device_array || n_vars <= 0 || n_vars > 5 || n_steps < 0 || rows <= 0 || cols <= 0 || !out) return HB_ERR_INVALID;
    if (n_steps == 0) return HB_OK;
    if (n_steps > 65535 || (rows + 31) / 32 > 65535) return HB_ERR_INVALID;   // grid.z = steps; callers chunk the time axis
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)n_steps);
    const int smem = n_vars * (32 * 33 + 1) * 8;       // <= 42 KB for 5 variables: inside the default dynamic limit
    hb_pack_dense_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(vars_device_array, n_vars, t_stride, row_stride, col_stride, rows, cols, out);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// =============================================================================================
// On-device calibration: differential evolution DE/rand/1/bin around the ensemble
// objective kernel (hb_stepper, objective mode).  The reference builds an Optimization.jl problem whose objective
// runs the model once per candidate on the host (ext/HydroModelsOptimizationExt/...jl:153-175) and hands it to
// BlackBoxOptim's DE (docs/src/tutorials/optimimal_parameters.md:189-195); here a whole population is one launch
// of the ensemble kernel and mutation / crossover / selection are two small kernels in between, so a generation
// never leaves the device.  Populations are stored exactly as the stepper reads its parameters: dimension d of
// member i at base[dim_off[d] + i] (parameter-major, member-fastest).
// Random numbers are counter based (SplitMix64 of seed, generation, member, draw index), so the sequence does not
// depend on the launch geometry and the numpy restatement in oracle/ reproduces it bit for bit; the arithmetic
// uses explicit round-to-nearest operations (no FMA contraction) for the same reason.
// =============================================================================================
namespace {

__device__ __forceinline__ double hb_de_uniform(unsigned long long seed, unsigned gen, unsigned member, unsigned k) {
    unsigned long long c = ((unsigned long long)gen << 40) ^ ((unsigned long long)member << 8) ^ (unsigned long long)k;
    unsigned long long z = seed + (c + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * 1.1102230246251565e-16;   // [0, 1)
}

__global__ void __launch_bounds__(256) hb_de_trial_kernel(const double *__restrict__ pop, double *__restrict__ trial,
                                                          const long long *__restrict__ dim_off, int D, int N,
                                                          const double *__restrict__ lb, const double *__restrict__ ub, double F,
                                                          double CR, unsigned long long seed, unsigned gen) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    // three distinct members, all different from i: offsets o1, o2, o3 in 1..N-1, pairwise distinct
    int o1 = 1 + (int)(hb_de_uniform(seed, gen, i, 0) * (N - 1));
    int o2 = 1 + (int)(hb_de_uniform(seed, gen, i, 1) * (N - 2));
    int o3 = 1 + (int)(hb_de_uniform(seed, gen, i, 2) * (N - 3));
    if (o1 > N - 1) o1 = N - 1;
    if (o2 > N - 2) o2 = N - 2;
    if (o3 > N - 3) o3 = N - 3;
    if (o2 >= o1) ++o2;
    const int lo = o1 < o2 ? o1 : o2, hi = o1 < o2 ? o2 : o1;
    if (o3 >= lo) ++o3;
    if (o3 >= hi) ++o3;
    const int r1 = (i + o1) % N, r2 = (i + o2) % N, r3 = (i + o3) % N;
    int jrand = (int)(hb_de_uniform(seed, gen, i, 3) * D);
    if (jrand > D - 1) jrand = D - 1;
    for (int d = 0; d < D; ++d) {
        const double *p = pop + dim_off[d];
        double v = p[i];
        if (hb_de_uniform(seed, gen, i, 4 + d) < CR || d == jrand) v = __dadd_rn(p[r1], __dmul_rn(F, __dsub_rn(p[r2], p[r3])));
        v = v < lb[d] ? lb[d] : v;
        v = v > ub[d] ? ub[d] : v;
        trial[dim_off[d] + i] = v;
    }
}

__global__ void __launch_bounds__(256) hb_de_select_kernel(double *__restrict__ pop, const double *__restrict__ trial, double *__restrict__ fit,
                                                           const double *__restrict__ trial_fit, const long long *__restrict__ dim_off, int D,
                                                           int N) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    // a NaN trial never replaces a member; a member whose own objective is NaN (KGE of a constant series, log of a
    // negative flow at generation 0) is replaced by any trial with a real objective, otherwise it would stay for the whole run
    const double tf = trial_fit[i], f = fit[i];
    if (tf <= f || (f != f && tf == tf)) {
        fit[i] = trial_fit[i];
        for (int d = 0; d < D; ++d) pop[dim_off[d] + i] = trial[dim_off[d] + i];
    }
}

// best (lowest) objective of the population and its first index -> hist[2*slot], hist[2*slot+1]; one block
__global__ void __launch_bounds__(1024) hb_de_best_kernel(const double *__restrict__ fit, int N, double *__restrict__ hist, int slot) {
    __shared__ double sv[1024];
    __shared__ int si[1024];
    double bv = __longlong_as_double(0x7ff0000000000000ll);
    int bi = -1;
    for (int i = threadIdx.x; i < N; i += 1024) {
        const double v = fit[i];
        if (v < bv) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double v = sv[threadIdx.x + s];
            const int j = si[threadIdx.x + s];
            if (j >= 0 && (v < sv[threadIdx.x] || si[threadIdx.x] < 0 || (v == sv[threadIdx.x] && j < si[threadIdx.x]))) {
                sv[threadIdx.x] = v;
                si[threadIdx.x] = j;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        hist[2 * slot] = sv[0];
        hist[2 * slot + 1] = (double)si[0];
    }
}

}  // namespace

// ---- population sharded over the ranks of a communicator ("parameter sets sharded; no communication beyond an
// NCCL gather") --------------------------------------------------------------------------------------------------------
// Global member g lives on rank g / n_local at local index g % n_local.  Every rank holds the WHOLE population as
// pop_all[world][D][n_local] — the ranks' blocks, refreshed by one all-gather per generation (hb_comm_all_gather, D * N * 8
// bytes: 3 MB for 65 536 members x 6 parameters) — because DE/rand/1 draws its parents from all members; trials, objectives
// and the selection stay local.  Draws are keyed by the GLOBAL member index, so the trajectory is the single-GPU one, bit for bit.
namespace {

__global__ void __launch_bounds__(256) hb_de_trial_sharded_kernel(const double *__restrict__ pop_all, double *__restrict__ trial,
                                                                  const long long *__restrict__ dim_off, int D, int n_local, int N, int member0,
                                                                  const double *__restrict__ lb, const double *__restrict__ ub, double F,
                                                                  double CR, unsigned long long seed, unsigned gen) {
    const int il = blockIdx.x * 256 + threadIdx.x;
    if (il >= n_local) return;
    const int i = member0 + il;
    int o1 = 1 + (int)(hb_de_uniform(seed, gen, i, 0) * (N - 1));
    int o2 = 1 + (int)(hb_de_uniform(seed, gen, i, 1) * (N - 2));
    int o3 = 1 + (int)(hb_de_uniform(seed, gen, i, 2) * (N - 3));
    if (o1 > N - 1) o1 = N - 1;
    if (o2 > N - 2) o2 = N - 2;
    if (o3 > N - 3) o3 = N - 3;
    if (o2 >= o1) ++o2;
    const int lo = o1 < o2 ? o1 : o2, hi = o1 < o2 ? o2 : o1;
    if (o3 >= lo) ++o3;
    if (o3 >= hi) ++o3;
    const int r1 = (i + o1) % N, r2 = (i + o2) % N, r3 = (i + o3) % N;
    int jrand = (int)(hb_de_uniform(seed, gen, i, 3) * D);
    if (jrand > D - 1) jrand = D - 1;
    const size_t blk = (size_t)D * n_local;
    auto at = [&](int g, int d) { return pop_all[(size_t)(g / n_local) * blk + (size_t)d * n_local + (g % n_local)]; };
    for (int d = 0; d < D; ++d) {
        double v = at(i, d);
        if (hb_de_uniform(seed, gen, i, 4 + d) < CR || d == jrand) v = __dadd_rn(at(r1, d), __dmul_rn(F, __dsub_rn(at(r2, d), at(r3, d))));
        v = v < lb[d] ? lb[d] : v;
        v = v > ub[d] ? ub[d] : v;
        trial[dim_off[d] + il] = v;
    }
}

__global__ void __launch_bounds__(256) hb_de_select_sharded_kernel(double *__restrict__ pop_block, const double *__restrict__ trial,
                                                                   double *__restrict__ fit, const double *__restrict__ trial_fit,
                                                                   const long long *__restrict__ dim_off, int D, int n_local) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_local) return;
    const double tf = trial_fit[i], f = fit[i];
    if (tf <= f || (f != f && tf == tf)) {
        fit[i] = tf;
        for (int d = 0; d < D; ++d) pop_block[(size_t)d * n_local + i] = trial[dim_off[d] + i];
    }
}

}  // namespace

extern "C" int hb_de_trial_sharded_device(const double *pop_all, double *trial, const int64_t *dim_offset, int n_dims, int n_local, int n_total,
                                          int member0, const double *lower, const double *upper, double F, double CR, uint64_t seed,
                                          uint32_t generation, void *stream) {
    if (!pop_all || !trial || !dim_offset || !lower || !upper || n_dims <= 0 || n_local <= 0 || n_total < 4 || n_total % n_local ||
        member0 < 0 || member0 % n_local || member0 + n_local > n_total)
        return HB_ERR_INVALID;
    hb_de_trial_sharded_kernel<<<(n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop_all, trial, (const long long *)dim_offset, n_dims,
                                                                                      n_local, n_total, member0, lower, upper, F, CR, seed,
                                                                                      generation);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_de_select_sharded_device(double *pop_block, const double *trial, double *fitness, const double *trial_fitness,
                                           const int64_t *dim_offset, int n_dims, int n_local, double *history, int history_slot,
                                           void *stream) {
    if (!pop_block || !trial || !fitness || !trial_fitness || !dim_offset || n_dims <= 0 || n_local <= 0 || fitness == trial_fitness)
        return HB_ERR_INVALID;
    hb_de_select_sharded_kernel<<<(n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop_block, trial, fitness, trial_fitness,
                                                                                       (const long long *)dim_offset, n_dims, n_local);
    if (history) hb_de_best_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(fitness, n_local, history, history_slot);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

// ---- adaptive DE: BlackBoxOptim.jl's `adaptive_de_rand_1_bin_radiuslimited` (the optimiser of every example of the
// reference: ext/HydroModelsOptimizationExt/examples.jl:28, docs/src/tutorials/optimimal_parameters.md:189-195; the package
// itself is a dependency that is not vendored under /root/reference — its published operators are restated here),
// made generation-synchronous: all members are targets once per generation instead of one per step.
//   * every member carries its own F and CR; a member whose trial did not improve it draws new ones from the bimodal Cauchy
//     mixtures  F ~ {C(0.65, 0.1) | C(1.0, 0.1)},  CR ~ {C(0.1, 0.1) | C(0.95, 0.1)}  (above 1 -> 1, below 0 -> redrawn);
//   * the three parents come from a window of `radius` consecutive members (ring order) that contains the target;
//   * a mutant component outside its bounds lands at a random point between the target's value and the bound.
// A standard Cauchy variate is X / Y of a point uniform in the unit disk (rejection from the square): only IEEE + - * /, so
// the numpy restatement reproduces every draw bit for bit (a tangent would not).
namespace {

constexpr unsigned kAdeDrawParam = 1u << 20;     // draw indices of the parameter redraws (trial draws stay below)

__device__ double hb_ade_cauchy01(unsigned long long seed, unsigned gen, unsigned member, unsigned k0, double loc1, double loc2, double scale) {
    // up to 8 attempts of (mixture pick, disk point); 16 draws each
    for (unsigned a = 0; a < 8; ++a) {
        const unsigned k = k0 + 64u * a;
        const double loc = hb_de_uniform(seed, gen, member, k) < 0.5 ? loc1 : loc2;
        double c = 0.0;
        bool ok = false;
        for (unsigned r = 0; r < 16 && !ok; ++r) {
            const double x = __dsub_rn(__dmul_rn(2.0, hb_de_uniform(seed, gen, member, k + 1 + 2 * r)), 1.0);
            const double y = __dsub_rn(__dmul_rn(2.0, hb_de_uniform(seed, gen, member, k + 2 + 2 * r)), 1.0);
            if (__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)) <= 1.0 && y != 0.0) { c = __ddiv_rn(x, y); ok = true; }
        }
        const double v = __dadd_rn(loc, __dmul_rn(scale, c));
        if (ok && v >= 0.0) return v > 1.0 ? 1.0 : v;
    }
    return loc1;
}

__global__ void __launch_bounds__(256) hb_ade_params_kernel(double *__restrict__ f, double *__restrict__ cr, const unsigned char *__restrict__ redraw,
                                                            int N, unsigned long long seed, unsigned gen) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N || (redraw && !redraw[i])) return;
    f[i] = hb_ade_cauchy01(seed, gen, i, kAdeDrawParam, 0.65, 1.0, 0.1);
    cr[i] = hb_ade_cauchy01(seed, gen, i, kAdeDrawParam + 1024u, 0.1, 0.95, 0.1);
}

__global__ void __launch_bounds__(256) hb_ade_trial_kernel(const double *__restrict__ pop, double *__restrict__ trial,
                                                           const long long *__restrict__ dim_off, int D, int N,
                                                           const double *__restrict__ lb, const double *__restrict__ ub,
                                                           const double *__restrict__ f, const double *__restrict__ cr, int radius,
                                                           unsigned long long seed, unsigned gen) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    // the window [i - w, i - w + radius) holds the target at position w; three distinct other positions
    int w = (int)(hb_de_uniform(seed, gen, i, 0) * radius);
    if (w > radius - 1) w = radius - 1;
    const int M = radius - 1;
    int o1 = (int)(hb_de_uniform(seed, gen, i, 1) * M);
    int o2 = (int)(hb_de_uniform(seed, gen, i, 2) * (M - 1));
    int o3 = (int)(hb_de_uniform(seed, gen, i, 3) * (M - 2));
    if (o1 > M - 1) o1 = M - 1;
    if (o2 > M - 2) o2 = M - 2;
    if (o3 > M - 3) o3 = M - 3;
    if (o2 >= o1) ++o2;
    const int lo = o1 < o2 ? o1 : o2, hi = o1 < o2 ? o2 : o1;
    if (o3 >= lo) ++o3;
    if (o3 >= hi) ++o3;
    o1 += o1 >= w;
    o2 += o2 >= w;
    o3 += o3 >= w;
    const int base = i - w + N;
    const int r1 = (base + o1) % N, r2 = (base + o2) % N, r3 = (base + o3) % N;
    int jrand = (int)(hb_de_uniform(seed, gen, i, 4) * D);
    if (jrand > D - 1) jrand = D - 1;
    const double F = f[i], CR = cr[i];
    for (int d = 0; d < D; ++d) {
        const double *p = pop + dim_off[d];
        const double own = p[i];
        double v = own;
        if (hb_de_uniform(seed, gen, i, 8 + 2 * d) < CR || d == jrand) {
            v = __dadd_rn(p[r1], __dmul_rn(F, __dsub_rn(p[r2], p[r3])));
            const double u = hb_de_uniform(seed, gen, i, 9 + 2 * d);
            if (v < lb[d]) v = __dadd_rn(own, __dmul_rn(u, __dsub_rn(lb[d], own)));
            else if (v > ub[d]) v = __dadd_rn(own, __dmul_rn(u, __dsub_rn(ub[d], own)));
        }
        trial[dim_off[d] + i] = v;
    }
}

// strict improvement (or a real-valued trial against a NaN incumbent) replaces the member; everyone else redraws F and CR
__global__ void __launch_bounds__(256) hb_ade_select_kernel(double *__restrict__ pop, const double *__restrict__ trial, double *__restrict__ fit,
                                                            const double *__restrict__ trial_fit, const long long *__restrict__ dim_off, int D,
                                                            int N, double *__restrict__ f, double *__restrict__ cr, unsigned long long seed,
                                                            unsigned gen) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const double tf = trial_fit[i], fi = fit[i];
    if (tf < fi || (fi != fi && tf == tf)) {
        fit[i] = tf;
        for (int d = 0; d < D; ++d) pop[dim_off[d] + i] = trial[dim_off[d] + i];
    } else {
        f[i] = hb_ade_cauchy01(seed, gen, i, kAdeDrawParam, 0.65, 1.0, 0.1);
        cr[i] = hb_ade_cauchy01(seed, gen, i, kAdeDrawParam + 1024u, 0.1, 0.95, 0.1);
    }
}

}  // namespace

extern "C" int hb_ade_init_device(double *f, double *cr, int n_members, uint64_t seed, void *stream) {
    if (!f || !cr || n_members <= 0) return HB_ERR_INVALID;
    hb_ade_params_kernel<<<(n_members + 255) / 256, 256, 0, (cudaStream_t)stream>>>(f, cr, nullptr, n_members, seed, 0u);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_ade_trial_device(const double *pop, double *trial, const int64_t *dim_offset, int n_dims, int n_members,
                                   const double *lower, const double *upper, const double *f, const double *cr, int radius, uint64_t seed,
                                   uint32_t generation, void *stream) {
    if (!pop || !trial || !dim_offset || !lower || !upper || !f || !cr || n_dims <= 0 || radius < 4 || n_members < radius || n_dims > (1 << 18))
        return HB_ERR_INVALID;
    hb_ade_trial_kernel<<<(n_members + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop, trial, (const long long *)dim_offset, n_dims, n_members,
                                                                                 lower, upper, f, cr, radius, seed, generation);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_ade_select_device(double *pop, const double *trial, double *fitness, const double *trial_fitness,
                                    const int64_t *dim_offset, int n_dims, int n_members, double *f, double *cr, uint64_t seed,
                                    uint32_t generation, double *history, int history_slot, void *stream) {
    if (!pop || !trial || !fitness || !trial_fitness || !dim_offset || !f || !cr || n_dims <= 0 || n_members <= 0) return HB_ERR_INVALID;
    if (pop == trial || fitness == trial_fitness) return HB_ERR_INVALID;
    hb_ade_select_kernel<<<(n_members + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop, trial, fitness, trial_fitness,
                                                                                  (const long long *)dim_offset, n_dims, n_members, f, cr, seed,
                                                                                  generation);
    if (history) hb_de_best_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(fitness, n_members, history, history_slot);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_de_trial_device(const double *pop, double *trial, const int64_t *dim_offset, int n_dims, int n_members,
                                  const double *lower, const double *upper, double F, double CR, uint64_t seed, uint32_t generation,
                                  void *stream) {
    if (!pop || !trial || !dim_offset || !lower || !upper || n_dims <= 0 || n_members < 4) return HB_ERR_INVALID;
    hb_de_trial_kernel<<<(n_members + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop, trial, (const long long *)dim_offset, n_dims, n_members,
                                                                                lower, upper, F, CR, seed, generation);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_de_best_device(const double *fitness, int n_members, double *history, int history_slot, void *stream) {
    if (!fitness || !history || n_members <= 0 || history_slot < 0) return HB_ERR_INVALID;
    hb_de_best_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(fitness, n_members, history, history_slot);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}

extern "C" int hb_de_select_device(double *pop, const double *trial, double *fitness, const double *trial_fitness,
                                   const int64_t *dim_offset, int n_dims, int n_members, double *history, int history_slot, void *stream) {
    if (!pop || !trial || !fitness || !trial_fitness || !dim_offset || n_dims <= 0 || n_members <= 0) return HB_ERR_INVALID;
    if (pop == trial || fitness == trial_fitness) return HB_ERR_INVALID;      // the kernel's arguments are __restrict__
    hb_de_select_kernel<<<(n_members + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pop, trial, fitness, trial_fitness,
                                                                                 (const long long *)dim_offset, n_dims, n_members);
    if (history) hb_de_best_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(fitness, n_members, history, history_slot);
    return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
}
