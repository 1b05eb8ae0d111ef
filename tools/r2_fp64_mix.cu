// Does a DMMA / DFMA hold the sub-partition's issue port, or only the FP64 pipe?  One warp per sub-partition (and three) runs
// 8 independent DMMA (or DFMA) chains interleaved with K independent FFMA + K IMAD per FP64 instruction; if the other pipes
// issue underneath the FP64 instruction the time stays flat until K reaches the FP64 instruction's pipe occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int K, bool MMA>
__global__ void k_mix(double *o, int it, double a, double b, float fa, int ia) {
    double c[8][2];
    float f[8]; int n[8];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; f[i] = i + threadIdx.x; n[i] = i * threadIdx.x; }
    for (int k = 0; k < it; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MMA) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i][0]) : "d"(a), "d"(b));
#pragma unroll
            for (int j = 0; j < K; ++j) {
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[(i + j) & 7]) : "f"(fa));
                asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(n[(i + j) & 7]) : "r"(ia));
            }
        }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + f[i] + n[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
template <int K, bool MMA> void run(double *o, int sms, int warps, double clocks) {
    const int it = 4000;
    float t = timeit([&] { k_mix<K, MMA><<<sms, warps * 32>>>(o, it, 1.0000001, 1e-9, 1.0000001f, 3); });
    printf("  %s + %2d FFMA + %2d IMAD each, %2d warps/SM: %.1f clk per FP64 instruction per sub-partition\n", MMA ? "DMMA" : "DFMA", K, K, warps,
           t * 1e-3 * clocks / ((double)it * 8 * warps / 4.0));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clocks = clk_khz * 1e3;
    double *o; cudaMalloc(&o, sizeof(double) * p.multiProcessorCount * 1024);
    for (int warps = 4; warps <= 12; warps += 8) {
        run<0, true>(o, p.multiProcessorCount, warps, clocks);
        run<2, true>(o, p.multiProcessorCount, warps, clocks);
        run<4, true>(o, p.multiProcessorCount, warps, clocks);
        run<6, true>(o, p.multiProcessorCount, warps, clocks);
        run<8, true>(o, p.multiProcessorCount, warps, clocks);
        run<12, true>(o, p.multiProcessorCount, warps, clocks);
        run<0, false>(o, p.multiProcessorCount, warps, clocks);
        run<1, false>(o, p.multiProcessorCount, warps, clocks);
        run<2, false>(o, p.multiProcessorCount, warps, clocks);
        run<4, false>(o, p.multiProcessorCount, warps, clocks);
    }
    return 0;
}
