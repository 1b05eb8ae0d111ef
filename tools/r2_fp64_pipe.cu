// FP64 pipe micro-benchmark (B200): DFMA vs DMMA m8n8k4 issue cost per SM sub-partition, and the XU-pipe instructions the
// 10-operation tanh leans on (MUFU.RCP64H, F2F both ways).  Prints FMA per clock per SM for each.  tools/ only, not product.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *o, int it, double a, double b) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x + i;
    for (int k = 0; k < it; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double *o, int it, double a, double b) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x + i; c[i][1] = i; }
    for (int k = 0; k < it; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(double *o, int it, double a) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = a * (threadIdx.x + i + 1);
    for (int k = 0; k < it; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) { float f = (float)c[i]; f = f * 1.0000001f; c[i] = (double)f; }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rcp(double *o, int it, double a) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = a * (threadIdx.x + i + 1);
    for (int k = 0; k < it; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int hi = __double2hiint(c[i]), r;
            asm volatile("{.reg .f64 t, u; mov.b64 t, {%2, %1}; rcp.approx.ftz.f64 u, t; mov.b64 {%2, %0}, u;}" : "=r"(r) : "r"(hi), "r"(0));
            c[i] = __hiloint2double(r, 0);
        }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, it = 20000;
    double *o; cudaMalloc(&o, sizeof(double) * sms * 1024);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int block = warps * 32;
        float t1 = timeit([&] { k_dfma<<<sms, block>>>(o, it, 1.0000001, 1e-9); });
        float t2 = timeit([&] { k_dmma<<<sms, block>>>(o, it, 1.0000001, 1e-9); });
        float t3 = timeit([&] { k_cvt<<<sms, block>>>(o, it, 1.0000001); });
        float t4 = timeit([&] { k_rcp<<<sms, block>>>(o, it, 1.0000001); });
        const double clocks = (double)clk_khz * 1e3;
        printf("warps/SM %2d: DFMA %.1f FMA/clk/SM (%.2f ms)  DMMA m8n8k4 %.1f FMA/clk/SM = %.1f clk per DMMA per sub-partition (%.2f ms)  "
               "F2F pair %.1f lanes/clk/SM (%.2f ms)  RCP64H %.1f lanes/clk/SM (%.2f ms)\n", warps,
               (double)it * 8 * block / (t1 * 1e-3 * clocks), t1,
               (double)it * 8 * warps * 256 / (t2 * 1e-3 * clocks), (t2 * 1e-3 * clocks) / ((double)it * 8 * warps / 4.0), t2,
               (double)it * 8 * block / (t3 * 1e-3 * clocks), t3, (double)it * 8 * block / (t4 * 1e-3 * clocks), t4);
    }
    printf("clock rate used %d kHz (nominal max; SM clocks under load may differ)\n", clk_khz);
    return 0;
}
