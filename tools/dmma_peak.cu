// Microbenchmark: peak DMMA.8x8x4 rate on this GPU (register-resident operands), with and
// without interleaved vector DFMA, for several warps-per-SM.  Prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC, int NFMA>
__global__ void k(double* out, int iters, double x) {
    double c[NACC > 0 ? NACC : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) f[i] = threadIdx.x * 1e-3 + i;
    double a = threadIdx.x * 1e-3 + x, b = threadIdx.x * 2e-3 - x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < NFMA; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int NFMA>
void run(int threads, int ctas_per_sm, int sms, double* d) {
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC, NFMA><<<sms * ctas_per_sm, threads>>>(d, 100, 0.5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NACC, NFMA><<<sms * ctas_per_sm, threads>>>(d, iters, 0.5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)sms * ctas_per_sm * threads / 32;
    double dm = warps * iters * NACC * 512.0 / ms / 1e9;          // 8*8*4*2 flops per DMMA
    double fm = warps * iters * NFMA * 64.0 / ms / 1e9;           // 32 lanes * 2 flops
    printf("acc=%2d fma=%2d threads=%4d ctas/sm=%d : %.2f ms  DMMA %.2f TF  DFMA %.2f TF\n", NACC, NFMA, threads, ctas_per_sm, ms, dm, fm);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* d; cudaMalloc(&d, 148 * 8 * 1024 * 8);
    int sms = p.multiProcessorCount;
    run<8, 0>(128, 1, sms, d); run<8, 0>(256, 1, sms, d); run<8, 0>(512, 1, sms, d); run<8, 0>(1024, 1, sms, d);
    run<32, 0>(256, 1, sms, d); run<32, 0>(128, 1, sms, d);
    run<1, 0>(128, 1, sms, d); run<2, 0>(128, 1, sms, d); run<4, 0>(128, 1, sms, d);
    run<0, 8>(256, 1, sms, d); run<0, 8>(1024, 1, sms, d);
    run<8, 2>(256, 1, sms, d); run<8, 8>(256, 1, sms, d); run<8, 16>(256, 1, sms, d);
    return 0;
}
