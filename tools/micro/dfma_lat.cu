// Micro-benchmark: dependent-issue latency and throughput of FP64 ops on this GPU (one warp per block).
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k_chain(double *out, long long *cyc, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    if (s == 1234.5) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc, int iters) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) x += __shfl_xor_sync(0xffffffffu, x, 1 << (u & 3));
    }
    long long t1 = clock64();
    if (x == 1234.5) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_mufu(double *out, long long *cyc, int iters) {
    double x = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
            x = y + 1.5;
        }
    }
    long long t1 = clock64();
    if (x == 1234.5) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double *out, long long *cyc, int iters) {
    __shared__ double s[64];
    s[threadIdx.x] = threadIdx.x; s[threadIdx.x + 32] = 1;
    __syncthreads();
    int idx = threadIdx.x & 31;
    long long t0 = clock64();
    double acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { double v = s[idx]; idx = ((int)v + u) & 31; acc += v; }
    }
    long long t1 = clock64();
    if (acc == 1234.5) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH> void run(const char *name, int warps, double *d_out, long long *d_cyc) {
    const int iters = 2000;
    k_chain<CH><<<1, 32 * warps>>>(d_out, d_cyc, iters, 0.999, 1e-3);
    k_chain<CH><<<1, 32 * warps>>>(d_out, d_cyc, iters, 0.999, 1e-3);
    long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s chains=%d warps=%d: %.2f cycles per DFMA per warp (%.2f cycles per dependent step)\n", name, CH, warps, (double)c / (iters * 4.0 * CH), (double)c / (iters * 4.0));
}
int main() {
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 64); cudaMalloc(&d_cyc, 64);
    run<1>("dfma", 1, d_out, d_cyc); run<2>("dfma", 1, d_out, d_cyc); run<4>("dfma", 1, d_out, d_cyc); run<8>("dfma", 1, d_out, d_cyc); run<16>("dfma", 1, d_out, d_cyc);
    run<1>("dfma", 4, d_out, d_cyc); run<4>("dfma", 4, d_out, d_cyc); run<4>("dfma", 8, d_out, d_cyc); run<4>("dfma", 16, d_out, d_cyc); run<8>("dfma", 16, d_out, d_cyc);
    long long c;
    k_shfl<<<1, 32>>>(d_out, d_cyc, 2000); k_shfl<<<1, 32>>>(d_out, d_cyc, 2000); cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("shfl64 + dadd dependent: %.1f cycles per step\n", (double)c / 8000);
    k_mufu<<<1, 32>>>(d_out, d_cyc, 2000); k_mufu<<<1, 32>>>(d_out, d_cyc, 2000); cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("mufu.rsq64h + dadd dependent: %.1f cycles per step\n", (double)c / 8000);
    k_lds<<<1, 32>>>(d_out, d_cyc, 2000); k_lds<<<1, 32>>>(d_out, d_cyc, 2000); cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("lds.64 -> f2i -> address dependent (+dadd): %.1f cycles per step\n", (double)c / 8000);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
