#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int iters, double b, double c) {
    // 1. dependent DFMA chain
    double a = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    // 2. ILP 8 independent chains
    double x[8]; for (int j = 0; j < 8; j++) x[j] = a + j;
    long long t2 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = fma(x[j], b, c);
    }
    long long t3 = clock64();
    // 3. dependent DADD chain
    double d = a;
    long long t4 = clock64();
    for (int i = 0; i < iters; i++) { d = d + b; d = d + c; d = d + b; d = d + c; }
    long long t5 = clock64();
    double s = a + d; for (int j = 0; j < 8; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMallocManaged(&cyc, 64);
    int iters = 4096;
    for (int warps = 1; warps <= 16; warps *= 2) {
        for (int rep = 0; rep < 2; rep++) { k<<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9); cudaDeviceSynchronize(); }
        printf("warps/SM %2d: dep DFMA %.2f cyc/op ; ILP8 %.2f cyc per 8 DFMA (%.2f/op) ; dep DADD %.2f cyc/op\n", warps,
               (double)cyc[0] / (iters * 4), (double)cyc[1] / iters, (double)cyc[1] / iters / 8, (double)cyc[2] / (iters * 4));
    }
    return 0;
}
