// Do FP64 math and shared-memory traffic overlap on B200, or do they share an issue path?
#include <cstdio>
#include <cuda_runtime.h>
template <int DP, int LD>
__global__ void __launch_bounds__(256) k(double* out, int iters, double b, double c) {
    __shared__ __align__(16) double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double x[8]; for (int j = 0; j < 8; j++) x[j] = threadIdx.x + j;
    unsigned long long acc = 0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; i++) {
        if (LD) {
#pragma unroll
            for (int j = 0; j < LD; j++) {
                double2 v = buf[(idx + j * 256) & 2047];
                acc ^= (unsigned long long)__double_as_longlong(v.x) + (unsigned long long)__double_as_longlong(v.y);
            }
            idx = (idx + 32) & 2047;
        }
        if (DP) {
#pragma unroll
            for (int r = 0; r < DP; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = fma(x[j], b, c);
        }
    }
    double s = (double)acc; for (int j = 0; j < 8; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int DP, int LD> float run(double* out, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<DP, LD><<<148, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<DP, LD><<<148, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 1 << 22);
    int iters = 20000;
    // per iteration per warp: DP*8 DFMA (each 2 cycles of the fp64 pipe per SMSP) ; LD LDS.128 (4 wavefronts each, SM-wide)
    float a = run<2, 0>(out, iters), b = run<0, 8>(out, iters), c = run<2, 8>(out, iters);
    printf("8 warps/SM: DP only %.2f ms, LDS only %.2f ms, both %.2f ms (sum %.2f, max %.2f)\n", a, b, c, a + b, a > b ? a : b);
    float a2 = run<4, 0>(out, iters), b2 = run<0, 4>(out, iters), c2 = run<4, 4>(out, iters);
    printf("8 warps/SM: DP only %.2f ms, LDS only %.2f ms, both %.2f ms (sum %.2f, max %.2f)\n", a2, b2, c2, a2 + b2, a2 > b2 ? a2 : b2);
    return 0;
}
