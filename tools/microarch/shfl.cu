// Warp shuffles vs shared-memory exchange on B200: cost per warp instruction, and whether shuffles overlap FP64 math.
// An exchange of 8 complex doubles per lane is 16 LDS/STS.128 (64 shared-memory wavefronts) or 32 SHFL.32.
#include <cstdio>
#include <cuda_runtime.h>
template <int DP, int SH, int LD>
__global__ void __launch_bounds__(256) k(double* out, int iters, double b, double c) {
    __shared__ __align__(16) double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double x[8]; for (int j = 0; j < 8; j++) x[j] = threadIdx.x + j;
    unsigned v[8]; for (int j = 0; j < 8; j++) v[j] = threadIdx.x * 7 + j;
    unsigned long long acc = 0;
    int idx = threadIdx.x;
    const int src = (threadIdx.x * 5 + 3) & 31;
    for (int i = 0; i < iters; i++) {
        if (SH) {
#pragma unroll
            for (int r = 0; r < SH; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = __shfl_sync(0xffffffffu, v[j], src) + j;
        }
        if (LD) {
#pragma unroll
            for (int j = 0; j < LD; j++) {
                double2 w = buf[(idx + j * 256) & 2047];
                acc ^= (unsigned long long)__double_as_longlong(w.x) + (unsigned long long)__double_as_longlong(w.y);
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
    double s = (double)acc; for (int j = 0; j < 8; j++) s += x[j] + v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int DP, int SH, int LD> float run(double* out, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<DP, SH, LD><<<148, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e0); k<DP, SH, LD><<<148, 256>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 1 << 22);
    int iters = 20000;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const double ghz = p.clockRate / 1e6;
    // per iteration per SM (8 warps): DP*8*8 warp-DFMA, SH*8*8 warp-SHFL, LD*8 warp-LDS.128 (4 wavefronts each)
    float dp = run<2, 0, 0>(out, iters), sh = run<0, 2, 0>(out, iters), ld = run<0, 0, 4>(out, iters);
    printf("clock %.3f GHz\n", ghz);
    printf("DFMA only : %.3f ms  -> %.3f cycles per warp instruction per SM\n", dp, dp * 1e-3 * ghz * 1e9 / (iters * 128.0));
    printf("SHFL only : %.3f ms  -> %.3f cycles per warp instruction per SM\n", sh, sh * 1e-3 * ghz * 1e9 / (iters * 128.0));
    printf("LDS.128   : %.3f ms  -> %.3f cycles per wavefront per SM\n", ld, ld * 1e-3 * ghz * 1e9 / (iters * 128.0));
    float a = run<2, 2, 0>(out, iters);
    printf("DFMA+SHFL : %.3f ms (sum %.3f, max %.3f)\n", a, dp + sh, dp > sh ? dp : sh);
    float b = run<0, 2, 4>(out, iters);
    printf("SHFL+LDS  : %.3f ms (sum %.3f, max %.3f)\n", b, sh + ld, sh > ld ? sh : ld);
    float c = run<2, 0, 4>(out, iters);
    printf("DFMA+LDS  : %.3f ms (sum %.3f, max %.3f)\n", c, dp + ld, dp > ld ? dp : ld);
    return 0;
}
