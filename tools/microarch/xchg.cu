// Warp-local 8x8 transpose of 8 complex doubles per lane (the exchange between two radix-8 passes of the PBS transform):
// through shared memory (8 STS.128 + 8 LDS.128 = 64 wavefronts per warp) or through shuffles (pre-rotate, 7 complex
// shuffles = 28 SHFL.32, post-rotate), each interleaved with the FP64 work of one radix-8 node (~84 FP64 instructions).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 3) & 7); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ void node(double2 (&y)[8], const double2 (&w)[7]) {
#pragma unroll
    for (int q = 1; q < 8; q++) y[q] = cmul(y[q], w[q - 1]);
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1)
#pragma unroll
        for (int q = 0; q < 8; q++)
            if ((q & s) == 0) {
                double2 a = y[q], b = y[q + s];
                y[q] = make_double2(a.x + b.x, a.y + b.y);
                y[q + s] = make_double2(a.x - b.x, a.y - b.y);
            }
}

__device__ __forceinline__ double2 sel(bool c, double2 a, double2 b) { return c ? a : b; }
__device__ __forceinline__ double2 shfl2(double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// new x[j] of lane g (within its group of 8) = old x[g] of lane j
__device__ __forceinline__ void transpose8_shfl(double2 (&x)[8], int lane) {
    const int g = lane & 7, base = lane & ~7;
    double2 y[8];
    // y[r] = x[(g + r) & 7]: barrel rotate left by g
#pragma unroll
    for (int r = 0; r < 8; r++) y[r] = sel(g & 1, x[(r + 1) & 7], x[r]);
#pragma unroll
    for (int r = 0; r < 8; r++) x[r] = sel(g & 2, y[(r + 2) & 7], y[r]);
#pragma unroll
    for (int r = 0; r < 8; r++) y[r] = sel(g & 4, x[(r + 4) & 7], x[r]);
    // round r: lane g sends y[r] = old x[(g+r)&7] to lane (g+r)&7; lane j receives from lane (j-r)&7
#pragma unroll
    for (int r = 1; r < 8; r++) y[r] = shfl2(y[r], base | ((g - r) & 7));
    // received z[r] = old[(g-r)&7][g] belongs in new x[(g-r)&7]: x[m] = z[(g-m)&7]; with w[q] = z[(-q)&7]: x[m] = w[(m-g)&7]
    double2 w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) w[q] = y[(8 - q) & 7];
    // rotate right by g
#pragma unroll
    for (int m = 0; m < 8; m++) y[m] = sel(g & 1, w[(m + 7) & 7], w[m]);
#pragma unroll
    for (int m = 0; m < 8; m++) w[m] = sel(g & 2, y[(m + 6) & 7], y[m]);
#pragma unroll
    for (int m = 0; m < 8; m++) x[m] = sel(g & 4, w[(m + 4) & 7], w[m]);
}

__device__ __forceinline__ void transpose8_smem(double2 (&x)[8], int t, double2* buf) {
    // pass-1 layout: idx = (t>>3)<<6 | e<<3 | (t&7); pass-2 layout: idx = t<<3 | e   (M = 512 per 64 threads)
#pragma unroll
    for (int e = 0; e < 8; e++) buf[swz(((t >> 3) << 6) | (e << 3) | (t & 7))] = x[e];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = buf[swz((t << 3) | e)];
    __syncwarp();
}

template <int MODE>   // 0: math only, 1: math + smem exchange, 2: math + shuffle exchange, 3: verify
__global__ void __launch_bounds__(64, 4) k(double* out, int iters, const double2* tw, int* bad) {
    __shared__ __align__(16) double2 buf[512];
    const int t = threadIdx.x;
    double2 x[8], w[7];
    for (int e = 0; e < 8; e++) x[e] = make_double2(t * 8 + e, -(t * 8 + e) * 0.5);
    for (int q = 0; q < 7; q++) w[q] = tw[(t * 7 + q) & 1023];
    if (MODE == 3) {
        double2 a[8], b[8];
        for (int e = 0; e < 8; e++) a[e] = b[e] = x[e];
        transpose8_smem(a, t, buf);
        transpose8_shfl(b, t & 31);
        for (int e = 0; e < 8; e++) if (a[e].x != b[e].x || a[e].y != b[e].y) atomicAdd(bad, 1);
        return;
    }
    for (int i = 0; i < iters; i++) {
        node(x, w);
        if (MODE == 1) transpose8_smem(x, t, buf);
        if (MODE == 2) transpose8_shfl(x, t & 31);
    }
    double s = 0; for (int e = 0; e < 8; e++) s += x[e].x + x[e].y;
    out[blockIdx.x * blockDim.x + t] = s;
}

template <int MODE> float run(double* out, int iters, const double2* tw, int* bad) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 64>>>(out, iters, tw, bad);
    cudaEventRecord(e0); k<MODE><<<148 * 4, 64>>>(out, iters, tw, bad); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    double* out; cudaMalloc(&out, 1 << 22);
    double2* tw; cudaMalloc(&tw, 1024 * sizeof(double2));
    double2 h[1024]; for (int i = 0; i < 1024; i++) h[i] = make_double2(1.0 - 1e-7 * i, 1e-4 * i);
    cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    k<3><<<4, 64>>>(out, 1, tw, bad);
    int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("shuffle transpose == shared-memory transpose: %s (%d mismatches)\n", hb == 0 ? "yes" : "NO", hb);
    const int iters = 4000;
    float a = run<0>(out, iters, tw, bad), b = run<1>(out, iters, tw, bad), c = run<2>(out, iters, tw, bad);
    const double per = 1.965e9 * 1e-3 / (iters * 8.0);   // cycles per warp-iteration per SM (8 warps per SM)
    printf("8 warps/SM, per warp and iteration (radix-8 node + exchange), SM cycles:\n");
    printf("  math only          %.3f ms  %.1f cycles\n", a, a * per);
    printf("  math + smem xchg   %.3f ms  %.1f cycles\n", b, b * per);
    printf("  math + shfl xchg   %.3f ms  %.1f cycles\n", c, c * per);
    return 0;
}
