// tfx_probe.cu — machine probes used by bench.py for roofline denominators that MEASURED_PEAKS.json does not carry:
// the FP64 FMA peak (the pipe that bounds the PBS FFT/MAC work) and the u64 x u32 multiply-accumulate peak
// (the pipe that bounds the keyswitch and the leveled conv).
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double b, double c) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = (double)(threadIdx.x + j) * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fma(a[j], b, c);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) imac_probe_kernel(uint64_t* out, int iters, uint32_t d0) {
    // same instruction pattern as the keyswitch inner product: lo64 += d * k_lo (IMAD.WIDE.U32), hi32 += d * k_hi (IMAD)
    uint64_t lo[8]; uint32_t hi[8];
    uint32_t klo = 0x9E3779B9u * (threadIdx.x + 1), khi = 0x7F4A7C15u ^ threadIdx.x;
    uint32_t d = d0 + (threadIdx.x & 7);
#pragma unroll
    for (int j = 0; j < 8; j++) { lo[j] = j; hi[j] = j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) { lo[j] += (uint64_t)d * klo; hi[j] += d * khi; }
        klo += d; khi ^= klo;
    }
    uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += lo[j] + ((uint64_t)hi[j] << 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// which: 0 = DFMA (returns FLOP/s), 1 = u64 += u32*u64 (returns MAC/s)
int probe_rate(int which, int sm_count, cudaStream_t stream, double* rate) {
    const int blocks = sm_count * 8, threads = 256, iters = 1 << 14;
    void* buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)blocks * threads * 8);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(probe)");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, stream);
        if (which == 0) dfma_probe_kernel<<<blocks, threads, 0, stream>>>((double*)buf, iters, 1.0000001, 1e-9);
        else imac_probe_kernel<<<blocks, threads, 0, stream>>>((uint64_t*)buf, iters, 3u);
        count_launch();
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
    if (e != cudaSuccess) return set_cuda_error(e, "probe kernel");
    const double ops = (double)blocks * threads * 8.0 * iters * (which == 0 ? 2.0 : 1.0);
    *rate = ops / (best * 1e-3);
    return TFX_OK;
}

}  // namespace tfx
