// tfx_leveled.cu — K3: leveled LWE x integer-weight conv2d (and depthwise sum-pool) plus elementwise a*sa + b*sb.
// Every tensor element is an LWE vector of `words` torus words; the contraction is over (ic, ky, kx) with clear
// integer weights, arithmetic mod 2^64.  Replaces (upstream) the MLIR-lowered leveled conv/matmul/add of
// Server.run behind reference homomorphic_eval.py:70 (work induced by models/backbone.py:67,69,81,102,232-239,276).
//
// Thread = one torus word (coalesced along the LWE vector), register tile = CV_OC output channels x CV_PX output
// pixels of one output row.  Grid order puts the word chunk slowest so that all CTAs working on one word chunk
// run together and the input slice they share ([Cin][H][W][128 words]) stays L2-resident: HBM sees the input
// once and the output once.
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

constexpr int CV_THREADS = 128;   // words per CTA
constexpr int CV_OC = 8;
constexpr int CV_PX = 4;

struct ConvArgs {
    const uint64_t* in; const int32_t* w; const uint64_t* bias; uint64_t* out;
    uint32_t Cin, H, W, words, Cout, kh, kw, stride, pad, Ho, Wo, oc_begin, oc_end, depthwise;
    uint32_t px_tiles, oc_tiles;
};

__global__ void __launch_bounds__(CV_THREADS) conv2d_kernel(ConvArgs a) {
    extern __shared__ int32_t s_w[];                                   // [CV_OC][Cin_eff][kh][kw]
    const uint32_t oc_tile = blockIdx.x % a.oc_tiles;
    const uint32_t px_tile = blockIdx.x / a.oc_tiles;
    const uint32_t oy = px_tile / a.px_tiles, ox0 = (px_tile % a.px_tiles) * CV_PX;
    const uint32_t oc0 = a.oc_begin + oc_tile * CV_OC;
    const uint32_t t = blockIdx.y * CV_THREADS + threadIdx.x;
    const uint32_t cin_eff = a.depthwise ? 1 : a.Cin;
    const uint32_t wsz = cin_eff * a.kh * a.kw;
    for (uint32_t i = threadIdx.x; i < CV_OC * wsz; i += CV_THREADS) {
        uint32_t o = i / wsz, rem = i - o * wsz;
        s_w[i] = (oc0 + o < a.oc_end) ? a.w[(size_t)(oc0 + o) * wsz + rem] : 0;
    }
    __syncthreads();
    if (t >= a.words) return;

    uint64_t acc[CV_OC][CV_PX];
#pragma unroll
    for (int o = 0; o < CV_OC; o++)
#pragma unroll
        for (int p = 0; p < CV_PX; p++) acc[o][p] = 0;

    if (!a.depthwise) {
        for (uint32_t ic = 0; ic < a.Cin; ic++)
            for (uint32_t ky = 0; ky < a.kh; ky++) {
                const int iy = (int)(oy * a.stride + ky) - (int)a.pad;
                if (iy < 0 || iy >= (int)a.H) continue;
                const uint64_t* rowp = a.in + ((size_t)ic * a.H + iy) * a.W * a.words + t;
                for (uint32_t kx = 0; kx < a.kw; kx++) {
                    uint64_t x[CV_PX];
#pragma unroll
                    for (int p = 0; p < CV_PX; p++) {
                        const int ix = (int)((ox0 + p) * a.stride + kx) - (int)a.pad;
                        x[p] = (ix >= 0 && ix < (int)a.W && ox0 + p < a.Wo) ? __ldg(rowp + (size_t)ix * a.words) : 0;
                    }
                    const int32_t* wp = s_w + (ic * a.kh + ky) * a.kw + kx;
#pragma unroll
                    for (int o = 0; o < CV_OC; o++) {
                        const uint64_t wv = (uint64_t)(int64_t)wp[o * wsz];
#pragma unroll
                        for (int p = 0; p < CV_PX; p++) acc[o][p] += wv * x[p];
                    }
                }
            }
    } else {
#pragma unroll
        for (int o = 0; o < CV_OC; o++) {
            const uint32_t oc = oc0 + o;
            if (oc >= a.oc_end) continue;
            for (uint32_t ky = 0; ky < a.kh; ky++) {
                const int iy = (int)(oy * a.stride + ky) - (int)a.pad;
                if (iy < 0 || iy >= (int)a.H) continue;
                const uint64_t* rowp = a.in + ((size_t)oc * a.H + iy) * a.W * a.words + t;
                for (uint32_t kx = 0; kx < a.kw; kx++) {
                    const uint64_t wv = (uint64_t)(int64_t)s_w[o * wsz + ky * a.kw + kx];
#pragma unroll
                    for (int p = 0; p < CV_PX; p++) {
                        const int ix = (int)((ox0 + p) * a.stride + kx) - (int)a.pad;
                        if (ix >= 0 && ix < (int)a.W && ox0 + p < a.Wo) acc[o][p] += wv * __ldg(rowp + (size_t)ix * a.words);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < CV_OC; o++) {
        const uint32_t oc = oc0 + o;
        if (oc >= a.oc_end) continue;
        const uint64_t bias = (a.bias && t == a.words - 1) ? a.bias[oc] : 0;
#pragma unroll
        for (int p = 0; p < CV_PX; p++) {
            if (ox0 + p >= a.Wo) continue;
            a.out[(((size_t)(oc - a.oc_begin) * a.Ho + oy) * a.Wo + ox0 + p) * a.words + t] = acc[o][p] + bias;
        }
    }
}

// no __restrict__: include/tfx.h allows out to alias a or b (element i is read before it is written, by the same thread)
__global__ void axpby_kernel(const uint64_t* a, uint64_t sa, const uint64_t* b, uint64_t sb,
                             uint64_t body_const, size_t count, uint32_t words, uint64_t* out) {
    const size_t total = count * words;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = a[i] * sa;
        if (b) v += b[i] * sb;
        if ((i + 1) % words == 0) v += body_const;
        out[i] = v;
    }
}

int launch_conv2d(const uint64_t* in, uint32_t Cin, uint32_t H, uint32_t W, uint32_t words, const int32_t* w, uint32_t Cout,
                  uint32_t kh, uint32_t kw, uint32_t stride, uint32_t pad, const uint64_t* bias_pt, uint32_t oc_begin,
                  uint32_t oc_end, uint32_t depthwise, uint64_t* out, int sm_count, cudaStream_t s) {
    (void)sm_count;
    if (oc_end > Cout || oc_begin >= oc_end) return set_error(TFX_ERR_ARG, "conv2d: bad output-channel range");
    if (stride == 0 || H + 2 * pad < kh || W + 2 * pad < kw) return set_error(TFX_ERR_ARG, "conv2d: bad geometry");
    if (depthwise && Cin != Cout) return set_error(TFX_ERR_ARG, "conv2d: depthwise needs Cin == Cout");
    ConvArgs a;
    a.in = in; a.w = w; a.bias = bias_pt; a.out = out; a.Cin = Cin; a.H = H; a.W = W; a.words = words; a.Cout = Cout;
    a.kh = kh; a.kw = kw; a.stride = stride; a.pad = pad; a.Ho = (H + 2 * pad - kh) / stride + 1; a.Wo = (W + 2 * pad - kw) / stride + 1;
    a.oc_begin = oc_begin; a.oc_end = oc_end; a.depthwise = depthwise;
    a.px_tiles = (a.Wo + CV_PX - 1) / CV_PX; a.oc_tiles = (oc_end - oc_begin + CV_OC - 1) / CV_OC;
    size_t smem = (size_t)CV_OC * (depthwise ? 1 : Cin) * kh * kw * 4;
    if (smem > 200 * 1024) return set_error(TFX_ERR_UNSUPPORTED, "conv2d: weight tile exceeds shared memory");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv2d)");
    }
    dim3 grid(a.oc_tiles * a.px_tiles * a.Ho, (words + CV_THREADS - 1) / CV_THREADS);
    conv2d_kernel<<<grid, CV_THREADS, smem, s>>>(a);
    count_launch();
    return check_launch("conv2d_kernel");
}

int launch_axpby(const uint64_t* a, int64_t sa, const uint64_t* b, int64_t sb, uint64_t body_const, size_t count,
                 uint32_t words, uint64_t* out, int sm_count, cudaStream_t s) {
    if (count == 0) return TFX_OK;
    size_t total = count * words;
    unsigned grid = (unsigned)((total + 255) / 256 < (size_t)sm_count * 16 ? (total + 255) / 256 : (size_t)sm_count * 16);
    axpby_kernel<<<grid, 256, 0, s>>>(a, (uint64_t)sa, b, (uint64_t)sb, body_const, count, words, out);
    count_launch();
    return check_launch("axpby_kernel");
}

}  // namespace tfx
