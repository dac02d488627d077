// tfx_keyswitch.cu — K2: batched LWE keyswitch big -> small key as a dense u64 contraction
//   out[b][j] = init[b][j] - sum_{i<big, lvl<l} d(b,i,lvl) * KSK[i][lvl][j]        (mod 2^64)
// with d the signed gadget digits of (in[b][i] << shift).  Digits are biased to unsigned (d' = d + B/2) so the
// inner product is IMAD.WIDE.U32 + IMAD per multiply-accumulate; the bias is undone with the precomputed
// column sums corr[j] = (B/2) * sum_rows KSK[row][j].
// Replaces (upstream) concrete-cpu's keyswitch behind reference homomorphic_eval.py:70.
//
// Device KSK layout: u64 [big*l][npad], npad = round_up(n+1, KS_TJ), zero padded; corr: u64 [npad].
// Tiling: CTA = KS_TB ciphertexts x KS_TJ output words, 256 threads, 4x4 register tile per thread,
// K dimension (big*l rows) staged through shared memory in chunks of KS_KC rows.
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

constexpr int KS_TB = 64;     // ciphertexts per CTA
constexpr int KS_TJ = 64;     // output words per CTA
constexpr int KS_KC = 128;    // contraction rows per stage
constexpr int KS_THREADS = 256;
constexpr size_t KS_SMEM = (size_t)KS_KC * KS_TB * 4 + (size_t)KS_KC * KS_TJ * 8;

struct KsArgs {
    const uint64_t* ksk; const uint64_t* corr; const uint64_t* in; uint64_t* out;
    uint32_t big_dim, n, npad; int base_log, level; uint32_t shift; uint64_t body_offset; uint32_t count;
};

__global__ void __launch_bounds__(KS_THREADS)
keyswitch_kernel(KsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t (*s_dig)[KS_TB] = reinterpret_cast<uint32_t (*)[KS_TB]>(smem_raw);                       // [KC][TB]
    uint64_t (*s_ksk)[KS_TJ] = reinterpret_cast<uint64_t (*)[KS_TJ]>(smem_raw + (size_t)KS_KC * KS_TB * 4);  // [KC][TJ]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;                 // 16 x 16 threads
    const uint32_t b0 = blockIdx.y * KS_TB, j0 = blockIdx.x * KS_TJ;
    const int l = a.level, bl = a.base_log;
    const uint32_t words_per_chunk = KS_KC / l;             // rows used per stage = words_per_chunk * l <= KS_KC
    const int rows_used = (int)words_per_chunk * l;
    const uint32_t rows_total = a.big_dim * l;
    const uint32_t half = 1u << (bl - 1), mask = (1u << bl) - 1;
    const int total = bl * l;

    // split accumulators: lo64 += d * K_lo (IMAD.WIDE.U32, carries stay inside the 64-bit word), hi32 += d * K_hi (IMAD);
    // result mod 2^64 = lo64 + (hi32 << 32).  Two integer multiply-adds per u64 MAC, no 64-bit pair shuffling.
    uint64_t acc_lo[4][4];
    uint32_t acc_hi[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { acc_lo[r][c] = 0; acc_hi[r][c] = 0; }

    for (uint32_t w0 = 0; w0 < a.big_dim; w0 += words_per_chunk) {
        __syncthreads();
        // stage digits: (ct, word) pairs of this chunk; lanes run along the word index (coalesced)
        for (uint32_t p = tid; p < (uint32_t)KS_TB * words_per_chunk; p += KS_THREADS) {
            const uint32_t wl = p % words_per_chunk, cl = p / words_per_chunk;
            const uint32_t b = b0 + cl, w = w0 + wl;
            uint64_t x = 0;
            if (b < a.count && w < a.big_dim) x = a.in[(size_t)b * (a.big_dim + 1) + w] << a.shift;
            uint64_t v = (total < 64) ? ((x + (1ULL << (63 - total))) >> (64 - total)) : x;
            for (int lvl = l - 1; lvl >= 0; lvl--) {
                uint32_t r = (uint32_t)v & mask;
                v >>= bl;
                if (r >= half) v += 1;                                     // signed digit r - B (carry); biased: r - B + B/2
                const uint32_t row = wl * l + lvl;                          // column groups of 4 XOR-swizzled by the row:
                s_dig[row][cl ^ ((row & 15) << 2)] = (r + half) & mask;    // d + B/2 in [0, B); lanes walk rows -> spread banks
            }
        }
        // stage KSK rows [w0*l, w0*l + KC) x [j0, j0 + TJ)
        for (uint32_t p = tid; p < (uint32_t)rows_used * KS_TJ; p += KS_THREADS) {
            const uint32_t jl = p % KS_TJ, rl = p / KS_TJ;
            const uint32_t row = w0 * l + rl;
            s_ksk[rl][jl] = (row < rows_total) ? a.ksk[(size_t)row * a.npad + j0 + jl] : 0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < rows_used; kk++) {
            const uint4 dg = *reinterpret_cast<const uint4*>(&s_dig[kk][(ty * 4) ^ ((kk & 15) << 2)]);
            const ulonglong2 ka = *reinterpret_cast<const ulonglong2*>(&s_ksk[kk][tx * 2]);
            const ulonglong2 kb = *reinterpret_cast<const ulonglong2*>(&s_ksk[kk][32 + tx * 2]);
            const uint32_t d[4] = {dg.x, dg.y, dg.z, dg.w};
            const uint64_t kv[4] = {ka.x, ka.y, kb.x, kb.y};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t klo = (uint32_t)kv[c], khi = (uint32_t)(kv[c] >> 32);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    acc_lo[r][c] += (uint64_t)d[r] * klo;
                    acc_hi[r][c] += d[r] * khi;
                }
            }
        }
    }
    // epilogue: out = init - (acc - corr)
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t b = b0 + ty * 4 + r;
        if (b >= a.count) continue;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const uint32_t j = j0 + (c < 2 ? tx * 2 + c : 32 + tx * 2 + (c - 2));
            if (j > a.n) continue;
            uint64_t init = 0;
            if (j == a.n) init = (a.in[(size_t)b * (a.big_dim + 1) + a.big_dim] << a.shift) + a.body_offset;
            const uint64_t acc = acc_lo[r][c] + ((uint64_t)acc_hi[r][c] << 32);
            a.out[(size_t)b * (a.n + 1) + j] = init - acc + a.corr[j];
        }
    }
}

// corr[j] = (B/2) * sum_rows ksk[row][j]
__global__ void ksk_corr_kernel(const uint64_t* __restrict__ ksk, uint64_t* __restrict__ corr, uint32_t rows, uint32_t npad,
                                int base_log) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    uint64_t s = 0;
    for (uint32_t r = 0; r < rows; r++) s += ksk[(size_t)r * npad + j];
    corr[j] = s << (base_log - 1);
}

// canonical [rows][n+1] <-> padded device layout [rows][npad]
__global__ void ksk_repack_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint32_t rows, uint32_t n1,
                                  uint32_t npad, int to_padded) {
    size_t total = (size_t)rows * npad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t r = (uint32_t)(i / npad), j = (uint32_t)(i - (size_t)r * npad);
        if (to_padded) dst[i] = (j < n1) ? src[(size_t)r * n1 + j] : 0;
        else if (j < n1) dst[(size_t)r * n1 + j] = src[i];
    }
}

uint32_t ksk_npad(uint32_t n) { return (n + 1 + KS_TJ - 1) / KS_TJ * KS_TJ; }

int launch_ksk_repack(const uint64_t* src, uint64_t* dst, uint32_t rows, uint32_t n, int to_padded, cudaStream_t s) {
    ksk_repack_kernel<<<1024, 256, 0, s>>>(src, dst, rows, n + 1, ksk_npad(n), to_padded);
    count_launch();
    return check_launch("ksk_repack_kernel");
}

int launch_ksk_corr(const uint64_t* ksk_padded, uint64_t* corr, uint32_t rows, uint32_t n, int base_log, cudaStream_t s) {
    uint32_t npad = ksk_npad(n);
    ksk_corr_kernel<<<(npad + 63) / 64, 64, 0, s>>>(ksk_padded, corr, rows, npad, base_log);
    count_launch();
    return check_launch("ksk_corr_kernel");
}

int launch_keyswitch(const KsLaunch& p, cudaStream_t stream) {
    if (p.count == 0) return TFX_OK;
    if (p.level < 1 || p.level > 16) return set_error(TFX_ERR_UNSUPPORTED, "keyswitch: ksk_level must be in 1..16");
    if (p.base_log < 1 || p.base_log > 31) return set_error(TFX_ERR_ARG, "keyswitch: base_log out of range");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(keyswitch)");
        configured = true;
    }
    KsArgs a;
    a.ksk = p.ksk; a.corr = p.ksk + (size_t)p.big_dim * p.level * ksk_npad(p.n);
    a.in = p.in; a.out = p.out; a.big_dim = p.big_dim; a.n = p.n; a.npad = ksk_npad(p.n);
    a.base_log = p.base_log; a.level = p.level; a.shift = p.shift; a.body_offset = p.body_offset; a.count = (uint32_t)p.count;
    dim3 grid(a.npad / KS_TJ, (unsigned)((p.count + KS_TB - 1) / KS_TB));
    keyswitch_kernel<<<grid, KS_THREADS, KS_SMEM, stream>>>(a);
    count_launch();
    return check_launch("keyswitch_kernel");
}

}  // namespace tfx
