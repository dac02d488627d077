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

// =====================================================================================================
// Tensor-core variant (IMMA): the contraction is limb-split — digits are u8 (biased, < 2^base_log), every KSK word is
// 8 unsigned bytes — and evaluated as a u8 x u8 -> s32 GEMM  [cts x K] x [K x 8*(n+1)]  with mma.sync m16n8k32; the 8
// byte columns of an output word are recombined with shifts mod 2^64 in the epilogue.  Justification (SURVEY 7.4-f asks
// for ncu evidence): the IMAD kernel above is integer-pipe bound (top stall math_pipe_throttle, 60 % of the measured
// IMAD.WIDE+IMAD rate) and was 11.6 % of the encrypted-image step (profiles/r01_bench_launches_summary.json).
// Layouts: digits u8 [B][Kp] (Kp = big*l, row (i, lvl) at i*l+lvl); key bytes u8 [npad*8][Kp] (column j*8+byte).
// CTA = 128 ciphertexts x 16 output words (128 byte columns), 8 warps as 2 x 4, warp tile 64 x 32.
// =====================================================================================================
constexpr int KI_TM = 128, KI_TN = 128, KI_THREADS = 256;

__global__ void ks_decompose_kernel(const uint64_t* __restrict__ in, uint8_t* __restrict__ dig, uint32_t count, uint32_t big_dim,
                                    int base_log, int level, uint32_t shift) {
    const size_t total = (size_t)count * big_dim;
    const uint32_t half = 1u << (base_log - 1), mask = (1u << base_log) - 1;
    const int tot = base_log * level;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t ct = p / big_dim; const uint32_t w = (uint32_t)(p - ct * big_dim);
        const uint64_t x = in[ct * (big_dim + 1) + w] << shift;
        uint64_t v = (tot < 64) ? ((x + (1ULL << (63 - tot))) >> (64 - tot)) : x;
        uint8_t* dst = dig + ct * ((size_t)big_dim * level) + (size_t)w * level;
        for (int lvl = level - 1; lvl >= 0; lvl--) {
            const uint32_t r = (uint32_t)v & mask;
            v >>= base_log;
            if (r >= half) v += 1;
            dst[lvl] = (uint8_t)((r + half) & mask);
        }
    }
}

// key bytes from the padded u64 layout: kb[(j*8 + b)*Kp + row] = byte b of ksk[row][j]
__global__ void ksk_bytes_kernel(const uint64_t* __restrict__ ksk, uint8_t* __restrict__ kb, uint32_t rows, uint32_t npad) {
    const size_t total = (size_t)rows * npad;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(p / rows), row = (uint32_t)(p - (size_t)j * rows);
        const uint64_t v = ksk[(size_t)row * npad + j];
#pragma unroll
        for (int b = 0; b < 8; b++) kb[((size_t)j * 8 + b) * rows + row] = (uint8_t)(v >> (8 * b));
    }
}

struct KiArgs {
    const uint8_t* dig; const uint8_t* kb; const uint64_t* corr; const uint64_t* in; uint64_t* out;
    uint32_t big_dim, n, Kp, kc; uint32_t shift; uint64_t body_offset; uint32_t count;
};

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(KI_THREADS)
keyswitch_imma_kernel(KiArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t kc = a.kc, kcp = kc + 16;                 // bytes of K per stage, padded row stride (conflict-free LDS.32)
    uint8_t* As = smem_raw;                                   // [128][kcp]
    uint8_t* Bs = smem_raw + (size_t)KI_TM * kcp;             // [128][kcp]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;                  // 2 x 4 warps
    const int g = lane >> 2, t = lane & 3;
    const uint32_t ct0 = blockIdx.y * KI_TM, col0 = blockIdx.x * KI_TN;     // col = word*8 + byte
    const uint32_t vec_per_row = kc / 16, vecs = KI_TM * vec_per_row;

    int acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[i][j][q] = 0;

    for (uint32_t k0 = 0; k0 < a.Kp; k0 += kc) {
        __syncthreads();
        for (uint32_t v = tid; v < vecs; v += KI_THREADS) {
            const uint32_t row = v / vec_per_row, part = v - row * vec_per_row;
            uint4 av = make_uint4(0, 0, 0, 0);
            if (ct0 + row < a.count) av = *reinterpret_cast<const uint4*>(a.dig + (size_t)(ct0 + row) * a.Kp + k0 + part * 16);
            *reinterpret_cast<uint4*>(As + (size_t)row * kcp + part * 16) = av;
            const uint4 bv = *reinterpret_cast<const uint4*>(a.kb + (size_t)(col0 + row) * a.Kp + k0 + part * 16);
            *reinterpret_cast<uint4*>(Bs + (size_t)row * kcp + part * 16) = bv;
        }
        __syncthreads();
        for (uint32_t ks = 0; ks < kc; ks += 32) {
            uint32_t af[4][4], bf[4][2];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint8_t* base = As + (size_t)(wm * 64 + i * 16 + g) * kcp + ks + t * 4;
                af[i][0] = *reinterpret_cast<const uint32_t*>(base);
                af[i][1] = *reinterpret_cast<const uint32_t*>(base + 8 * kcp);
                af[i][2] = *reinterpret_cast<const uint32_t*>(base + 16);
                af[i][3] = *reinterpret_cast<const uint32_t*>(base + 8 * kcp + 16);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint8_t* base = Bs + (size_t)(wn * 32 + j * 8 + g) * kcp + ks + t * 4;
                bf[j][0] = *reinterpret_cast<const uint32_t*>(base);
                bf[j][1] = *reinterpret_cast<const uint32_t*>(base + 16);
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) mma_u8(acc[i][j], af[i], bf[j]);
        }
    }
    // epilogue: byte columns (2t, 2t+1) of word j live in this thread; recombine over the 4 threads of the group
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t word = (col0 >> 3) + wn * 4 + j;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int hrow = 0; hrow < 2; hrow++) {
                uint64_t v = ((uint64_t)(uint32_t)acc[i][j][hrow * 2] << (16 * t)) + ((uint64_t)(uint32_t)acc[i][j][hrow * 2 + 1] << (16 * t + 8));
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                const uint32_t ct = ct0 + wm * 64 + i * 16 + hrow * 8 + g;
                if (t == hrow && ct < a.count && word <= a.n) {
                    uint64_t init = 0;
                    if (word == a.n) init = (a.in[(size_t)ct * (a.big_dim + 1) + a.big_dim] << a.shift) + a.body_offset;
                    a.out[(size_t)ct * (a.n + 1) + word] = init - v + a.corr[word];
                }
            }
        }
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

int launch_ksk_bytes(const uint64_t* ksk_padded, uint8_t* kb, uint32_t rows, uint32_t n, cudaStream_t s) {
    ksk_bytes_kernel<<<2048, 256, 0, s>>>(ksk_padded, kb, rows, ksk_npad(n));
    count_launch();
    return check_launch("ksk_bytes_kernel");
}

// tensor-core path usable?  digits must fit a byte and the s32 accumulators must not overflow
bool keyswitch_imma_ok(uint32_t big_dim, int base_log, int level) {
    if (base_log > 8 || level < 1 || level > 8) return false;
    const double worst = (double)big_dim * level * 255.0 * (double)((1u << base_log) - 1);
    return worst < 2147483647.0 && ((size_t)big_dim * level) % (32 * level) == 0;
}

static int launch_keyswitch_imma(const KsLaunch& p, cudaStream_t stream) {
    const uint32_t Kp = p.big_dim * p.level, kc = 32 * p.level, npad = ksk_npad(p.n);
    const size_t smem = (size_t)2 * KI_TM * (kc + 16);
    {   // per launch: the attribute is per device, and a process may drive several devices / host threads
        cudaError_t e = cudaFuncSetAttribute(keyswitch_imma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(keyswitch_imma)");
        // same shared-memory carve-out as the PBS kernels: the chains of a lookup layer run on several streams, and an SM only
        // hosts CTAs of kernels that agree on the L1 / shared-memory split (a different split waits for the SM to drain)
        cudaFuncSetAttribute(keyswitch_imma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(ks_decompose_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    ks_decompose_kernel<<<(unsigned)p.sm_count * 16, 256, 0, stream>>>(p.in, p.digits, (uint32_t)p.count, p.big_dim, p.base_log, p.level, p.shift);
    count_launch();
    int rc = check_launch("ks_decompose_kernel"); if (rc) return rc;
    // Blackwell tensor path (tcgen05.mma kind::i8, TMA operands, TMEM accumulators); TFX_KS_IMMA=1 keeps the mma.sync kernel
    if (keyswitch_umma_ok(p.big_dim, p.n, p.level)) return launch_keyswitch_umma(p, stream);
    KiArgs a;
    a.dig = p.digits; a.kb = p.ksk_bytes; a.corr = p.ksk + (size_t)p.big_dim * p.level * npad; a.in = p.in; a.out = p.out;
    a.big_dim = p.big_dim; a.n = p.n; a.Kp = Kp; a.kc = kc; a.shift = p.shift; a.body_offset = p.body_offset; a.count = (uint32_t)p.count;
    dim3 grid(npad * 8 / KI_TN, (unsigned)((p.count + KI_TM - 1) / KI_TM));
    keyswitch_imma_kernel<<<grid, KI_THREADS, smem, stream>>>(a);
    count_launch();
    return check_launch("keyswitch_imma_kernel");
}

int launch_keyswitch(const KsLaunch& p, cudaStream_t stream) {
    if (p.count == 0) return TFX_OK;
    if (p.ksk_bytes && p.digits && keyswitch_imma_ok(p.big_dim, p.base_log, p.level)) return launch_keyswitch_imma(p, stream);
    if (p.level < 1 || p.level > 16) return set_error(TFX_ERR_UNSUPPORTED, "keyswitch: ksk_level must be in 1..16");
    if (p.base_log < 1 || p.base_log > 31) return set_error(TFX_ERR_ARG, "keyswitch: base_log out of range");
    {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(keyswitch)");
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
