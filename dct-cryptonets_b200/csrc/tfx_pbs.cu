// tfx_pbs.cu — K1: batched programmable bootstrap (mod-switch, blind rotation, sample extract fused),
// plus the BSK -> Fourier conversion and the FFT test hooks that share its transform.
//
// One CTA (M/8 threads, 8 complex points per thread) owns one ciphertext at a time (grid-stride over the
// batch); two CTAs share an SM for N <= 2048.  The GLWE accumulator ((k+1) x N torus words) lives in shared memory
// for the whole blind rotation.  Per CMux step the CTA runs, for every accumulator component r and gadget level,
// the forward FFT of the digit polynomial (rotation and decomposition fused into the first register pass; the transform needs no twist),
// multiplies the spectrum with the Fourier bootstrapping-key rows (r, lvl, *) streamed from HBM/L2 in a
// thread-major layout (coalesced 16 B per lane) and accumulates the (k+1) output spectra in registers; then the
// (k+1) inverse FFTs run from those registers and add the rounded result into the accumulator.
// Replaces (upstream) concrete-cpu's bootstrap behind reference homomorphic_eval.py:70.
#include <stdio.h>
#include <stdlib.h>
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

template <int LOGN, int K> struct PbsCfg {
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGM = LOGN - 1;
    static constexpr int M = N / 2;
    static constexpr int TPF = M / 8;                    // threads per ciphertext (8 complex points each)
    static constexpr int G = K + 1;
    static constexpr int THREADS = TPF;
    // accumulator + two transform buffers (one per interleaved transform) + node twiddle table + the hand-out slot
    // (kept in the dynamic allocation: a static __shared__ word on top of a 227 KB dynamic limit is rejected)
    // N = 8192 (7-bit lookups): accumulator 128 KB + one transform buffer 64 KB; the 64 KB twiddle table stays in global memory
    static constexpr bool TW_SMEM = (LOGN <= 12);
    static constexpr size_t smem_bytes(int) {
        return (size_t)G * N * 8 + (size_t)M * 16 * (DUAL ? 2 : 1) + (TW_SMEM ? (size_t)M * 16 : 0) + 16;
    }
    // Two transforms interleaved per thread (more ILP, one more buffer) or one at a time.  Measured (profiles/r01_pbs_experiments.md):
    // interleaving wins 13 % for (k=1, N=2048, l=2) — four forward and two inverse transforms pair up — and loses 6 % for
    // (k=2, N=1024, l=1), whose three forward / three inverse transforms leave an odd one out.
#if defined(TFX_PBS_SINGLE)
    static constexpr bool DUAL = false;
#elif defined(TFX_PBS_DUAL)
    static constexpr bool DUAL = true;
#else
    static constexpr bool DUAL = !(K == 2 && LOGN <= 10) && LOGN <= 12;
#endif
    static constexpr int MIN_BLOCKS = (LOGN <= 11) ? 2 : 1;
};

struct PbsArgs {
    const double2* bsk;        // [n][G][l][G][8][TPF]
    const double2* tw;         // [M] node twiddles, flat per pass
    const uint64_t* in;        // [B][n+1]
    const uint64_t* luts;      // [T][N]
    const uint32_t* lut_index; // [B]
    uint64_t* out;             // [B][big_dim+1], big_dim >= kN (extra mask words are zero)
    uint32_t n, big_dim;
    int base_log, level, mode;
    uint64_t body_const;
    uint32_t count;
    uint32_t* work_counter;    // device counter (zeroed before the launch): ciphertexts are handed out dynamically
};

// signed digit `lvl` (1-based) of x in O(1): raw digit plus the carry of the balanced representation of the
// lower levels (carry iff lower part >= thr[lvl]); identical to the sequential rule of decompose_digit().
struct DigitCtx { uint64_t round_add; int top_shift; int base_log; uint64_t mask, half; };

// single-level gadget: the balanced digit is the arithmetic top base_log bits of the rounded word
__device__ __forceinline__ double digit1_as_double(uint64_t x, const DigitCtx& dc) {
    const int64_t sd = (int64_t)(x + dc.round_add) >> dc.top_shift;
    return __longlong_as_double(0x4338000000000000LL + sd) - 6755399441055744.0;
}

__device__ __forceinline__ double digit_as_double(uint64_t x, const DigitCtx& dc, int shift_in_top, uint64_t low_mask, uint64_t thr) {
    const uint64_t v = (x + dc.round_add) >> dc.top_shift;            // top base_log*level bits, rounded
    uint64_t d = (v >> shift_in_top) & dc.mask;
    d += ((v & low_mask) >= thr) ? 1 : 0;
    int64_t sd = (int64_t)d;
    if (d >= dc.half) sd -= (int64_t)(dc.mask + 1);
    // exact int -> double without the conversion pipe: 1.5 * 2^52 + sd as a bit pattern, minus 1.5 * 2^52
    return __longlong_as_double(0x4338000000000000LL + sd) - 6755399441055744.0;
}

template <int LOGN, int K>
__global__ void __launch_bounds__(PbsCfg<LOGN, K>::THREADS, PbsCfg<LOGN, K>::MIN_BLOCKS)
pbs_kernel(PbsArgs a) {
    using C = PbsCfg<LOGN, K>;
    constexpr int N = C::N, M = C::M, LOGM = C::LOGM, TPF = C::TPF, G = C::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);                    // [G][N]
    double2* bufs = reinterpret_cast<double2*>(acc + (size_t)G * N);          // [2][M] swizzled, alternate per transform
    double2* s_twbuf = bufs + (C::DUAL ? 2 : 1) * M;                          // [M] (absent when the table stays in global memory)
    const double2* s_tw = C::TW_SMEM ? s_twbuf : a.tw;
    volatile uint32_t* s_ct = reinterpret_cast<volatile uint32_t*>(s_twbuf + (C::TW_SMEM ? M : 0));  // next ciphertext index (dynamic hand-out)

    const int t = threadIdx.x;
    auto sync = [] { __syncthreads(); };
    auto wsync = [] { __syncwarp(); };
    double2* bufa = bufs;
    double2* bufb = C::DUAL ? bufs + M : bufs;

    if constexpr (C::TW_SMEM) for (int i = t; i < M; i += TPF) s_twbuf[i] = a.tw[i];

    DigitCtx dc;
    {
        const int total = a.base_log * a.level;
        dc.round_add = (total < 64) ? (1ULL << (63 - total)) : 0;
        dc.top_shift = 64 - total;
        dc.base_log = a.base_log;
        dc.mask = (1ULL << a.base_log) - 1;
        dc.half = 1ULL << (a.base_log - 1);
    }
    const int n_fwd = G * a.level;                 // forward transforms per CMux step, index f = r * level + lvl

    // coefficient pair (j, j+M) of X^ahat * acc_r - acc_r
    auto diff_pair = [&](const uint64_t* ar, int jc, uint32_t ahat, uint64_t& d0, uint64_t& d1) {
        const uint32_t s0 = (uint32_t)(jc - (int)ahat) & (2 * N - 1);
        const uint64_t r0 = s0 < N ? ar[s0] : (uint64_t)0 - ar[s0 - N];
        d0 = r0 - ar[jc];
        const uint32_t s1 = (s0 + M) & (2 * N - 1);
        const uint64_t r1 = s1 < N ? ar[s1] : (uint64_t)0 - ar[s1 - N];
        d1 = r1 - ar[jc + M];
    };
    // digit selector of level index lvl (0 = most significant): shift, low mask, carry threshold of the balanced
    // representation of the m = level-1-lvl lower digits: (B/2 - 1) * (B^m - 1) / (B - 1) + 1 (== B/2 for m = 1)
    auto level_sel = [&](int lvl, int& shift_in_top, uint64_t& low_mask, uint64_t& thr) {
        shift_in_top = a.base_log * (a.level - 1 - lvl);
        low_mask = (shift_in_top > 0) ? ((1ULL << shift_in_top) - 1) : 0;
        thr = ~0ULL;
        if (shift_in_top > 0) {
            uint64_t rep = 0;
            for (int q = 0; q < a.level - 1 - lvl; q++) rep = (rep << a.base_log) | 1ULL;
            thr = (dc.half - 1) * rep + 1;
        }
    };
    // pass-0 input of forward transform f = r * level + lvl: digit polynomial of X^ahat * acc_r - acc_r
    auto load_digits = [&](double2 (&x)[8], int f, uint32_t ahat) {
        const int r = f / a.level, lvl = f - r * a.level;
        const uint64_t* ar = acc + (size_t)r * N;
#ifdef TFX_EXP_NODIGITS
        for (int e = 0; e < 8; e++) x[e] = make_double2((double)(t + e + f), (double)ahat);
        return;
#endif
        if (a.level == 1) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int jc = first_pass_index<LOGM>(t, e);
                uint64_t d0, d1;
                diff_pair(ar, jc, ahat, d0, d1);
                x[e] = make_double2(digit1_as_double(d0, dc), digit1_as_double(d1, dc));
            }
            return;
        }
        int sh; uint64_t lm, thr;
        level_sel(lvl, sh, lm, thr);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int jc = first_pass_index<LOGM>(t, e);
            uint64_t d0, d1;
            diff_pair(ar, jc, ahat, d0, d1);
            x[e] = make_double2(digit_as_double(d0, dc, sh, lm, thr), digit_as_double(d1, dc, sh, lm, thr));
        }
    };
    // two consecutive levels of the same component: the accumulator reads and the rotation are shared
    auto load_digits_2levels = [&](double2 (&xa)[8], double2 (&xb)[8], int f, uint32_t ahat) {
        const int r = f / a.level, lvl = f - r * a.level;
        const uint64_t* ar = acc + (size_t)r * N;
        int sha, shb; uint64_t lma, lmb, thra, thrb;
        level_sel(lvl, sha, lma, thra);
        level_sel(lvl + 1, shb, lmb, thrb);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int jc = first_pass_index<LOGM>(t, e);
            uint64_t d0, d1;
            diff_pair(ar, jc, ahat, d0, d1);
            xa[e] = make_double2(digit_as_double(d0, dc, sha, lma, thra), digit_as_double(d1, dc, sha, lma, thra));
            xb[e] = make_double2(digit_as_double(d0, dc, shb, lmb, thrb), digit_as_double(d1, dc, shb, lmb, thrb));
        }
    };

    // dynamic hand-out of ciphertexts: with small batches (multi-GPU shards) a static stride leaves SMs unevenly loaded
    for (;;) {
        __syncthreads();
        if (t == 0) *s_ct = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const uint32_t ct = *s_ct;
        if (ct >= a.count) break;
        const uint64_t* in = a.in + (size_t)ct * (a.n + 1);
        {   // acc = X^{-bhat} * (0, .., 0, LUT)
            const uint64_t* lut = a.luts + (size_t)a.lut_index[ct] * N;
            const uint32_t bhat = mod_switch(__ldg(in + a.n), LOGN + 1);
            for (int j = t; j < K * N; j += TPF) acc[j] = 0;
            for (int j = t; j < N; j += TPF) {
                uint32_t idx = (j + bhat) & (2 * N - 1);
                acc[(size_t)K * N + j] = idx < N ? lut[idx] : (uint64_t)0 - lut[idx - N];
            }
        }
        __syncthreads();

        uint64_t a_next = __ldg(in);                                   // mask word of the next step, fetched one step ahead
        for (uint32_t i = 0; i < a.n; i++) {
            const uint32_t ahat = mod_switch(a_next, LOGN + 1);
            a_next = __ldg(in + i + 1);                                // (word n is the body: harmless, unused)
            if (ahat == 0) continue;                                   // CTA-uniform
            const double2* key_i = a.bsk + (size_t)i * G * a.level * G * M + t;
            double2 part[G][8];
#pragma unroll
            for (int c = 0; c < G; c++)
#pragma unroll
                for (int e = 0; e < 8; e++) part[c][e] = make_double2(0.0, 0.0);

            // Fourier MAC of spectrum x with BSK_i rows (f, c): one fma chain over f ascending
            auto mac = [&](const double2 (&x)[8], int f) {
#ifdef TFX_EXP_NOMAC
                for (int e = 0; e < 8; e++) part[f % G][e] = x[e];
                return;
#endif
                const double2* krow = key_i + (size_t)f * G * M;
#pragma unroll
                for (int c = 0; c < G; c++) {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const double2 kv = __ldg(krow + (size_t)c * M + e * TPF);
                        double re = part[c][e].x, im = part[c][e].y;
                        re = fma(x[e].x, kv.x, re); re = fma(-x[e].y, kv.y, re);
                        im = fma(x[e].x, kv.y, im); im = fma(x[e].y, kv.x, im);
                        part[c][e] = make_double2(re, im);
                    }
                }
            };

            // forward transforms two at a time (independent streams interleaved in every thread)
            if constexpr (!C::DUAL) {
#pragma unroll 1
                for (int f = 0; f < n_fwd; f++) {
                    double2 xa[8], w[7];
                    load_tw<LOGM, 0>(w, t, s_tw);
                    load_digits(xa, f, ahat);
                    fft_forward_regs2<LOGM, false>(xa, xa, w, t, bufa, bufb, s_tw, sync, wsync);
                    mac(xa, f);
                }
            } else {
#pragma unroll 1
            for (int f = 0; f + 1 < n_fwd; f += 2) {
                double2 xa[8], xb[8], w[7];
                load_tw<LOGM, 0>(w, t, s_tw);
                if ((a.level & 1) == 0) {                              // f even and level even: (f, f+1) are two levels of one component
                    load_digits_2levels(xa, xb, f, ahat);
                } else {
                    load_digits(xa, f, ahat);
                    load_digits(xb, f + 1, ahat);
                }
                fft_forward_regs2<LOGM, true>(xa, xb, w, t, bufa, bufb, s_tw, sync, wsync);
                mac(xa, f);
                mac(xb, f + 1);
            }
            if (n_fwd & 1) {
                double2 xa[8], w[7];
                load_tw<LOGM, 0>(w, t, s_tw);
                load_digits(xa, n_fwd - 1, ahat);
                fft_forward_regs2<LOGM, false>(xa, xa, w, t, bufa, bufb, s_tw, sync, wsync);
                mac(xa, n_fwd - 1);
            }
            }

            auto add_back = [&](const double2 (&x)[8], int c) {
                uint64_t* ac = acc + (size_t)c * N;
#ifdef TFX_EXP_NOADDBACK
                if (x[0].x == 123.456) ac[t] = 1;
                return;
#endif
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int jc = first_pass_index<LOGM>(t, e);                          // the key carries the 1/M of the inverse transform
                    ac[jc] += double_to_torus(x[e].x);
                    ac[jc + M] += double_to_torus(x[e].y);
                }
            };
            // inverse transforms, two output components at a time
            if constexpr (!C::DUAL) {
#pragma unroll
                for (int c = 0; c < G; c++) {
                    double2 xa[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) xa[e] = part[c][e];
                    fft_inverse_regs2<LOGM, false>(xa, xa, t, bufa, bufb, s_tw, sync, wsync);
                    add_back(xa, c);
                }
            } else {
#pragma unroll
            for (int c = 0; c + 1 < G; c += 2) {
                double2 xa[8], xb[8];
#pragma unroll
                for (int e = 0; e < 8; e++) { xa[e] = part[c][e]; xb[e] = part[c + 1][e]; }
                fft_inverse_regs2<LOGM, true>(xa, xb, t, bufa, bufb, s_tw, sync, wsync);
                add_back(xa, c);
                add_back(xb, c + 1);
            }
            if (G & 1) {
                double2 xa[8];
#pragma unroll
                for (int e = 0; e < 8; e++) xa[e] = part[G - 1][e];
                fft_inverse_regs2<LOGM, false>(xa, xa, t, bufa, bufb, s_tw, sync, wsync);
                add_back(xa, G - 1);
            }
            }
            __syncthreads();                                           // accumulator complete before the next rotation reads
        }

        // sample extract coefficient 0 -> LWE under the big key (its first kN bits are this set's GLWE key)
        uint64_t* o = a.out + (size_t)ct * ((size_t)a.big_dim + 1);
        for (int j = t; j < K * N; j += TPF) {
            int comp = j / N, tt = j - comp * N;
            const uint64_t* ar = acc + (size_t)comp * N;
            uint64_t v = (tt == 0) ? ar[0] : (uint64_t)0 - ar[N - tt];
            if (a.mode == 0) o[j] = v; else o[j] -= v;
        }
        if (a.mode == 0) for (uint32_t j = K * N + t; j < a.big_dim; j += TPF) o[j] = 0;
        if (t == 0) {
            uint64_t bv = acc[(size_t)K * N];
            if (a.mode == 0) o[a.big_dim] = bv; else o[a.big_dim] -= bv + a.body_const;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// v8 (round 2): all (k+1)*l forward transforms of a CMux step run interleaved in one instruction stream (one buffer per
// transform, node twiddles fetched once per pass for all of them, independent dependency chains for the scheduler), the
// Fourier MAC produces one output spectrum at a time and runs its first inverse pass at once, and the remaining inverse
// passes of the (k+1) outputs run interleaved again.  Three CTA barriers per step instead of seven.  Rotation and gadget
// decomposition work on the high 32 bits only (base_log * level <= 31: the rounding constant has no low word).
// Arithmetic per value is identical to pbs_kernel (same node sequence, same FMA chain over (r, level) ascending).
// ---------------------------------------------------------------------------------------------------
template <int LOGN, int K, int LEVEL> struct PbsCfg8 {
    static constexpr int N = 1 << LOGN, LOGM = LOGN - 1, M = N / 2, TPF = M / 8, G = K + 1, NF = G * LEVEL;
    static constexpr size_t SMEM = (size_t)G * N * 8 + (size_t)NF * M * 16 + (size_t)M * 16;
    static constexpr int MIN_BLOCKS = (LOGN <= 10) ? 4 : (LOGN == 11 ? 2 : 1);
};

template <int LOGN, int K, int LEVEL>
__global__ void __launch_bounds__(PbsCfg8<LOGN, K, LEVEL>::TPF, PbsCfg8<LOGN, K, LEVEL>::MIN_BLOCKS)
pbs_kernel_v8(PbsArgs a) {
    using C = PbsCfg8<LOGN, K, LEVEL>;
    constexpr int N = C::N, M = C::M, LOGM = C::LOGM, TPF = C::TPF, G = C::G, NF = C::NF;
    constexpr int LAST = FftPlan<LOGM>::P - 1;
    // the table-lookup shape: first two (radix-4) levels of the M = 1024 plan fused into one 16-point pass, 2 exchanges per transform
    constexpr bool FUSED16 = (LOGM == 10 && K == 1 && LEVEL == 2);
    constexpr int MAC_UNROLL = (K >= 2) ? 1 : G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);                    // [G][N]
    double2* bufs = reinterpret_cast<double2*>(acc + (size_t)G * N);          // [NF][M] swizzled, one per interleaved transform
    double2* s_tw = bufs + (size_t)NF * M;                                    // [M]; entry M-1 is not a twiddle:
    volatile uint32_t* s_ct = reinterpret_cast<volatile uint32_t*>(s_tw + (M - 1));   // it holds the hand-out slot

    const int t = threadIdx.x;
    auto sync = [] { __syncthreads(); };
    auto wsync = [] { __syncwarp(); };
    for (int i = t; i < M - 1; i += TPF) s_tw[i] = a.tw[i];

    // gadget constants, 32-bit: v = top (base_log * LEVEL) bits of the rounded word
    const int total = a.base_log * LEVEL;
    const int rsh = 32 - total;
    const uint32_t radd_hi = 1u << (31 - total);
    const uint32_t dmask = (1u << a.base_log) - 1u, dhalf = 1u << (a.base_log - 1);
    uint32_t lvl_shift[LEVEL], lvl_lowmask[LEVEL], lvl_thr[LEVEL];
#pragma unroll
    for (int lvl = 0; lvl < LEVEL; lvl++) {
        const int sh = a.base_log * (LEVEL - 1 - lvl);
        lvl_shift[lvl] = (uint32_t)sh;
        lvl_lowmask[lvl] = sh > 0 ? ((1u << sh) - 1u) : 0u;
        uint32_t rep = 0;
        for (int q = 0; q < LEVEL - 1 - lvl; q++) rep = (rep << a.base_log) | 1u;
        lvl_thr[lvl] = sh > 0 ? (dhalf - 1u) * rep + 1u : 0xffffffffu;
    }
    // exact int32 -> double: bits (0x43300000, sd ^ 0x80000000) are 2^52 + 2^31 + sd
    auto i2d = [](int32_t sd) { return __hiloint2double(0x43300000, sd ^ (int)0x80000000) - 4503601774854144.0; };
    auto digits_of = [&](uint64_t d, double (&out)[LEVEL]) {
        const uint32_t hi = (uint32_t)(d >> 32) + radd_hi;
        if constexpr (LEVEL == 1) {
            out[0] = i2d((int32_t)hi >> rsh);
        } else {
            const uint32_t v = hi >> rsh;
#pragma unroll
            for (int lvl = 0; lvl < LEVEL; lvl++) {
                uint32_t dg = (v >> lvl_shift[lvl]) & dmask;
                dg += ((v & lvl_lowmask[lvl]) >= lvl_thr[lvl]) ? 1u : 0u;
                int32_t sd = (int32_t)dg;
                if (dg >= dhalf) sd -= (int32_t)(dmask + 1u);
                out[lvl] = i2d(sd);
            }
        }
    };

    for (;;) {
        __syncthreads();
        if (t == 0) *s_ct = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const uint32_t ct = *s_ct;
        if (ct >= a.count) break;
        const uint64_t* in = a.in + (size_t)ct * (a.n + 1);
        {   // acc = X^{-bhat} * (0, .., 0, LUT)
            const uint64_t* lut = a.luts + (size_t)a.lut_index[ct] * N;
            const uint32_t bhat = mod_switch(__ldg(in + a.n), LOGN + 1);
            for (int j = t; j < K * N; j += TPF) acc[j] = 0;
            for (int j = t; j < N; j += TPF) {
                uint32_t idx = (j + bhat) & (2 * N - 1);
                acc[(size_t)K * N + j] = idx < N ? lut[idx] : (uint64_t)0 - lut[idx - N];
            }
        }
        __syncthreads();

        uint64_t a_next = __ldg(in);
        for (uint32_t i = 0; i < a.n; i++) {
            const uint32_t ahat = mod_switch(a_next, LOGN + 1);
            a_next = __ldg(in + i + 1);
            if (ahat == 0) continue;                                   // CTA-uniform
            const double2* key_i = a.bsk + (size_t)i * NF * G * M + t;

            double2 x[NF][8], w[7];
            if constexpr (FUSED16) {
                // M = 1024, two levels, k = 1: a 64-thread group owns one accumulator component; its threads hold 16 points of both
                // gadget-level transforms of that component, run the plan's two radix-4 levels without an exchange, and store
                const int grp = t >> 6, tl = t & 63;
                const uint64_t* ar = acc + (size_t)grp * N;
                double2 z[LEVEL][16];
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int jc = tl + e * 64;
                    const uint32_t s0 = (uint32_t)(jc - (int)ahat) & (2 * N - 1);
                    const uint32_t s1 = s0 + M;
                    const uint64_t m0 = (uint64_t)0 - (uint64_t)((s0 >> LOGN) & 1u), m1 = (uint64_t)0 - (uint64_t)((s1 >> LOGN) & 1u);
                    const uint64_t d0 = ((ar[s0 & (N - 1)] ^ m0) - m0) - ar[jc];
                    const uint64_t d1 = ((ar[s1 & (N - 1)] ^ m1) - m1) - ar[jc + M];
                    double g0[LEVEL], g1[LEVEL];
                    digits_of(d0, g0);
                    digits_of(d1, g1);
#pragma unroll
                    for (int lvl = 0; lvl < LEVEL; lvl++) z[lvl][e] = make_double2(g0[lvl], g1[lvl]);
                }
#pragma unroll
                for (int lvl = 0; lvl < LEVEL; lvl++) {
                    fused16_forward(z[lvl], s_tw);
                    double2* b = bufs + (size_t)(grp * LEVEL + lvl) * M;
#pragma unroll
                    for (int e = 0; e < 16; e++) b[swz(tl + e * 64)] = z[lvl][e];
                }
                __syncthreads();
                fft_forward_multi_from2<LOGM, NF>(x, w, t, bufs, s_tw, wsync);
            } else {
            load_tw<LOGM, 0>(w, t, s_tw);
            // rotation + decomposition: coefficient pairs (j, j + M) of X^ahat * acc_r - acc_r, all levels at once
#pragma unroll
            for (int r = 0; r < G; r++) {
                const uint64_t* ar = acc + (size_t)r * N;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int jc = first_pass_index<LOGM>(t, e);
                    const uint32_t s0 = (uint32_t)(jc - (int)ahat) & (2 * N - 1);
                    const uint32_t s1 = s0 + M;                        // bit LOGN of s0 / s1 = sign of the wrapped coefficient
                    const uint64_t m0 = (uint64_t)0 - (uint64_t)((s0 >> LOGN) & 1u), m1 = (uint64_t)0 - (uint64_t)((s1 >> LOGN) & 1u);
                    const uint64_t d0 = ((ar[s0 & (N - 1)] ^ m0) - m0) - ar[jc];
                    const uint64_t d1 = ((ar[s1 & (N - 1)] ^ m1) - m1) - ar[jc + M];
                    double g0[LEVEL], g1[LEVEL];
                    digits_of(d0, g0);
                    digits_of(d1, g1);
#pragma unroll
                    for (int lvl = 0; lvl < LEVEL; lvl++) x[r * LEVEL + lvl][e] = make_double2(g0[lvl], g1[lvl]);
                }
            }
            fft_forward_multi<LOGM, NF>(x, w, t, bufs, s_tw, sync, wsync);   // w now holds the last pass's twiddles
            }

            // Fourier MAC, one output component at a time (fma chain over f = r * l + lvl ascending), then that spectrum's
            // first inverse pass (the same node twiddles, conjugated) and its store for the interleaved remainder
            // k = 2: the three output components run as a rolled loop (61 KB -> 45 KB of straight-line code, no spill: +3 % measured);
            // k = 1 keeps both components unrolled (rolled: -14 %)
#pragma unroll (MAC_UNROLL)
            for (int c = 0; c < G; c++) {
                double2 o[8];
#pragma unroll
                for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
#pragma unroll
                for (int f = 0; f < NF; f++) {
                    const double2* krow = key_i + ((size_t)f * G + c) * M;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const double2 kv = __ldg(krow + e * TPF);
                        double re = o[e].x, im = o[e].y;
                        re = fma(x[f][e].x, kv.x, re); re = fma(-x[f][e].y, kv.y, re);
                        im = fma(x[f][e].x, kv.y, im); im = fma(x[f][e].y, kv.x, im);
                        o[e] = make_double2(re, im);
                    }
                }
                pass_nodes<LOGM, LAST, true>(o, w);
                pass_store<LOGM, LAST>(o, t, bufs + (size_t)c * M);
            }
            if constexpr (FUSED16) {
                fft_inverse_multi_pass2<LOGM, G>(t, bufs, s_tw, wsync);
                __syncthreads();
                const int grp = t >> 6, tl = t & 63;                        // group c finishes output component c
                const double2* b = bufs + (size_t)grp * M;
                double2 z[16];
#pragma unroll
                for (int e = 0; e < 16; e++) z[e] = b[swz(tl + e * 64)];
                fused16_inverse(z, s_tw);
                uint64_t* ac = acc + (size_t)grp * N;
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int jc = tl + e * 64;
                    ac[jc] += double_to_torus(z[e].x);
                    ac[jc + M] += double_to_torus(z[e].y);
                }
            } else {
            double2 y[G][8];
            fft_inverse_multi_rest<LOGM, G>(y, t, bufs, s_tw, sync, wsync);
#pragma unroll
            for (int c = 0; c < G; c++) {
                uint64_t* ac = acc + (size_t)c * N;
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int jc = first_pass_index<LOGM>(t, e);                          // the key carries the 1/M of the inverse transform
                    ac[jc] += double_to_torus(y[c][e].x);
                    ac[jc + M] += double_to_torus(y[c][e].y);
                }
            }
            }
            __syncthreads();                                           // accumulator complete, buffers free
        }

        uint64_t* o = a.out + (size_t)ct * ((size_t)a.big_dim + 1);
        for (int j = t; j < K * N; j += TPF) {
            int comp = j / N, tt = j - comp * N;
            const uint64_t* ar = acc + (size_t)comp * N;
            uint64_t v = (tt == 0) ? ar[0] : (uint64_t)0 - ar[N - tt];
            if (a.mode == 0) o[j] = v; else o[j] -= v;
        }
        if (a.mode == 0) for (uint32_t j = K * N + t; j < a.big_dim; j += TPF) o[j] = 0;
        if (t == 0) {
            uint64_t bv = acc[(size_t)K * N];
            if (a.mode == 0) o[a.big_dim] = bv; else o[a.big_dim] -= bv + a.body_const;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Plain transforms: one group of TPF threads per polynomial.  MODE 0: u64 (signed) input -> thread-major
// spectrum scaled by 1/M (BSK conversion).  MODE 1: double input -> canonical order, unscaled (test hook).
// ---------------------------------------------------------------------------------------------------
template <int LOGN, int MODE>
__global__ void __launch_bounds__(1 << (LOGN - 4))
fft_forward_kernel(const void* __restrict__ in, double2* __restrict__ out, const double2* __restrict__ tw, size_t polys) {
    constexpr int N = 1 << LOGN, M = N / 2, LOGM = LOGN - 1, TPF = M / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    double2* s_tw = buf + M;
    const int t = threadIdx.x;
    for (int i = t; i < M; i += TPF) s_tw[i] = tw[i];
    auto sync = [] { __syncthreads(); };
    for (size_t p = blockIdx.x; p < polys; p += gridDim.x) {
        __syncthreads();
        double2 x[8], w[7];
        load_tw<LOGM, 0>(w, t, s_tw);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int jc = first_pass_index<LOGM>(t, e);
            if (MODE == 0) {
                const uint64_t* src = reinterpret_cast<const uint64_t*>(in) + p * N;
                x[e] = make_double2((double)(int64_t)src[jc], (double)(int64_t)src[jc + M]);
            } else {
                const double* src = reinterpret_cast<const double*>(in) + p * N;
                x[e] = make_double2(src[jc], src[jc + M]);
            }
        }
        fft_forward_regs2<LOGM, false>(x, x, w, t, buf, buf, s_tw, sync, sync);
        double2* dst = out + p * M;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            if (MODE == 0) dst[e * TPF + t] = make_double2(x[e].x * (1.0 / M), x[e].y * (1.0 / M));
            else dst[last_pass_index<LOGM>(t, e)] = x[e];
        }
    }
}

// canonical spectrum -> torus polynomial: inverse transform, exact 1/M, round (test hook)
template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 4))
fft_inverse_kernel(const double2* __restrict__ in, uint64_t* __restrict__ out, const double2* __restrict__ tw, size_t polys) {
    constexpr int N = 1 << LOGN, M = N / 2, LOGM = LOGN - 1, TPF = M / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    double2* s_tw = buf + M;
    const int t = threadIdx.x;
    for (int i = t; i < M; i += TPF) s_tw[i] = tw[i];
    auto sync = [] { __syncthreads(); };
    for (size_t p = blockIdx.x; p < polys; p += gridDim.x) {
        __syncthreads();
        double2 x[8];
        const double2* src = in + p * M;
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = src[last_pass_index<LOGM>(t, e)];
        fft_inverse_regs2<LOGM, false>(x, x, t, buf, buf, s_tw, sync, sync);
        uint64_t* dst = out + p * N;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int jc = first_pass_index<LOGM>(t, e);
            dst[jc] = double_to_torus(x[e].x * (1.0 / M));
            dst[jc + M] = double_to_torus(x[e].y * (1.0 / M));
        }
    }
}

// thread-major <-> canonical permutation of one Fourier polynomial (export / import of the BSK)
template <int LOGN>
__global__ void bsk_permute_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t polys, int to_canonical) {
    constexpr int M = 1 << (LOGN - 1), LOGM = LOGN - 1, TPF = M / 8;
    size_t total = polys * M;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t p = i / M; int r = (int)(i - p * M);
        int e = r / TPF, t = r - e * TPF;
        int canon = last_pass_index<LOGM>(t, e);
        if (to_canonical) out[p * M + canon] = in[i]; else out[i] = in[p * M + canon];
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------------
template <int LOGN, int K>
static int launch_pbs_t(const PbsArgs& a, int sm_count, cudaStream_t stream) {
    using C = PbsCfg<LOGN, K>;
    size_t smem = C::smem_bytes((int)a.n);
    int blocks_per_sm = 1;
    if (smem > 227 * 1024) return set_error(TFX_ERR_UNSUPPORTED, "pbs: shared memory footprint exceeds 227 KB");
    // per launch: the attribute is per device, and a process may drive several devices / host threads
    cudaError_t e = cudaFuncSetAttribute(pbs_kernel<LOGN, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(pbs)");
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, pbs_kernel<LOGN, K>, C::THREADS, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "occupancy(pbs)");
    if (blocks_per_sm < 1) return set_error(TFX_ERR_UNSUPPORTED, "pbs: kernel does not fit on an SM");
    unsigned grid = (unsigned)sm_count * blocks_per_sm;
    if (grid > a.count) grid = a.count;
    // measurement knob (profiles/r01_pbs_experiments.md): cap the grid, e.g. to one CTA per SM
    if (const char* cap = getenv("TFX_PBS_GRID_CAP")) { unsigned g = (unsigned)atoi(cap); if (g && g < grid) grid = g; }
    pbs_kernel<LOGN, K><<<grid, C::THREADS, smem, stream>>>(a);
    count_launch();
    return check_launch("pbs_kernel");
}

template <int LOGN, int K, int LEVEL>
static int launch_pbs8_t(const PbsArgs& a, int sm_count, cudaStream_t stream) {
    using C = PbsCfg8<LOGN, K, LEVEL>;
    const size_t smem = C::SMEM;
    auto kern = pbs_kernel_v8<LOGN, K, LEVEL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(pbs v8)");
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(pbs v8 carveout)");
    int blocks_per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, C::TPF, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "occupancy(pbs v8)");
    if (blocks_per_sm < 1) return set_error(TFX_ERR_UNSUPPORTED, "pbs v8: kernel does not fit on an SM");
    if (getenv("TFX_PBS_VERBOSE")) fprintf(stderr, "[tfx] pbs_kernel_v8<%d,%d,%d>: %d CTAs/SM, %zu B smem\n", LOGN, K, LEVEL, blocks_per_sm, smem);
    unsigned grid = (unsigned)sm_count * blocks_per_sm;
    if (grid > a.count) grid = a.count;
    if (const char* cap = getenv("TFX_PBS_GRID_CAP")) { unsigned g = (unsigned)atoi(cap); if (g && g < grid) grid = g; }
    kern<<<grid, C::TPF, smem, stream>>>(a);
    count_launch();
    return check_launch("pbs_kernel_v8");
}

// v8 instantiations: the shipped sets and their neighbours; everything else runs the general kernel
static bool pbs8_dispatch(const PbsLaunch& p, const PbsArgs& a, cudaStream_t stream, int* rc) {
    if (p.base_log * p.level > 31 || getenv("TFX_PBS_V7")) return false;
#define TFX_PBS8_CASE(LN, KK, LL) if (p.N == (1u << LN) && p.k == KK && p.level == LL) { *rc = launch_pbs8_t<LN, KK, LL>(a, p.sm_count, stream); return true; }
    TFX_PBS8_CASE(10, 2, 1) TFX_PBS8_CASE(11, 1, 2) TFX_PBS8_CASE(11, 1, 1) TFX_PBS8_CASE(10, 1, 1) TFX_PBS8_CASE(10, 1, 2)
    TFX_PBS8_CASE(9, 2, 1) TFX_PBS8_CASE(9, 1, 2)
#undef TFX_PBS8_CASE
    return false;
}

int pbs_supported(uint32_t N, uint32_t k) {
    if (k == 1) return N == 512 || N == 1024 || N == 2048 || N == 4096 || N == 8192;
    if (k == 2) return N == 512 || N == 1024 || N == 2048;
    return 0;
}

int launch_pbs(const PbsLaunch& p, cudaStream_t stream) {
    PbsArgs a;
    a.bsk = reinterpret_cast<const double2*>(p.bsk);
    a.tw = reinterpret_cast<const double2*>(p.tw);
    a.in = p.in; a.luts = p.luts; a.lut_index = p.lut_index; a.out = p.out;
    a.work_counter = p.work_counter;
    {
        cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaMemsetAsync(pbs work counter)");
    }
    a.n = p.n; a.big_dim = p.big_dim; a.base_log = p.base_log; a.level = p.level; a.mode = p.mode; a.body_const = p.body_const;
    a.count = (uint32_t)p.count;
    { int rc8 = 0; if (pbs8_dispatch(p, a, stream, &rc8)) return rc8; }
#define TFX_PBS_CASE(LN, KK) if (p.N == (1u << LN) && p.k == KK) return launch_pbs_t<LN, KK>(a, p.sm_count, stream);
    TFX_PBS_CASE(9, 1) TFX_PBS_CASE(10, 1) TFX_PBS_CASE(11, 1) TFX_PBS_CASE(12, 1) TFX_PBS_CASE(13, 1)
    TFX_PBS_CASE(9, 2) TFX_PBS_CASE(10, 2) TFX_PBS_CASE(11, 2)
#undef TFX_PBS_CASE
    return set_error(TFX_ERR_UNSUPPORTED, "pbs: no kernel compiled for this (N, k)");
}

template <int LOGN>
static int launch_fft_t(int which, const void* in, void* out, const double* tw, size_t polys,
                        int sm_count, cudaStream_t stream) {
    constexpr int M = 1 << (LOGN - 1), TPF = M / 8;
    size_t smem = (size_t)M * 16 * 2;
    unsigned grid = (unsigned)(polys < (size_t)sm_count * 8 ? polys : (size_t)sm_count * 8);
    if (grid == 0) return TFX_OK;
    const double2* twd = reinterpret_cast<const double2*>(tw);
    cudaError_t e;
    if (which == 0) {
        e = cudaFuncSetAttribute(fft_forward_kernel<LOGN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "attr");
        fft_forward_kernel<LOGN, 0><<<grid, TPF, smem, stream>>>(in, reinterpret_cast<double2*>(out), twd, polys);
    } else if (which == 1) {
        e = cudaFuncSetAttribute(fft_forward_kernel<LOGN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "attr");
        fft_forward_kernel<LOGN, 1><<<grid, TPF, smem, stream>>>(in, reinterpret_cast<double2*>(out), twd, polys);
    } else if (which == 2) {
        e = cudaFuncSetAttribute(fft_inverse_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_cuda_error(e, "attr");
        fft_inverse_kernel<LOGN><<<grid, TPF, smem, stream>>>(reinterpret_cast<const double2*>(in),
                                                                reinterpret_cast<uint64_t*>(out), twd, polys);
    } else {
        bsk_permute_kernel<LOGN><<<(unsigned)sm_count * 4, 256, 0, stream>>>(
            reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), polys, which == 3 ? 1 : 0);
    }
    count_launch();
    return check_launch("fft kernel");
}

// which: 0 = u64 -> thread-major Fourier, 1 = double -> canonical Fourier, 2 = canonical Fourier -> torus,
//        3 = thread-major -> canonical, 4 = canonical -> thread-major
int launch_fft(int which, uint32_t N, const void* in, void* out, const double* tw, size_t polys,
               int sm_count, cudaStream_t stream) {
    switch (N) {
        case 512:  return launch_fft_t<9>(which, in, out, tw, polys, sm_count, stream);
        case 1024: return launch_fft_t<10>(which, in, out, tw, polys, sm_count, stream);
        case 2048: return launch_fft_t<11>(which, in, out, tw, polys, sm_count, stream);
        case 4096: return launch_fft_t<12>(which, in, out, tw, polys, sm_count, stream);
        case 8192: return launch_fft_t<13>(which, in, out, tw, polys, sm_count, stream);
        default: return set_error(TFX_ERR_UNSUPPORTED, "fft: unsupported N");
    }
}

}  // namespace tfx
