// tfx_common.cuh — device primitives shared by the TFHE kernels (sm_100a).
// Floating point discipline: this library is compiled with -fmad=false; every fused multiply-add is an
// explicit fma().  The butterfly dataflow is radix-2 (DIF forward / DIT inverse) regrouped into register
// passes of 2-3 stages; regrouping does not change any floating-point operation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfx {

// ------------------------------------------------------------------------------------------------
// Counter-mode PRF: ChaCha20 block function, key = seed || seed ^ 0xA5.., nonce = stream, counter = block.
// ------------------------------------------------------------------------------------------------
struct Seed { uint32_t w[4]; };

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

#define TFX_QR(a, b, c, d)                                                                         \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);                          \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

__device__ __forceinline__ void chacha_block(const Seed& seed, uint64_t stream, uint64_t block, uint64_t out[8]) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 4; i++) { s[4 + i] = seed.w[i]; s[8 + i] = seed.w[i] ^ 0xA5A5A5A5u; }
    s[12] = (uint32_t)block; s[13] = (uint32_t)(block >> 32);
    s[14] = (uint32_t)stream; s[15] = (uint32_t)(stream >> 32);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        TFX_QR(x[0], x[4], x[8], x[12]) TFX_QR(x[1], x[5], x[9], x[13])
        TFX_QR(x[2], x[6], x[10], x[14]) TFX_QR(x[3], x[7], x[11], x[15])
        TFX_QR(x[0], x[5], x[10], x[15]) TFX_QR(x[1], x[6], x[11], x[12])
        TFX_QR(x[2], x[7], x[8], x[13]) TFX_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t lo = x[2 * i] + s[2 * i], hi = x[2 * i + 1] + s[2 * i + 1];
        out[i] = ((uint64_t)hi << 32) | lo;
    }
}

__device__ __forceinline__ uint64_t prf_u64(const Seed& seed, uint64_t stream, uint64_t idx) {
    uint64_t blk[8];
    chacha_block(seed, stream, idx >> 3, blk);
    uint64_t r = blk[0];
#pragma unroll
    for (int i = 1; i < 8; i++) if ((idx & 7) == (uint64_t)i) r = blk[i];
    return r;
}

enum { ST_BIGKEY = 1, ST_SMALLKEY = 2, ST_KSK_MASK = 3, ST_KSK_NOISE = 4, ST_BSK_MASK = 5, ST_BSK_NOISE = 6,
       ST_ENC_MASK = 7, ST_ENC_NOISE = 8 };
__host__ __device__ __forceinline__ uint64_t stream_id(int purpose, uint64_t set, uint64_t index) {
    return ((uint64_t)purpose << 56) | (set << 48) | index;
}

// ------------------------------------------------------------------------------------------------
// Deterministic Gaussian (Box-Muller; ln and sincos from IEEE basic operations only).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double det_ln(double x) {
    uint64_t bits = (uint64_t)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = __longlong_as_double((long long)bits);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 27.0;
#pragma unroll
    for (int k = 25; k >= 1; k -= 2) p = fma(p, s2, 1.0 / (double)k);
    return fma((double)e, 0.6931471805599453, 2.0 * (s * p));
}

__device__ __forceinline__ double det_cos_turn(double u) {
    double t = u * 4.0;
    int q = (int)t;
    double g = t - (double)q;
    double a = g * 1.5707963267948966;
    double a2 = a * a;
    double cs = 0.0, sn = 0.0;
#pragma unroll
    for (int k = 14; k >= 1; k--) {
        cs = fma(cs, a2, 1.0) * (-1.0 / (double)((2 * k - 1) * (2 * k)));
        sn = fma(sn, a2, 1.0) * (-1.0 / (double)((2 * k) * (2 * k + 1)));
    }
    cs = fma(cs, a2, 1.0);
    sn = fma(sn, a2, 1.0) * a;
    switch (q & 3) {
        case 0: return cs;
        case 1: return -sn;
        case 2: return -cs;
        default: return sn;
    }
}

__device__ __forceinline__ double prf_gauss(const Seed& seed, uint64_t stream, uint64_t idx) {
    uint64_t blk[8];
    uint64_t w = 2 * idx;               // words w, w+1 live in the same block (w even)
    chacha_block(seed, stream, w >> 3, blk);
    uint64_t x = blk[0], y = blk[1];
#pragma unroll
    for (int i = 2; i < 8; i += 2) if ((w & 7) == (uint64_t)i) { x = blk[i]; y = blk[i + 1]; }
    double u1 = (double)((x >> 11) + 1) * 0x1p-53;
    double u2 = (double)(y >> 11) * 0x1p-53;
    double r = sqrt(-2.0 * det_ln(u1));
    return r * det_cos_turn(u2);
}

__device__ __forceinline__ uint64_t prf_noise(const Seed& seed, uint64_t stream, uint64_t idx, double std) {
    double v = prf_gauss(seed, stream, idx) * (std * 0x1p64);
    return (uint64_t)__double2ll_rn(v);
}

// ------------------------------------------------------------------------------------------------
// Integer primitives
// ------------------------------------------------------------------------------------------------
// signed gadget digit `lvl` (1-based, weight q/B^lvl) of x; tie rule: raw digit == B/2 -> -B/2 with carry.
__device__ __forceinline__ int64_t decompose_digit(uint64_t x, int base_log, int level, int lvl) {
    int total = base_log * level;
    uint64_t v = (total < 64) ? ((x + (1ULL << (63 - total))) >> (64 - total)) : x;
    uint64_t B = 1ULL << base_log, half = B >> 1, mask = B - 1;
    int64_t d = 0;
    for (int q = level; q >= lvl; q--) {
        uint64_t r = v & mask;
        v >>= base_log;
        if (r >= half) { d = (int64_t)r - (int64_t)B; v += 1; }
        else d = (int64_t)r;
    }
    return d;
}

__device__ __forceinline__ uint32_t mod_switch(uint64_t x, int log2_2N) {
    return (uint32_t)((((x >> (64 - log2_2N - 1)) + 1) >> 1) & ((1u << log2_2N) - 1));
}

// double (integer valued, |v| < 2^117) -> torus word mod 2^64
__device__ __forceinline__ uint64_t double_to_torus(double v) {
    double r = rint(v * 0x1p-64);
    double y = fma(-r, 0x1p64, v);
    if (y >= 0x1p63) y -= 0x1p64;
    if (y < -0x1p63) y += 0x1p64;
    return (uint64_t)__double2ll_rn(y);
}

// ------------------------------------------------------------------------------------------------
// Complex helpers (definitions fixed by the parity contract, see oracle/tfhe_oracle.c §6)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// ------------------------------------------------------------------------------------------------
// FFT plan: M = 2^LOGM complex points, TPF = M/8 threads, 8 points per thread per pass.
// Pass p works on an index bit field of width WD[p] (2 or 3) whose lowest bit is LO[p].
// Field layouts are chosen so the XOR swizzle below keeps every quarter-warp access conflict free.
// ------------------------------------------------------------------------------------------------
template <int LOGM> struct FftPlan;
template <> struct FftPlan<8>  { static constexpr int P = 3; static constexpr int WD[4] = {3, 3, 2, 0}; };
template <> struct FftPlan<9>  { static constexpr int P = 3; static constexpr int WD[4] = {3, 3, 3, 0}; };
template <> struct FftPlan<10> { static constexpr int P = 4; static constexpr int WD[4] = {3, 2, 3, 2}; };
template <> struct FftPlan<11> { static constexpr int P = 4; static constexpr int WD[4] = {3, 3, 3, 2}; };
template <> struct FftPlan<12> { static constexpr int P = 4; static constexpr int WD[4] = {3, 3, 3, 3}; };

template <int LOGM, int PASS> struct PassInfo {
    static constexpr int WD = FftPlan<LOGM>::WD[PASS];
    static constexpr int lo_calc() { int s = 0; for (int q = 0; q <= PASS; q++) s += FftPlan<LOGM>::WD[q]; return LOGM - s; }
    static constexpr int LO = lo_calc();
};

__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 3) & 7); }

// "rest" index of sub-group u (0/1) of thread t in a 2-bit pass: [warp][u][lane].  With this placement every pass
// after the first one touches only the 256 points owned by the thread's warp (index >> 8 == warp), so those passes
// need warp-level synchronisation only.
__device__ __forceinline__ int pass2_rest(int t, int u) { return ((t >> 5) << 6) | (u << 5) | (t & 31); }

// index (complex position in the M-array) of element e of thread t in a pass
template <int LOGM, int LO, int WD>
__device__ __forceinline__ int elem_index(int t, int e) {
    constexpr int TPF = 1 << (LOGM - 3);
    int rest, f;
    if (WD == 3) { rest = t; f = e; }
    else { rest = pass2_rest(t, e >> 2); f = e & 3; }
    (void)TPF;
    return ((rest >> LO) << (LO + WD)) | (f << LO) | (rest & ((1 << LO) - 1));
}

// radix-2 stages of one pass on the 8 register values.  tw: flat twiddle table (stage with half h at offset M-2h).
template <int LOGM, int LO, int WD, bool INV>
__device__ __forceinline__ void pass_butterflies(double2 (&x)[8], int t, const double2* __restrict__ tw) {
    constexpr int M = 1 << LOGM;
#pragma unroll
    for (int q = 0; q < WD; q++) {
        const int fb = INV ? q : (WD - 1 - q);            // field bit handled by this stage
        const int half = 1 << (LO + fb);
        const int off = M - 2 * half;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int f = (WD == 3) ? e : (e & 3);
            if (f & (1 << fb)) continue;                    // e is the 'a' element of its pair
            const int eb = e | (1 << fb);
            if (half == 1) {
                double2 a = x[e], b = x[eb];
                x[e] = cadd(a, b); x[eb] = csub(a, b);
            } else {
                const int rest = (WD == 3) ? t : pass2_rest(t, e >> 2);
                const int j = ((f & ((1 << fb) - 1)) << LO) | (rest & ((1 << LO) - 1));
                const double2 w = tw[off + j];
                if (!INV) {
                    double2 a = x[e], b = x[eb];
                    x[e] = cadd(a, b);
                    x[eb] = cmul(csub(a, b), w);
                } else {
                    double2 a = x[e], b = cmulc(x[eb], w);
                    x[e] = cadd(a, b); x[eb] = csub(a, b);
                }
            }
        }
    }
}

template <int LOGM, int PASS>
__device__ __forceinline__ void pass_store(const double2 (&x)[8], int t, double2* __restrict__ buf) {
#ifdef TFX_EXP_NOSMEM
    return;
#endif
    using PI = PassInfo<LOGM, PASS>;
#pragma unroll
    for (int e = 0; e < 8; e++) buf[swz(elem_index<LOGM, PI::LO, PI::WD>(t, e))] = x[e];
}
template <int LOGM, int PASS>
__device__ __forceinline__ void pass_load(double2 (&x)[8], int t, const double2* __restrict__ buf) {
#ifdef TFX_EXP_NOSMEM
    return;
#endif
    using PI = PassInfo<LOGM, PASS>;
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = buf[swz(elem_index<LOGM, PI::LO, PI::WD>(t, e))];
}

// Forward passes 1..P-1 given pass-0 INPUT values already in x (pass-0 element order: index = t + e*TPF).
// On return x holds the frequency-domain values of the last pass's elements (index elem_index<LAST>(t, e)).
// SYNC is a functor performing the barrier for the threads sharing `buf`.
// SYNC_CTA orders the first (cross-warp) exchange, SYNC_WARP the later, warp-local ones.
template <int LOGM, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_forward_regs(double2 (&x)[8], int t, double2* __restrict__ buf,
                                                 const double2* __restrict__ tw, SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    pass_butterflies<LOGM, PassInfo<LOGM, 0>::LO, PassInfo<LOGM, 0>::WD, false>(x, t, tw);
    pass_store<LOGM, 0>(x, t, buf);
    sync();
    pass_load<LOGM, 1>(x, t, buf);
    pass_butterflies<LOGM, PassInfo<LOGM, 1>::LO, PassInfo<LOGM, 1>::WD, false>(x, t, tw);
    if constexpr (PL::P >= 3) {
        pass_store<LOGM, 1>(x, t, buf);                 // in place: a thread rewrites exactly the elements it loaded
        wsync();
        pass_load<LOGM, 2>(x, t, buf);
        pass_butterflies<LOGM, PassInfo<LOGM, 2>::LO, PassInfo<LOGM, 2>::WD, false>(x, t, tw);
    }
    if constexpr (PL::P >= 4) {
        pass_store<LOGM, 2>(x, t, buf);
        wsync();
        pass_load<LOGM, 3>(x, t, buf);
        pass_butterflies<LOGM, PassInfo<LOGM, 3>::LO, PassInfo<LOGM, 3>::WD, false>(x, t, tw);
    }
}
template <int LOGM, typename SYNC>
__device__ __forceinline__ void fft_forward_regs(double2 (&x)[8], int t, double2* __restrict__ buf,
                                                 const double2* __restrict__ tw, SYNC sync) {
    fft_forward_regs<LOGM>(x, t, buf, tw, sync, sync);
}

// ---- twiddles of a pass preloaded into registers (issued before the preceding barrier so their shared-memory
//      latency overlaps the barrier wait and the data loads) ----------------------------------------------------
// slot layout: stage with field bit fb uses slots [(1 << WD) - (2 << fb), ...) indexed by the low fb bits of f
template <int WD> struct TwSlots { static constexpr int N = (1 << WD) - 1; };

template <int LOGM, int LO, int WD>
__device__ __forceinline__ void load_tw(double2 (&w)[7], int t, const double2* __restrict__ tw) {
    constexpr int M = 1 << LOGM;
    const int rest_lo = ((WD == 3) ? t : pass2_rest(t, 0)) & ((1 << LO) - 1);
#pragma unroll
    for (int fb = WD - 1; fb >= 0; fb--) {
        const int half = 1 << (LO + fb);
        if (half == 1) continue;
        const int off = M - 2 * half;
        const int base = (1 << WD) - (2 << fb);
#pragma unroll
        for (int fl = 0; fl < (1 << fb); fl++) {
            if (LO == 0 && (fl == 0 || 2 * fl == half)) continue;           // exactly 1 and i: handled without a multiply
            w[base + fl] = tw[off + ((fl << LO) | rest_lo)];
        }
    }
}

template <int LOGM, int LO, int WD, bool INV>
__device__ __forceinline__ void pass_butterflies_w(double2 (&x)[8], const double2 (&w)[7]) {
#ifdef TFX_EXP_NOMATH
    return;
#endif
#pragma unroll
    for (int q = 0; q < WD; q++) {
        const int fb = INV ? q : (WD - 1 - q);
        const int half = 1 << (LO + fb);
        const int base = (1 << WD) - (2 << fb);
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int f = (WD == 3) ? e : (e & 3);
            if (f & (1 << fb)) continue;
            const int eb = e | (1 << fb);
            const int fl = f & ((1 << fb) - 1);
            const bool is_one = (half == 1) || (LO == 0 && fl == 0);
            const bool is_i = (half != 1) && (LO == 0) && (2 * fl == half);
            if (!INV) {
                const double2 a = x[e], b = x[eb];
                x[e] = cadd(a, b);
                const double2 d = csub(a, b);
                if (is_one) x[eb] = d;
                else if (is_i) x[eb] = make_double2(-d.y, d.x);              // d * i
                else x[eb] = cmul(d, w[base + fl]);
            } else {
                double2 b = x[eb];
                if (is_one) {}
                else if (is_i) b = make_double2(b.y, -b.x);                  // b * conj(i)
                else b = cmulc(b, w[base + fl]);
                const double2 a = x[e];
                x[e] = cadd(a, b); x[eb] = csub(a, b);
            }
        }
    }
}

// ---- two transforms interleaved in one thread (independent instruction streams hide shared-memory latency) ----
#define TFX_TW(P) load_tw<LOGM, PassInfo<LOGM, P>::LO, PassInfo<LOGM, P>::WD>(w, t, tw);
#define TFX_PASS2(P, INV) \
    pass_butterflies_w<LOGM, PassInfo<LOGM, P>::LO, PassInfo<LOGM, P>::WD, INV>(xa, w); \
    if (DUAL) pass_butterflies_w<LOGM, PassInfo<LOGM, P>::LO, PassInfo<LOGM, P>::WD, INV>(xb, w);
#define TFX_STORE2(P) pass_store<LOGM, P>(xa, t, bufa); if (DUAL) pass_store<LOGM, P>(xb, t, bufb);
#define TFX_LOAD2(P) pass_load<LOGM, P>(xa, t, bufa); if (DUAL) pass_load<LOGM, P>(xb, t, bufb);

// forward: xa/xb hold pass-0 inputs, w the pass-0 twiddles (load_tw<.., pass 0>); on return xa/xb hold the spectra
template <int LOGM, bool DUAL, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_forward_regs2(double2 (&xa)[8], double2 (&xb)[8], double2 (&w)[7], int t,
                                                  double2* __restrict__ bufa, double2* __restrict__ bufb,
                                                  const double2* __restrict__ tw, SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    TFX_PASS2(0, false)
    TFX_TW(1)
    sync();                                   // every warp is done reading the buffers' previous contents
    TFX_STORE2(0)
    sync();
    TFX_LOAD2(1)
    TFX_PASS2(1, false)
    if constexpr (PL::P >= 3) {
        TFX_STORE2(1)
        TFX_TW(2)
        wsync();
        TFX_LOAD2(2)
        TFX_PASS2(2, false)
    }
    if constexpr (PL::P >= 4) {
        TFX_STORE2(2)
        TFX_TW(3)
        wsync();
        TFX_LOAD2(3)
        TFX_PASS2(3, false)
    }
}

// inverse: xa/xb hold spectra (last-pass element order); on return pass-0 elements before untwist / scaling
template <int LOGM, bool DUAL, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_inverse_regs2(double2 (&xa)[8], double2 (&xb)[8], int t, double2* __restrict__ bufa,
                                                  double2* __restrict__ bufb, const double2* __restrict__ tw,
                                                  SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    constexpr int LAST = PL::P - 1;
    double2 w[7];
    TFX_TW(LAST)
    TFX_PASS2(LAST, true)
    sync();                                   // every warp is done reading the buffers' previous contents
    TFX_STORE2(LAST)
    if constexpr (PL::P >= 4) {
        TFX_TW(2)
        wsync();
        TFX_LOAD2(2)
        TFX_PASS2(2, true)
        TFX_STORE2(2)
    }
    TFX_TW(1)
    wsync();
    TFX_LOAD2(1)
    TFX_PASS2(1, true)
    TFX_STORE2(1)
    TFX_TW(0)
    sync();
    TFX_LOAD2(0)
    TFX_PASS2(0, true)
}
#undef TFX_TW
#undef TFX_PASS2
#undef TFX_STORE2
#undef TFX_LOAD2

// Inverse: x holds last-pass elements in the frequency domain; on return x holds pass-0 elements
// (index t + e*TPF) BEFORE the untwist / 1/M scaling.
template <int LOGM, typename SYNC>
__device__ __forceinline__ void fft_inverse_regs(double2 (&x)[8], int t, double2* __restrict__ buf,
                                                 const double2* __restrict__ tw, SYNC sync) {
    using PL = FftPlan<LOGM>;
    if constexpr (PL::P >= 4) {
        pass_butterflies<LOGM, PassInfo<LOGM, 3>::LO, PassInfo<LOGM, 3>::WD, true>(x, t, tw);
        pass_store<LOGM, 3>(x, t, buf);
        sync();
        pass_load<LOGM, 2>(x, t, buf);
    }
    if constexpr (PL::P >= 3) {
        pass_butterflies<LOGM, PassInfo<LOGM, 2>::LO, PassInfo<LOGM, 2>::WD, true>(x, t, tw);
        pass_store<LOGM, 2>(x, t, buf);
        sync();
        pass_load<LOGM, 1>(x, t, buf);
    }
    pass_butterflies<LOGM, PassInfo<LOGM, 1>::LO, PassInfo<LOGM, 1>::WD, true>(x, t, tw);
    pass_store<LOGM, 1>(x, t, buf);
    sync();
    pass_load<LOGM, 0>(x, t, buf);
    pass_butterflies<LOGM, PassInfo<LOGM, 0>::LO, PassInfo<LOGM, 0>::WD, true>(x, t, tw);
}

// Inverse transform after the caller has run the LAST pass's butterflies from registers and stored them:
// loads the next pass and finishes; on return x holds pass-0 elements before untwist / scaling.
// Precondition: the last pass's values were stored and a warp-level sync done.  Warp-local passes use wsync,
// the final cross-warp exchange uses sync.
template <int LOGM, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_inverse_tail(double2 (&x)[8], int t, double2* __restrict__ buf,
                                                 const double2* __restrict__ tw, SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    if constexpr (PL::P >= 4) {
        pass_load<LOGM, 2>(x, t, buf);
        pass_butterflies<LOGM, PassInfo<LOGM, 2>::LO, PassInfo<LOGM, 2>::WD, true>(x, t, tw);
        pass_store<LOGM, 2>(x, t, buf);
        wsync();
    }
    pass_load<LOGM, 1>(x, t, buf);
    pass_butterflies<LOGM, PassInfo<LOGM, 1>::LO, PassInfo<LOGM, 1>::WD, true>(x, t, tw);
    pass_store<LOGM, 1>(x, t, buf);
    sync();
    pass_load<LOGM, 0>(x, t, buf);
    pass_butterflies<LOGM, PassInfo<LOGM, 0>::LO, PassInfo<LOGM, 0>::WD, true>(x, t, tw);
}

// canonical FFT position of element e of thread t after the last forward pass
template <int LOGM>
__device__ __forceinline__ int last_pass_index(int t, int e) {
    constexpr int LAST = FftPlan<LOGM>::P - 1;
    return elem_index<LOGM, PassInfo<LOGM, LAST>::LO, PassInfo<LOGM, LAST>::WD>(t, e);
}

}  // namespace tfx
