// tfx_common.cuh — device primitives shared by the TFHE kernels (sm_100a).
// Floating point discipline: this library is compiled with -fmad=false; every fused multiply-add is an
// explicit fma().  The transform dataflow (radix-8 / radix-4 nodes of the negacyclic factor tree) is defined by
// oracle/tfhe_oracle.c section 6 and mirrored here operation for operation.
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

// double (integer valued, |v| < 2^117) -> torus word mod 2^64.
// y = v - 2^64 * rint(v / 2^64) lies in [-2^63, 2^63] (exact); round-to-nearest-even conversion.  The only value the
// saturating conversion gets wrong is y == +2^63 (must wrap to -2^63): it comes back as INT64_MAX, which no in-range y
// can produce (doubles near 2^63 are multiples of 1024), so it is patched with integer ops (off the FP64 pipe).
__device__ __forceinline__ uint64_t double_to_torus(double v) {
    // rint(v * 2^-64) without the conversion pipe: adding 1.5 * 2^52 rounds to nearest even at the unit place (|v| < 2^115)
    const double r = fma(v, 0x1p-64, 6755399441055744.0) - 6755399441055744.0;
    const double y = fma(-r, 0x1p64, v);
    const long long q = __double2ll_rn(y);
    return (uint64_t)q + (q == 0x7fffffffffffffffLL ? 1ULL : 0ULL);
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
// Negacyclic transform (definition: oracle/tfhe_oracle.c section 6): z_j = p_j + i p_{j+M} evaluated at the roots of
// X^M = i by a tree of radix-8 / radix-4 nodes, no separate twist.  M = 2^LOGM complex points, TPF = M/8 threads,
// 8 points per thread per pass.  Pass p works on an index bit field of width WD[p] (2 or 3) whose lowest bit is
// LO[p]; a 3-bit pass is one radix-8 node per thread, a 2-bit pass two radix-4 nodes.  Field layouts are chosen so
// the XOR swizzle below keeps every quarter-warp access conflict free.
// ------------------------------------------------------------------------------------------------
template <int LOGM> struct FftPlan;
template <> struct FftPlan<8>  { static constexpr int P = 3; static constexpr int WD[4] = {3, 3, 2, 0}; };
template <> struct FftPlan<9>  { static constexpr int P = 3; static constexpr int WD[4] = {3, 3, 3, 0}; };
template <> struct FftPlan<10> { static constexpr int P = 4; static constexpr int WD[4] = {2, 2, 3, 3}; };   // the two radix-4 levels fuse into one 16-point pass (pbs_kernel_v8)
template <> struct FftPlan<11> { static constexpr int P = 4; static constexpr int WD[4] = {3, 3, 3, 2}; };
template <> struct FftPlan<12> { static constexpr int P = 4; static constexpr int WD[4] = {3, 3, 3, 3}; };

template <int LOGM, int PASS> struct PassInfo {
    static constexpr int WD = FftPlan<LOGM>::WD[PASS];
    static constexpr int done_calc() { int s = 0; for (int q = 0; q < PASS; q++) s += FftPlan<LOGM>::WD[q]; return s; }
    static constexpr int S = done_calc();                       // index bits already split off = bits of the node id h
    static constexpr int LO = LOGM - S - WD;
    static constexpr int off_calc() {                            // offset of this pass's twiddles in the flat table
        int off = 0, done = 0;
        for (int q = 0; q < PASS; q++) { off += (1 << done) * ((1 << FftPlan<LOGM>::WD[q]) - 1); done += FftPlan<LOGM>::WD[q]; }
        return off;
    }
    static constexpr int OFF = off_calc();
};

__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 3) & 7); }

// "rest" index of sub-group u (0/1) of thread t in a 2-bit pass: [warp][u][lane].  With this placement every pass
// after the first one touches only the 256 points owned by the thread's warp (index >> 8 == warp), so those passes
// need warp-level synchronisation only.
__device__ __forceinline__ int pass2_rest(int t, int u) { return ((t >> 5) << 6) | (u << 5) | (t & 31); }

// index (complex position in the M-array) of element e of thread t in a pass
template <int LOGM, int LO, int WD>
__device__ __forceinline__ int elem_index(int t, int e) {
    int rest, f;
    if (WD == 3) { rest = t; f = e; }
    else { rest = pass2_rest(t, e >> 2); f = e & 3; }
    return ((rest >> LO) << (LO + WD)) | (f << LO) | (rest & ((1 << LO) - 1));
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

// node twiddles rho^q (q = 1..R-1) of the thread's node(s) in pass PASS: w[0..6] (radix 8) or w[0..2], w[3..5] (two radix-4)
template <int LOGM, int PASS>
__device__ __forceinline__ void load_tw(double2 (&w)[7], int t, const double2* __restrict__ tw) {
    using PI = PassInfo<LOGM, PASS>;
    if (PI::WD == 3) {
        const double2* p = tw + PI::OFF + (t >> PI::LO) * 7;
#pragma unroll
        for (int q = 0; q < 7; q++) w[q] = p[q];
    } else {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const double2* p = tw + PI::OFF + (pass2_rest(t, u) >> PI::LO) * 3;
#pragma unroll
            for (int q = 0; q < 3; q++) w[u * 3 + q] = p[q];
        }
    }
}

#define TFX_SQRT_HALF 0.70710678118654757
__device__ __forceinline__ double2 mul_i(double2 a) { return make_double2(-a.y, a.x); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }
__device__ __forceinline__ double2 mul_w8(double2 a) { return make_double2(TFX_SQRT_HALF * (a.x - a.y), TFX_SQRT_HALF * (a.x + a.y)); }
__device__ __forceinline__ double2 mul_w83(double2 a) { return make_double2(-(TFX_SQRT_HALF * (a.x + a.y)), TFX_SQRT_HALF * (a.x - a.y)); }
__device__ __forceinline__ double2 mul_w8c(double2 a) { return make_double2(TFX_SQRT_HALF * (a.x + a.y), TFX_SQRT_HALF * (a.y - a.x)); }
__device__ __forceinline__ double2 mul_w83c(double2 a) { return make_double2(TFX_SQRT_HALF * (a.y - a.x), -(TFX_SQRT_HALF * (a.x + a.y))); }

// forward radix-R node on y[B .. B+R) (B = 0 or 4): premultiply by rho^q, then the constant-twiddle DIF network
template <int WD, int B>
__device__ __forceinline__ void node_forward(double2 (&y)[8], const double2 (&w)[7], int wbase) {
    constexpr int R = 1 << WD;
#pragma unroll
    for (int q = 1; q < R; q++) y[B + q] = cmul(y[B + q], w[wbase + q - 1]);
    if (WD == 3) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double2 a = y[B + q], b = y[B + q + 4], d = csub(a, b);
            y[B + q] = cadd(a, b);
            y[B + q + 4] = (q == 0) ? d : (q == 1) ? mul_w8(d) : (q == 2) ? mul_i(d) : mul_w83(d);
        }
    }
#pragma unroll
    for (int base = 0; base < R; base += 4)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double2 a = y[B + base + q], b = y[B + base + q + 2], d = csub(a, b);
            y[B + base + q] = cadd(a, b);
            y[B + base + q + 2] = (q == 0) ? d : mul_i(d);
        }
#pragma unroll
    for (int base = 0; base < R; base += 2) {
        const double2 a = y[B + base], b = y[B + base + 1];
        y[B + base] = cadd(a, b); y[B + base + 1] = csub(a, b);
    }
}

template <int WD, int B>
__device__ __forceinline__ void node_inverse(double2 (&y)[8], const double2 (&w)[7], int wbase) {
    constexpr int R = 1 << WD;
#pragma unroll
    for (int base = 0; base < R; base += 2) {
        const double2 a = y[B + base], b = y[B + base + 1];
        y[B + base] = cadd(a, b); y[B + base + 1] = csub(a, b);
    }
#pragma unroll
    for (int base = 0; base < R; base += 4)
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const double2 a = y[B + base + q];
            double2 b = y[B + base + q + 2];
            if (q == 1) b = mul_mi(b);
            y[B + base + q] = cadd(a, b); y[B + base + q + 2] = csub(a, b);
        }
    if (WD == 3) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const double2 a = y[B + q];
            double2 b = y[B + q + 4];
            b = (q == 0) ? b : (q == 1) ? mul_w8c(b) : (q == 2) ? mul_mi(b) : mul_w83c(b);
            y[B + q] = cadd(a, b); y[B + q + 4] = csub(a, b);
        }
    }
#pragma unroll
    for (int q = 1; q < R; q++) y[B + q] = cmulc(y[B + q], w[wbase + q - 1]);
}

template <int LOGM, int PASS, bool INV>
__device__ __forceinline__ void pass_nodes(double2 (&x)[8], const double2 (&w)[7]) {
#ifdef TFX_EXP_NOMATH
    return;
#endif
    constexpr int WD = PassInfo<LOGM, PASS>::WD;
    if (WD == 3) {
        if (!INV) node_forward<3, 0>(x, w, 0); else node_inverse<3, 0>(x, w, 0);
    } else {
        if (!INV) { node_forward<2, 0>(x, w, 0); node_forward<2, 4>(x, w, 3); }
        else { node_inverse<2, 0>(x, w, 0); node_inverse<2, 4>(x, w, 3); }
    }
}

// ---- one or two transforms interleaved in one thread; twiddles of the next pass are loaded before the barrier ----
#define TFX_TW(P) load_tw<LOGM, P>(w, t, tw);
#define TFX_PASS2(P, INV) pass_nodes<LOGM, P, INV>(xa, w); if (DUAL) pass_nodes<LOGM, P, INV>(xb, w);
#define TFX_STORE2(P) pass_store<LOGM, P>(xa, t, bufa); if (DUAL) pass_store<LOGM, P>(xb, t, bufb);
#define TFX_LOAD2(P) pass_load<LOGM, P>(xa, t, bufa); if (DUAL) pass_load<LOGM, P>(xb, t, bufb);

// forward: xa/xb hold pass-0 inputs (element e of thread t = index t + e*TPF), w the pass-0 twiddles;
// on return xa/xb hold the spectra in the last pass's element order.  SYNC_CTA orders the first (cross-warp)
// exchange and protects the buffers' previous contents, SYNC_WARP the later, warp-local exchanges.
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
        if constexpr (LOGM >= 12) sync(); else wsync();   // M = 4096: pass 1 spans the points of a warp pair
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

// inverse (no 1/M factor): xa/xb hold spectra (last-pass element order); on return element e = index t + e*TPF
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
    if constexpr (LOGM >= 12) sync(); else wsync();
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

// ---- NT transforms interleaved in one thread (round 2): every transform has its own buffer bufs + q*M, twiddles are
// loaded once per pass for all NT.  The arithmetic per transform is exactly that of fft_forward_regs2 / fft_inverse_regs2.
// Buffer protocol (the caller owns the CTA-level ordering): on entry to fft_forward_multi nobody reads or writes the
// buffers any more (the caller's end-of-step barrier); after the one CTA barrier inside, a warp only touches the points it
// owns (index >> 8 == warp), so everything up to and including fft_inverse_multi's last warp-local pass needs __syncwarp only.
template <int LOGM, int NT, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_forward_multi(double2 (&x)[NT][8], double2 (&w)[7], int t, double2* __restrict__ bufs,
                                                  const double2* __restrict__ tw, SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    constexpr int M = 1 << LOGM;
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 0, false>(x[q], w);
    load_tw<LOGM, 1>(w, t, tw);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_store<LOGM, 0>(x[q], t, bufs + q * M);
    sync();
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 1>(x[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 1, false>(x[q], w);
    if constexpr (PL::P >= 3) {
#pragma unroll
        for (int q = 0; q < NT; q++) pass_store<LOGM, 1>(x[q], t, bufs + q * M);
        load_tw<LOGM, 2>(w, t, tw);
        if constexpr (LOGM >= 12) sync(); else wsync();
#pragma unroll
        for (int q = 0; q < NT; q++) pass_load<LOGM, 2>(x[q], t, bufs + q * M);
#pragma unroll
        for (int q = 0; q < NT; q++) pass_nodes<LOGM, 2, false>(x[q], w);
    }
    if constexpr (PL::P >= 4) {
#pragma unroll
        for (int q = 0; q < NT; q++) pass_store<LOGM, 2>(x[q], t, bufs + q * M);
        load_tw<LOGM, 3>(w, t, tw);
        wsync();
#pragma unroll
        for (int q = 0; q < NT; q++) pass_load<LOGM, 3>(x[q], t, bufs + q * M);
#pragma unroll
        for (int q = 0; q < NT; q++) pass_nodes<LOGM, 3, false>(x[q], w);
    }
}

// forward tail for the fused plan: the caller ran passes 0 and 1 (fused16_forward), stored every transform at its natural
// index and passed a CTA barrier; this runs passes 2.. on all NT transforms interleaved
template <int LOGM, int NT, typename SYNC_WARP>
__device__ __forceinline__ void fft_forward_multi_from2(double2 (&x)[NT][8], double2 (&w)[7], int t, double2* __restrict__ bufs,
                                                        const double2* __restrict__ tw, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    constexpr int M = 1 << LOGM;
    static_assert(PL::P == 4, "fused plan has four passes");
    load_tw<LOGM, 2>(w, t, tw);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 2>(x[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 2, false>(x[q], w);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_store<LOGM, 2>(x[q], t, bufs + q * M);
    load_tw<LOGM, 3>(w, t, tw);
    wsync();
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 3>(x[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 3, false>(x[q], w);
}

// inverse counterpart: runs pass 2 on all NT spectra (stored by the caller with pass_store<LOGM, 3>) and stores them; the caller
// then passes a CTA barrier and finishes with fused16_inverse
template <int LOGM, int NT, typename SYNC_WARP>
__device__ __forceinline__ void fft_inverse_multi_pass2(int t, double2* __restrict__ bufs, const double2* __restrict__ tw, SYNC_WARP wsync) {
    constexpr int M = 1 << LOGM;
    double2 y[NT][8], w[7];
    load_tw<LOGM, 2>(w, t, tw);
    wsync();
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 2>(y[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 2, true>(y[q], w);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_store<LOGM, 2>(y[q], t, bufs + q * M);
}

// inverse, all passes BELOW the last one: the caller has run the last pass's nodes on every spectrum and stored it with
// pass_store<LOGM, LAST> into bufs + q*M (lane-private positions).  On return y[q][e] = coefficient pair t + e*TPF (no 1/M).
template <int LOGM, int NT, typename SYNC_CTA, typename SYNC_WARP>
__device__ __forceinline__ void fft_inverse_multi_rest(double2 (&y)[NT][8], int t, double2* __restrict__ bufs,
                                                       const double2* __restrict__ tw, SYNC_CTA sync, SYNC_WARP wsync) {
    using PL = FftPlan<LOGM>;
    constexpr int M = 1 << LOGM;
    double2 w[7];
    if constexpr (PL::P >= 4) {
        load_tw<LOGM, 2>(w, t, tw);
        wsync();
#pragma unroll
        for (int q = 0; q < NT; q++) pass_load<LOGM, 2>(y[q], t, bufs + q * M);
#pragma unroll
        for (int q = 0; q < NT; q++) pass_nodes<LOGM, 2, true>(y[q], w);
#pragma unroll
        for (int q = 0; q < NT; q++) pass_store<LOGM, 2>(y[q], t, bufs + q * M);
    }
    load_tw<LOGM, 1>(w, t, tw);
    if constexpr (LOGM >= 12) sync(); else wsync();
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 1>(y[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 1, true>(y[q], w);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_store<LOGM, 1>(y[q], t, bufs + q * M);
    load_tw<LOGM, 0>(w, t, tw);
    sync();
#pragma unroll
    for (int q = 0; q < NT; q++) pass_load<LOGM, 0>(y[q], t, bufs + q * M);
#pragma unroll
    for (int q = 0; q < NT; q++) pass_nodes<LOGM, 0, true>(y[q], w);
}

// ---- M = 1024, plan {2,2,3,3}: the two radix-4 levels of the plan fused into ONE 16-point pass per thread (no exchange
// between them).  A thread of a 64-thread group holds z[e] = point tl + 64 e (e = 0..15) of its transform: level 0 nodes are
// {q, q+4, q+8, q+12} (stride 256, node 0), level 1 nodes are {4h .. 4h+3} (stride 64, node h = top two index bits).  The
// arithmetic per node is node_forward<2,.> / node_inverse<2,.> exactly (same order of operations, same twiddle entries).
__device__ __forceinline__ void r4_forward(double2& y0, double2& y1, double2& y2, double2& y3, double2 w1, double2 w2, double2 w3) {
    y1 = cmul(y1, w1); y2 = cmul(y2, w2); y3 = cmul(y3, w3);
    { const double2 a = y0, b = y2, d = csub(a, b); y0 = cadd(a, b); y2 = d; }
    { const double2 a = y1, b = y3, d = csub(a, b); y1 = cadd(a, b); y3 = mul_i(d); }
    { const double2 a = y0, b = y1; y0 = cadd(a, b); y1 = csub(a, b); }
    { const double2 a = y2, b = y3; y2 = cadd(a, b); y3 = csub(a, b); }
}
__device__ __forceinline__ void r4_inverse(double2& y0, double2& y1, double2& y2, double2& y3, double2 w1, double2 w2, double2 w3) {
    { const double2 a = y0, b = y1; y0 = cadd(a, b); y1 = csub(a, b); }
    { const double2 a = y2, b = y3; y2 = cadd(a, b); y3 = csub(a, b); }
    { const double2 a = y0, b = y2; y0 = cadd(a, b); y2 = csub(a, b); }
    { const double2 a = y1, b = mul_mi(y3); y1 = cadd(a, b); y3 = csub(a, b); }
    y1 = cmulc(y1, w1); y2 = cmulc(y2, w2); y3 = cmulc(y3, w3);
}
// tw: the flat node-twiddle table of M = 1024 (pass 0 at offset 0: node 0; pass 1 at offset 3: nodes 0..3, three entries each)
__device__ __forceinline__ void fused16_forward(double2 (&z)[16], const double2* __restrict__ tw) {
    {
        const double2 w1 = tw[0], w2 = tw[1], w3 = tw[2];
#pragma unroll
        for (int q = 0; q < 4; q++) r4_forward(z[q], z[q + 4], z[q + 8], z[q + 12], w1, w2, w3);
    }
#pragma unroll
    for (int h = 0; h < 4; h++) {
        const double2 w1 = tw[3 + 3 * h], w2 = tw[4 + 3 * h], w3 = tw[5 + 3 * h];
        r4_forward(z[4 * h], z[4 * h + 1], z[4 * h + 2], z[4 * h + 3], w1, w2, w3);
    }
}
__device__ __forceinline__ void fused16_inverse(double2 (&z)[16], const double2* __restrict__ tw) {
#pragma unroll
    for (int h = 0; h < 4; h++) {
        const double2 w1 = tw[3 + 3 * h], w2 = tw[4 + 3 * h], w3 = tw[5 + 3 * h];
        r4_inverse(z[4 * h], z[4 * h + 1], z[4 * h + 2], z[4 * h + 3], w1, w2, w3);
    }
    {
        const double2 w1 = tw[0], w2 = tw[1], w3 = tw[2];
#pragma unroll
        for (int q = 0; q < 4; q++) r4_inverse(z[q], z[q + 4], z[q + 8], z[q + 12], w1, w2, w3);
    }
}

// coefficient-pair index (j, j + M) that element e of thread t holds before the first forward pass / after the last inverse pass
// (t + e * TPF when the first pass is a radix-8 pass; a leading radix-4 pass places its two nodes per thread differently)
template <int LOGM>
__device__ __forceinline__ int first_pass_index(int t, int e) {
    return elem_index<LOGM, PassInfo<LOGM, 0>::LO, PassInfo<LOGM, 0>::WD>(t, e);
}

// canonical transform position of element e of thread t after the last forward pass
template <int LOGM>
__device__ __forceinline__ int last_pass_index(int t, int e) {
    constexpr int LAST = FftPlan<LOGM>::P - 1;
    return elem_index<LOGM, PassInfo<LOGM, LAST>::LO, PassInfo<LOGM, LAST>::WD>(t, e);
}

}  // namespace tfx
