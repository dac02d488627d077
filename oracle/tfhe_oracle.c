/*
 * tfhe_oracle.c — CPU restatement of the TFHE hot path (TEST INFRASTRUCTURE ONLY).
 *
 * PARITY UNPINNED: the arithmetic the reference delegates to lives in concrete-python==2.7.0 /
 * concrete-ml==1.6.1 (env.yml:35-36 of the reference), which is neither under /root/reference nor
 * installable here, and the reference ships no tests / golden vectors for this path.
 * This file restates the public TFHE algorithms that sit behind the call sites
 * dct-cryptonets/homomorphic_eval.py:70 (forward(fhe='execute')), :315 (keygen) of the reference.
 * Where a convention is a free choice (tie rule of the decomposer, FFT dataflow, RNG) THIS FILE
 * DEFINES it and the CUDA kernels in dct-cryptonets_b200/csrc must match it bit for bit.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Nothing under dct-cryptonets_b200/ links, imports or calls it.
 *
 * Build: see oracle/Makefile (gcc -O3 -ffp-contract=off -mavx2 -mfma -fopenmp).  -ffp-contract=off
 * matters: every fused multiply-add below is an explicit fma(); everything else rounds separately,
 * exactly like the CUDA side compiled with -fmad=false and explicit __fma_rn.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct { double re, im; } cplx;

/* ------------------------------------------------------------------------------------------------
 * 1. Counter-mode PRF (ChaCha20 block function).  u64 number `idx` of stream `stream` under `seed`
 *    is word (idx & 7) of block (idx >> 3).  key = seed || (seed ^ 0xA5A5..), nonce = stream.
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define QR(a, b, c, d)                                                                             \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);                          \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

static void chacha_block(const uint32_t seed[4], uint64_t stream, uint64_t block, uint64_t out[8]) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
    for (int i = 0; i < 4; i++) { s[4 + i] = seed[i]; s[8 + i] = seed[i] ^ 0xA5A5A5A5u; }
    s[12] = (uint32_t)block; s[13] = (uint32_t)(block >> 32);
    s[14] = (uint32_t)stream; s[15] = (uint32_t)(stream >> 32);
    memcpy(x, s, sizeof x);
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 8; i++) {
        uint32_t lo = x[2 * i] + s[2 * i], hi = x[2 * i + 1] + s[2 * i + 1];
        out[i] = ((uint64_t)hi << 32) | lo;
    }
}

static uint64_t prf_u64(const uint32_t seed[4], uint64_t stream, uint64_t idx) {
    uint64_t blk[8];
    chacha_block(seed, stream, idx >> 3, blk);
    return blk[idx & 7];
}

ORC_API void orc_prf_fill(const uint8_t seed16[16], uint64_t stream, uint64_t first, uint64_t count, uint64_t* out) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
    for (uint64_t i = 0; i < count; i++) out[i] = prf_u64(seed, stream, first + i);
}

/* ------------------------------------------------------------------------------------------------
 * 2. Deterministic Gaussian: Box-Muller with hand-rolled ln / sincos built only from IEEE + - * /
 *    sqrt and explicit fma, so the CUDA side reproduces every bit.
 * ---------------------------------------------------------------------------------------------- */
static double det_ln(double x) { /* x in (0, 1], normal */
    uint64_t bits; memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m; memcpy(&m, &bits, 8);          /* m in [1,2) */
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 27.0;
    for (int k = 25; k >= 1; k -= 2) p = fma(p, s2, 1.0 / (double)k);
    return fma((double)e, 0.6931471805599453, 2.0 * (s * p));
}

static void det_sincos_turn(double u, double* c, double* s) { /* angle = 2*pi*u, u in [0,1) */
    double t = u * 4.0;
    int q = (int)t;                    /* quadrant 0..3 */
    double g = t - (double)q;          /* exact */
    double a = g * 1.5707963267948966; /* [0, pi/2) */
    double a2 = a * a;
    /* Taylor, Horner with fma; 15 terms each */
    double cs = 0.0, sn = 0.0;
    for (int k = 14; k >= 1; k--) {
        cs = fma(cs, a2, 1.0) * (-1.0 / (double)((2 * k - 1) * (2 * k)));
        sn = fma(sn, a2, 1.0) * (-1.0 / (double)((2 * k) * (2 * k + 1)));
    }
    cs = fma(cs, a2, 1.0);
    sn = fma(sn, a2, 1.0) * a;
    switch (q & 3) {
        case 0: *c = cs; *s = sn; break;
        case 1: *c = -sn; *s = cs; break;
        case 2: *c = -cs; *s = -sn; break;
        default: *c = sn; *s = -cs; break;
    }
}

/* standard normal number `idx` of a stream: consumes PRF words 2*idx, 2*idx+1 */
static double prf_gauss(const uint32_t seed[4], uint64_t stream, uint64_t idx) {
    uint64_t x = prf_u64(seed, stream, 2 * idx), y = prf_u64(seed, stream, 2 * idx + 1);
    double u1 = (double)((x >> 11) + 1) * 0x1p-53; /* (0,1] */
    double u2 = (double)(y >> 11) * 0x1p-53;       /* [0,1) */
    double r = sqrt(-2.0 * det_ln(u1));
    double c, s; det_sincos_turn(u2, &c, &s);
    (void)s;
    return r * c;
}

/* torus noise word: round(g * std * 2^64) as two's complement u64 */
static uint64_t prf_noise(const uint32_t seed[4], uint64_t stream, uint64_t idx, double std) {
    double v = prf_gauss(seed, stream, idx) * (std * 0x1p64);
    return (uint64_t)(int64_t)llrint(v);
}

ORC_API void orc_gauss_fill(const uint8_t seed16[16], uint64_t stream, uint64_t first, uint64_t count, double* out) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
    for (uint64_t i = 0; i < count; i++) out[i] = prf_gauss(seed, stream, first + i);
}

/* stream ids (purpose << 40 | index).  Shared convention with the CUDA side. */
enum { ST_BIGKEY = 1, ST_SMALLKEY = 2, ST_KSK_MASK = 3, ST_KSK_NOISE = 4, ST_BSK_MASK = 5, ST_BSK_NOISE = 6,
       ST_ENC_MASK = 7, ST_ENC_NOISE = 8 };
static inline uint64_t stream_id(int purpose, uint64_t set, uint64_t index) {
    return ((uint64_t)purpose << 56) | (set << 48) | index;
}

/* binary secret key: bit i = PRF word i & 1 */
ORC_API void orc_gen_binary_key(const uint8_t seed16[16], int purpose, uint32_t set, uint32_t dim, uint64_t* key) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
    for (uint32_t i = 0; i < dim; i++) key[i] = prf_u64(seed, stream_id(purpose, set, 0), i) & 1;
}

/* ------------------------------------------------------------------------------------------------
 * 3. LWE encrypt / decrypt.  ct = (a_0..a_{dim-1}, b), b = <a,s> + pt + e.
 *    ct number c of a call uses mask stream (ST_ENC_MASK, c0 + c) and noise stream (ST_ENC_NOISE, c0 + c).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_lwe_encrypt(const uint64_t* key, uint32_t dim, double std, const uint64_t* pts, uint64_t count,
                             const uint8_t seed16[16], uint64_t first_index, uint64_t* out) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)count; c++) {
        uint64_t* ct = out + (uint64_t)c * (dim + 1);
        uint64_t ms = stream_id(ST_ENC_MASK, 0, first_index + c);
        uint64_t dot = 0;
        for (uint32_t i = 0; i < dim; i++) {
            uint64_t a = prf_u64(seed, ms, i);
            ct[i] = a;
            dot += a * key[i];
        }
        ct[dim] = dot + pts[c] + prf_noise(seed, stream_id(ST_ENC_NOISE, 0, first_index + c), 0, std);
    }
}

ORC_API void orc_lwe_phase(const uint64_t* key, uint32_t dim, const uint64_t* cts, uint64_t count, uint64_t* phases) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)count; c++) {
        const uint64_t* ct = cts + (uint64_t)c * (dim + 1);
        uint64_t dot = 0;
        for (uint32_t i = 0; i < dim; i++) dot += ct[i] * key[i];
        phases[c] = ct[dim] - dot;
    }
}

/* ------------------------------------------------------------------------------------------------
 * 4. Signed gadget decomposition (tie rule fixed HERE: raw digit == B/2 maps to -B/2 + carry).
 *    digits[lvl-1], lvl = 1..l, weight q / B^lvl; all in [-B/2, B/2).
 * ---------------------------------------------------------------------------------------------- */
static inline void decompose(uint64_t x, int base_log, int level, int64_t* digits) {
    int total = base_log * level;
    uint64_t v = (total < 64) ? ((x + (1ULL << (63 - total))) >> (64 - total)) : x;
    uint64_t B = 1ULL << base_log, half = B >> 1, mask = B - 1;
    for (int lvl = level; lvl >= 1; lvl--) {
        uint64_t d = v & mask;
        v >>= base_log;
        if (d >= half) { digits[lvl - 1] = (int64_t)d - (int64_t)B; v += 1; }
        else digits[lvl - 1] = (int64_t)d;
    }
}

ORC_API void orc_decompose(const uint64_t* xs, uint64_t count, int base_log, int level, int64_t* out) {
    for (uint64_t i = 0; i < count; i++) decompose(xs[i], base_log, level, out + i * level);
}

/* ------------------------------------------------------------------------------------------------
 * 5. Keyswitch key + keyswitch.
 *    KSK[i][lvl-1][0..n] = LWE_small( S_big[i] * q / B^lvl ), layout [kN][l][n+1].
 *    keyswitch input word is first scaled: x -> x << shift ; body additionally += body_offset
 *    (used by the rounding chain, A.7 step 1; shift = 0 / offset = 0 for a plain keyswitch).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_gen_ksk(const uint64_t* big_key, uint32_t big_dim, const uint64_t* small_key, uint32_t n,
                         int base_log, int level, double std, const uint8_t seed16[16], uint32_t set, uint64_t* ksk) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < (int64_t)big_dim * level; row++) {
        uint32_t i = (uint32_t)(row / level); int lvl = (int)(row % level) + 1;
        uint64_t* ct = ksk + (uint64_t)row * (n + 1);
        uint64_t ms = stream_id(ST_KSK_MASK, set, row);
        uint64_t dot = 0;
        for (uint32_t j = 0; j < n; j++) {
            uint64_t a = prf_u64(seed, ms, j);
            ct[j] = a; dot += a * small_key[j];
        }
        uint64_t pt = big_key[i] << (64 - base_log * lvl);
        ct[n] = dot + pt + prf_noise(seed, stream_id(ST_KSK_NOISE, set, 0), (uint64_t)row, std);
    }
}

ORC_API void orc_keyswitch(const uint64_t* ksk, uint32_t big_dim, uint32_t n, int base_log, int level,
                           const uint64_t* in, uint64_t count, int shift, uint64_t body_offset, uint64_t* out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t c = 0; c < (int64_t)count; c++) {
        const uint64_t* ct = in + (uint64_t)c * (big_dim + 1);
        uint64_t* o = out + (uint64_t)c * (n + 1);
        int64_t dg[64];
        for (uint32_t j = 0; j < n; j++) o[j] = 0;
        o[n] = (ct[big_dim] << shift) + body_offset;
        for (uint32_t i = 0; i < big_dim; i++) {
            decompose(ct[i] << shift, base_log, level, dg);
            for (int lvl = 0; lvl < level; lvl++) {
                uint64_t d = (uint64_t)dg[lvl];
                if (!d) continue;
                const uint64_t* row = ksk + ((uint64_t)i * level + lvl) * (n + 1);
                for (uint32_t j = 0; j <= n; j++) o[j] -= d * row[j];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * 6. Negacyclic transform.  N real coefficients <-> M = N/2 complex values.
 *    z_j = p_j + i p_{j+M} is evaluated at the M roots of X^M = i (all of them roots of X^N + 1), WITHOUT a
 *    separate twist: the factor tree  X^(R h) - r  ->  prod_m (X^h - rho w_R^m),  rho = r^(1/R),  is walked with
 *    radix-8 / radix-4 nodes ("passes").  DEFINITION of one forward node on x_0..x_{R-1} (stride 2^LO):
 *        y_q = x_q * rho^q (q >= 1, complex multiply cmul below, rho^q from the table)
 *        then an R-point DIF network with constant twiddles, output in bit-reversed order:
 *          R = 8:  (q,q+4): a+b, (a-b)*w8^q   with w8^0 = 1, w8^1 = (c,c), w8^2 = i, w8^3 = (-c,c), c = sqrt(1/2)
 *                  (q,q+2): a+b, (a-b)*{1,i} ;  (q,q+1): a+b, a-b
 *          R = 4:  (q,q+2): a+b, (a-b)*{1,i} ;  (q,q+1): a+b, a-b
 *        multiplication by i is (re,im) -> (-im,re); by (c,c): (c*(re-im), c*(re+im)); by (-c,c): (-(c*(re+im)), c*(re-im)).
 *    The inverse node is the exact mirror (DIT with conjugated constants, then x_q = y_q * conj(rho^q)) and carries no
 *    1/R: the total factor 1/M is folded into the Fourier bootstrapping key (orc_bsk_to_fourier scales by 1/M).
 *    Pass plan (bits per pass, first pass = top index bits): log2(M)=8:{3,3,2} 9:{3,3,3} 10:{2,2,3,3} 11:{3,3,3,2}
 *    12:{3,3,3,3}; other sizes: 3s then the remainder.  Node (pass p, high index h with S bits already split off) uses
 *        rho = exp(i*pi*(1 + 4*bitrev_S(h)) / (R * 2^(S+1))),
 *    rho^q = exp(i*2*pi * q*(1+4*bitrev_S(h)) / (R * 2^(S+2))) taken from cosl/sinl on the first octant + exact symmetries.
 *    Position b of the result holds the evaluation at root number bitrev(b).
 *      cmul (a*b)      : re = fma(ar, br, -(ai*bi)) ; im = fma(ar, bi, ai*br)
 *      cmulc(a*conj b) : re = fma(ar, br,   ai*bi ) ; im = fma(ai, br, -(ar*bi))
 * ---------------------------------------------------------------------------------------------- */
static inline cplx cmul(cplx a, cplx b) {
    cplx r; r.re = fma(a.re, b.re, -(a.im * b.im)); r.im = fma(a.re, b.im, a.im * b.re); return r;
}
static inline cplx cmulc(cplx a, cplx b) {
    cplx r; r.re = fma(a.re, b.re, a.im * b.im); r.im = fma(a.im, b.re, -(a.re * b.im)); return r;
}
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cplx csub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
#define SQRT_HALF 0.70710678118654757

/* exp(i*2*pi*num/den), den a power of two; exact symmetries */
static cplx unit_root(uint64_t num, uint64_t den) {
    num %= den;
    uint64_t oct8 = (8 * num) / den;            /* 0..7 */
    uint64_t rnum = 8 * num - oct8 * den;       /* angle within octant = 2*pi*rnum/(8*den) */
    int flip = (int)(oct8 & 1);
    long double theta;
    const long double PI_L = 3.14159265358979323846264338327950288L;
    if (!flip) theta = 2.0L * PI_L * (long double)rnum / (long double)(8 * den);
    else       theta = 2.0L * PI_L * (long double)(den - rnum) / (long double)(8 * den);
    double c, s;
    if (rnum == 0 && !flip) { c = 1.0; s = 0.0; }
    else if (rnum == 0 && flip) { c = s = (double)sqrtl(0.5L); }    /* theta = pi/4 exactly */
    else { c = (double)cosl(theta); s = (double)sinl(theta); }
    double x, y;
    switch (oct8) {
        case 0: x = c;  y = s;  break;
        case 1: x = s;  y = c;  break;
        case 2: x = -s; y = c;  break;
        case 3: x = -c; y = s;  break;
        case 4: x = -c; y = -s; break;
        case 5: x = -s; y = -c; break;
        case 6: x = s;  y = -c; break;
        default: x = c; y = -s; break;
    }
    cplx r = { x, y }; return r;
}

static uint32_t bitrev(uint32_t v, int bits) { uint32_t r = 0; for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i); return r; }

typedef struct { uint32_t N, M, logM; int npass; int wd[8]; int lo[8]; uint32_t off[8]; cplx* tw; } fft_plan;

static fft_plan* plan_cache[32];

static void plan_passes(int logM, int* npass, int* wd) {
    static const int P8[] = {3, 3, 2}, P9[] = {3, 3, 3}, P10[] = {2, 2, 3, 3}, P11[] = {3, 3, 3, 2}, P12[] = {3, 3, 3, 3};
    const int* src = 0; int n = 0;
    switch (logM) { case 8: src = P8; n = 3; break; case 9: src = P9; n = 3; break; case 10: src = P10; n = 4; break;
                    case 11: src = P11; n = 4; break; case 12: src = P12; n = 4; break; default: break; }
    if (src) { for (int i = 0; i < n; i++) wd[i] = src[i]; *npass = n; return; }
    n = 0; int left = logM;
    while (left >= 3) { wd[n++] = 3; left -= 3; }
    if (left) wd[n++] = left;
    *npass = n;
}

static fft_plan* get_plan(uint32_t N) {
    int lg = 0; while ((1u << lg) < N) lg++;
    fft_plan* p;
#pragma omp critical(orc_plan)
    {
        p = plan_cache[lg];
        if (!p) {
            p = (fft_plan*)malloc(sizeof *p);
            p->N = N; p->M = N / 2; p->logM = lg - 1;
            plan_passes((int)p->logM, &p->npass, p->wd);
            p->tw = (cplx*)calloc(p->M, sizeof(cplx));
            uint32_t off = 0; int done = 0;
            for (int q = 0; q < p->npass; q++) {
                int wd = p->wd[q], R = 1 << wd, S = done;
                p->lo[q] = (int)p->logM - done - wd;
                p->off[q] = off;
                for (uint32_t h = 0; h < (1u << S); h++)
                    for (int e = 1; e < R; e++)
                        p->tw[off + h * (R - 1) + (e - 1)] = unit_root((uint64_t)e * (1 + 4ULL * bitrev(h, S)), (uint64_t)R << (S + 2));
                off += (1u << S) * (R - 1);
                done += wd;
            }
            plan_cache[lg] = p;
        }
    }
    return p;
}

/* table exported so the product's own table can be compared in tests: tw [M][2] (last entry zero) */
ORC_API void orc_fft_tables(uint32_t N, double* tw) {
    fft_plan* p = get_plan(N);
    memcpy(tw, p->tw, sizeof(cplx) * p->M);
}

static inline cplx mul_i(cplx a) { cplx r = { -a.im, a.re }; return r; }
static inline cplx mul_mi(cplx a) { cplx r = { a.im, -a.re }; return r; }                       /* * conj(i) */
static inline cplx mul_w8(cplx a) { cplx r = { SQRT_HALF * (a.re - a.im), SQRT_HALF * (a.re + a.im) }; return r; }
static inline cplx mul_w83(cplx a) { cplx r = { -(SQRT_HALF * (a.re + a.im)), SQRT_HALF * (a.re - a.im) }; return r; }
static inline cplx mul_w8c(cplx a) { cplx r = { SQRT_HALF * (a.re + a.im), SQRT_HALF * (a.im - a.re) }; return r; }   /* * conj(w8)   */
static inline cplx mul_w83c(cplx a) { cplx r = { SQRT_HALF * (a.im - a.re), -(SQRT_HALF * (a.re + a.im)) }; return r; } /* * conj(w8^3) */

static void node_forward(cplx* y, int wd, const cplx* rho) {
    int R = 1 << wd;
    for (int q = 1; q < R; q++) y[q] = cmul(y[q], rho[q - 1]);
    if (wd == 3) {
        for (int q = 0; q < 4; q++) {
            cplx a = y[q], b = y[q + 4], d = csub(a, b);
            y[q] = cadd(a, b);
            y[q + 4] = (q == 0) ? d : (q == 1) ? mul_w8(d) : (q == 2) ? mul_i(d) : mul_w83(d);
        }
    }
    if (wd >= 2) {
        for (int base = 0; base < R; base += 4)
            for (int q = 0; q < 2; q++) {
                cplx a = y[base + q], b = y[base + q + 2], d = csub(a, b);
                y[base + q] = cadd(a, b);
                y[base + q + 2] = (q == 0) ? d : mul_i(d);
            }
    }
    for (int base = 0; base < R; base += 2) { cplx a = y[base], b = y[base + 1]; y[base] = cadd(a, b); y[base + 1] = csub(a, b); }
}

static void node_inverse(cplx* y, int wd, const cplx* rho) {
    int R = 1 << wd;
    for (int base = 0; base < R; base += 2) { cplx a = y[base], b = y[base + 1]; y[base] = cadd(a, b); y[base + 1] = csub(a, b); }
    if (wd >= 2) {
        for (int base = 0; base < R; base += 4)
            for (int q = 0; q < 2; q++) {
                cplx a = y[base + q], b = y[base + q + 2];
                if (q == 1) b = mul_mi(b);
                y[base + q] = cadd(a, b); y[base + q + 2] = csub(a, b);
            }
    }
    if (wd == 3) {
        for (int q = 0; q < 4; q++) {
            cplx a = y[q], b = y[q + 4];
            b = (q == 0) ? b : (q == 1) ? mul_w8c(b) : (q == 2) ? mul_mi(b) : mul_w83c(b);
            y[q] = cadd(a, b); y[q + 4] = csub(a, b);
        }
    }
    for (int q = 1; q < R; q++) y[q] = cmulc(y[q], rho[q - 1]);
}

static void fft_forward(const fft_plan* p, const double* poly /*[N]*/, cplx* z /*[M]*/) {
    uint32_t M = p->M;
    for (uint32_t j = 0; j < M; j++) { z[j].re = poly[j]; z[j].im = poly[j + M]; }
    cplx y[16];
    for (int q = 0; q < p->npass; q++) {
        int wd = p->wd[q], lo = p->lo[q], R = 1 << wd;
        for (uint32_t h = 0; h < (M >> (lo + wd)); h++)
            for (uint32_t l = 0; l < (1u << lo); l++) {
                uint32_t base = (h << (lo + wd)) | l;
                for (int e = 0; e < R; e++) y[e] = z[base + ((uint32_t)e << lo)];
                node_forward(y, wd, p->tw + p->off[q] + h * (R - 1));
                for (int e = 0; e < R; e++) z[base + ((uint32_t)e << lo)] = y[e];
            }
    }
}

/* inverse WITHOUT the 1/M factor: poly_j = Re z_j, poly_{j+M} = Im z_j after the mirrored passes */
static void fft_inverse(const fft_plan* p, cplx* z /*[M], destroyed*/, double* poly /*[N]*/) {
    uint32_t M = p->M;
    cplx y[16];
    for (int q = p->npass - 1; q >= 0; q--) {
        int wd = p->wd[q], lo = p->lo[q], R = 1 << wd;
        for (uint32_t h = 0; h < (M >> (lo + wd)); h++)
            for (uint32_t l = 0; l < (1u << lo); l++) {
                uint32_t base = (h << (lo + wd)) | l;
                for (int e = 0; e < R; e++) y[e] = z[base + ((uint32_t)e << lo)];
                node_inverse(y, wd, p->tw + p->off[q] + h * (R - 1));
                for (int e = 0; e < R; e++) z[base + ((uint32_t)e << lo)] = y[e];
            }
    }
    for (uint32_t j = 0; j < M; j++) { poly[j] = z[j].re; poly[j + M] = z[j].im; }
}

/* double (integer valued up to rounding, any magnitude < 2^117) -> torus word, mod 2^64 */
static inline uint64_t double_to_torus(double v) {
    double r = nearbyint(v * 0x1p-64);
    double y = fma(-r, 0x1p64, v);
    if (y >= 0x1p63) y -= 0x1p64;
    if (y < -0x1p63) y += 0x1p64;
    return (uint64_t)(int64_t)llrint(y);
}

ORC_API void orc_fft_forward(uint32_t N, const double* poly, double* out /*[M][2]*/) {
    fft_forward(get_plan(N), poly, (cplx*)out);
}
ORC_API void orc_fft_inverse(uint32_t N, const double* in /*[M][2]*/, double* poly) {
    fft_plan* p = get_plan(N);
    cplx* z = (cplx*)malloc(sizeof(cplx) * p->M);
    memcpy(z, in, sizeof(cplx) * p->M);
    fft_inverse(p, z, poly);
    for (uint32_t j = 0; j < N; j++) poly[j] *= 1.0 / (double)p->M;      /* helper returns the true inverse */
    free(z);
}
ORC_API void orc_double_to_torus(const double* v, uint64_t count, uint64_t* out) {
    for (uint64_t i = 0; i < count; i++) out[i] = double_to_torus(v[i]);
}

/* exact negacyclic product a * s for a binary polynomial s (keygen), mod 2^64 */
static void poly_mul_binary_acc(const uint64_t* a, const uint64_t* sbits, uint32_t N, uint64_t* acc) {
    for (uint32_t j = 0; j < N; j++) {
        if (!sbits[j]) continue;
        for (uint32_t i = 0; i < N - j; i++) acc[i + j] += a[i];
        for (uint32_t i = N - j; i < N; i++) acc[i + j - N] -= a[i];
    }
}

/* exact schoolbook negacyclic product of two u64 polynomials (tests) */
ORC_API void orc_poly_mul_negacyclic(const uint64_t* a, const uint64_t* b, uint32_t N, uint64_t* out) {
    memset(out, 0, 8ULL * N);
    for (uint32_t i = 0; i < N; i++)
        for (uint32_t j = 0; j < N; j++) {
            uint64_t v = a[i] * b[j];
            if (i + j < N) out[i + j] += v; else out[i + j - N] -= v;
        }
}

/* ------------------------------------------------------------------------------------------------
 * 7. Bootstrapping key.  BSK_i = GGSW(s_i) under the GLWE key; row (r, lvl) is a GLWE
 *    encryption of 0 with s_i * q/B^lvl added to coefficient 0 of component r (r < k: mask r, r = k: body).
 *    Standard-domain layout u64 [n][k+1 (r)][l (lvl)][k+1 (component c)][N].
 *    Randomness: row index R = (i*(k+1) + r)*l + (lvl-1); mask poly c from stream (ST_BSK_MASK, set, R*k + c)
 *    words 0..N-1; noise coefficient t from stream (ST_BSK_NOISE, set, R) gaussian index t.
 *    The GLWE key is the big LWE key split into k polynomials of N bits.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_gen_bsk(const uint64_t* small_key, uint32_t n, const uint64_t* big_key, uint32_t k, uint32_t N,
                         int base_log, int level, double std, const uint8_t seed16[16], uint32_t set, uint64_t* bsk) {
    uint32_t seed[4]; memcpy(seed, seed16, 16);
    int64_t rows = (int64_t)n * (k + 1) * level;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t R = 0; R < rows; R++) {
        uint32_t i = (uint32_t)(R / ((k + 1) * level));
        uint32_t r = (uint32_t)((R / level) % (k + 1));
        int lvl = (int)(R % level) + 1;
        uint64_t* row = bsk + (uint64_t)R * (k + 1) * N;
        uint64_t* body = row + (uint64_t)k * N;
        uint64_t ns = stream_id(ST_BSK_NOISE, set, R);
        for (uint32_t t = 0; t < N; t++) body[t] = prf_noise(seed, ns, t, std);
        for (uint32_t c = 0; c < k; c++) {
            uint64_t ms = stream_id(ST_BSK_MASK, set, (uint64_t)R * k + c);
            uint64_t* a = row + (uint64_t)c * N;
            for (uint32_t t = 0; t < N; t++) a[t] = prf_u64(seed, ms, t);
            poly_mul_binary_acc(a, big_key + (uint64_t)c * N, N, body);
        }
        uint64_t g = small_key[i] << (64 - base_log * lvl);
        row[(uint64_t)r * N] += g;
    }
}

/* Fourier BSK, canonical layout double [n][k+1][l][k+1][M][2], position p = FFT output position
 * (bit-reversed evaluation order, see section 6).  Coefficients enter as signed i64 -> double. */
ORC_API void orc_bsk_to_fourier(const uint64_t* bsk, uint32_t n, uint32_t k, uint32_t N, int level, double* out) {
    fft_plan* p = get_plan(N);
    int64_t polys = (int64_t)n * (k + 1) * level * (k + 1);
#pragma omp parallel
    {
        double* tmp = (double*)malloc(sizeof(double) * N);
#pragma omp for schedule(static)
        for (int64_t q = 0; q < polys; q++) {
            const uint64_t* src = bsk + (uint64_t)q * N;
            for (uint32_t t = 0; t < N; t++) tmp[t] = (double)(int64_t)src[t];
            double* dst = out + (uint64_t)q * N;
            fft_forward(p, tmp, (cplx*)dst);
            for (uint32_t t = 0; t < N; t++) dst[t] *= 1.0 / (double)p->M;   /* the inverse transform carries no 1/M */
        }
        free(tmp);
    }
}

/* ------------------------------------------------------------------------------------------------
 * 8. Programmable bootstrap.
 *    in  : u64 [B][n+1] under the small key; luts u64 [T][N]; lut_index u32 [B]
 *    out : u64 [B][big_dim+1] under the big key, big_dim >= kN: the GLWE key is the first kN bits of the big key, the
 *          remaining mask words are zero (mode 0) / untouched (mode 1).
 *    mode 0: out = result ; mode 1: out -= (result + (0,..,0,body_const))   (rounding chain, A.7 step 4)
 *    MAC order (fixed): F_c = one fma chain starting from 0 over (r ascending, lvl ascending).
 * ---------------------------------------------------------------------------------------------- */
static inline uint32_t mod_switch(uint64_t x, uint32_t log2_2N) {
    return (uint32_t)((((x >> (64 - log2_2N - 1)) + 1) >> 1) & ((1u << log2_2N) - 1));
}

static inline uint64_t rot_coeff(const uint64_t* p, uint32_t N, uint32_t j, uint32_t shift) {
    /* coefficient j of X^shift * p, shift in [0, 2N) */
    uint32_t idx = (j - shift) & (2 * N - 1);
    return idx < N ? p[idx] : (uint64_t)0 - p[idx - N];
}

ORC_API void orc_pbs(const double* bsk_f, uint32_t n, uint32_t k, uint32_t N, uint32_t big, int base_log, int level,
                     const uint64_t* in, const uint64_t* luts, const uint32_t* lut_index, uint64_t count,
                     int mode, uint64_t body_const, uint64_t* out) {
    fft_plan* p = get_plan(N);
    uint32_t M = N / 2, log2_2N = p->logM + 2;
#pragma omp parallel
    {
        uint64_t* acc = (uint64_t*)malloc(8ULL * (k + 1) * N);
        double* dpoly = (double*)malloc(8ULL * N);
        cplx* D = (cplx*)malloc(sizeof(cplx) * M);
        cplx* F = (cplx*)malloc(sizeof(cplx) * (k + 1) * M);
        uint64_t* diff = (uint64_t*)malloc(8ULL * N);
        int64_t dg[64];
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < (int64_t)count; b++) {
            const uint64_t* ct = in + (uint64_t)b * (n + 1);
            const uint64_t* lut = luts + (uint64_t)lut_index[b] * N;
            uint32_t bhat = mod_switch(ct[n], log2_2N);
            memset(acc, 0, 8ULL * k * N);
            for (uint32_t j = 0; j < N; j++) acc[(uint64_t)k * N + j] = rot_coeff(lut, N, j, (2 * N - bhat) & (2 * N - 1));
            for (uint32_t i = 0; i < n; i++) {
                uint32_t ahat = mod_switch(ct[i], log2_2N);
                if (ahat == 0) continue;
                const double* key_i = bsk_f + (uint64_t)i * (k + 1) * level * (k + 1) * N;
                for (uint32_t q = 0; q < (k + 1) * M; q++) { F[q].re = 0.0; F[q].im = 0.0; }
                for (uint32_t r = 0; r <= k; r++) {
                    const uint64_t* ar = acc + (uint64_t)r * N;
                    for (uint32_t j = 0; j < N; j++) diff[j] = rot_coeff(ar, N, j, ahat) - ar[j];
                    for (int lvl = 0; lvl < level; lvl++) {
                        for (uint32_t j = 0; j < N; j++) { decompose(diff[j], base_log, level, dg); dpoly[j] = (double)dg[lvl]; }
                        fft_forward(p, dpoly, D);
                        for (uint32_t c = 0; c <= k; c++) {
                            const cplx* K = (const cplx*)(key_i + (((uint64_t)r * level + lvl) * (k + 1) + c) * N);
                            cplx* pc = F + (uint64_t)c * M;
                            for (uint32_t q = 0; q < M; q++) {
                                double re = pc[q].re, im = pc[q].im;
                                re = fma(D[q].re, K[q].re, re); re = fma(-D[q].im, K[q].im, re);
                                im = fma(D[q].re, K[q].im, im); im = fma(D[q].im, K[q].re, im);
                                pc[q].re = re; pc[q].im = im;
                            }
                        }
                    }
                }
                for (uint32_t c = 0; c <= k; c++) {
                    fft_inverse(p, F + (uint64_t)c * M, dpoly);
                    uint64_t* ac = acc + (uint64_t)c * N;
                    for (uint32_t j = 0; j < N; j++) ac[j] += double_to_torus(dpoly[j]);
                }
            }
            /* sample extract coefficient 0 */
            uint64_t* o = out + (uint64_t)b * (big + 1);
            for (uint32_t r = 0; r < k; r++) {
                const uint64_t* ar = acc + (uint64_t)r * N;
                for (uint32_t t = 0; t < N; t++) {
                    uint64_t v = (t == 0) ? ar[0] : (uint64_t)0 - ar[N - t];
                    if (mode == 0) o[r * N + t] = v; else o[r * N + t] -= v;
                }
            }
            if (mode == 0) for (uint32_t t = k * N; t < big; t++) o[t] = 0;
            uint64_t bv = acc[(uint64_t)k * N];
            if (mode == 0) o[big] = bv; else o[big] -= bv + body_const;
        }
        free(acc); free(dpoly); free(D); free(F); free(diff);
    }
}

/* ------------------------------------------------------------------------------------------------
 * 9. Leveled ops on big-key ciphertext tensors: integer-weight conv2d, add, scalar ops.
 *    in  u64 [Cin][H][W][dim+1] ; w int32 [Cout][Cin][kh][kw] ; out u64 [Cout][Ho][Wo][dim+1]
 *    bias_pt (optional, u64 [Cout]) is added to the body word.  depthwise: w is [Cout][1][kh][kw], ic = oc.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_conv2d(const uint64_t* in, uint32_t Cin, uint32_t H, uint32_t W, uint32_t words,
                        const int32_t* w, uint32_t Cout, uint32_t kh, uint32_t kw, uint32_t stride, uint32_t pad,
                        const uint64_t* bias_pt, uint32_t depthwise, uint64_t* out) {
    uint32_t Cin_eff = depthwise ? 1 : Cin;
    uint32_t Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (int64_t oc = 0; oc < (int64_t)Cout; oc++)
        for (int64_t oy = 0; oy < (int64_t)Ho; oy++)
            for (uint32_t ox = 0; ox < Wo; ox++) {
                uint64_t* o = out + (((uint64_t)oc * Ho + oy) * Wo + ox) * words;
                memset(o, 0, 8ULL * words);
                for (uint32_t icw = 0; icw < Cin_eff; icw++)
                    for (uint32_t ky = 0; ky < kh; ky++)
                        for (uint32_t kx = 0; kx < kw; kx++) {
                            int64_t iy = (int64_t)oy * stride + ky - pad, ix = (int64_t)ox * stride + kx - pad;
                            if (iy < 0 || iy >= (int64_t)H || ix < 0 || ix >= (int64_t)W) continue;
                            uint32_t ic = depthwise ? (uint32_t)oc : icw;
                            int32_t wv = w[((oc * Cin_eff + icw) * kh + ky) * kw + kx];
                            if (!wv) continue;
                            const uint64_t* src = in + (((uint64_t)ic * H + iy) * W + ix) * words;
                            uint64_t wu = (uint64_t)(int64_t)wv;
                            for (uint32_t t = 0; t < words; t++) o[t] += wu * src[t];
                        }
                if (bias_pt) o[words - 1] += bias_pt[oc];
            }
}

/* out = a * sa + b * sb (+ const on body); sa/sb signed integers; b may be NULL */
ORC_API void orc_axpby(const uint64_t* a, int64_t sa, const uint64_t* b, int64_t sb, uint64_t body_const,
                       uint64_t count, uint32_t words, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)count; c++) {
        for (uint32_t t = 0; t < words; t++) {
            uint64_t i = (uint64_t)c * words + t;
            uint64_t v = a[i] * (uint64_t)sa;
            if (b) v += b[i] * (uint64_t)sb;
            if (t == words - 1) v += body_const;
            out[i] = v;
        }
    }
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 when it starts more than one process per node: the timing legs set the team size
 * explicitly instead of inheriting it */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
