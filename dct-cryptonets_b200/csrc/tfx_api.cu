// tfx_api.cu — C ABI of libtfx_b200.so (declared in include/tfx.h): contexts, keysets, argument checking.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <new>
#include <vector>
#include "tfx_internal.h"

namespace tfx {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return code; }
int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof g_err, "%s: %s", where, cudaGetErrorString(e));
    return TFX_ERR_CUDA;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? TFX_OK : set_cuda_error(e, what);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

uint32_t ksk_npad(uint32_t n);
int launch_ksk_repack(const uint64_t* src, uint64_t* dst, uint32_t rows, uint32_t n, int to_padded, cudaStream_t s);
int launch_ksk_corr(const uint64_t* ksk_padded, uint64_t* corr, uint32_t rows, uint32_t n, int base_log, cudaStream_t s);
int launch_ksk_bytes(const uint64_t* ksk_padded, uint8_t* kb, uint32_t rows, uint32_t n, cudaStream_t s);
bool keyswitch_imma_ok(uint32_t big_dim, int base_log, int level);

// ---- host node-twiddle table of the negacyclic transform (definition: oracle/tfhe_oracle.c section 6): roots from
//      cosl/sinl on the first octant, the other octants by exact symmetry
static void unit_root(uint64_t num, uint64_t den, double* re, double* im) {
    const long double PI_L = 3.14159265358979323846264338327950288L;
    num %= den;
    const uint64_t oct = (8 * num) / den;
    const uint64_t rem = 8 * num - oct * den;
    const bool mirrored = (oct & 1) != 0;
    double c, s;
    if (rem == 0) {
        if (mirrored) { c = (double)sqrtl(0.5L); s = c; } else { c = 1.0; s = 0.0; }
    } else {
        const long double theta = 2.0L * PI_L * (long double)(mirrored ? den - rem : rem) / (long double)(8 * den);
        c = (double)cosl(theta); s = (double)sinl(theta);
    }
    static const int sx[8] = {+1, +1, -1, -1, -1, -1, +1, +1};
    static const int sy[8] = {+1, +1, +1, +1, -1, -1, -1, -1};
    static const bool swap_cs[8] = {false, true, true, false, false, true, true, false};
    const double x = swap_cs[oct] ? s : c, y = swap_cs[oct] ? c : s;
    *re = sx[oct] > 0 ? x : -x;
    *im = sy[oct] > 0 ? y : -y;
}

static uint32_t bit_reverse(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
}

// bits per pass for log2(M); must match FftPlan<> in tfx_common.cuh
static int pass_plan(int logM, int wd[8]) {
    switch (logM) {
        case 8: wd[0] = 3; wd[1] = 3; wd[2] = 2; return 3;
        case 9: wd[0] = 3; wd[1] = 3; wd[2] = 3; return 3;
        case 10: wd[0] = 2; wd[1] = 2; wd[2] = 3; wd[3] = 3; return 4;
        case 11: wd[0] = 3; wd[1] = 3; wd[2] = 3; wd[3] = 2; return 4;
        case 12: wd[0] = 3; wd[1] = 3; wd[2] = 3; wd[3] = 3; return 4;
        default: { int n = 0, left = logM; while (left >= 3) { wd[n++] = 3; left -= 3; } if (left) wd[n++] = left; return n; }
    }
}

// flat table: pass p, node h (S_p bits), power q = 1..R-1 at offset_p + h*(R-1) + q-1:
//   rho^q = exp(i*2*pi * q*(1 + 4*bitrev(h)) / (R * 2^(S+2)))
static void make_tables(uint32_t N, std::vector<double>& tw) {
    const uint32_t M = N / 2;
    int logM = 0; while ((1u << logM) < M) logM++;
    tw.assign((size_t)M * 2, 0.0);
    int wd[8];
    const int np = pass_plan(logM, wd);
    uint32_t off = 0; int done = 0;
    for (int p = 0; p < np; p++) {
        const int R = 1 << wd[p], S = done;
        for (uint32_t h = 0; h < (1u << S); h++)
            for (int q = 1; q < R; q++) {
                const size_t idx = off + (size_t)h * (R - 1) + (q - 1);
                unit_root((uint64_t)q * (1 + 4ULL * bit_reverse(h, S)), (uint64_t)R << (S + 2), &tw[2 * idx], &tw[2 * idx + 1]);
            }
        off += (1u << S) * (R - 1);
        done += wd[p];
    }
}

struct FftTables { uint32_t N = 0; double* tw_d = nullptr; };

}  // namespace tfx

using namespace tfx;

struct tfx_ctx {
    int device; cudaStream_t stream; bool own_stream; int sm_count;
    std::vector<FftTables> tables;
    uint8_t* scratch = nullptr; size_t scratch_bytes = 0;      // keyswitch digit matrix (tensor-core path), grown on demand
    uint32_t* work_counter = nullptr;                          // dynamic ciphertext hand-out of the PBS kernel
};

struct KeySet1 {
    tfx_pbs_params p;
    uint64_t* small_key_d = nullptr;   // [n]
    uint64_t* ksk_d = nullptr;         // padded [big*l][npad] + corr[npad]
    uint8_t* ksk_bytes_d = nullptr;    // byte-split [npad*8][big*l] (tensor-core keyswitch), null if not applicable
    double* bsk_d = nullptr;           // thread-major Fourier
    uint64_t* bsk_std_d = nullptr;     // optional
    bool has_ksk = false, has_bsk = false, has_secret = false;
};

struct tfx_keyset {
    tfx_ctx* ctx; uint32_t big_dim; uint64_t* big_key_d = nullptr; bool has_big = false;
    std::vector<KeySet1> sets; size_t bytes = 0;
};

static int use_device(tfx_ctx* ctx) {
    cudaError_t e = cudaSetDevice(ctx->device);
    return e == cudaSuccess ? TFX_OK : set_cuda_error(e, "cudaSetDevice");
}

static int get_tables(tfx_ctx* ctx, uint32_t N, FftTables** out) {
    for (auto& t : ctx->tables) if (t.N == N) { *out = &t; return TFX_OK; }
    std::vector<double> tw;
    make_tables(N, tw);
    FftTables t; t.N = N;
    const size_t bytes = (size_t)N / 2 * 16;
    cudaError_t e = cudaMalloc(&t.tw_d, bytes); if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(tw)");
    e = cudaMemcpyAsync(t.tw_d, tw.data(), bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);       // host vectors die at scope exit
    if (e != cudaSuccess) { cudaFree(t.tw_d); return set_cuda_error(e, "upload fft tables"); }
    ctx->tables.push_back(t);
    *out = &ctx->tables.back();
    return TFX_OK;
}

static size_t bsk_doubles(const tfx_pbs_params& p) { return (size_t)p.n * (p.k + 1) * p.bsk_level * (p.k + 1) * p.N; }
static size_t bsk_polys(const tfx_pbs_params& p) { return (size_t)p.n * (p.k + 1) * p.bsk_level * (p.k + 1); }
static size_t ksk_words_padded(uint32_t big, const tfx_pbs_params& p) { return ((size_t)big * p.ksk_level + 1) * ksk_npad(p.n); }

static int check_params(uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets) {
    if (!sets || nsets == 0 || nsets > 16) return set_error(TFX_ERR_ARG, "keyset: need 1..16 parameter sets");
    for (uint32_t s = 0; s < nsets; s++) {
        const tfx_pbs_params& p = sets[s];
        if ((uint64_t)p.k * p.N > big_dim) return set_error(TFX_ERR_ARG, "keyset: k*N must not exceed big_dim");
        if (!pbs_supported(p.N, p.k)) return set_error(TFX_ERR_UNSUPPORTED, "keyset: no PBS kernel for this (N, k)");
        if (p.n < 1 || p.n > 4096) return set_error(TFX_ERR_ARG, "keyset: n out of range");
        if (p.bsk_level < 1 || p.bsk_base_log < 1 || p.bsk_base_log * p.bsk_level > 64) return set_error(TFX_ERR_ARG, "keyset: bad BSK gadget");
        if (p.bsk_base_log > 52) return set_error(TFX_ERR_ARG, "keyset: BSK digits must be exact in fp64");
        if (p.ksk_level < 1 || p.ksk_level > 16 || p.ksk_base_log < 1 || p.ksk_base_log > 31 || p.ksk_base_log * p.ksk_level > 64)
            return set_error(TFX_ERR_ARG, "keyset: bad KSK gadget");
    }
    return TFX_OK;
}

static int alloc_keyset(tfx_ctx* ctx, uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets, tfx_keyset** out) {
    tfx_keyset* ks = new (std::nothrow) tfx_keyset();
    if (!ks) return set_error(TFX_ERR_STATE, "out of host memory");
    ks->ctx = ctx; ks->big_dim = big_dim;
    cudaError_t e = cudaMalloc(&ks->big_key_d, (size_t)big_dim * 8);
    if (e != cudaSuccess) { delete ks; return set_cuda_error(e, "cudaMalloc(big key)"); }
    ks->bytes += (size_t)big_dim * 8;
    ks->sets.resize(nsets);
    for (uint32_t s = 0; s < nsets; s++) {
        KeySet1& k1 = ks->sets[s]; k1.p = sets[s];
        e = cudaMalloc(&k1.small_key_d, (size_t)k1.p.n * 8);
        if (e == cudaSuccess) e = cudaMalloc(&k1.ksk_d, ksk_words_padded(big_dim, k1.p) * 8);
        if (e == cudaSuccess) e = cudaMalloc(&k1.bsk_d, bsk_doubles(k1.p) * 8);
        if (e != cudaSuccess) { tfx_keyset_destroy(ks); return set_cuda_error(e, "cudaMalloc(keys)"); }
        ks->bytes += (size_t)k1.p.n * 8 + ksk_words_padded(big_dim, k1.p) * 8 + bsk_doubles(k1.p) * 8;
    }
    *out = ks;
    return TFX_OK;
}

// byte-split copy of the KSK for the tensor-core keyswitch (call after the padded layout + column sums are in place)
static int build_ksk_bytes(tfx_keyset* ks, KeySet1& k1) {
    const tfx_pbs_params& p = k1.p;
    if (!keyswitch_imma_ok(ks->big_dim, (int)p.ksk_base_log, (int)p.ksk_level)) return TFX_OK;
    const size_t rows = (size_t)ks->big_dim * p.ksk_level, bytes = rows * ksk_npad(p.n) * 8;
    if (!k1.ksk_bytes_d) {
        cudaError_t e = cudaMalloc(&k1.ksk_bytes_d, bytes);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(ksk bytes)");
        ks->bytes += bytes;
    }
    return launch_ksk_bytes(k1.ksk_d, k1.ksk_bytes_d, (uint32_t)rows, p.n, ks->ctx->stream);
}

extern "C" {

const char* tfx_last_error(void) { return g_err; }
const char* tfx_version(void) { return "tfx_b200 0.1 (sm_100a)"; }
int tfx_pbs_supported(uint32_t N, uint32_t k) { return pbs_supported(N, k); }
uint64_t tfx_launch_count(void) { return g_launches.load(); }

int tfx_ctx_create(int device_ordinal, void* stream, int private_stream, tfx_ctx** out) {
    if (!out) return set_error(TFX_ERR_ARG, "ctx_create: null out");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return set_error(TFX_ERR_CUDA, "no CUDA device: tfx_b200 has no CPU fallback");
    if (device_ordinal < 0 || device_ordinal >= ndev) return set_error(TFX_ERR_ARG, "ctx_create: bad device ordinal");
    e = cudaSetDevice(device_ordinal); if (e != cudaSuccess) return set_cuda_error(e, "cudaSetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device_ordinal); if (e != cudaSuccess) return set_cuda_error(e, "cudaGetDeviceProperties");
    if (prop.major < 10) return set_error(TFX_ERR_UNSUPPORTED, "tfx_b200 is built for sm_100a (Blackwell) only");
    tfx_ctx* c = new (std::nothrow) tfx_ctx();
    if (!c) return set_error(TFX_ERR_STATE, "out of host memory");
    c->device = device_ordinal; c->sm_count = prop.multiProcessorCount;
    c->tables.reserve(8);
    if (!private_stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return set_cuda_error(e, "cudaStreamCreate"); }
        c->own_stream = true;
    }
    *out = c;
    return TFX_OK;
}

void tfx_ctx_destroy(tfx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& t : ctx->tables) cudaFree(t.tw_d);
    cudaFree(ctx->scratch);
    cudaFree(ctx->work_counter);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int tfx_ctx_set_stream(tfx_ctx* ctx, void* stream) {
    if (!ctx) return set_error(TFX_ERR_ARG, "null ctx");
    if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)stream;
    return TFX_OK;
}

int tfx_ctx_synchronize(tfx_ctx* ctx) {
    if (!ctx) return set_error(TFX_ERR_ARG, "null ctx");
    int rc = use_device(ctx); if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    return e == cudaSuccess ? TFX_OK : set_cuda_error(e, "cudaStreamSynchronize");
}

void tfx_keyset_destroy(tfx_keyset* ks) {
    if (!ks) return;
    cudaSetDevice(ks->ctx->device);
    cudaStreamSynchronize(ks->ctx->stream);
    cudaFree(ks->big_key_d);
    for (auto& k1 : ks->sets) { cudaFree(k1.small_key_d); cudaFree(k1.ksk_d); cudaFree(k1.ksk_bytes_d); cudaFree(k1.bsk_d); cudaFree(k1.bsk_std_d); }
    delete ks;
}

size_t tfx_keyset_device_bytes(tfx_keyset* ks) { return ks ? ks->bytes : 0; }

int tfx_keyset_create_empty(tfx_ctx* ctx, uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets, tfx_keyset** out) {
    if (!ctx || !out) return set_error(TFX_ERR_ARG, "keyset_create_empty: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    rc = check_params(big_dim, sets, nsets); if (rc) return rc;
    return alloc_keyset(ctx, big_dim, sets, nsets, out);
}

int tfx_keyset_generate(tfx_ctx* ctx, uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets, const uint8_t seed[16],
                        int keep_standard_bsk, tfx_keyset** out) {
    if (!ctx || !out || !seed) return set_error(TFX_ERR_ARG, "keyset_generate: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    rc = check_params(big_dim, sets, nsets); if (rc) return rc;
    tfx_keyset* ks = nullptr;
    rc = alloc_keyset(ctx, big_dim, sets, nsets, &ks); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    rc = launch_gen_binary_key(seed, 1 /*ST_BIGKEY*/, 0, big_dim, ks->big_key_d, st);
    ks->has_big = true;
    for (uint32_t s = 0; s < nsets && rc == TFX_OK; s++) {
        KeySet1& k1 = ks->sets[s]; const tfx_pbs_params& p = k1.p;
        rc = launch_gen_binary_key(seed, 2 /*ST_SMALLKEY*/, s, p.n, k1.small_key_d, st); if (rc) break;
        k1.has_secret = true;
        // KSK: generate canonical rows into a temp, repack to the padded layout, column sums
        uint64_t* tmp = nullptr;
        const size_t rows = (size_t)big_dim * p.ksk_level;
        cudaError_t e = cudaMalloc(&tmp, rows * (p.n + 1) * 8);
        if (e != cudaSuccess) { rc = set_cuda_error(e, "cudaMalloc(ksk tmp)"); break; }
        rc = launch_gen_ksk(seed, s, ks->big_key_d, big_dim, k1.small_key_d, p.n, p.ksk_base_log, p.ksk_level, p.lwe_std, tmp, st);
        if (!rc) rc = launch_ksk_repack(tmp, k1.ksk_d, (uint32_t)rows, p.n, 1, st);
        if (!rc) rc = launch_ksk_corr(k1.ksk_d, k1.ksk_d + rows * ksk_npad(p.n), (uint32_t)rows, p.n, p.ksk_base_log, st);
        if (!rc) rc = build_ksk_bytes(ks, k1);
        cudaStreamSynchronize(st); cudaFree(tmp);
        if (rc) break;
        k1.has_ksk = true;
        // BSK: standard domain then forward FFT into the thread-major Fourier layout
        uint64_t* std_d = nullptr;
        e = cudaMalloc(&std_d, bsk_doubles(p) * 8);
        if (e != cudaSuccess) { rc = set_cuda_error(e, "cudaMalloc(bsk standard)"); break; }
        rc = launch_gen_bsk(seed, s, k1.small_key_d, p.n, ks->big_key_d, p.k, p.N, p.bsk_base_log, p.bsk_level, p.glwe_std, std_d, st);
        FftTables* tb = nullptr;
        if (!rc) rc = get_tables(ctx, p.N, &tb);
        if (!rc) rc = launch_fft(0, p.N, std_d, k1.bsk_d, tb->tw_d, bsk_polys(p), ctx->sm_count, st);
        cudaError_t es = cudaStreamSynchronize(st);
        if (!rc && es != cudaSuccess) rc = set_cuda_error(es, "keygen");
        if (keep_standard_bsk && !rc) { k1.bsk_std_d = std_d; ks->bytes += bsk_doubles(p) * 8; } else cudaFree(std_d);
        if (rc) break;
        k1.has_bsk = true;
    }
    if (rc) { tfx_keyset_destroy(ks); return rc; }
    *out = ks;
    return TFX_OK;
}

int tfx_keyset_drop_secret(tfx_keyset* ks) {
    if (!ks) return set_error(TFX_ERR_ARG, "null keyset");
    int rc = use_device(ks->ctx); if (rc) return rc;
    cudaMemsetAsync(ks->big_key_d, 0, (size_t)ks->big_dim * 8, ks->ctx->stream); ks->has_big = false;
    for (auto& k1 : ks->sets) { cudaMemsetAsync(k1.small_key_d, 0, (size_t)k1.p.n * 8, ks->ctx->stream); k1.has_secret = false; }
    return TFX_OK;
}

static int copy_sync(tfx_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    return e == cudaSuccess ? TFX_OK : set_cuda_error(e, "cudaMemcpy");
}

int tfx_keyset_get_secret(tfx_keyset* ks, int set, uint64_t* key_h) {
    if (!ks || !key_h) return set_error(TFX_ERR_ARG, "get_secret: null argument");
    int rc = use_device(ks->ctx); if (rc) return rc;
    if (set < 0) {
        if (!ks->has_big) return set_error(TFX_ERR_STATE, "keyset holds no secret key");
        return copy_sync(ks->ctx, key_h, ks->big_key_d, (size_t)ks->big_dim * 8, cudaMemcpyDeviceToHost);
    }
    if ((size_t)set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "get_secret: bad set");
    if (!ks->sets[set].has_secret) return set_error(TFX_ERR_STATE, "keyset holds no secret key");
    return copy_sync(ks->ctx, key_h, ks->sets[set].small_key_d, (size_t)ks->sets[set].p.n * 8, cudaMemcpyDeviceToHost);
}

int tfx_keyset_set_secret(tfx_keyset* ks, int set, const uint64_t* key_h) {
    if (!ks || !key_h) return set_error(TFX_ERR_ARG, "set_secret: null argument");
    int rc = use_device(ks->ctx); if (rc) return rc;
    if (set < 0) {
        rc = copy_sync(ks->ctx, ks->big_key_d, key_h, (size_t)ks->big_dim * 8, cudaMemcpyHostToDevice);
        if (!rc) ks->has_big = true;
        return rc;
    }
    if ((size_t)set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "set_secret: bad set");
    rc = copy_sync(ks->ctx, ks->sets[set].small_key_d, key_h, (size_t)ks->sets[set].p.n * 8, cudaMemcpyHostToDevice);
    if (!rc) ks->sets[set].has_secret = true;
    return rc;
}

int tfx_keyset_get_ksk(tfx_keyset* ks, uint32_t set, uint64_t* ksk_h) {
    if (!ks || !ksk_h || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "get_ksk: bad argument");
    KeySet1& k1 = ks->sets[set];
    if (!k1.has_ksk) return set_error(TFX_ERR_STATE, "keyset holds no KSK for this set");
    int rc = use_device(ks->ctx); if (rc) return rc;
    const size_t rows = (size_t)ks->big_dim * k1.p.ksk_level, words = rows * (k1.p.n + 1);
    uint64_t* tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, words * 8); if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc");
    rc = launch_ksk_repack(k1.ksk_d, tmp, (uint32_t)rows, k1.p.n, 0, ks->ctx->stream);
    if (!rc) rc = copy_sync(ks->ctx, ksk_h, tmp, words * 8, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    return rc;
}

int tfx_keyset_set_ksk(tfx_keyset* ks, uint32_t set, const uint64_t* ksk_h) {
    if (!ks || !ksk_h || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "set_ksk: bad argument");
    KeySet1& k1 = ks->sets[set];
    int rc = use_device(ks->ctx); if (rc) return rc;
    const size_t rows = (size_t)ks->big_dim * k1.p.ksk_level, words = rows * (k1.p.n + 1);
    uint64_t* tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, words * 8); if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc");
    rc = copy_sync(ks->ctx, tmp, ksk_h, words * 8, cudaMemcpyHostToDevice);
    if (!rc) rc = launch_ksk_repack(tmp, k1.ksk_d, (uint32_t)rows, k1.p.n, 1, ks->ctx->stream);
    if (!rc) rc = launch_ksk_corr(k1.ksk_d, k1.ksk_d + rows * ksk_npad(k1.p.n), (uint32_t)rows, k1.p.n, k1.p.ksk_base_log, ks->ctx->stream);
    if (!rc) rc = build_ksk_bytes(ks, k1);
    cudaStreamSynchronize(ks->ctx->stream);
    cudaFree(tmp);
    if (!rc) k1.has_ksk = true;
    return rc;
}

static int bsk_xfer(tfx_keyset* ks, uint32_t set, double* host, bool to_host) {
    KeySet1& k1 = ks->sets[set];
    int rc = use_device(ks->ctx); if (rc) return rc;
    const size_t bytes = bsk_doubles(k1.p) * 8;
    double* tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, bytes); if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc");
    if (to_host) {
        rc = launch_fft(3, k1.p.N, k1.bsk_d, tmp, nullptr, bsk_polys(k1.p), ks->ctx->sm_count, ks->ctx->stream);
        if (!rc) rc = copy_sync(ks->ctx, host, tmp, bytes, cudaMemcpyDeviceToHost);
    } else {
        rc = copy_sync(ks->ctx, tmp, host, bytes, cudaMemcpyHostToDevice);
        if (!rc) rc = launch_fft(4, k1.p.N, tmp, k1.bsk_d, nullptr, bsk_polys(k1.p), ks->ctx->sm_count, ks->ctx->stream);
        cudaStreamSynchronize(ks->ctx->stream);
    }
    cudaFree(tmp);
    return rc;
}

int tfx_keyset_get_bsk_fourier(tfx_keyset* ks, uint32_t set, double* bsk_h) {
    if (!ks || !bsk_h || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "get_bsk_fourier: bad argument");
    if (!ks->sets[set].has_bsk) return set_error(TFX_ERR_STATE, "keyset holds no BSK for this set");
    return bsk_xfer(ks, set, bsk_h, true);
}

int tfx_keyset_set_bsk_fourier(tfx_keyset* ks, uint32_t set, const double* bsk_h) {
    if (!ks || !bsk_h || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "set_bsk_fourier: bad argument");
    int rc = bsk_xfer(ks, set, const_cast<double*>(bsk_h), false);
    if (!rc) ks->sets[set].has_bsk = true;
    return rc;
}

int tfx_keyset_get_bsk_standard(tfx_keyset* ks, uint32_t set, uint64_t* bsk_h) {
    if (!ks || !bsk_h || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "get_bsk_standard: bad argument");
    KeySet1& k1 = ks->sets[set];
    if (!k1.bsk_std_d) return set_error(TFX_ERR_STATE, "standard-domain BSK was not kept (keep_standard_bsk = 0)");
    int rc = use_device(ks->ctx); if (rc) return rc;
    return copy_sync(ks->ctx, bsk_h, k1.bsk_std_d, bsk_doubles(k1.p) * 8, cudaMemcpyDeviceToHost);
}

static int pick_key(tfx_keyset* ks, int key_sel, const uint64_t** key, uint32_t* dim) {
    if (key_sel < 0) {
        if (!ks->has_big) return set_error(TFX_ERR_STATE, "keyset holds no secret key");
        *key = ks->big_key_d; *dim = ks->big_dim; return TFX_OK;
    }
    if ((size_t)key_sel >= ks->sets.size()) return set_error(TFX_ERR_ARG, "bad key selector");
    if (!ks->sets[key_sel].has_secret) return set_error(TFX_ERR_STATE, "keyset holds no secret key");
    *key = ks->sets[key_sel].small_key_d; *dim = ks->sets[key_sel].p.n; return TFX_OK;
}

int tfx_lwe_encrypt(tfx_ctx* ctx, tfx_keyset* ks, int key_sel, double std, const uint64_t* pts_d, size_t count,
                    const uint8_t enc_seed[16], uint64_t first_index, uint64_t* out_d) {
    if (!ctx || !ks || !enc_seed || (count && (!pts_d || !out_d))) return set_error(TFX_ERR_ARG, "lwe_encrypt: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    const uint64_t* key; uint32_t dim;
    rc = pick_key(ks, key_sel, &key, &dim); if (rc) return rc;
    return launch_lwe_encrypt(enc_seed, key, dim, std, pts_d, count, first_index, out_d, ctx->stream);
}

int tfx_lwe_phase(tfx_ctx* ctx, tfx_keyset* ks, int key_sel, const uint64_t* cts_d, size_t count, uint64_t* phases_d) {
    if (!ctx || !ks || (count && (!cts_d || !phases_d))) return set_error(TFX_ERR_ARG, "lwe_phase: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    const uint64_t* key; uint32_t dim;
    rc = pick_key(ks, key_sel, &key, &dim); if (rc) return rc;
    return launch_lwe_phase(key, dim, cts_d, count, phases_d, ctx->stream);
}

int tfx_keyswitch_batch(tfx_ctx* ctx, tfx_keyset* ks, uint32_t set, const uint64_t* in_d, uint64_t* out_d, size_t B,
                        uint32_t shift, uint64_t body_offset) {
    if (!ctx || !ks || set >= ks->sets.size() || (B && (!in_d || !out_d))) return set_error(TFX_ERR_ARG, "keyswitch_batch: bad argument");
    if (shift > 63) return set_error(TFX_ERR_ARG, "keyswitch_batch: shift > 63");
    KeySet1& k1 = ks->sets[set];
    if (!k1.has_ksk) return set_error(TFX_ERR_STATE, "keyset holds no KSK for this set");
    int rc = use_device(ctx); if (rc) return rc;
    KsLaunch p;
    p.ksk_bytes = k1.ksk_bytes_d; p.digits = nullptr;
    if (k1.ksk_bytes_d && !getenv("TFX_KS_IMAD")) {          // TFX_KS_IMAD=1: measurement knob, forces the integer-pipe kernel
        const size_t need = B * (size_t)ks->big_dim * k1.p.ksk_level;
        if (need > ctx->scratch_bytes) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0;
            cudaError_t e = cudaMalloc(&ctx->scratch, need);
            if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(keyswitch digits)");
            ctx->scratch_bytes = need;
        }
        p.digits = ctx->scratch;
    }
    p.ksk = k1.ksk_d; p.in = in_d; p.out = out_d; p.big_dim = ks->big_dim; p.n = k1.p.n; p.base_log = (int)k1.p.ksk_base_log;
    p.level = (int)k1.p.ksk_level; p.shift = shift; p.body_offset = body_offset; p.count = B; p.sm_count = ctx->sm_count;
    return launch_keyswitch(p, ctx->stream);
}

int tfx_pbs_batch(tfx_ctx* ctx, tfx_keyset* ks, uint32_t set, const uint64_t* in_d, const uint64_t* luts_d,
                  const uint32_t* lut_index_d, uint64_t* out_d, size_t B, int mode, uint64_t body_const) {
    if (!ctx || !ks || set >= ks->sets.size()) return set_error(TFX_ERR_ARG, "pbs_batch: bad argument");
    if (B == 0) return TFX_OK;
    if (!in_d || !luts_d || !lut_index_d || !out_d) return set_error(TFX_ERR_ARG, "pbs_batch: null buffer");
    if (mode != 0 && mode != 1) return set_error(TFX_ERR_ARG, "pbs_batch: mode must be 0 or 1");
    if (B > 0xffffffffull) return set_error(TFX_ERR_ARG, "pbs_batch: batch too large");
    KeySet1& k1 = ks->sets[set];
    if (!k1.has_bsk) return set_error(TFX_ERR_STATE, "keyset holds no BSK for this set");
    int rc = use_device(ctx); if (rc) return rc;
    FftTables* tb = nullptr;
    rc = get_tables(ctx, k1.p.N, &tb); if (rc) return rc;
    if (!ctx->work_counter) {
        cudaError_t e = cudaMalloc(&ctx->work_counter, 64);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaMalloc(pbs work counter)");
    }
    PbsLaunch p;
    p.work_counter = ctx->work_counter;
    p.bsk = k1.bsk_d; p.tw = tb->tw_d; p.in = in_d; p.luts = luts_d; p.lut_index = lut_index_d; p.out = out_d;
    p.n = k1.p.n; p.k = k1.p.k; p.N = k1.p.N; p.big_dim = ks->big_dim; p.base_log = (int)k1.p.bsk_base_log; p.level = (int)k1.p.bsk_level;
    p.mode = mode; p.body_const = body_const; p.count = B; p.sm_count = ctx->sm_count;
    return launch_pbs(p, ctx->stream);
}

int tfx_linear_conv2d(tfx_ctx* ctx, const uint64_t* in_d, uint32_t Cin, uint32_t H, uint32_t W, uint32_t words,
                      const int32_t* w_d, uint32_t Cout, uint32_t kh, uint32_t kw, uint32_t stride, uint32_t pad,
                      const uint64_t* bias_pt_d, uint32_t oc_begin, uint32_t oc_end, uint32_t depthwise, uint64_t* out_d) {
    if (!ctx || !in_d || !w_d || !out_d) return set_error(TFX_ERR_ARG, "conv2d: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    return launch_conv2d(in_d, Cin, H, W, words, w_d, Cout, kh, kw, stride, pad, bias_pt_d, oc_begin, oc_end, depthwise, out_d,
                         ctx->sm_count, ctx->stream);
}

int tfx_linear_axpby(tfx_ctx* ctx, const uint64_t* a_d, int64_t sa, const uint64_t* b_d, int64_t sb, uint64_t body_const,
                     size_t count, uint32_t words, uint64_t* out_d) {
    if (!ctx || !a_d || !out_d || words == 0) return set_error(TFX_ERR_ARG, "axpby: bad argument");
    int rc = use_device(ctx); if (rc) return rc;
    return launch_axpby(a_d, sa, b_d, sb, body_const, count, words, out_d, ctx->sm_count, ctx->stream);
}

int tfx_probe_rate(tfx_ctx* ctx, int which, double* rate_out) {
    if (!ctx || !rate_out || which < 0 || which > 1) return set_error(TFX_ERR_ARG, "probe_rate: bad argument");
    int rc = use_device(ctx); if (rc) return rc;
    return probe_rate(which, ctx->sm_count, ctx->stream, rate_out);
}

int tfx_fft_tables(uint32_t N, double* tw_h) {
    if (!tw_h || N < 16 || (N & (N - 1))) return set_error(TFX_ERR_ARG, "fft_tables: bad argument");
    std::vector<double> tw;
    make_tables(N, tw);
    memcpy(tw_h, tw.data(), tw.size() * 8);
    return TFX_OK;
}

int tfx_fft_forward(tfx_ctx* ctx, uint32_t N, const double* polys_d, size_t P, double* freq_d) {
    if (!ctx || !polys_d || !freq_d) return set_error(TFX_ERR_ARG, "fft_forward: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    FftTables* tb = nullptr;
    rc = get_tables(ctx, N, &tb); if (rc) return rc;
    return launch_fft(1, N, polys_d, freq_d, tb->tw_d, P, ctx->sm_count, ctx->stream);
}

int tfx_fft_inverse(tfx_ctx* ctx, uint32_t N, const double* freq_d, size_t P, uint64_t* torus_d) {
    if (!ctx || !freq_d || !torus_d) return set_error(TFX_ERR_ARG, "fft_inverse: null argument");
    int rc = use_device(ctx); if (rc) return rc;
    FftTables* tb = nullptr;
    rc = get_tables(ctx, N, &tb); if (rc) return rc;
    return launch_fft(2, N, freq_d, torus_d, tb->tw_d, P, ctx->sm_count, ctx->stream);
}

}  // extern "C"
