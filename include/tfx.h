/*
 * tfx.h — C ABI of the B200-native TFHE execution backend (libtfx_b200.so).
 *
 * This is the drop-in boundary for the hot path behind the reference's Concrete-ML call sites
 * (reference dct-cryptonets/homomorphic_eval.py:70 forward(fhe='execute'), :315 fhe_circuit.keygen()).
 * The reference itself has no native interface (it is pure Python over concrete-python 2.7.0, pinned at
 * env.yml:36); the entry points below are what a concrete-cpu style FFI for this path binds
 * (keygen / encrypt / decrypt / keyswitch / bootstrap / leveled linear ops), batched over ciphertexts.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; tfx_last_error() gives the message (thread local).
 *  - no exceptions, no ownership transfer of caller buffers; handles are created/destroyed by the library.
 *  - pointers named *_d are DEVICE pointers on the context's device; *_h are HOST pointers.
 *  - all work is enqueued on the context's CUDA stream; functions with *_h outputs synchronise that stream.
 *  - torus = uint64_t with wrap-around; an LWE ciphertext of dimension d is d+1 words (a_0..a_{d-1}, b).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with TFX_ERR_CUDA.
 */
#ifndef TFX_H
#define TFX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TFX_API __attribute__((visibility("default")))
#else
#define TFX_API
#endif

#define TFX_OK 0
#define TFX_ERR_ARG (-1)
#define TFX_ERR_CUDA (-2)
#define TFX_ERR_STATE (-3)
#define TFX_ERR_UNSUPPORTED (-4)

typedef struct tfx_ctx tfx_ctx;
typedef struct tfx_keyset tfx_keyset;

/* One PBS flavour: small LWE key of dimension n, GLWE (k, N) with k*N <= big_dim of the keyset (its GLWE key is the
 * first k*N bits of the big LWE key, so a sample-extracted ciphertext, zero-padded to big_dim, is a ciphertext under
 * the big key), bootstrapping-key gadget (2^bsk_base_log, bsk_level), keyswitch-key gadget (2^ksk_base_log,
 * ksk_level; the KSK always maps all big_dim mask words to the small key).  Noise standard deviations are fractions
 * of the torus. */
typedef struct tfx_pbs_params {
    uint32_t n, k, N;
    uint32_t bsk_base_log, bsk_level;
    uint32_t ksk_base_log, ksk_level;
    uint32_t reserved;
    double lwe_std, glwe_std;
} tfx_pbs_params;

TFX_API const char* tfx_last_error(void);
TFX_API const char* tfx_version(void);
/* 1 if the (N, k) pair has a compiled PBS kernel */
TFX_API int tfx_pbs_supported(uint32_t N, uint32_t k);

/* stream: the cudaStream_t to enqueue on (e.g. torch's current stream; NULL is the CUDA default stream).
 * private_stream != 0 ignores `stream` and creates a non-blocking stream owned by the context. */
TFX_API int tfx_ctx_create(int device_ordinal, void* stream, int private_stream, tfx_ctx** out);
TFX_API void tfx_ctx_destroy(tfx_ctx* ctx);
TFX_API int tfx_ctx_set_stream(tfx_ctx* ctx, void* stream);
TFX_API int tfx_ctx_synchronize(tfx_ctx* ctx);

/* ---- keys (replaces Circuit.keygen(), reference homomorphic_eval.py:315) --------------------------------
 * Generates on the device: big LWE key (big_dim bits), per set a small key, KSK big->small and the Fourier
 * BSK.  Deterministic in `seed`.  keep_standard_bsk != 0 also keeps the standard-domain BSK (tests). */
TFX_API int tfx_keyset_generate(tfx_ctx* ctx, uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets,
                        const uint8_t seed[16], int keep_standard_bsk, tfx_keyset** out);
/* evaluation-only keyset filled by the tfx_keyset_set_* calls (server side of FHEModelServer.run) */
TFX_API int tfx_keyset_create_empty(tfx_ctx* ctx, uint32_t big_dim, const tfx_pbs_params* sets, uint32_t nsets, tfx_keyset** out);
TFX_API void tfx_keyset_destroy(tfx_keyset* keys);
TFX_API int tfx_keyset_drop_secret(tfx_keyset* keys);
/* set < 0 : big key (big_dim words of 0/1); set >= 0 : that set's small key (n words) */
TFX_API int tfx_keyset_get_secret(tfx_keyset* keys, int set, uint64_t* key_h);
TFX_API int tfx_keyset_set_secret(tfx_keyset* keys, int set, const uint64_t* key_h);
/* KSK layout u64 [big_dim][ksk_level][n+1] */
TFX_API int tfx_keyset_get_ksk(tfx_keyset* keys, uint32_t set, uint64_t* ksk_h);
TFX_API int tfx_keyset_set_ksk(tfx_keyset* keys, uint32_t set, const uint64_t* ksk_h);
/* Fourier BSK in canonical layout double [n][k+1][bsk_level][k+1][N/2][2] (transform output order, pre-scaled by 2/N) */
TFX_API int tfx_keyset_get_bsk_fourier(tfx_keyset* keys, uint32_t set, double* bsk_h);
TFX_API int tfx_keyset_set_bsk_fourier(tfx_keyset* keys, uint32_t set, const double* bsk_h);
/* standard-domain BSK u64 [n][k+1][bsk_level][k+1][N]; only if generated with keep_standard_bsk */
TFX_API int tfx_keyset_get_bsk_standard(tfx_keyset* keys, uint32_t set, uint64_t* bsk_h);
TFX_API size_t tfx_keyset_device_bytes(tfx_keyset* keys);

/* ---- client ops (replace Client.encrypt / Client.decrypt) -------------------------------------------------
 * key_sel < 0 : big key, else small key of that set.  std: noise standard deviation (torus fraction).
 * ciphertext c uses PRF streams indexed first_index + c under enc_seed. */
TFX_API int tfx_lwe_encrypt(tfx_ctx* ctx, tfx_keyset* keys, int key_sel, double std, const uint64_t* plaintexts_d, size_t count,
                    const uint8_t enc_seed[16], uint64_t first_index, uint64_t* out_d);
/* phases[c] = b - <a, s>; decoding (rounding to the message grid) is the caller's */
TFX_API int tfx_lwe_phase(tfx_ctx* ctx, tfx_keyset* keys, int key_sel, const uint64_t* cts_d, size_t count, uint64_t* phases_d);

/* ---- server ops (replace Server.run's per-op runtime calls) ----------------------------------------------
 * keyswitch big -> small key of `set`:  in [B][big_dim+1] -> out [B][n+1].  Every input word is first
 * scaled by 2^shift and body_offset is added to the body (rounding chain; 0/0 for a plain keyswitch). */
TFX_API int tfx_keyswitch_batch(tfx_ctx* ctx, tfx_keyset* keys, uint32_t set, const uint64_t* in_d, uint64_t* out_d, size_t B,
                        uint32_t shift, uint64_t body_offset);
/* programmable bootstrap: in [B][n+1], luts [T][N], lut_index [B] -> out [B][big_dim+1] (mask words k*N..big_dim-1 are zero).
 * mode 0: out = PBS(in);  mode 1: out -= PBS(in) + (0,..,0,body_const)  (bit-extraction step fused). */
TFX_API int tfx_pbs_batch(tfx_ctx* ctx, tfx_keyset* keys, uint32_t set, const uint64_t* in_d, const uint64_t* luts_d,
                  const uint32_t* lut_index_d, uint64_t* out_d, size_t B, int mode, uint64_t body_const);
/* leveled conv: in [Cin][H][W][words], w int32 [Cout][Cin][kh][kw] ([Cout][1][kh][kw] when depthwise != 0, which
 * needs Cin == Cout; used for the sum-pool), bias_pt [Cout] or NULL (added to body),
 * out [oc_end-oc_begin][Ho][Wo][words] holding output channels oc_begin..oc_end-1. */
TFX_API int tfx_linear_conv2d(tfx_ctx* ctx, const uint64_t* in_d, uint32_t Cin, uint32_t H, uint32_t W, uint32_t words,
                      const int32_t* w_d, uint32_t Cout, uint32_t kh, uint32_t kw, uint32_t stride, uint32_t pad,
                      const uint64_t* bias_pt_d, uint32_t oc_begin, uint32_t oc_end, uint32_t depthwise, uint64_t* out_d);
/* out = a*sa + b*sb (+ body_const on the body word); b_d may be NULL; out may alias a or b */
TFX_API int tfx_linear_axpby(tfx_ctx* ctx, const uint64_t* a_d, int64_t sa, const uint64_t* b_d, int64_t sb, uint64_t body_const,
                     size_t count, uint32_t words, uint64_t* out_d);

/* ---- test / introspection hooks ----------------------------------------------------------------------------*/
/* host node-twiddle table of the negacyclic transform exactly as uploaded to the device: tw [N/2][2] (last entry zero) */
TFX_API int tfx_fft_tables(uint32_t N, double* tw_h);
/* negacyclic FFT of P real polynomials: polys [P][N] (double) -> freq [P][N/2][2] canonical order, and back */
TFX_API int tfx_fft_forward(tfx_ctx* ctx, uint32_t N, const double* polys_d, size_t P, double* freq_d);
TFX_API int tfx_fft_inverse(tfx_ctx* ctx, uint32_t N, const double* freq_d, size_t P, uint64_t* torus_d);
/* machine probes for the roofline denominators: which 0 -> FP64 FMA rate in FLOP/s, 1 -> u64 += u32*u64 rate in MAC/s */
TFX_API int tfx_probe_rate(tfx_ctx* ctx, int which, double* rate_out);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
TFX_API uint64_t tfx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TFX_H */
