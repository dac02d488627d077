// tfx_internal.h — host-side declarations shared by the translation units of libtfx_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/tfx.h"

namespace tfx {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* what);
void count_launch();

struct PbsLaunch {
    const double* bsk; const double* tw;
    const uint64_t* in; const uint64_t* luts; const uint32_t* lut_index; uint64_t* out;
    uint32_t n, k, N, big_dim; int base_log, level, mode; uint64_t body_const; size_t count; int sm_count;
    uint32_t* work_counter;
};
int pbs_supported(uint32_t N, uint32_t k);
int launch_pbs(const PbsLaunch& p, cudaStream_t stream);
int launch_fft(int which, uint32_t N, const void* in, void* out, const double* tw, size_t polys,
               int sm_count, cudaStream_t stream);

struct KsLaunch {
    const uint64_t* ksk;       // device, padded layout + column sums (see tfx_keyswitch.cu)
    const uint8_t* ksk_bytes;  // device, byte-split layout for the tensor-core path (or null)
    uint8_t* digits;           // device scratch [count][big_dim*level] for the tensor-core path (or null)
    const uint64_t* in; uint64_t* out;
    uint32_t big_dim, n; int base_log, level; uint32_t shift; uint64_t body_offset; size_t count; int sm_count;
};
int launch_keyswitch(const KsLaunch& p, cudaStream_t stream);
uint32_t ksk_npad(uint32_t n);
// tcgen05 / TMA variant of the limb-split contraction (tfx_keyswitch_umma.cu); digits and key bytes as for the IMMA kernel
bool keyswitch_umma_ok(uint32_t big_dim, uint32_t n, int level);
int launch_keyswitch_umma(const KsLaunch& p, cudaStream_t stream);

// keygen / client kernels (tfx_keygen.cu)
int launch_gen_binary_key(const uint8_t seed[16], int purpose, uint32_t set, uint32_t dim, uint64_t* key_d, cudaStream_t s);
int launch_gen_ksk(const uint8_t seed[16], uint32_t set, const uint64_t* big_key_d, uint32_t big_dim, const uint64_t* small_key_d,
                   uint32_t n, int base_log, int level, double std, uint64_t* ksk_d, cudaStream_t s);
int launch_gen_bsk(const uint8_t seed[16], uint32_t set, const uint64_t* small_key_d, uint32_t n, const uint64_t* big_key_d,
                   uint32_t k, uint32_t N, int base_log, int level, double std, uint64_t* bsk_std_d, cudaStream_t s);
int launch_lwe_encrypt(const uint8_t seed[16], const uint64_t* key_d, uint32_t dim, double std, const uint64_t* pts_d,
                       size_t count, uint64_t first_index, uint64_t* out_d, cudaStream_t s);
int launch_lwe_phase(const uint64_t* key_d, uint32_t dim, const uint64_t* cts_d, size_t count, uint64_t* phases_d, cudaStream_t s);

// leveled ops (tfx_leveled.cu)
int launch_conv2d(const uint64_t* in, uint32_t Cin, uint32_t H, uint32_t W, uint32_t words, const int32_t* w, uint32_t Cout,
                  uint32_t kh, uint32_t kw, uint32_t stride, uint32_t pad, const uint64_t* bias_pt, uint32_t oc_begin,
                  uint32_t oc_end, uint32_t depthwise, uint64_t* out, int sm_count, cudaStream_t s);
int launch_axpby(const uint64_t* a, int64_t sa, const uint64_t* b, int64_t sb, uint64_t body_const, size_t count,
                 uint32_t words, uint64_t* out, int sm_count, cudaStream_t s);

int probe_rate(int which, int sm_count, cudaStream_t stream, double* rate);

}  // namespace tfx
