// tfx_keygen.cu — K6: key generation and client-side LWE encrypt / phase (decrypt) kernels.
// Replaces (upstream) Circuit.keygen() / Client.encrypt / Client.decrypt behind reference
// homomorphic_eval.py:315 and :70.  All randomness comes from the counter-mode PRF in tfx_common.cuh, so a
// seed reproduces every key and ciphertext word (and the oracle reproduces them on the CPU).
#include <string.h>
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

static Seed make_seed(const uint8_t s[16]) { Seed r; memcpy(r.w, s, 16); return r; }

__global__ void gen_binary_key_kernel(Seed seed, uint64_t stream, uint32_t dim, uint64_t* __restrict__ key) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < dim) key[i] = prf_u64(seed, stream, i) & 1;
}

// block-wide sum of u64 (mod 2^64), result valid in thread 0
__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v, uint64_t* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    uint64_t r = 0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)((blockDim.x + 31) >> 5); w++) r += s_red[w];
    return r;
}

// one CTA per LWE row: mask from PRF stream `mask_stream0 + row`, body = <a, key> + pt + noise.
// pt: mode 0 -> pts[row]; mode 1 (KSK) -> big_key[row / level] << (64 - base_log * (row % level + 1))
struct LweRowArgs {
    Seed seed; int mask_purpose, noise_purpose; uint32_t set; uint64_t first_index;
    const uint64_t* key; uint32_t dim; double std;
    const uint64_t* pts; const uint64_t* big_key; int base_log, level; int mode;
    uint64_t* out; uint32_t out_stride;
};

__global__ void __launch_bounds__(128) lwe_rows_kernel(LweRowArgs a, size_t rows) {
    __shared__ uint64_t s_red[4];
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        uint64_t* ct = a.out + row * a.out_stride;
        const uint64_t ms = stream_id(a.mask_purpose, a.set, a.first_index + row);
        uint64_t dot = 0;
        for (uint32_t blk = threadIdx.x; blk * 8 < a.dim; blk += blockDim.x) {
            uint64_t w[8];
            chacha_block(a.seed, ms, blk, w);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t i = blk * 8 + q;
                if (i < a.dim) { ct[i] = w[q]; dot += w[q] * a.key[i]; }
            }
        }
        dot = block_sum_u64(dot, s_red);
        if (threadIdx.x == 0) {
            uint64_t pt, noise;
            if (a.mode == 0) {
                pt = a.pts[row];
                noise = prf_noise(a.seed, stream_id(a.noise_purpose, a.set, a.first_index + row), 0, a.std);
            } else {
                uint32_t i = (uint32_t)(row / a.level); int lvl = (int)(row % a.level) + 1;
                pt = a.big_key[i] << (64 - a.base_log * lvl);
                noise = prf_noise(a.seed, stream_id(a.noise_purpose, a.set, 0), row, a.std);
            }
            ct[a.dim] = dot + pt + noise;
        }
    }
}

__global__ void __launch_bounds__(128) lwe_phase_kernel(const uint64_t* __restrict__ key, uint32_t dim,
                                                        const uint64_t* __restrict__ cts, size_t count,
                                                        uint64_t* __restrict__ phases) {
    __shared__ uint64_t s_red[4];
    for (size_t c = blockIdx.x; c < count; c += gridDim.x) {
        const uint64_t* ct = cts + c * (dim + 1);
        uint64_t dot = 0;
        for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) dot += ct[i] * key[i];
        dot = block_sum_u64(dot, s_red);
        if (threadIdx.x == 0) phases[c] = ct[dim] - dot;
    }
}

// Standard-domain BSK rows.  One CTA per GGSW row R = (i*(k+1) + r)*l + (lvl-1):
//   mask polys a_c from PRF, body = noise + sum_c a_c * S_c (exact negacyclic, binary S_c) ; gadget on component r.
// Output layout u64 [rows][k+1][N].
struct BskArgs {
    Seed seed; uint32_t set; const uint64_t* small_key; const uint64_t* big_key;
    uint32_t n, k, N; int base_log, level; double std; uint64_t* out;
};

__global__ void __launch_bounds__(256) gen_bsk_kernel(BskArgs a, size_t rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_a = reinterpret_cast<uint64_t*>(smem_raw);          // [N]
    uint16_t* s_pos = reinterpret_cast<uint16_t*>(s_a + a.N);       // [N] positions of set key bits
    __shared__ uint32_t s_cnt;
    const uint32_t N = a.N, k = a.k;
    for (size_t R = blockIdx.x; R < rows; R += gridDim.x) {
        const uint32_t i = (uint32_t)(R / ((k + 1) * a.level));
        const uint32_t r = (uint32_t)((R / a.level) % (k + 1));
        const int lvl = (int)(R % a.level) + 1;
        uint64_t* row = a.out + R * (size_t)(k + 1) * N;
        uint64_t* body = row + (size_t)k * N;
        const uint64_t ns = stream_id(ST_BSK_NOISE, a.set, R);
        // per-thread body accumulators live in global (body[t]) — each thread owns coefficients t = tid + q*256
        for (uint32_t t = threadIdx.x; t < N; t += blockDim.x) body[t] = prf_noise(a.seed, ns, t, a.std);
        for (uint32_t c = 0; c < k; c++) {
            __syncthreads();
            if (threadIdx.x == 0) s_cnt = 0;
            const uint64_t ms = stream_id(ST_BSK_MASK, a.set, R * k + c);
            uint64_t* ac = row + (size_t)c * N;
            for (uint32_t blk = threadIdx.x; blk * 8 < N; blk += blockDim.x) {
                uint64_t w[8];
                chacha_block(a.seed, ms, blk, w);
#pragma unroll
                for (int q = 0; q < 8; q++) { s_a[blk * 8 + q] = w[q]; ac[blk * 8 + q] = w[q]; }
            }
            __syncthreads();
            // compact list of set bits of S_c (order irrelevant: integer sums are exact)
            for (uint32_t j = threadIdx.x; j < N; j += blockDim.x)
                if (a.big_key[(size_t)c * N + j]) s_pos[atomicAdd(&s_cnt, 1u)] = (uint16_t)j;
            __syncthreads();
            const uint32_t cnt = s_cnt;
            for (uint32_t t = threadIdx.x; t < N; t += blockDim.x) {
                uint64_t sum = 0;
                for (uint32_t q = 0; q < cnt; q++) {
                    const uint32_t j = s_pos[q];
                    // coefficient t of X^j * a : a[t-j] if t >= j else -a[t-j+N]
                    sum += (t >= j) ? s_a[t - j] : (uint64_t)0 - s_a[t - j + N];
                }
                body[t] += sum;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) row[(size_t)r * N] += a.small_key[i] << (64 - a.base_log * lvl);
        __syncthreads();
    }
}

int launch_gen_binary_key(const uint8_t seed[16], int purpose, uint32_t set, uint32_t dim, uint64_t* key_d, cudaStream_t s) {
    gen_binary_key_kernel<<<(dim + 127) / 128, 128, 0, s>>>(make_seed(seed), stream_id(purpose, set, 0), dim, key_d);
    count_launch();
    return check_launch("gen_binary_key_kernel");
}

int launch_gen_ksk(const uint8_t seed[16], uint32_t set, const uint64_t* big_key_d, uint32_t big_dim, const uint64_t* small_key_d,
                   uint32_t n, int base_log, int level, double std, uint64_t* ksk_d, cudaStream_t s) {
    LweRowArgs a;
    a.seed = make_seed(seed); a.mask_purpose = ST_KSK_MASK; a.noise_purpose = ST_KSK_NOISE; a.set = set; a.first_index = 0;
    a.key = small_key_d; a.dim = n; a.std = std; a.pts = nullptr; a.big_key = big_key_d; a.base_log = base_log; a.level = level;
    a.mode = 1; a.out = ksk_d; a.out_stride = n + 1;
    size_t rows = (size_t)big_dim * level;
    lwe_rows_kernel<<<(unsigned)(rows < 8192 ? rows : 8192), 128, 0, s>>>(a, rows);
    count_launch();
    return check_launch("lwe_rows_kernel(ksk)");
}

int launch_lwe_encrypt(const uint8_t seed[16], const uint64_t* key_d, uint32_t dim, double std, const uint64_t* pts_d,
                       size_t count, uint64_t first_index, uint64_t* out_d, cudaStream_t s) {
    if (count == 0) return TFX_OK;
    LweRowArgs a;
    a.seed = make_seed(seed); a.mask_purpose = ST_ENC_MASK; a.noise_purpose = ST_ENC_NOISE; a.set = 0; a.first_index = first_index;
    a.key = key_d; a.dim = dim; a.std = std; a.pts = pts_d; a.big_key = nullptr; a.base_log = 0; a.level = 1;
    a.mode = 0; a.out = out_d; a.out_stride = dim + 1;
    lwe_rows_kernel<<<(unsigned)(count < 8192 ? count : 8192), 128, 0, s>>>(a, count);
    count_launch();
    return check_launch("lwe_rows_kernel(encrypt)");
}

int launch_lwe_phase(const uint64_t* key_d, uint32_t dim, const uint64_t* cts_d, size_t count, uint64_t* phases_d, cudaStream_t s) {
    if (count == 0) return TFX_OK;
    lwe_phase_kernel<<<(unsigned)(count < 8192 ? count : 8192), 128, 0, s>>>(key_d, dim, cts_d, count, phases_d);
    count_launch();
    return check_launch("lwe_phase_kernel");
}

int launch_gen_bsk(const uint8_t seed[16], uint32_t set, const uint64_t* small_key_d, uint32_t n, const uint64_t* big_key_d,
                   uint32_t k, uint32_t N, int base_log, int level, double std, uint64_t* bsk_std_d, cudaStream_t s) {
    BskArgs a;
    a.seed = make_seed(seed); a.set = set; a.small_key = small_key_d; a.big_key = big_key_d; a.n = n; a.k = k; a.N = N;
    a.base_log = base_log; a.level = level; a.std = std; a.out = bsk_std_d;
    size_t rows = (size_t)n * (k + 1) * level;
    size_t smem = (size_t)N * 8 + (size_t)N * 2;
    cudaError_t e = cudaFuncSetAttribute(gen_bsk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gen_bsk)");
    gen_bsk_kernel<<<(unsigned)(rows < 4096 ? rows : 4096), 256, smem, s>>>(a, rows);
    count_launch();
    return check_launch("gen_bsk_kernel");
}

}  // namespace tfx
