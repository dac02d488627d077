// tfx_keyswitch_umma.cu — K2 on the Blackwell tensor path: the limb-split keyswitch contraction
//     acc[ct][word*8 + byte] = sum_K digit[ct][K] * keybyte[word*8 + byte][K]        (u8 x u8 -> s32)
// as a tcgen05.mma kind::i8 GEMM with TMEM accumulators, operands streamed by TMA (cp.async.bulk.tensor, 128-byte swizzle)
// through a 4-stage mbarrier pipeline.  One CTA computes a 128-ciphertext x 32-word (256 byte-column) tile over the whole
// contraction; warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2-5 = epilogue (TMEM -> registers,
// byte columns recombined mod 2^64, out = init - acc + corr).  Same digit matrix / key-byte layouts and the same exact
// integer result as keyswitch_imma_kernel (tfx_keyswitch.cu), which stays as the fallback.
// Replaces (upstream) concrete-cpu's keyswitch behind reference homomorphic_eval.py:70.
#include <cuda.h>
#include <stdlib.h>
#include "tfx_common.cuh"
#include "tfx_internal.h"

namespace tfx {

constexpr int KU_M = 128, KU_N = 256, KU_K = 128, KU_UMMA_K = 32;
constexpr int KU_THREADS = 192;                                    // 6 warps
constexpr uint32_t KU_A_BYTES = KU_M * KU_K, KU_B_BYTES = KU_N * KU_K, KU_STAGE_BYTES = KU_A_BYTES + KU_B_BYTES;
// Pipeline depth: 4 stages (193 KB, one CTA per SM) for full-size batches; 1 stage (50 KB) for the small batches of a rank's
// wave-sized chains, which run beside PBS kernels of other streams: a 193 KB CTA only fits once ALL PBS CTAs of an SM have retired
// (measured: the keyswitch of a 592-row chunk waited 2-8 ms for an SM, profiles/r02_experiments.md 5), a 50 KB CTA takes the
// place of the first PBS CTA that exits.
constexpr size_t ku_smem(int stages) { return (size_t)stages * KU_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */; }

struct KuArgs {
    const uint64_t* corr; const uint64_t* in; uint64_t* out;
    uint32_t big_dim, n, k_stages; uint32_t shift; uint64_t body_offset; uint32_t count;
};

__device__ __forceinline__ uint32_t ku_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ku_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor:
// start >> 4 in bits [0,14), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64))
__device__ __forceinline__ uint64_t ku_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

template <int KU_STAGES>
__global__ void __launch_bounds__(KU_THREADS, 1)
keyswitch_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, KuArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t base = (ku_smem(smem_raw) + 1023u) & ~1023u;             // swizzle-128B tiles need 1024-byte alignment
    unsigned char* gen_base = smem_raw + (base - ku_smem(smem_raw));
    const uint32_t bars = base + KU_STAGES * KU_STAGE_BYTES;                 // full[4], empty[4], accum, tmem slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + (size_t)KU_STAGES * KU_STAGE_BYTES + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row0 = blockIdx.y * KU_M, col0 = blockIdx.x * KU_N;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (KU_STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (2 * KU_STAGES);

    if (threadIdx.x == 0) {
        for (int s = 0; s < KU_STAGES; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(full(s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(empty(s)));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(accum_bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
    }
    if (warp == 2) {                                                         // 256 accumulator columns (s32, 128 lanes)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(ku_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                                     // ===== TMA producer =====
            for (uint32_t ks = 0; ks < a.k_stages; ks++) {
                const int s = ks % KU_STAGES; const uint32_t it = ks / KU_STAGES;
                if (it > 0) ku_mbar_wait(empty(s), (it - 1) & 1);            // the MMAs that read this slot have retired
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(full(s)), "r"(KU_STAGE_BYTES) : "memory");
                const uint32_t dst_a = base + s * KU_STAGE_BYTES, dst_b = dst_a + KU_A_BYTES;
                const int k0 = (int)(ks * KU_K);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             :: "r"(dst_a), "l"(&tmap_a), "r"(full(s)), "r"(k0), "r"((int)row0) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             :: "r"(dst_b), "l"(&tmap_b), "r"(full(s)), "r"(k0), "r"((int)col0) : "memory");
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                                     // ===== MMA issuer =====
            // instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 (2 << 4), A = B = unsigned 8 bit (0), both K-major,
            // N >> 3 in bits [17,23), M >> 4 in bits [24,29)
            const uint32_t idesc = (2u << 4) | ((uint32_t)(KU_N >> 3) << 17) | ((uint32_t)(KU_M >> 4) << 24);
            for (uint32_t ks = 0; ks < a.k_stages; ks++) {
                const int s = ks % KU_STAGES; const uint32_t it = ks / KU_STAGES;
                ku_mbar_wait(full(s), it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = base + s * KU_STAGE_BYTES, sb = sa + KU_A_BYTES;
#pragma unroll
                for (int k = 0; k < KU_K / KU_UMMA_K; k++) {
                    const uint64_t da = ku_desc(sa + k * KU_UMMA_K), db = ku_desc(sb + k * KU_UMMA_K);
                    const uint32_t accumulate = (ks > 0 || k > 0) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                                 :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                }
                // arrives on empty[s] once the MMAs issued so far have finished reading shared memory
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(empty(s)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(accum_bar) : "memory");
        }
    } else {                                                                 // ===== epilogue: warps 2..5, TMEM lane quarter warp % 4 =====
        ku_mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t q = (uint32_t)(warp & 3);
        const uint32_t ct = row0 + q * 32 + lane;
        const uint32_t word0 = col0 >> 3;
        uint64_t body = 0;
        if (ct < a.count) body = (a.in[(size_t)ct * (a.big_dim + 1) + a.big_dim] << a.shift) + a.body_offset;
#pragma unroll 1
        for (int c = 0; c < KU_N / 32; c++) {                                // 32 columns = 4 output words per load
            uint32_t v[32];
            const uint32_t taddr = tmem + ((q * 32u) << 16) + (uint32_t)(c * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint64_t acc = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) acc += (uint64_t)v[j * 8 + b] << (8 * b);      // byte column b weighs 2^(8b), mod 2^64
                const uint32_t word = word0 + c * 4 + j;
                if (ct < a.count && word <= a.n)
                    a.out[(size_t)ct * (a.n + 1) + word] = (word == a.n ? body : 0) - acc + a.corr[word];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(tmem) : "memory");
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// u8 matrix [rows][kp] (kp contiguous) -> tiles of box_rows x 128 bytes, 128-byte swizzle, zero fill outside
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t kp, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {kp, rows};
    const cuuint64_t strides[1] = {kp};
    const cuuint32_t box[2] = {(cuuint32_t)KU_K, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool keyswitch_umma_ok(uint32_t big_dim, uint32_t n, int level) {
    const uint64_t kp = (uint64_t)big_dim * level;
    return !getenv("TFX_KS_IMMA") && kp % KU_K == 0 && (ksk_npad(n) * 8) % KU_N == 0 && encode_fn() != nullptr;
}

// digits: u8 [count][kp] (ks_decompose_kernel), key bytes: u8 [npad*8][kp]; returns TFX_ERR_UNSUPPORTED if the maps cannot be built
int launch_keyswitch_umma(const KsLaunch& p, cudaStream_t stream) {
    const uint64_t kp = (uint64_t)p.big_dim * p.level;
    const uint32_t npad = ksk_npad(p.n);
    CUtensorMap ma, mb;
    if (!make_map(&ma, p.digits, p.count, kp, KU_M) || !make_map(&mb, p.ksk_bytes, (uint64_t)npad * 8, kp, KU_N))
        return set_error(TFX_ERR_UNSUPPORTED, "keyswitch: cuTensorMapEncodeTiled failed");
    const bool small = p.count <= 2048 && !getenv("TFX_KS_DEEP");          // TFX_KS_DEEP=1: measurement knob, always 4 stages
    const size_t smem = ku_smem(small ? 1 : 4);
    cudaError_t e = small ? cudaFuncSetAttribute(keyswitch_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(keyswitch_umma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(keyswitch_umma)");
    KuArgs a;
    a.corr = p.ksk + (size_t)p.big_dim * p.level * npad; a.in = p.in; a.out = p.out; a.big_dim = p.big_dim; a.n = p.n;
    a.k_stages = (uint32_t)(kp / KU_K); a.shift = p.shift; a.body_offset = p.body_offset; a.count = (uint32_t)p.count;
    dim3 grid(npad * 8 / KU_N, (unsigned)((p.count + KU_M - 1) / KU_M));
    if (small) keyswitch_umma_kernel<1><<<grid, KU_THREADS, smem, stream>>>(ma, mb, a);
    else keyswitch_umma_kernel<4><<<grid, KU_THREADS, smem, stream>>>(ma, mb, a);
    count_launch();
    return check_launch("keyswitch_umma_kernel");
}

}  // namespace tfx
