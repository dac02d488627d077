// Which B200 data paths run beside the FP64 / shared-memory issue path, and which compete with it?
// Round-2 probes behind DESIGN.md §6 ("what bounds the PBS kernel").  Every experiment runs one CTA per SM with ROLE
// warps: a warp executes exactly one instruction stream (DFMA, LDS.128, STS.128, LDTM, STTM, DMMA, IMAD, LDG) so that
// overlap between two streams is overlap between *different* warps, not a scheduling artefact of one in-order warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o mio_paths mio_paths.cu && ./mio_paths
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

enum Role { R_NONE = 0, R_DFMA, R_LDS, R_STS, R_LDTM, R_STTM, R_DMMA, R_IMAD, R_LDG, R_TMA };

struct Cfg { int role[16]; int warps; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldtm16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void sttm16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}

// per iteration and warp: DFMA 64 instructions; LDS/STS 8 x 128-bit (32 wavefronts); LDTM/STTM 4 x (32 lanes x 16 words) = 8 KB;
// DMMA 16 x m8n8k4; IMAD 64; LDG 8 x 128-bit from an L2-resident buffer; TMA one 8 KB bulk copy per iteration (lane 0)
__global__ void __launch_bounds__(512, 1) roles_kernel(Cfg cfg, double* out, int iters, const double2* gbuf, double b, double c) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* buf = reinterpret_cast<double2*>(smem);                  // 2048 double2 = 32 KB
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t mbar_all[16];                    // one barrier and one 8 KB landing buffer per warp
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int i = t; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (t < 16) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar_all[t])));
    if (t == 0) asm volatile("fence.mbarrier_init.release.cluster;");
    uint64_t& mbar = mbar_all[warp];
    unsigned char* tma_dst = smem + 32768 + warp * 8192;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    const int role = warp < cfg.warps ? cfg.role[warp] : R_NONE;
    double acc = 0;
    if (role == R_DFMA) {
        double x[8];
        for (int j = 0; j < 8; j++) x[j] = t + j;
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = fma(x[j], b, c);
        for (int j = 0; j < 8; j++) acc += x[j];
    } else if (role == R_LDS) {
        unsigned long long a = 0; int idx = t;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double2 v = buf[(idx + j * 256) & 2047];
                a ^= (unsigned long long)__double_as_longlong(v.x) + (unsigned long long)__double_as_longlong(v.y);
            }
            idx = (idx + 32) & 2047;
        }
        acc = (double)a;
    } else if (role == R_STS) {
        int idx = t; double2 v = make_double2(t, 1);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) buf[(idx + j * 256) & 2047] = v;
            idx = (idx + 32) & 2047; v.x += 1.0;
        }
        acc = v.x;
    } else if (role == R_LDTM) {
        uint32_t v[16]; uint32_t a = 0;
        for (int i = 0; i < iters; i++) {
            uint32_t v1[16], v2[16], v3[16];
            ldtm16(tbase + ((i + 0) & 7) * 16, v);
            ldtm16(tbase + ((i + 1) & 7) * 16, v1);
            ldtm16(tbase + ((i + 2) & 7) * 16, v2);
            ldtm16(tbase + ((i + 3) & 7) * 16, v3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; q++) a ^= v[q] ^ v1[q] ^ v2[q] ^ v3[q];
        }
        acc = a;
    } else if (role == R_STTM) {
        uint32_t v[16];
        for (int q = 0; q < 16; q++) v[q] = t * 16 + q;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 4; j++) { sttm16(tbase + ((i + j) & 7) * 16, v); v[j] += 1; }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        acc = v[0];
    } else if (role == R_DMMA) {
        double d0[4] = {1, 2, 3, 4}, d1[4] = {5, 6, 7, 8};
        double a0 = t * 1e-3, b0 = 1.0 + lane * 1e-9;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(d0[j]), "+d"(d0[j + 1]) : "d"(a0), "d"(b0));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(d1[j]), "+d"(d1[j + 1]) : "d"(a0), "d"(b0));
                }
            }
        }
        for (int j = 0; j < 4; j++) acc += d0[j] + d1[j];
    } else if (role == R_IMAD) {
        uint32_t x[8]; uint32_t m = 0x9E3779B9u + t;
        for (int j = 0; j < 8; j++) x[j] = t + j;
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = x[j] * m + (uint32_t)i;
        uint32_t a = 0; for (int j = 0; j < 8; j++) a ^= x[j];
        acc = a;
    } else if (role == R_LDG) {
        unsigned long long a = 0; size_t idx = (size_t)blockIdx.x * 4096 + t;
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double2 v = __ldg(gbuf + ((idx + j * 512) & ((1u << 20) - 1)));
                a ^= (unsigned long long)__double_as_longlong(v.x) + (unsigned long long)__double_as_longlong(v.y);
            }
            idx += 4096;
        }
        acc = (double)a;
    } else if (role == R_TMA) {
        if (lane == 0) {
            uint32_t phase = 0;
            for (int i = 0; i < iters; i++) {
                const double2* src = gbuf + (((size_t)(blockIdx.x * 16 + warp) * 8192 + (size_t)i * 512) & ((1u << 20) - 1));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8192;" :: "r"(smem_u32(&mbar)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 8192, [%2];"
                             :: "r"(smem_u32(tma_dst)), "l"(src), "r"(smem_u32(&mbar)) : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                phase ^= 1;
            }
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + t] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base_s));
}

static const char* NAME[] = {"-", "DFMA", "LDS", "STS", "LDTM", "STTM", "DMMA", "IMAD", "LDG", "TMA"};
static double* g_out; static double2* g_buf;

static float run(const Cfg& c, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = c.warps * 32; const size_t smem = 32768 + 16 * 8192;
    roles_kernel<<<148, threads, smem>>>(c, g_out, iters / 8, g_buf, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    roles_kernel<<<148, threads, smem>>>(c, g_out, iters, g_buf, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

// nA warps of role A and nB warps of role B, interleaved over the 4 sub-partitions (warp w lives on SMSP w % 4)
static Cfg mix(int a, int nA, int b, int nB) {
    Cfg c{}; c.warps = 16;
    int ia = 0, ib = 0;
    for (int w = 0; w < 16; w++) c.role[w] = R_NONE;
    // fill sub-partition by sub-partition so both roles are spread over all four
    for (int w = 0; w < 16 && ia < nA; w++) { c.role[w] = a; ia++; }
    for (int w = 15; w >= 0 && ib < nB; w--) if (c.role[w] == R_NONE) { c.role[w] = b; ib++; }
    return c;
}

static void pair(int a, int nA, int b, int nB, int iters) {
    float ta = run(mix(a, nA, R_NONE, 0), iters), tb = run(mix(R_NONE, 0, b, nB), iters), tab = run(mix(a, nA, b, nB), iters);
    printf("%-4s x%-2d %8.3f ms | %-4s x%-2d %8.3f ms | both %8.3f ms | sum %8.3f max %8.3f | overlap %.2f\n",
           NAME[a], nA, ta, NAME[b], nB, tb, tab, ta + tb, ta > tb ? ta : tb,
           (ta + tb - tab) / (ta < tb ? ta : tb));   // 1.0 = perfectly parallel paths, 0.0 = one serialised path
}

// ---- TMEM layout discovery: store with 32x32b (thread i = lane i), load with 16x256b, print the mapping ----
__global__ void tmem_layout_kernel(uint32_t* out) {
    __shared__ uint32_t base_s;
    const int t = threadIdx.x;
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(smem_u32(&base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = base_s;
    uint32_t v[16];
    for (int q = 0; q < 16; q++) v[q] = (t << 8) | q;                 // value = (lane, column)
    sttm16(tb, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncwarp();
    uint32_t r[8];
    // 16x256b.x1: 4 registers per thread; .x2: 8 registers (16 columns)
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 8; q++) out[t * 16 + q] = r[q];
    uint32_t s[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]) : "r"(tb + (16u << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 8; q++) out[t * 16 + 8 + q] = s[q];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(tb));
}

int main() {
    cudaMalloc(&g_out, 148 * 512 * 8);
    cudaMalloc(&g_buf, (size_t)(1 << 20) * 16);
    cudaMemset(g_buf, 0, (size_t)(1 << 20) * 16);
    cudaFuncSetAttribute(roles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 16 * 8192);
    const int it = 20000;
    printf("one CTA of 16 warps per SM, 148 SMs, %d iterations; role warps are spread over the four sub-partitions\n", it);
    printf("per warp-iteration: DFMA 64 instr | LDS/STS 8x128b = 32 wavefronts | LDTM/STTM 4 x 2 KB | DMMA 16 | IMAD 64 | LDG 8x128b | TMA 8 KB\n");
    puts("--- single streams, scaling with the number of warps");
    for (int r : {R_DFMA, R_LDS, R_STS, R_LDTM, R_STTM, R_DMMA, R_IMAD, R_LDG, R_TMA})
        for (int n : {4, 8}) {
            float ms = run(mix(r, n, R_NONE, 0), it);
            double cyc = ms * 1e-3 * 1.965e9 / ((double)it * n);           // SM cycles per warp-iteration
            printf("%-4s x%-2d %8.3f ms  %8.2f SM-cycles per warp-iteration\n", NAME[r], n, ms, cyc);
        }
    puts("--- pairs on different warps (overlap 1.0 = independent paths, 0.0 = one shared path)");
    pair(R_DFMA, 8, R_LDS, 8, it);
    pair(R_DFMA, 4, R_LDS, 4, it);
    pair(R_DFMA, 8, R_STS, 8, it);
    pair(R_DFMA, 8, R_LDG, 8, it);
    pair(R_DFMA, 8, R_IMAD, 8, it);
    pair(R_LDS, 8, R_IMAD, 8, it);
    pair(R_DFMA, 8, R_LDTM, 4, it);
    pair(R_LDS, 8, R_LDTM, 4, it);
    pair(R_DFMA, 8, R_STTM, 4, it);
    pair(R_LDS, 8, R_STTM, 4, it);
    pair(R_DMMA, 8, R_LDS, 8, it);
    pair(R_DMMA, 8, R_DFMA, 8, it);
    pair(R_DFMA, 8, R_TMA, 4, it);
    pair(R_LDS, 8, R_TMA, 4, it);
    pair(R_LDS, 8, R_LDG, 8, it);
    pair(R_LDTM, 4, R_STTM, 4, it);
    {   // three streams at once: DFMA x6 + LDS x6 + LDTM x4
        Cfg c{}; c.warps = 16;
        for (int w = 0; w < 16; w++) c.role[w] = w < 6 ? R_DFMA : w < 12 ? R_LDS : R_LDTM;
        Cfg d = c, l = c, m = c;
        for (int w = 0; w < 16; w++) { if (d.role[w] != R_DFMA) d.role[w] = R_NONE; if (l.role[w] != R_LDS) l.role[w] = R_NONE; if (m.role[w] != R_LDTM) m.role[w] = R_NONE; }
        printf("DFMA x6 %.3f ms | LDS x6 %.3f ms | LDTM x4 %.3f ms | all three %.3f ms\n", run(d, it), run(l, it), run(m, it), run(c, it));
    }

    uint32_t* lay; cudaMalloc(&lay, 32 * 16 * 4);
    tmem_layout_kernel<<<1, 32>>>(lay);
    uint32_t h[512];
    if (cudaMemcpy(h, lay, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
        puts("--- TMEM layout: st 32x32b.x16 (thread = lane, 16 columns), ld 16x256b.x2 at lanes 0-15 then 16-31: thread: (lane,col)...");
        for (int t = 0; t < 32; t++) {
            printf("t%02d:", t);
            for (int q = 0; q < 16; q++) printf(" (%2u,%2u)", h[t * 16 + q] >> 8, h[t * 16 + q] & 255);
            puts("");
        }
    } else puts("TMEM layout kernel failed");
    return 0;
}
