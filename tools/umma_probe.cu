// umma_probe.cu — hardware probe of the tcgen05 conventions the tensor-core UDE kernels rely on
// (sm_100a).  Not part of the product: a diagnostic that pins, on a real B200,
//   * the shared-memory matrix-descriptor fields (LBO / SBO) for un-swizzled K-major and MN-major
//     bf16 operands, for the SAME physical buffer viewed both ways,
//   * the instruction descriptor (kind::f16, bf16 inputs, fp32 accumulate, M = 128 / 64),
//   * the TMEM accumulator mapping (row -> lane, column) for M = 128 and M = 64 (incl. lane offset 16),
//   * tcgen05.ld / tcgen05.st 32x32b round trips,
//   * issue-to-completion cycles of back-to-back MMAs of the shapes the kernels use,
//   * the accuracy of the 3-way bf16 split (6 products) against fp64.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    return d;                    // layout type 0 = no swizzle
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                // D format f32
    d |= 1u << 7;                // A format bf16
    d |= 1u << 10;               // B format bf16
    d |= (uint32_t)a_mn_major << 15;
    d |= (uint32_t)b_mn_major << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
        :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
           "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct ProbeCfg {
    int M, N;                 // instruction shape
    int n_mma;                // instructions issued back to back (k-steps or product terms)
    int a_mn, b_mn;           // 1 = MN-major
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_step, b_step;  // byte advance of the start address per instruction
    uint32_t a_bytes, b_bytes;
    uint32_t d_lane_off;      // lane offset of the accumulator (0 or 16)
    int repeat;               // timing: repeat the whole sequence this many times (accumulating)
    int zero_first;           // first instruction overwrites D
    int d_alt;                // column offset of a second accumulator used by odd instructions (0 = same D)
};

// One CTA, 128 threads.  A/B images are copied verbatim into shared memory.
__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeCfg c, const uint8_t* a_img, const uint8_t* b_img, float* d_out /*[128][256]*/,
                                                       long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((c.a_bytes + 1023) / 1024) * 1024;
    for (uint32_t i = threadIdx.x; i < 200 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < c.a_bytes / 4; i += 128) ((uint32_t*)sa)[i] = ((const uint32_t*)a_img)[i];
    for (uint32_t i = threadIdx.x; i < c.b_bytes / 4; i += 128) ((uint32_t*)sb)[i] = ((const uint32_t*)b_img)[i];
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s;
    // clear the accumulator region (all 128 lanes x 256 columns) so untouched cells read 0
    {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0;
        const uint32_t lane_base = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
        for (int col = 0; col < 512; col += 16) tmem_st16(tbase + lane_base + col, z);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(c.M, c.N, c.a_mn, c.b_mn);
        const uint32_t d_addr = tbase + (c.d_lane_off << 16);
        // descriptors are loop invariant: build them before the timed region (the real kernels do the same)
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ad[i] = make_desc(smem_u32(sa) + i * c.a_step, c.a_lbo, c.a_sbo);
            bd[i] = make_desc(smem_u32(sb) + i * c.b_step, c.b_lbo, c.b_sbo);
        }
        t0 = clock64();
        if (c.n_mma == 4) {
            mma_bf16(d_addr, ad[0], bd[0], idesc, c.zero_first ? 0u : 1u);
            mma_bf16(d_addr, ad[1], bd[1], idesc, 1u);
            mma_bf16(d_addr, ad[2], bd[2], idesc, 1u);
            mma_bf16(d_addr, ad[3], bd[3], idesc, 1u);
#pragma unroll 1
            for (int rep = 1; rep < c.repeat; ++rep) {
                mma_bf16(d_addr, ad[0], bd[0], idesc, 1u);
                mma_bf16(d_addr + c.d_alt, ad[1], bd[1], idesc, 1u);
                mma_bf16(d_addr + 2 * c.d_alt, ad[2], bd[2], idesc, 1u);
                mma_bf16(d_addr + 3 * c.d_alt, ad[3], bd[3], idesc, 1u);
            }
        } else {
            for (int i = 0; i < c.n_mma; ++i) {
                const uint64_t a2 = make_desc(smem_u32(sa) + i * c.a_step, c.a_lbo, c.a_sbo);
                const uint64_t b2 = make_desc(smem_u32(sb) + i * c.b_step, c.b_lbo, c.b_sbo);
                mma_bf16(d_addr, a2, b2, idesc, (c.zero_first && i == 0) ? 0u : 1u);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    tc_fence_after();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int col = 0; col < 256; col += 16) {
            uint32_t r[16];
            tmem_ld16(tbase + lane_base + col, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) d_out[(warp * 32 + lane) * 256 + col + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// host side: physical images of the "dual view" layouts
// ------------------------------------------------------------------------------------------------
static uint16_t f2bf(float f) {       // exact for the small integers used here
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);
}
static uint16_t f2bf_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
    return (uint16_t)(r >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// operand buffer: element (row, feat) at (feat/8)*S_f + (row/8)*128 + (row%8)*16 + (feat%8)*2, S_f = n_rows*16
static size_t ob_off(int row, int feat, int n_rows) { return (size_t)(feat / 8) * n_rows * 16 + (row / 8) * 128 + (row % 8) * 16 + (feat % 8) * 2; }

struct Run {
    std::vector<float> d;
    long long cycles;
};
static Run run(const ProbeCfg& c, const std::vector<uint8_t>& a, const std::vector<uint8_t>& b) {
    uint8_t *da, *db;
    float* dd;
    long long* dc;
    CK(cudaMalloc(&da, a.size()));
    CK(cudaMalloc(&db, b.size()));
    CK(cudaMalloc(&dd, 128 * 256 * 4));
    CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice));
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(c, da, db, dd, dc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    Run r;
    r.d.resize(128 * 256);
    CK(cudaMemcpy(r.d.data(), dd, 128 * 256 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&r.cycles, dc, 8, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    return r;
}

static int lane_of_row(int M, int m, int lane_off) { return M == 128 ? m : (m % 16) + 32 * (m / 16) + lane_off; }

int main() {
    srand(1);
    // ---------------- test 1: chain product, K-major A (traj x feat) and K-major B (out x in)
    // A: 128 traj x K feat (operand buffer, rows = traj); B: N out x K in (weight layout, rows = out)
    for (int variant = 0; variant < 1; ++variant) {
        const int M = 128, N = 64, K = 64;
        std::vector<float> A(M * K), B(N * K);
        for (auto& v : A) v = (float)(rand() % 7 - 3);
        for (auto& v : B) v = (float)(rand() % 5 - 2);
        std::vector<uint8_t> ai(M * K * 2), bi(N * K * 2);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) *(uint16_t*)&ai[ob_off(m, k, M)] = f2bf(A[m * K + k]);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) *(uint16_t*)&bi[ob_off(n, k, N)] = f2bf(B[n * K + k]);
        ProbeCfg c{};
        c.M = M; c.N = N; c.n_mma = K / 16; c.a_mn = 0; c.b_mn = 0;
        // believed: K-major: SBO = stride between 8-row groups (128 B), LBO = stride between 8-feature groups
        c.a_sbo = 128; c.a_lbo = M * 16; c.b_sbo = 128; c.b_lbo = N * 16;
        if (variant == 1) { std::swap(c.a_sbo, c.a_lbo); std::swap(c.b_sbo, c.b_lbo); }
        c.a_step = 2 * M * 16; c.b_step = 2 * N * 16;
        c.a_bytes = ai.size(); c.b_bytes = bi.size(); c.repeat = 1; c.zero_first = 1;
        Run r = run(c, ai, bi);
        double err = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
                err = fmax(err, fabs(ref - r.d[m * 256 + n]));
            }
        printf("test1 K-major A x K-major B, M128 N64 K64, variant %d (%s): max err %.3f  cycles %lld\n", variant,
               variant == 0 ? "SBO=128,LBO=group" : "swapped", err, r.cycles);
    }
    // ---------------- test 2: delta chain: A K-major (traj x out), B = W[out][in] viewed MN-major (N' = in, K' = out)
    for (int variant = 0; variant < 2; ++variant) {
        const int M = 128, NO = 64 /*out = K'*/, NI = 32 /*in = N'*/;
        std::vector<float> A(M * NO), W(NO * NI);
        for (auto& v : A) v = (float)(rand() % 7 - 3);
        for (auto& v : W) v = (float)(rand() % 5 - 2);
        std::vector<uint8_t> ai(M * NO * 2), bi(NO * NI * 2);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < NO; ++k) *(uint16_t*)&ai[ob_off(m, k, M)] = f2bf(A[m * NO + k]);
        for (int o = 0; o < NO; ++o)
            for (int i = 0; i < NI; ++i) *(uint16_t*)&bi[ob_off(o, i, NO)] = f2bf(W[o * NI + i]);   // rows = out, feat = in
        ProbeCfg c{};
        c.M = M; c.N = NI; c.n_mma = NO / 16; c.a_mn = 0; c.b_mn = 1;
        c.a_sbo = 128; c.a_lbo = M * 16;
        // believed MN-major: SBO = stride between 8-element MN groups (in/8) = NO*16; LBO = stride between 8-row K groups = 128
        c.b_sbo = NO * 16; c.b_lbo = 128;
        if (variant == 1) std::swap(c.b_sbo, c.b_lbo);
        c.a_step = 2 * M * 16; c.b_step = 2 * 128;     // K' advances by 16 out rows = 2 row groups
        c.a_bytes = ai.size(); c.b_bytes = bi.size(); c.repeat = 1; c.zero_first = 1;
        Run r = run(c, ai, bi);
        double err = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < NI; ++n) {
                double ref = 0;
                for (int k = 0; k < NO; ++k) ref += (double)A[m * NO + k] * W[k * NI + n];
                err = fmax(err, fabs(ref - r.d[m * 256 + n]));
            }
        printf("test2 delta chain (B MN-major), M128 N32 K64, variant %d: max err %.3f  cycles %lld\n", variant, err, r.cycles);
    }
    // ---------------- test 3: dW: A = delta^T (MN-major, M = out feat, K = traj), B = act (MN-major, N = in feat, K = traj), M = 64
    for (int variant = 0; variant < 4; ++variant) {
        const int T = 128, NO = 64, NI = 72;
        std::vector<float> Dl(T * NO), Ac(T * NI);
        for (auto& v : Dl) v = (float)(rand() % 7 - 3);
        for (auto& v : Ac) v = (float)(rand() % 5 - 2);
        std::vector<uint8_t> ai(T * NO * 2), bi(T * NI * 2);
        for (int t = 0; t < T; ++t)
            for (int f = 0; f < NO; ++f) *(uint16_t*)&ai[ob_off(t, f, T)] = f2bf(Dl[t * NO + f]);
        for (int t = 0; t < T; ++t)
            for (int f = 0; f < NI; ++f) *(uint16_t*)&bi[ob_off(t, f, T)] = f2bf(Ac[t * NI + f]);
        ProbeCfg c{};
        c.M = 64; c.N = NI; c.n_mma = T / 16; c.a_mn = 1; c.b_mn = 1;
        c.a_sbo = T * 16; c.a_lbo = 128; c.b_sbo = T * 16; c.b_lbo = 128;
        if (variant & 1) { std::swap(c.a_sbo, c.a_lbo); std::swap(c.b_sbo, c.b_lbo); }
        c.d_lane_off = (variant & 2) ? 16 : 0;
        c.a_step = 256; c.b_step = 256;
        c.a_bytes = ai.size(); c.b_bytes = bi.size(); c.repeat = 1; c.zero_first = 1;
        Run r = run(c, ai, bi);
        double err = 0;
        for (int m = 0; m < NO; ++m)
            for (int n = 0; n < NI; ++n) {
                double ref = 0;
                for (int t = 0; t < T; ++t) ref += (double)Dl[t * NO + m] * Ac[t * NI + n];
                err = fmax(err, fabs(ref - r.d[lane_of_row(64, m, c.d_lane_off) * 256 + n]));
            }
        // untouched lanes must stay 0
        double stray = 0;
        for (int lane = 0; lane < 128; ++lane) {
            const bool used = ((lane % 32) / 16) == (int)(c.d_lane_off / 16);
            if (!used)
                for (int n = 0; n < NI; ++n) stray = fmax(stray, fabs(r.d[lane * 256 + n]));
        }
        printf("test3 dW (A,B MN-major) M64 N72 K128, variant %d (swap=%d lane_off=%u): max err %.3f stray %.3f cycles %lld\n",
               variant, variant & 1, c.d_lane_off, err, stray, r.cycles);
    }
    // ---------------- timing: cycles per instruction for the shapes the kernels use
    {
        struct Shape { const char* name; int M, N, a_mn, b_mn; };
        const Shape shapes[] = {
            {"chain  M128 N64  K-major/K-major", 128, 64, 0, 0}, {"chain  M128 N32  K-major/K-major", 128, 32, 0, 0},
            {"chain  M128 N16  K-major/K-major", 128, 16, 0, 0}, {"chain  M128 N128 K-major/K-major", 128, 128, 0, 0},
            {"delta  M128 N64  K-major/MN-major", 128, 64, 0, 1}, {"dW     M64  N72  MN/MN", 64, 72, 1, 1},
            {"dW     M64  N40  MN/MN", 64, 40, 1, 1},             {"dW     M64  N16  MN/MN", 64, 16, 1, 1},
            {"dW     M128 N72  MN/MN", 128, 72 + 8, 1, 1},
            {"dW     M64  N136 MN/MN", 64, 136, 1, 1},            {"dW     M128 N144 MN/MN", 128, 144, 1, 1},
            {"dW     M128 N128 MN/MN", 128, 128, 1, 1},           {"dW     M128 N144 K-major A/MN B", 128, 144, 0, 1},
            {"chain  M128 N192 K-major/K-major", 128, 192, 0, 0}, {"chain  M128 N256 K-major/K-major", 128, 256, 0, 0},
            {"chain  M64 N64 K-major/K-major", 64, 64, 0, 0},{"chain  M64 N256 K-major/K-major", 64, 256, 0, 0},
        };
        std::vector<uint8_t> ai(128 * 128 * 2, 0), bi(160 * 128 * 2, 0);
        for (const Shape& s : shapes) {
            ProbeCfg c{};
            c.M = s.M; c.N = s.N; c.n_mma = 4; c.a_mn = s.a_mn; c.b_mn = s.b_mn;
            c.a_sbo = s.a_mn ? 2048 : 128; c.a_lbo = s.a_mn ? 128 : 2048;
            c.b_sbo = s.b_mn ? 2048 : 128; c.b_lbo = s.b_mn ? 128 : 2048;
            c.a_step = s.a_mn ? 256 : 4096; c.b_step = s.b_mn ? 256 : 4096;
            c.a_bytes = ai.size(); c.b_bytes = bi.size(); c.zero_first = 1;
            c.repeat = 1;
            Run r1 = run(c, ai, bi);
            c.repeat = 65;
            Run r2 = run(c, ai, bi);
            c.d_alt = 64;
            Run r3 = run(c, ai, bi);
            printf("timing %-36s: %7.1f cycles / instruction, %7.1f with 4 rotating accumulators (4 instr: %lld cycles incl. commit+wait)\n", s.name,
                   (double)(r2.cycles - r1.cycles) / (64.0 * 4), (double)(r3.cycles - r1.cycles) / (64.0 * 4), r1.cycles);
        }
    }
    // ---------------- accuracy of the 3-way bf16 split, 6 products, K = 64
    {
        const int M = 128, N = 64, K = 64;
        std::vector<float> A(M * K), B(N * K);
        for (auto& v : A) v = (float)((rand() / (double)RAND_MAX) * 2 - 1);
        for (auto& v : B) v = (float)((rand() / (double)RAND_MAX) * 0.5 - 0.25);
        auto split3 = [](float x, uint16_t (&h)[3]) {
            float r = x;
            for (int i = 0; i < 3; ++i) {
                h[i] = f2bf_rn(r);
                r -= bf2f(h[i]);
            }
        };
        // images: three arrays each for A and B, stacked
        const size_t asz = (size_t)M * K * 2, bsz = (size_t)N * K * 2;
        std::vector<uint8_t> ai(3 * asz), bi(3 * bsz);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                uint16_t h[3];
                split3(A[m * K + k], h);
                for (int s = 0; s < 3; ++s) *(uint16_t*)&ai[s * asz + ob_off(m, k, M)] = h[s];
            }
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                uint16_t h[3];
                split3(B[n * K + k], h);
                for (int s = 0; s < 3; ++s) *(uint16_t*)&bi[s * bsz + ob_off(n, k, N)] = h[s];
            }
        // The probe kernel advances both start addresses linearly, so issue the product terms as separate launches
        // is not possible; instead emulate the accumulation order with one launch per (sa, sb) pair accumulating on
        // the host in fp32 (the MMA-internal accumulation over K is what we want to see).
        const int pairs[6][2] = {{0, 0}, {0, 1}, {1, 0}, {1, 1}, {0, 2}, {2, 0}};
        std::vector<float> acc(M * N, 0.f);
        for (int p = 0; p < 6; ++p) {
            std::vector<uint8_t> a1(ai.begin() + pairs[p][0] * asz, ai.begin() + (pairs[p][0] + 1) * asz);
            std::vector<uint8_t> b1(bi.begin() + pairs[p][1] * bsz, bi.begin() + (pairs[p][1] + 1) * bsz);
            ProbeCfg c{};
            c.M = M; c.N = N; c.n_mma = K / 16; c.a_mn = 0; c.b_mn = 0;
            c.a_sbo = 128; c.a_lbo = M * 16; c.b_sbo = 128; c.b_lbo = N * 16;
            c.a_step = 2 * M * 16; c.b_step = 2 * N * 16;
            c.a_bytes = a1.size(); c.b_bytes = b1.size(); c.repeat = 1; c.zero_first = 1;
            Run r = run(c, a1, b1);
            double e3 = 0, scale = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    acc[m * N + n] += r.d[m * 256 + n];
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
                    e3 = fmax(e3, fabs(ref - acc[m * N + n]));
                    scale = fmax(scale, fabs(ref));
                }
            printf("split3 after %d product terms: max abs err %.3e (max |ref| %.3f) -> rel %.3e\n", p + 1, e3, scale, e3 / scale);
        }
        // fp32 FFMA reference error for comparison
        double ef = 0, scale = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                float f = 0.f;
                for (int k = 0; k < K; ++k) {
                    ref += (double)A[m * K + k] * B[n * K + k];
                    f = fmaf(A[m * K + k], B[n * K + k], f);
                }
                ef = fmax(ef, fabs(ref - f));
                scale = fmax(scale, fabs(ref));
            }
        printf("fp32 fmaf chain for comparison: rel %.3e\n", ef / scale);
    }
    printf("probe done\n");
    return 0;
}
