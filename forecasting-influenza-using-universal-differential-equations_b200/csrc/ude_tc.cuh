// ude_tc.cuh — tcgen05 (5th-gen tensor core) path of the fused RK4 forward / discrete-adjoint
// kernels (sm_100a).  Used when batch x hidden width makes the MLP layers a real dense
// contraction (BASELINE config 4: 2^20 trajectories, 64-wide nets); everything else runs the
// FFMA engine of ude_engine.cuh.
//
// Replaces, like ude_rk4.cuh, the Python loop of torchdiffeq's FixedGridODESolver + rk4_alt_step_func
// driven from /root/reference/lib/VAE.py:137 and the autograd replay of /root/reference/lib/VAE.py:203,
// with the RHS of /root/reference/lib/models.py:129-146,177-188,230-254 fused in.
//
// One CTA = one tile of 128 trajectories for a whole integration.
//   * a trajectory is TMEM lane t / operand row t; thread t (warps 0-3) also keeps its RK4 state,
//     adjoint state and side-channel sums in REGISTERS;
//   * layer products   z_l = a_l W_l^T      : M = 128 traj, N = out, K = in   (A, B from smem)
//     delta products   dA_l = G_l W_l       : M = 128 traj, N = in,  K = out  (B = same image, MN-major view)
//     weight gradients dW_l = G_l^T [a_l|1] : M = 64 (out), N = in + 8, K = 128 traj (both MN-major views
//                                             of the SAME operand buffers the chain uses K-major)
//   * operands are 3-way bf16 splits (ude_tc_model.h); accumulators are fp32 in TMEM;
//   * epilogues (bias, ELU, ELU', splits) run thread-per-trajectory straight out of TMEM with
//     tcgen05.ld 32x32b; the 8 warps split the columns (warp w: lane quarter w&3, column chunks w>>2);
//   * weight gradients leave TMEM once per RHS evaluation through per-warp staging slices and bulk reductions
//     (cp.reduce.async.bulk.add.f32) into a per-CTA slab (per tile with weight-set groups); slabs are summed in a
//     fixed order afterwards (deterministic).
// Descriptor conventions (LBO/SBO of un-swizzled K-major and MN-major operands, M = 64 lane mapping)
// are pinned on hardware by tools/umma_probe.cu.
#pragma once
#include <cuda_bf16.h>

#include "ude_engine.cuh"
#include "ude_tc_model.h"

namespace ude {

struct TcFwdArgs {
    const uint8_t* blob;    // packed weight images + biases (TcModel::blob_bytes)
    const float* y0;        // [B,R,L]
    GridTables g;
    float* y_out;           // [n_out,B,R,L]
    float* tape;            // [n_steps,4,3R,B] or null (same format as the FFMA path)
    double* side_partial;   // [gridDim.x, 8]
    long long B;
    ReparamArgs rp;         // rp.mean != nullptr: y0 is produced in the prologue (ude_engine.cuh) and written as output 0
    DecArgs dec;            // dec.n_dec == 1: decoded outputs (ude_engine.cuh DecArgs; the tensor path is R = 1, one decoder output)
    // Weight-set groups (UdeDesc::n_groups = G > 1, BASELINE configs[4]): ONE launch walks the tiles of all G group-major
    // sub-problems (B = trajectories of ONE group); the weight blob is re-staged when a CTA's next tile belongs to another group;
    // side partials are kept per tile and trajectory warp: side_partial [n_tiles][4][8].
    int n_groups;
    long long blob_stride;          // bytes between the packed blobs of two groups
    long long tape_group_floats;    // floats between the tapes of two groups
    int bayes;                      // 1: `blob` holds one packed weight set per RHS evaluation (ude_engine.cuh BayesArgs), blob_stride apart
};

struct TcBwdArgs {
    const uint8_t* blob;
    const float* y0;
    GridTables g;
    const float* tape;
    const float* g_out;     // [n_out,B,R,L]
    const float* g_side;    // [5] or null
    float* g_y0;            // [B,R,L]
    float* slabs;           // [gridDim.x, slab_floats], zeroed by the host wrapper
    long long B;
    ReparamArgs rp;         // rp.mean != nullptr: also dL/dz (rp.g_z) in the epilogue
    DecArgs dec;            // dec.n_dec == 1: cotangents arrive as dL/dy_pred + g_side[5] = dL/dreg (g_out unused)
    int n_groups;           // see TcFwdArgs; slabs are then per TILE: [n_tiles][slab_floats], g_side [G][5]
    long long blob_stride;
    long long tape_group_floats;
    // Bayes funcs: `blob` holds one weight set per evaluation; the gradients of evaluation e are reduced twice, as they are into
    // `slabs` (-> dL/dw_mean) and multiplied by the evaluation's noise into `bz_slabs` (-> dL/dw_std); bayes_z is the noise in SLAB
    // layout, [4 * n_steps][slab_floats]
    const float* bayes_z;
    float* bz_slabs;        // [gridDim.x][slab_floats]
};

// ----------------------------------------------------------------------------------
// tcgen05 primitives
// ----------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version; layout type 0 = no swizzle
    return d;
}
// kind::f16, bf16 x bf16 -> fp32
__host__ __device__ __forceinline__ uint32_t tc_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// three loads in flight before one wait (the forward accumulator is three column blocks)
__device__ __forceinline__ void tc_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[16], float (&b)[16], float (&c)[16]) {
    uint32_t r[48];
#define UDE_LD16(T, O)                                                                                              \
    asm volatile(                                                                                                   \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"   \
        : "=r"(r[O + 0]), "=r"(r[O + 1]), "=r"(r[O + 2]), "=r"(r[O + 3]), "=r"(r[O + 4]), "=r"(r[O + 5]),          \
          "=r"(r[O + 6]), "=r"(r[O + 7]), "=r"(r[O + 8]), "=r"(r[O + 9]), "=r"(r[O + 10]), "=r"(r[O + 11]),        \
          "=r"(r[O + 12]), "=r"(r[O + 13]), "=r"(r[O + 14]), "=r"(r[O + 15])                                       \
        : "r"(T))
    UDE_LD16(t0, 0);
    UDE_LD16(t1, 16);
    UDE_LD16(t2, 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[16 + i]);
        c[i] = __uint_as_float(r[32 + i]);
    }
}
__device__ __forceinline__ void tc_ld16x2(uint32_t t0, uint32_t t1, float (&a)[16], float (&b)[16]) {
    uint32_t r[32];
    UDE_LD16(t0, 0);
    UDE_LD16(t1, 16);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a[i] = __uint_as_float(r[i]);
        b[i] = __uint_as_float(r[16 + i]);
    }
}
#undef UDE_LD16
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"r"(
            __float_as_uint(v[0])),
        "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])),
        "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// split issue / wait variants: several TMEM accesses in flight, one wait
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st16_issue(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::"r"(
            __float_as_uint(v[0])),
        "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
        "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
        "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])),
        "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Packed fp32 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 on register pairs).  The epilogues are bound by issue slots
// on dependent instruction streams, not by pipe throughput: one instruction per PAIR of activations where the
// scalar code needs two.  Same IEEE round-to-nearest results as the scalar operations.
__device__ __forceinline__ void sub2(float& x, float& y, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void add2(float& x, float& y, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& x, float& y, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// (x, y) = (a0, a1) * (b0, b1) + (c0, c1)
__device__ __forceinline__ void fma2(float& x, float& y, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
// ELU'(h) from the ELU output a (a > 0 ? 1 : a + 1), two at a time
// (= min(a, 0) + 1: one FMNMX each and one FADD2 instead of add, compare and select; same values)
__device__ __forceinline__ void elu_grad2(float& g0, float& g1, float a0, float a1) {
    add2(g0, g1, fminf(a0, 0.f), fminf(a1, 0.f), 1.0f, 1.0f);
}

// ELU(alpha=1) with exp(h) = ex2.approx.ftz(h * log2 e): the same MUFU.EX2 the FFMA engine's __expf uses,
// minus its denormal-range rescaling (h < -87 flushes to exp = 0, i.e. ELU = -1 exactly as it rounds anyway).
__device__ __forceinline__ float tc_elu(float h) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(h * 1.4426950408889634f));
    return h > 0.f ? h : e - 1.0f;
}
// two at a time (same values as tc_elu)
// ELU(h) = max(h, 0) + (exp(min(h, 0)) - 1): for h > 0 the second term is exactly 0, for h <= 0 the first one --
// the same values as the select, with min / max (ALU) and packed arithmetic instead of compare + select
__device__ __forceinline__ void tc_elu2(float& h0, float& h1) {
    float s0, s1, e0, e1;
    mul2(s0, s1, fminf(h0, 0.f), fminf(h1, 0.f), 1.4426950408889634f, 1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
    add2(e0, e1, e0, e1, -1.0f, -1.0f);
    add2(h0, h1, fmaxf(h0, 0.f), fmaxf(h1, 0.f), e0, e1);
}

// ----------------------------------------------------------------------------------
// 3-way bf16 split of 8 consecutive features -> one 16-byte group per split array
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
// (a, b) - (bf16 pair h as fp32): the residual of one split level, both lanes in one FADD2
__device__ __forceinline__ void split_residual(float& ra, float& rb, float a, float b, uint32_t h) {
    sub2(ra, rb, a, b, __uint_as_float(h << 16), __uint_as_float(h & 0xFFFF0000u));
}
__device__ __forceinline__ void split3_store8(const float* v /*8 values*/, uint8_t* p0, uint8_t* p1, uint8_t* p2) {
    uint32_t h0[4], h1[4], h2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        float ra, rb, sa, sb;
        h0[j] = pack_bf16x2(a, b);
        split_residual(ra, rb, a, b, h0[j]);
        h1[j] = pack_bf16x2(ra, rb);
        split_residual(sa, sb, ra, rb, h1[j]);
        h2[j] = pack_bf16x2(sa, sb);
    }
    *reinterpret_cast<uint4*>(p0) = make_uint4(h0[0], h0[1], h0[2], h0[3]);
    *reinterpret_cast<uint4*>(p1) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
    *reinterpret_cast<uint4*>(p2) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
}

// same split, 16 features -> registers: p[2*s + g] = group g (8 features) of split array s
__device__ __forceinline__ void split3_pack16(const float (&v)[16], uint4 (&p)[6]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint32_t h0[4], h1[4], h2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = v[8 * g + 2 * j], b = v[8 * g + 2 * j + 1];
            float ra, rb, sa, sb;
            h0[j] = pack_bf16x2(a, b);
            split_residual(ra, rb, a, b, h0[j]);
            h1[j] = pack_bf16x2(ra, rb);
            split_residual(sa, sb, ra, rb, h1[j]);
            h2[j] = pack_bf16x2(sa, sb);
        }
        p[0 + g] = make_uint4(h0[0], h0[1], h0[2], h0[3]);
        p[2 + g] = make_uint4(h1[0], h1[1], h1[2], h1[3]);
        p[4 + g] = make_uint4(h2[0], h2[1], h2[2], h2[3]);
    }
}

// Placement of the split arrays of an activation of width in_p inside Abuf, in 8-feature groups:
//   [ a2 | a1 | ones | a3 ]   (the weight-gradient product reads [a2|a1|ones] and [a1|ones] as ONE B operand)
__host__ __device__ __forceinline__ int abuf_group(int s, int in_p) {
    const int g = in_p >> 3;
    return s == 0 ? g : (s == 1 ? 0 : 2 * g + 1);
}
__host__ __device__ __forceinline__ int abuf_ones_group(int in_p) { return 2 * (in_p >> 3); }

// ----------------------------------------------------------------------------------
// MMA issue (one thread)
// ----------------------------------------------------------------------------------
// forward: D[128, out_p] = sum_{i+j<=4} a_i w_j, ONE fp32 accumulator block.  (Accumulating the small products into
// the same columns as a1 w1 costs nothing in accuracy -- every partial sum is rounded at the magnitude of the
// result either way -- and keeps the accumulator of a layer at out_p TMEM columns, so the rate and the aug chain
// can be in flight at the same time next to the saved activations of both nets.)
template <class Emit>
__host__ __device__ __forceinline__ void tc_issue_forward(const TcLayer& Lp, uint32_t abuf, uint32_t wimg, uint32_t d_addr, Emit&& tc_mma) {
    const int ksteps = Lp.in_p >> 4;
    const uint32_t w_kgrp = 3u * Lp.out_p * 16u;
    const uint32_t id = tc_idesc(128, Lp.out_p, 0, 0);
    uint32_t acc = 0;
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t ao = k * 2 * kTcGroupBytes;
#pragma unroll
        for (int p = 0; p < 6; ++p) {      // smallest products first
            const int i = (0x001210 >> (4 * p)) & 15;      // a split: 0,1,2,1,0,0
            const int j = (0x010012 >> (4 * p)) & 15;      // w split: 2,1,0,0,1,0
            tc_mma(d_addr, tc_desc(abuf + abuf_group(i, Lp.in_p) * kTcGroupBytes + ao, kTcGroupBytes, 128),
                   tc_desc(wimg + Lp.w_off + j * Lp.out_p * 16 + k * 2 * w_kgrp, w_kgrp, 128), id, acc);
            acc = 1u;
        }
    }
}
// delta chain: D[128, in_p] = sum_{i+j<=4} G_i W_j   (W_j: MN-major view of split array j)
template <class Emit>
__host__ __device__ __forceinline__ void tc_issue_delta(const TcLayer& Lp, uint32_t dbuf, uint32_t wimg, uint32_t d_addr, Emit&& tc_mma) {
    const int ksteps = Lp.out_p >> 4;
    const uint32_t g_arr = (Lp.out_p >> 3) * kTcGroupBytes;
    const uint32_t w_kgrp = 3u * Lp.out_p * 16u;
    const uint32_t id = tc_idesc(128, Lp.in_p, 0, 1);
    uint32_t acc = 0;
    for (int k = 0; k < ksteps; ++k) {
#pragma unroll
        for (int p = 0; p < 6; ++p) {
            const int i = (p == 2 || p == 3) ? 1 : (p == 5 ? 2 : 0);
            const int j = (p == 1 || p == 3) ? 1 : (p == 4 ? 2 : 0);
            tc_mma(d_addr, tc_desc(dbuf + i * g_arr + k * 2 * kTcGroupBytes, kTcGroupBytes, 128),
                   tc_desc(wimg + Lp.w_off + j * Lp.out_p * 16 + k * 256, 128, w_kgrp), id, acc);
            acc = 1u;
        }
    }
}
// weight gradients: Q[64, 2*in_p + 8]: cols [0,in_p) = G1 a2, [in_p,2in_p) = (G1+G2) a1, col 2in_p = sum_t (G1+G2)
// = bias gradient;  dW = cols[k] + cols[in_p + k].
// Only the first-order products of the 3-way split are formed: the second-order ones (G1 a3, G3 a1, G2 a2, each
// ~2^-16 relative) are zero-mean round-off residuals, independent across trajectories, and the sum over
// >= 128 trajectories x 4 stages averages them to ~1e-6 relative (measured 1.7e-6 at ONE tile, ONE step;
// tests pin it).  The activation images are [a2 | a1 | ones] in one of two alternating sweep buffers, so the
// B-operand addresses here are relative to the sweep buffer (tc_issue_table adds the buffer offset).
template <class Emit>
__host__ __device__ __forceinline__ void tc_issue_dw(const TcLayer& Lp, uint32_t dbuf, uint32_t abuf, uint32_t q_addr, int g_groups,
                                                     Emit&& tc_mma) {
    const uint32_t g_arr = (uint32_t)g_groups * kTcGroupBytes;      // distance of the G2 array from the G1 array
    const int in_p = Lp.in_p;
    const uint32_t id_full = tc_idesc(64, 2 * in_p + 8, 1, 1), id_a1 = tc_idesc(64, in_p + 8, 1, 1);
    const uint32_t b_a1 = abuf + abuf_group(0, in_p) * kTcGroupBytes;
    for (int k = 0; k < kTcTile / 16; ++k) {
        const uint32_t ko = k * 256;
        const uint64_t g1 = tc_desc(dbuf + ko, 128, kTcGroupBytes), g2 = tc_desc(dbuf + g_arr + ko, 128, kTcGroupBytes);
        tc_mma(q_addr, g1, tc_desc(abuf + ko, 128, kTcGroupBytes), id_full, k > 0);
        tc_mma(q_addr + in_p, g2, tc_desc(b_a1 + ko, 128, kTcGroupBytes), id_a1, 1u);
    }
}

// ----------------------------------------------------------------------------------
// MMA tables.  Every descriptor of a launch is loop invariant (the operand buffers never move), so each
// layer's instruction list is built ONCE, on the host, and travels as a __grid_constant__ kernel
// parameter: the issuing warp streams 32-byte entries from the constant bank through UNIFORM registers
// straight into tcgen05.mma (one elected lane).  With in-loop descriptor arithmetic in vector registers
// the single issuing thread was the bottleneck: 115-125 cycles per instruction against the 50-cycle
// hardware floor (tools/umma_probe.cu, tools/tc_prof.py).
// Shared-memory addresses in the table are relative to the dynamic shared-memory base.
// ----------------------------------------------------------------------------------
struct __align__(16) TcMma {
    uint32_t a_lo, a_hi, b_lo, b_hi;
    uint32_t idesc, d_col, acc, pad;     // d_col: TMEM column relative to the phase's accumulator base
};
constexpr int kTcMaxMma = 640;
struct TcTable {
    TcMma e[kTcMaxMma];
};

struct TcRecorder {
    TcMma* out;
    int n;
    __host__ __device__ void operator()(uint32_t d_col, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
        TcMma e;
        e.a_lo = (uint32_t)adesc; e.a_hi = (uint32_t)(adesc >> 32);
        e.b_lo = (uint32_t)bdesc; e.b_hi = (uint32_t)(bdesc >> 32);
        e.idesc = idesc; e.d_col = d_col; e.acc = acc; e.pad = 0;
        out[n++] = e;
    }
};

// one elected lane of a converged warp issues; everything else stays warp-uniform
__device__ __forceinline__ void tc_mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// executed by ALL lanes of warp 0
__device__ __forceinline__ void tc_issue_table(const TcTable& tab, int begin, int n, uint32_t smem_base16, uint32_t tmem,
                                               uint32_t b_extra16 = 0u, uint32_t a_extra16 = 0u) {
#pragma unroll 2
    for (int i = begin; i < begin + n; ++i) {
        const TcMma& e = tab.e[i];
        const uint64_t ad = ((uint64_t)e.a_hi << 32) | (uint64_t)(e.a_lo + smem_base16 + a_extra16);
        const uint64_t bd = ((uint64_t)e.b_hi << 32) | (uint64_t)(e.b_lo + smem_base16 + b_extra16);
        tc_mma_elect(tmem + e.d_col, ad, bd, e.idesc, e.acc);
    }
}

// optional cycle accounting (-DUDE_TC_PROFILE): CTA 0, threads 0 (issuer) and 128 (warp 4) accumulate the
// cycles they spend per category; read back with ude_debug_tc_prof().
#ifdef UDE_TC_PROFILE
__device__ unsigned long long g_tc_prof[2 * 8];
#define TCP_T0() const long long tcp_t0 = clock64()
#define TCP_ADD(c, id) (c).prof[id] += clock64() - tcp_t0
#define TCP_T1() const long long tcp_t1 = clock64()
#define TCP_ADD1(c, id) (c).prof[id] += clock64() - tcp_t1
#else
#define TCP_T0()
#define TCP_ADD(c, id)
#define TCP_T1()
#define TCP_ADD1(c, id)
#endif
enum { TCP_SYNC = 0, TCP_ISSUE = 1, TCP_WAIT = 2, TCP_EPI_FWD = 3, TCP_EPI_BWD = 4, TCP_TAIL = 5, TCP_TOTAL = 6 };

// ----------------------------------------------------------------------------------
// per-CTA context
// ----------------------------------------------------------------------------------
struct TcCtx {
#ifdef UDE_TC_PROFILE
    long long prof[8];
#endif
    uint8_t* abuf_p;        // operand buffers (generic pointers): forward image of the rate chain / sweep images
    uint8_t* abuf2_p;       // forward image of the aug chain
    uint8_t* dbuf_p;
    uint8_t* ghead_p;       // [G1 | G2] image of the head cotangent
    uint32_t abuf, dbuf, wimg;   // shared-space addresses
    const float* bias;      // bias block in smem
    uint32_t tmem;          // TMEM base address
    uint64_t* bar;          // MMA completion barrier (forward products, delta chain)
    uint64_t* bar_b;        // second completion barrier (weight-gradient products)
    uint64_t* bar2;         // completion barrier of the aug chain's forward products
    uint32_t phase_b, phase2;
    uint32_t phase_h[2];    // head hand-off barriers (trajectory warps)
    uint8_t* stage_p;       // backward: gradient staging buffer of the bulk reductions
    const float* bz_z;      // Bayes funcs: this evaluation's noise in slab layout (nullptr: off) and the second slab
    float* bz_slab;
    uint32_t sweep_par;     // sweep buffer of the current backward phase
    uint32_t sweep_buf_bytes;
    uint32_t smem_base16;   // dynamic shared-memory base >> 4 (MMA table addresses are relative)
    uint32_t phase;
    int warp;               // warp index (warp-uniform by construction)
    int q, h, lane, t;      // lane quarter, column half, lane, trajectory row of this thread
    uint32_t lane_addr;     // (32*q) << 16
};

// Operand hand-off between the 8 epilogue warps and the dedicated MMA-issuing warp (warp 8): the epilogue
// warps ARRIVE (non-blocking) once their shared-memory writes and TMEM reads of the phase are fenced, the
// issuer SYNCs on the same named barrier and fires the phase's instruction list.  The issuer blocks in the
// tensor queue for about as long as the MMAs run, which is exactly why it must not be an epilogue warp.
constexpr int kTcBarOperands = 1;      // rate chain / backward sweep
constexpr int kTcBarOperands2 = 2;     // aug chain
constexpr int kTcBarDrained = 3;       // backward sweep: weight-gradient accumulator read out
// register hand-over (setmaxnreg): 65536 >= 32 * (kTcEpiWarps * kTcEpiRegs + 4 * kTcIssuerRegs)
// (the pool is what the CTA got at launch: threads x launch registers; asking for more than the issuer warpgroup
// released blocks forever.  With 16 epilogue warps ptxas fits the epilogue code into the 96 launch registers and
// no hand-over is needed.)
constexpr int kTcIssuerRegs = 40;
constexpr int kTcEpiRegs = 232;
constexpr bool kTcRegHandOver = kTcEpiWarps == 8;
__device__ __forceinline__ void tc_operands_ready(TcCtx& c, int bar_id = kTcBarOperands) {
    TCP_T0();
    tc_proxy_fence();
    tc_fence_before();
    asm volatile("bar.arrive %0, %1;" ::"r"(bar_id), "n"(kTcBarThreads) : "memory");
    TCP_ADD(c, TCP_SYNC);
}
__device__ __forceinline__ void tc_issuer_wait_operands(int bar_id = kTcBarOperands) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kTcBarThreads) : "memory");
    tc_fence_after();
}
// issuer side of one forward phase / one backward phase
__device__ __forceinline__ void tc_issuer_forward(const TcTable& tab, const TcLayer& Lp, uint32_t smem_base16, uint32_t tmem_acc,
                                                  uint64_t* bar, int bar_id, uint32_t a_extra16) {
    tc_issuer_wait_operands(bar_id);
    tc_issue_table(tab, Lp.mma_fwd, Lp.n_fwd, smem_base16, tmem_acc, 0u, a_extra16);
    tc_commit_elect(bar);
}
// issuer side of the backward sweep of one net (mirrors tc_net_backward): the head's Wh^T g and the first layer's
// W0^T delta0 run on the FMA pipe, so the delta chain C_l exists for the hidden layers l = n-2 .. 1 only; the
// weight-gradient products W_l of every layer share one accumulator and wait for the previous one to be drained.
__device__ __forceinline__ void tc_issuer_net_backward(const TcTable& tab, const TcLayer* Ls, int n, uint32_t smem_base16, uint32_t tmem,
                                                       uint64_t* bar, uint64_t* bar_b, uint32_t buf16, uint32_t& par) {
    tc_issuer_wait_operands(kTcBarOperands2);                    // step H1: head image + image of a_{n-2}
    tc_issue_table(tab, Ls[n - 1].mma_w, Ls[n - 1].n_w, smem_base16, tmem, par ? buf16 : 0u);
    tc_commit_elect(bar_b);
    par ^= 1u;
    tc_issuer_wait_operands();                                   // step H2: G_{n-2}, image of a_{n-3}
    tc_issue_table(tab, Ls[n - 2].mma_c, Ls[n - 2].n_c, smem_base16, tmem);
    tc_commit_elect(bar);
    tc_issuer_wait_operands(kTcBarDrained);                      // the head's gradients have left the accumulator
    tc_issue_table(tab, Ls[n - 2].mma_w, Ls[n - 2].n_w, smem_base16, tmem, par ? buf16 : 0u);
    tc_commit_elect(bar_b);
    for (int l = n - 2; l >= 1; --l) {
        tc_issuer_wait_operands();                               // G_{l-1}
        if (l - 1 >= 1) {
            tc_issue_table(tab, Ls[l - 1].mma_c, Ls[l - 1].n_c, smem_base16, tmem);
            tc_commit_elect(bar);
        }
        par ^= 1u;
        tc_issuer_wait_operands(kTcBarDrained);
        tc_issue_table(tab, Ls[l - 1].mma_w, Ls[l - 1].n_w, smem_base16, tmem, par ? buf16 : 0u);
        tc_commit_elect(bar_b);
    }
    par ^= 1u;
}
__device__ __forceinline__ void tc_wait_mma(TcCtx& c) {
    TCP_T0();
    mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    tc_fence_after();
    TCP_ADD(c, TCP_WAIT);
}
__device__ __forceinline__ void tc_wait_mma_b(TcCtx& c) {
    TCP_T0();
    mbar_wait(c.bar_b, c.phase_b);
    c.phase_b ^= 1u;
    tc_fence_after();
    TCP_ADD(c, TCP_WAIT);
}
__device__ __forceinline__ void tc_prof_init(TcCtx& c) {
#ifdef UDE_TC_PROFILE
    for (int i = 0; i < 8; ++i) c.prof[i] = 0;
    c.prof[TCP_TOTAL] = -clock64();
#endif
}
__device__ __forceinline__ void tc_prof_dump(TcCtx& c) {
#ifdef UDE_TC_PROFILE
    c.prof[TCP_TOTAL] += clock64();
    if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 128))
        for (int i = 0; i < 8; ++i) g_tc_prof[(threadIdx.x == 0 ? 0 : 8) + i] = (unsigned long long)c.prof[i];
#endif
}

// write 16 consecutive features [16*chunk, 16*chunk+16) of an activation of width `width` into Abuf
__device__ __forceinline__ void tc_store_act16(const TcCtx& c, uint8_t* img, const float (&v)[16], int chunk, int width) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint8_t* base = img + (size_t)(2 * chunk + g) * kTcGroupBytes + c.t * 16;
        split3_store8(v + 8 * g, base + abuf_group(0, width) * kTcGroupBytes, base + abuf_group(1, width) * kTcGroupBytes,
                      base + abuf_group(2, width) * kTcGroupBytes);
    }
}
// write 16 consecutive features of a cotangent of width `width` into Dbuf ([G1 | G2 | G3])
__device__ __forceinline__ void tc_store_delta16(const TcCtx& c, const float (&v)[16], int chunk, int width) {
    const size_t arr = (size_t)(width >> 3) * kTcGroupBytes;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint8_t* base = c.dbuf_p + (size_t)(2 * chunk + g) * kTcGroupBytes + c.t * 16;
        split3_store8(v + 8 * g, base, base + arr, base + 2 * arr);
    }
}
// 2-way split of 8 consecutive features (the sweep images carry a1, a2 only)
__device__ __forceinline__ void split2_store8(const float* v, uint8_t* p0, uint8_t* p1) {
    uint32_t h0[4], h1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        float ra, rb;
        h0[j] = pack_bf16x2(a, b);
        split_residual(ra, rb, a, b, h0[j]);
        h1[j] = pack_bf16x2(ra, rb);
    }
    *reinterpret_cast<uint4*>(p0) = make_uint4(h0[0], h0[1], h0[2], h0[3]);
    *reinterpret_cast<uint4*>(p1) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
}
// sweep image [a2 | a1 | ones] of an activation of width `width` into sweep buffer `buf`
__device__ __forceinline__ void tc_store_sweep16(const TcCtx& c, uint32_t buf, const float (&v)[16], int chunk, int width) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint8_t* base = c.abuf_p + (size_t)buf * c.sweep_buf_bytes + (size_t)(2 * chunk + g) * kTcGroupBytes + c.t * 16;
        split2_store8(v + 8 * g, base + abuf_group(0, width) * kTcGroupBytes, base + abuf_group(1, width) * kTcGroupBytes);
    }
}
__device__ __forceinline__ void tc_store_sweep_ones(const TcCtx& c, uint32_t buf, int width) {
    uint8_t* p = c.abuf_p + (size_t)buf * c.sweep_buf_bytes + (size_t)abuf_ones_group(width) * kTcGroupBytes + c.t * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(0x00003F80u, 0u, 0u, 0u);      // feature 0 = bf16(1.0)
}
__device__ __forceinline__ void tc_store_pack_delta(const TcCtx& c, const uint4 (&p)[6], int chunk, int width) {
    const size_t arr = (size_t)(width >> 3) * kTcGroupBytes;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint8_t* base = c.dbuf_p + (size_t)(2 * chunk + g) * kTcGroupBytes + c.t * 16;
#pragma unroll
        for (int sidx = 0; sidx < 3; ++sidx) *reinterpret_cast<uint4*>(base + sidx * arr) = p[2 * sidx + g];
    }
}
__device__ __forceinline__ void tc_store_ones(const TcCtx& c, int width) {
    uint8_t* p = c.abuf_p + (size_t)abuf_ones_group(width) * kTcGroupBytes + c.t * 16;
    *reinterpret_cast<uint4*>(p) = make_uint4(0x00003F80u, 0u, 0u, 0u);      // feature 0 = bf16(1.0)
}

// X (the MLP input u = vec(x), lib/models.py:119 Flatten) -> Abuf as the width-16 input of layer 0
template <int D0>
__device__ __forceinline__ void tc_store_x(const TcCtx& c, uint8_t* img, const float (&x)[D0]) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = i < D0 ? x[i] : 0.f;
    tc_store_act16(c, img, v, 0, 16);
}

// ----------------------------------------------------------------------------------
// Forward of the nets over the tile.  The rate chain and the aug chain (lib/models.py:230-254: both nets eat the
// same input) advance in LOCK STEP: while the epilogue warps turn the accumulator of rate layer p into the operand
// image of rate layer p+1, the tensor pipe runs aug layer p, and vice versa.  Each chain has its own forward image,
// accumulator columns, completion mbarrier and operand hand-off barrier.
// ----------------------------------------------------------------------------------
struct TcChain {
    uint8_t* img;           // forward image of this chain
    uint64_t* bar;          // completion barrier of this chain's products
    uint32_t acc_col;       // TMEM accumulator (column offset)
    int bar_id;             // operand hand-off barrier
};
// The head layers (2R / 3R outputs, lib/models.py:124,172) are too thin for the tensor pipe (a 64 -> 3 product costs
// as many instructions as 64 -> 64): they run on the FMA pipe in fp32.  Every epilogue warp folds the 16 activations it
// just produced into NH partial dot products; the helper warps park theirs in shared memory, the trajectory warp of
// the same lane quarter adds them up one lock-step slot later.
constexpr int kTcBarSweep = 12;     // named barriers 12..15: backward sweep, lane quarter q
constexpr int kTcBarTile = 4;       // the 8 epilogue warps at a tile boundary (weight-set groups: before the blob is re-staged)
struct TcHeadScratch {
    uint64_t bar[2][4];             // helper warps -> trajectory warp of the lane quarter, per chain (count kTcSlots - 1);
                                    // the partial sums themselves travel through the chain's (idle) TMEM accumulator
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NH>
__device__ __forceinline__ void tc_head_partial(const float* hw, int in_p, int ch, const float (&v)[16], float (&acc)[NH]) {
#pragma unroll
    for (int o = 0; o < NH; ++o) {
        const float4* w4 = reinterpret_cast<const float4*>(hw + o * in_p + 16 * ch);
        float a0 = 0.f, a1 = 0.f;                  // even / odd features (packed FFMA2), summed at the end
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 w = w4[i];
            fma2(a0, a1, v[4 * i + 0], v[4 * i + 1], w.x, w.y, a0, a1);
            fma2(a0, a1, v[4 * i + 2], v[4 * i + 3], w.z, w.w, a0, a1);
        }
        acc[o] += a0 + a1;
    }
}
// epilogue of hidden layer Lp of one chain: accumulator -> bias, ELU -> (saved activation) -> image of the next
// layer, or (last hidden layer) the partial head products
template <bool SAVE, int NH>
__device__ __forceinline__ void tc_fwd_epilogue(const TcLayer& Lp, const TcLayer* Lhead, int chain, const TcChain& ch_, uint32_t& phase,
                                                TcCtx& c, TcHeadScratch* hs, float (&head)[NH]) {
    {
        TCP_T0();
        mbar_wait(ch_.bar, phase);
        phase ^= 1u;
        tc_fence_after();
        TCP_ADD(c, TCP_WAIT);
    }
    TCP_T0();
    const uint32_t d0 = c.tmem + c.lane_addr + ch_.acc_col;
    const float* bias = c.bias + Lp.b_off;
    float acc[NH];
#pragma unroll
    for (int o = 0; o < NH; ++o) acc[o] = 0.f;
    for (int ch = c.h; ch < (Lp.out_p >> 4); ch += kTcSlots) {
        float v0[16];
        tc_ld16(d0 + 16 * ch, v0);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias + 16 * ch + i);
            add2(v0[i + 0], v0[i + 1], v0[i + 0], v0[i + 1], b.x, b.y);
            add2(v0[i + 2], v0[i + 3], v0[i + 2], v0[i + 3], b.z, b.w);
        }
        if (Lp.elu) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) tc_elu2(v0[i], v0[i + 1]);
        }
        if (SAVE) tc_st16_issue(c.tmem + c.lane_addr + Lp.save_col + 16 * ch, v0);      // completion awaited once, below
        if (Lhead) tc_head_partial<NH>(c.bias + Lhead->hw_off, Lhead->in_p, ch, v0, acc);
        else tc_store_act16(c, ch_.img, v0, ch, Lp.out_p);
    }
    if (SAVE) tc_wait_st();
    TCP_ADD(c, TCP_EPI_FWD);
    if (!Lhead) {
        tc_operands_ready(c, ch_.bar_id);
    } else {
        if (c.h == 0) {
#pragma unroll
            for (int o = 0; o < NH; ++o) head[o] = acc[o];
        } else {
            float pv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pv[i] = 0.f;
#pragma unroll
            for (int o = 0; o < NH; ++o) pv[o] = acc[o];
            // into the columns of this warp's OWN first chunk (already consumed by itself): the trajectory warp of the
            // quarter may still be reading its chunks of the same accumulator
            tc_st16(d0 + 16 * c.h, pv);
            tc_fence_before();
            __syncwarp();
            if (c.lane == 0) mbar_arrive(&hs->bar[chain][c.q]);
        }
        // the accumulator and the saved activations are read: nothing else to hand over (the next products of this
        // chain are issued after the next evaluation's operands_ready, which these threads reach in program order)
    }
}
// trajectory warps: head = own partial + helpers' partials + bias
template <int NH>
__device__ __forceinline__ void tc_head_finish(const TcLayer& Lh, int chain, const TcChain& ch_, TcCtx& c, TcHeadScratch* hs, uint32_t& phase,
                                               float (&head)[NH]) {
    if (c.h != 0) return;
    if (kTcSlots > 1) {
        mbar_wait(&hs->bar[chain][c.q], phase);
        phase ^= 1u;
        tc_fence_after();
    }
    const float* bias = c.bias + Lh.b_off;
    float s[NH];
#pragma unroll
    for (int o = 0; o < NH; ++o) s[o] = 0.f;
#pragma unroll
    for (int k = 0; k + 1 < kTcSlots; ++k) {
        float pv[16];
        tc_ld16(c.tmem + c.lane_addr + ch_.acc_col + 16 * (k + 1), pv);
#pragma unroll
        for (int o = 0; o < NH; ++o) s[o] += pv[o];
    }
#pragma unroll
    for (int o = 0; o < NH; ++o) head[o] = (head[o] + s[o]) + bias[o];
}
// one RHS evaluation's MLPs: x (registers of the trajectory threads) -> head outputs
template <bool SAVE, int R, int L>
__device__ __forceinline__ void tc_nets_forward(const TcModel& M, const TcChain& cr, const TcChain& ca, TcCtx& c, TcHeadScratch* hs,
                                                bool is_traj, const float (&x)[R * L], float (&o_r)[2 * R], float (&o_a)[3 * R]) {
    static_assert(R == 1, "FFMA heads: 2R / 3R partials travel as one float4");
    if (is_traj) {
        if (M.n_rate > 0) tc_store_x<R * L>(c, cr.img, x);
        if (M.n_aug > 0) tc_store_x<R * L>(c, ca.img, x);
    }
    if (M.n_rate > 0) tc_operands_ready(c, cr.bar_id);
    if (M.n_aug > 0) tc_operands_ready(c, ca.bar_id);
    const int np = M.n_rate > M.n_aug ? M.n_rate : M.n_aug;
    for (int p = 0; p < np; ++p) {
        if (p < M.n_rate - 1) tc_fwd_epilogue<SAVE, 2 * R>(M.rate[p], p == M.n_rate - 2 ? &M.rate[p + 1] : nullptr, 0, cr, c.phase, c, hs, o_r);
        else if (p == M.n_rate - 1) tc_head_finish<2 * R>(M.rate[p], 0, cr, c, hs, c.phase_h[0], o_r);
        if (p < M.n_aug - 1) tc_fwd_epilogue<SAVE, 3 * R>(M.aug[p], p == M.n_aug - 2 ? &M.aug[p + 1] : nullptr, 1, ca, c.phase2, c, hs, o_a);
        else if (p == M.n_aug - 1) tc_head_finish<3 * R>(M.aug[p], 1, ca, c, hs, c.phase_h[1], o_a);
    }
}

// ----------------------------------------------------------------------------------
// RHS tail on registers (same arithmetic, same operation order as rhs_tail in ude_engine.cuh):
//   lib/models.py:135-146 (Fp), :183-188 (Fa), :236-254 (FaFp)
// ----------------------------------------------------------------------------------
template <int R, int L>
__device__ __forceinline__ void tc_rhs_tail(const TcModel& M, const float (&x)[R * L], const float (&o_r)[2 * R],
                                            const float (&o_a)[3 * R], bool valid, SideAcc& side, float (&f)[3 * R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float x0 = x[r * L], x1 = x[r * L + 1], x2 = x[r * L + 2];
        float f0 = 0.f, f1 = 0.f, f2 = 0.f;
        if (M.n_rate > 0) {
            const float beta = fabsf(o_r[2 * r]), gamma = fabsf(o_r[2 * r + 1]);
            const float plus_i = __fmul_rn(__fmul_rn(beta, x0), x1);
            const float minus_i = __fmul_rn(gamma, x1);
            f0 = -plus_i;
            f1 = __fsub_rn(plus_i, minus_i);
            f2 = minus_i;
            if (valid) {
                side.sb += beta; side.sg += gamma;
                side.sb2 += (double)beta * beta; side.sg2 += (double)gamma * gamma;
            }
        }
        if (M.n_aug > 0) {
            const float a0 = o_a[3 * r], a1 = o_a[3 * r + 1], a2 = o_a[3 * r + 2];
            if (M.n_rate > 0) {
                f0 = __fadd_rn(f0, __fmul_rn(M.fa_w, a0));
                f1 = __fadd_rn(f1, __fmul_rn(M.fa_w, a1));
                f2 = __fadd_rn(f2, __fmul_rn(M.fa_w, a2));
            } else {
                f0 = a0; f1 = a1; f2 = a2;
            }
            if (valid) side.sf2 += (double)a0 * a0 + (double)a1 * a1 + (double)a2 * a2;
        }
        if (valid) side.rows += 1;
        if (x0 > 2.f || x0 < -1.f) f0 = 0.f;
        if (x1 > 2.f || x1 < -1.f) f1 = 0.f;
        if (x2 > 2.f || x2 < -1.f) f2 = 0.f;
        f[3 * r] = f0; f[3 * r + 1] = f1; f[3 * r + 2] = f2;
    }
}

// VJP of the tail (same arithmetic as rhs_tail_vjp): g = dL/df -> head cotangents + direct dL/dx
template <int R, int L>
__device__ __forceinline__ void tc_rhs_tail_vjp(const TcModel& M, const float (&x)[R * L], const float (&o_r)[2 * R],
                                                const float (&o_a)[3 * R], bool valid, const float (&gs)[5],
                                                const float (&g)[3 * R], float (&d_r)[2 * R], float (&d_a)[3 * R],
                                                float (&xb)[3 * R]) {
    const float vm = valid ? 1.f : 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float x0 = x[r * L], x1 = x[r * L + 1], x2 = x[r * L + 2];
        float g0 = g[3 * r], g1 = g[3 * r + 1], g2 = g[3 * r + 2];
        if (x0 > 2.f || x0 < -1.f) g0 = 0.f;
        if (x1 > 2.f || x1 < -1.f) g1 = 0.f;
        if (x2 > 2.f || x2 < -1.f) g2 = 0.f;
        xb[3 * r] = xb[3 * r + 1] = xb[3 * r + 2] = 0.f;
        if (M.n_rate > 0) {
            const float o0 = o_r[2 * r], o1 = o_r[2 * r + 1];
            const float beta = fabsf(o0), gamma = fabsf(o1);
            const float gp = g1 - g0, gm = g2 - g1;
            const float dbeta = gp * x0 * x1 + vm * (gs[0] + 2.f * gs[2] * beta);
            const float dgamma = gm * x1 + vm * (gs[1] + 2.f * gs[3] * gamma);
            d_r[2 * r] = (o0 > 0.f ? dbeta : (o0 < 0.f ? -dbeta : 0.f));
            d_r[2 * r + 1] = (o1 > 0.f ? dgamma : (o1 < 0.f ? -dgamma : 0.f));
            xb[3 * r] = gp * beta * x1;
            xb[3 * r + 1] = gp * beta * x0 + gm * gamma;
        }
        if (M.n_aug > 0) {
            const float w = M.n_rate > 0 ? M.fa_w : 1.0f;
            const float cc = vm * 2.f * gs[4];
            d_a[3 * r] = w * g0 + cc * o_a[3 * r];
            d_a[3 * r + 1] = w * g1 + cc * o_a[3 * r + 1];
            d_a[3 * r + 2] = w * g2 + cc * o_a[3 * r + 2];
        }
    }
}

// ----------------------------------------------------------------------------------
// common kernel prologue / epilogue
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_copy_model(TcModel* dst, const TcModel& src) {
    const int* s = reinterpret_cast<const int*>(&src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(TcModel) / sizeof(int)); i += kTcThreads) d[i] = s[i];
}

__device__ __forceinline__ void tc_stage_blob(uint8_t* dst, const uint8_t* src, uint32_t total, uint64_t* bar, uint32_t& phase) {
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, total);
        uint32_t done = 0;
        while (done < total) {
            uint32_t chunk = total - done;
            if (chunk > 32768u) chunk = 32768u;
            tma_bulk_g2s(dst + done, src + done, chunk, bar);
            done += chunk;
        }
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
}

template <int NCOLS>
__device__ __forceinline__ void tc_tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tc_tmem_free(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS));
}

// block reduction of the side sums over the 128 trajectory threads (warps 0-3)
__device__ __forceinline__ void tc_side_reduce(const SideAcc& s, double n_evals, double* scratch /*[8*8]*/, double* out) {
    double v[7] = {s.sb, s.sg, s.sb2, s.sg2, s.sf2, (double)s.rows, s.reg};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 7; ++i) scratch[warp * 8 + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0.0;
        for (int w = 0; w < kTcThreads / 32; ++w) t += scratch[w * 8 + threadIdx.x];
        out[threadIdx.x == 6 ? 7 : threadIdx.x] = t;
    }
    if (threadIdx.x == 7) out[6] = n_evals;
}

// ----------------------------------------------------------------------------------
// the issuing warp replays the (data independent) phase sequence of its CTA
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_issuer_nets_forward(const TcModel& M, const TcTable& tab, uint32_t smem_base16, uint32_t tmem_r,
                                                       uint32_t tmem_a, uint64_t* bar, uint64_t* bar2) {
    const int np = M.n_rate > M.n_aug ? M.n_rate : M.n_aug;
    const uint32_t a2 = (uint32_t)(M.off_abuf2 - M.off_abuf) >> 4;
    for (int p = 0; p < np; ++p) {
        if (p < M.n_rate - 1) tc_issuer_forward(tab, M.rate[p], smem_base16, tmem_r, bar, kTcBarOperands, 0u);
        if (p < M.n_aug - 1) tc_issuer_forward(tab, M.aug[p], smem_base16, tmem_a, bar2, kTcBarOperands2, a2);
    }
}
__device__ __forceinline__ void tc_issuer_loop_fwd(const TcModel& M, const TcTable& tab, long long n_my_tiles, int n_steps,
                                                   uint32_t smem_base16, uint32_t tmem, uint64_t* bar, uint64_t* bar2) {
    for (long long tile = 0; tile < n_my_tiles; ++tile)
        for (int e = 0; e < 4 * n_steps; ++e)
            tc_issuer_nets_forward(M, tab, smem_base16, tmem + M.tm_acc_fwd, tmem + M.tm_acc_fwd + M.tm_acc_stride, bar, bar2);
}
__device__ __forceinline__ void tc_issuer_loop_bwd(const TcModel& M, const TcTable& tab, long long n_my_tiles, int n_steps,
                                                   uint32_t smem_base16, uint32_t tmem, uint64_t* bar, uint64_t* bar2, uint64_t* bar_b) {
    uint32_t par = 0;      // which of the two sweep buffers holds the current phase's activation image
    for (long long tile = 0; tile < n_my_tiles; ++tile)
        for (int e = 0; e < 4 * n_steps; ++e) {
            tc_issuer_nets_forward(M, tab, smem_base16, tmem + M.tm_fwd, tmem + M.tm_fwd + M.tm_acc_stride, bar, bar2);
            if (M.n_rate > 0) tc_issuer_net_backward(tab, M.rate, M.n_rate, smem_base16, tmem, bar, bar_b, (uint32_t)(M.sweep_buf_bytes >> 4), par);
            if (M.n_aug > 0) tc_issuer_net_backward(tab, M.aug, M.n_aug, smem_base16, tmem, bar, bar_b, (uint32_t)(M.sweep_buf_bytes >> 4), par);
        }
}

// ===================================================================================
// forward kernel
// ===================================================================================
// ---- decoded outputs (ude_engine.cuh DecArgs), tensor path: R = 1, ONE decoder output, thread = trajectory ------------
__device__ __forceinline__ void tc_decode_store(const DecArgs& d, int jj, long long B, long long traj, float v0, float v1, float v2,
                                                SideAcc& side) {
    float acc = __ldg(d.b);
    acc = fmaf(__ldg(d.W), v0, acc);
    acc = fmaf(__ldg(d.W + 1), v1, acc);
    acc = fmaf(__ldg(d.W + 2), v2, acc);
    d.y_pred[(long long)jj * B + traj] = acc;
    d.mask[(long long)jj * B + traj] = (unsigned char)reg_mask_bits(v0, v1, v2);
    side.reg += (double)reg_term(v0) + (double)reg_term(v1) + (double)reg_term(v2);
}
// adjoint: cotangent of the three live components of output jj, and the decoder-gradient terms of this trajectory summed over
// the warp into acc4[warp][0..3] = (dWd0, dWd1, dWd2, dbd)  (fixed shuffle tree: bit-reproducible)
__device__ __forceinline__ void tc_decode_vjp(const DecArgs& d, float g_reg, int jj, long long B, long long traj, bool valid, float v0, float v1,
                                              float v2, float (&gl)[3], float* acc4, int lane) {
    float gp = 0.f;
    unsigned m = 0u;
    if (valid) {
        gp = __ldg(d.g_pred + (long long)jj * B + traj);
        m = d.mask[(long long)jj * B + traj];
    }
    gl[0] = fmaf(g_reg, reg_sign(m, 0), __ldg(d.W) * gp);
    gl[1] = fmaf(g_reg, reg_sign(m, 1), __ldg(d.W + 1) * gp);
    gl[2] = fmaf(g_reg, reg_sign(m, 2), __ldg(d.W + 2) * gp);
    float p[4] = {gp * v0, gp * v1, gp * v2, gp};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p[i] += __shfl_xor_sync(0xffffffffu, p[i], o);
    }
    if (lane < 4) acc4[lane] += (lane == 0 ? p[0] : (lane == 1 ? p[1] : (lane == 2 ? p[2] : p[3])));
}

template <int R, int L, bool BAYES = false>
__global__ void __launch_bounds__(kTcThreads, 1) rk4_fwd_tc_kernel(const __grid_constant__ TcModel M, const __grid_constant__ TcTable tab, const TcFwdArgs a_in) {
    const TcFwdArgs& a = a_in;
    constexpr int D0 = R * L, DN = 3 * R;
    static_assert(D0 <= 16, "tensor path: R*L <= 16");
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* const smem = tc_smem;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(8) uint64_t mbar_b;
    __shared__ __align__(8) uint64_t mbar2;
    __shared__ __align__(16) TcHeadScratch head_scratch;
    __shared__ uint32_t tmem_slot;
    __shared__ double red[(kTcThreads / 32) * 8];

    if (threadIdx.x == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mbar, 1);
        mbar_init(&mbar_b, 1);
        mbar_init(&mbar2, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&head_scratch.bar[i >> 2][i & 3], kTcSlots > 1 ? kTcSlots - 1 : 1);
    }
    if (threadIdx.x < 32) tc_tmem_alloc<128>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t wphase = 0;
    const bool grouped = a.n_groups > 1;
    const long long tiles_per_group = (a.B + kTcTile - 1) / kTcTile;
    long long g_cur = grouped ? (long long)blockIdx.x / tiles_per_group : 0;       // group of this CTA's first tile
    tc_stage_blob(smem + M.off_w, a.blob + g_cur * a.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);

    TcCtx c;
    c.abuf_p = smem + M.off_abuf;
    c.dbuf_p = smem + M.off_dbuf;
    c.ghead_p = smem + M.off_ghead;
    c.abuf = smem_u32(smem + M.off_abuf);
    c.dbuf = smem_u32(smem + M.off_dbuf);
    c.wimg = smem_u32(smem + M.off_w);
    c.bias = reinterpret_cast<const float*>(smem + M.off_bias);
    c.tmem = tmem_slot;
    c.bar = &mbar;
    c.bar_b = &mbar_b;
    c.bar2 = &mbar2;
    c.abuf2_p = smem + M.off_abuf2;
    c.phase_b = 0;
    c.phase2 = 0;
    c.phase_h[0] = c.phase_h[1] = 0;
    c.stage_p = smem + M.off_stage;
    c.bz_z = nullptr;
    c.bz_slab = nullptr;
    c.sweep_par = 0;
    c.sweep_buf_bytes = (uint32_t)M.sweep_buf_bytes;
    c.smem_base16 = smem_u32(smem) >> 4;
    c.phase = 0;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform
    c.warp = warp;
    c.q = warp & 3; c.h = warp >> 2; c.lane = threadIdx.x & 31;
    c.t = c.q * 32 + c.lane;
    c.lane_addr = (uint32_t)(c.q * 32) << 16;
    const bool is_traj = warp < 4;
    tc_prof_init(c);
    const TcChain chain_r{c.abuf_p, c.bar, (uint32_t)M.tm_acc_fwd, kTcBarOperands};
    const TcChain chain_a{c.abuf2_p, c.bar2, (uint32_t)(M.tm_acc_fwd + M.tm_acc_stride), kTcBarOperands2};

    SideAcc side;
    side.clear();
    const long long B = a.B;
    const long long n_tiles = tiles_per_group * a.n_groups;
    const long long tape_stage = (long long)DN * B;

    if (warp >= kTcEpiWarps) {
        if (kTcRegHandOver) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcIssuerRegs));      // issuer + 3 idle warps of its warpgroup
        if (warp == kTcEpiWarps) {
        const long long n_my = n_tiles > (long long)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        tc_issuer_loop_fwd(M, tab, n_my, a.g.n_steps, c.smem_base16, c.tmem, c.bar, c.bar2);
        }
    } else {
    if (kTcRegHandOver) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcEpiRegs));
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long grp = grouped ? tile / tiles_per_group : 0;
        const long long b0 = (tile - grp * tiles_per_group) * kTcTile;
        // this group's views of the caller's arrays (group-major layouts, include/ude_b200.h); `a` shadows the launch arguments
        TcFwdArgs ag = a_in;
        if (grouped) {
            ag.y0 += grp * B * D0;
            ag.y_out += grp * (long long)a_in.g.n_out * B * D0;
            if (ag.tape) ag.tape += grp * a_in.tape_group_floats;
            if (grp != g_cur) {     // the issuer waits for this tile's first operands: no product reads the weights now
                asm volatile("bar.sync %0, %1;" ::"n"(kTcBarTile), "n"(32 * kTcEpiWarps) : "memory");
                tc_stage_blob(smem + M.off_w, a_in.blob + grp * a_in.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);
                g_cur = grp;
            }
        }
        const TcFwdArgs& a = ag;
        const bool valid = is_traj && (b0 + c.t < B);
        float x[D0], y[DN], k1[DN], k2[DN], k3[DN];
#pragma unroll
        for (int i = 0; i < D0; ++i) x[i] = 0.f;
        const bool decoded = a.dec.n_dec > 0;
        if (valid && a.rp.mean) {                    // fused reparam: this thread's row from (eps, mean, std)
            static_assert(R == 1, "tensor path: R = 1");
            const long long traj = b0 + c.t;
            float* dst = a.y_out + traj * (long long)D0;
            const float z0 = reparam_z(a.rp, traj, 0, 1, L - 1, 0), z1 = reparam_z(a.rp, traj, 0, 1, L - 1, 1);
#pragma unroll
            for (int l = 0; l < L; ++l) {
                x[l] = reparam_y(z0, z1, l >= 3 ? reparam_z(a.rp, traj, 0, 1, L - 1, l - 1) : 0.f, l);
                dst[l] = x[l];
            }
        } else if (valid) {
            const float* src = a.y0 + (b0 + c.t) * (long long)D0;
            float* dst = a.y_out + (b0 + c.t) * (long long)D0;          // output 0 = y0 (torchdiffeq: solution[0] = y0)
            if ((D0 & 3) == 0) {
#pragma unroll
                for (int i = 0; i < D0; i += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
                    x[i] = v.x; x[i + 1] = v.y; x[i + 2] = v.z; x[i + 3] = v.w;
                    if (!decoded) st4(dst + i, v);
                }
            } else {
#pragma unroll
                for (int i = 0; i < D0; ++i) { x[i] = __ldg(src + i); if (!decoded) dst[i] = x[i]; }
            }
        }
        if (valid && decoded) tc_decode_store(a.dec, 0, B, b0 + c.t, x[0], x[1], x[2], side);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            y[3 * r] = x[r * L]; y[3 * r + 1] = x[r * L + 1]; y[3 * r + 2] = x[r * L + 2];
        }
        int j = 1;
        for (int n = 0; n < a.g.n_steps; ++n) {
            const float dt = __ldg(a.g.dt + n);
#pragma unroll 1
            for (int stage = 0; stage < 4; ++stage) {
                if (BAYES) {        // this evaluation's weight set: every product and head loop of the previous one is consumed
                    asm volatile("bar.sync %0, %1;" ::"n"(kTcBarTile), "n"(32 * kTcEpiWarps) : "memory");
                    tc_stage_blob(smem + M.off_w, a_in.blob + (long long)(n * 4 + stage) * a_in.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);
                }
                float o_r[2 * R], o_a[3 * R];
#pragma unroll
                for (int i = 0; i < 2 * R; ++i) o_r[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 3 * R; ++i) o_a[i] = 0.f;
                tc_nets_forward<false, R, L>(M, chain_r, chain_a, c, &head_scratch, is_traj, x, o_r, o_a);
                if (is_traj) {
                    float f[DN];
                    tc_rhs_tail<R, L>(M, x, o_r, o_a, valid, side, f);
                    if (a.tape && valid) {
                        float* tp = a.tape + (long long)(n * 4 + stage) * tape_stage + b0 + c.t;
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) tp[(long long)(3 * r + cc) * B] = x[r * L + cc];   // stage input
                    }
                    float xn[DN];
#pragma unroll
                    for (int i = 0; i < DN; ++i) {
                        if (stage == 0) {
                            k1[i] = f[i];
                            xn[i] = __fadd_rn(y[i], __fmul_rn(__fmul_rn(dt, f[i]), 1.0f / 3.0f));
                        } else if (stage == 1) {
                            k2[i] = f[i];
                            xn[i] = __fadd_rn(y[i], __fmul_rn(dt, __fsub_rn(f[i], __fmul_rn(k1[i], 1.0f / 3.0f))));
                        } else if (stage == 2) {
                            k3[i] = f[i];
                            xn[i] = __fadd_rn(y[i], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[i], k2[i]), f[i])));
                        } else {
                            const float s = __fadd_rn(__fadd_rn(k1[i], __fmul_rn(3.0f, __fadd_rn(k2[i], k3[i]))), f[i]);
                            xn[i] = __fadd_rn(y[i], __fmul_rn(__fmul_rn(s, dt), 0.125f));
                        }
                    }
                    if (stage == 3) {
                        // outputs that fall into (t_n, t_{n+1}]  (FixedGridODESolver.integrate)
                        int jj = j;
                        while (jj < a.g.n_out && __ldg(a.g.out_step + jj) == n) {
                            if (valid) {
                                const int kind = __ldg(a.g.out_kind + jj);
                                const float w = __ldg(a.g.out_w + jj);
                                float row[D0];
#pragma unroll
                                for (int i = 0; i < D0; ++i) row[i] = x[i];
#pragma unroll
                                for (int r = 0; r < R; ++r)
#pragma unroll
                                    for (int cc = 0; cc < 3; ++cc) {
                                        const float yy = y[3 * r + cc], xx = xn[3 * r + cc];
                                        row[r * L + cc] = kind == 1 ? xx : (kind == 0 ? yy : __fadd_rn(yy, __fmul_rn(w, __fsub_rn(xx, yy))));
                                    }
                                float* dst = a.y_out + ((long long)jj * B + b0 + c.t) * (long long)D0;
                                if (decoded) {
                                    tc_decode_store(a.dec, jj, B, b0 + c.t, row[0], row[1], row[2], side);
                                } else if ((D0 & 3) == 0) {
#pragma unroll
                                    for (int i = 0; i < D0; i += 4) st4(dst + i, make_float4(row[i], row[i + 1], row[i + 2], row[i + 3]));
                                } else {
#pragma unroll
                                    for (int i = 0; i < D0; ++i) dst[i] = row[i];
                                }
                            }
                            ++jj;
                        }
                        j = jj;
                        if (decoded && valid && a.tape && n == a.g.n_steps - 1) {      // the extra tape record: the final state
#pragma unroll
                            for (int i = 0; i < DN; ++i) a.tape[(long long)(4 * a.g.n_steps) * tape_stage + (long long)i * B + b0 + c.t] = xn[i];
                        }
#pragma unroll
                        for (int i = 0; i < DN; ++i) y[i] = xn[i];
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        x[r * L] = xn[3 * r]; x[r * L + 1] = xn[3 * r + 1]; x[r * L + 2] = xn[3 * r + 2];
                    }
                }
            }
        }
        if (grouped && is_traj) {      // side sums per tile and trajectory warp (summed per group, in tile order, afterwards)
            double v[8] = {side.sb, side.sg, side.sb2, side.sg2, side.sf2, (double)side.rows, 0.0, side.reg};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
            }
            v[6] = 4.0 * a.g.n_steps;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c.lane == i) a.side_partial[(tile * 4 + c.q) * 8 + i] = v[i];
            side.clear();
        }
    }
    }
    tc_prof_dump(c);
    if (!grouped) tc_side_reduce(side, 4.0 * a.g.n_steps, red, a.side_partial + (long long)blockIdx.x * 8);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc_tmem_free<128>(c.tmem);
}

// ===================================================================================
// backward kernel (discrete adjoint)
// ===================================================================================
// Sweep one net backwards over the tile.  Pre: head cotangent in `g_head` (trajectory threads), saved activations
// in TMEM, x in registers.  Post: dx[0..D0) += W_0^T G_0 (trajectory threads), weight and bias gradients added to the
// CTA slab.
//
// The sweep is a dependency chain  G_l -> (delta products C_l) -> epilogue -> G_{l-1} -> ...  and every link costs a
// tensor-pipe round trip, so the two thin ends of the chain stay on the CUDA cores: the head's Wh^T g (NH <= 3
// outputs) and the first layer's W0^T G_0 (D0 <= 16 inputs) are FMA loops over fp32 weights in shared memory.
// Everything that is not on the chain runs behind it: the weight-gradient products W_l (own completion barrier, one
// shared accumulator guarded by the "drained" barrier), the activation image of the NEXT layer (two alternating
// sweep buffers), the ELU' factors, and the flush of the gradients into the per-CTA slab.
//
//   step H1        : G_{n-1} (registers) -> head image [G1|G2], image of a_{n-2};  ready -> issuer: W_{n-1}
//   step H2        : (behind W_{n-1}) G_{n-2} = (Wh^T g) * ELU'(a_{n-2}) -> Dbuf, image of a_{n-3};  ready -> issuer:
//                    C_{n-2}, then W_{n-2} once the head's gradients are flushed out of the shared accumulator
//   step l=n-2..1  : image for W_{l-1};  ELU'(a_{l-1});  wait C_l;  G_{l-1} -> registers (l = 1: partial W0^T G_0);
//                    wait W_l;  G_{l-1} -> Dbuf;  ready -> issuer: C_{l-1} (l >= 2), W_{l-1};  flush W_l
//   step 0         : wait W_0, flush;  partial dx of the helper warps -> trajectory warps (through TMEM)
// Gradients of one layer, TMEM -> per-warp staging slice -> bulk reductions into the CTA slab
// (cp.reduce.async.bulk.add.f32, SASS UBLKRED): the L2 atomics run on the copy engine in the background, and because
// every warp stages and hands over its OWN rows (16q .. 16q+15 of the M = 64 accumulator live in lanes 0..15 of
// quarter q) and column chunks, the flush needs no CTA-wide barrier at all.
// Slab layout of a layer: [quarter q][chunk][4][16 rows] float4 (one contiguous KB per warp and chunk), then [64] bias sums.
// Returns after the accumulator has been read; the caller releases it to the issuer ("drained").
constexpr int kTcStageSlice = kTcChunksPerWarp * 1024 + 64;      // bytes per warp: its chunks + 16 bias sums
__device__ __forceinline__ void tc_flush_grad(const TcModel& M, const TcLayer& Lp, TcCtx& c, float* slab) {
    if (16 * c.q >= Lp.out) return;              // no real output row in this lane quarter (heads, 32-wide layers)
    const bool act = c.lane < 16;
    const uint32_t q0 = c.tmem + c.lane_addr + M.tm_q;
    uint8_t* slice = c.stage_p + (size_t)c.warp * kTcStageSlice;
    // this warp's previous bulk reductions must have read the slice -- and must have LANDED: reductions of consecutive
    // evaluations hit the same slab addresses, and only completion order makes the fp32 sums bit-reproducible
    // (they were issued a whole phase ago: the wait is short)
    if (c.lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kTcChunksPerWarp; ++j) {
        const int ch = c.h + kTcSlots * j;
        if (ch < (Lp.in_p >> 4)) {
            float u[16], v[16];
            tc_ld16x2(q0 + 16 * ch, q0 + Lp.in_p + 16 * ch, u, v);
            if (act) {
                float4* st = reinterpret_cast<float4*>(slice + j * 1024);
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                {
                    float4 t;
                    add2(t.x, t.y, u[i], u[i + 1], v[i], v[i + 1]);
                    add2(t.z, t.w, u[i + 2], u[i + 3], v[i + 2], v[i + 3]);
                    st[(i >> 2) * 16 + c.lane] = t;
                }
            }
        }
    }
    if (c.h == 0) {
        float u[16];
        tc_ld16(q0 + 2 * Lp.in_p, u);            // 8 columns of the ones group (+8 unused columns of the region)
        if (act) reinterpret_cast<float*>(slice + kTcChunksPerWarp * 1024)[c.lane] = u[0];
    }
}
// second half: hand the staged gradients to the copy engine (call after tc_drained)
template <bool BAYES>
__device__ __forceinline__ void tc_flush_commit(const TcLayer& Lp, TcCtx& c, float* slab) {
    if (16 * c.q >= Lp.out) return;
    tc_proxy_fence();
    __syncwarp();
    if (c.lane == 0) {
        const int n_chunks = Lp.in_p >> 4;
        float* base = slab + Lp.slab_off;
        const uint32_t slice = smem_u32(c.stage_p + (size_t)c.warp * kTcStageSlice);
#pragma unroll
        for (int j = 0; j < kTcChunksPerWarp; ++j) {
            const int ch = c.h + kTcSlots * j;
            if (ch < n_chunks)
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                                 base + (size_t)(c.q * n_chunks + ch) * 256),
                             "r"(slice + j * 1024), "n"(1024)
                             : "memory");
        }
        if (c.h == 0)
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                             base + 64 * Lp.in_p + 16 * c.q),
                         "r"(slice + kTcChunksPerWarp * 1024), "n"(64)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (BAYES) {
        // Bayes funcs: the same staged gradients, multiplied by this evaluation's noise, reduced into the second slab.  The first
        // reductions must have READ the slice before it is scaled in place.
        if (c.lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        const int n_chunks = Lp.in_p >> 4;
        uint8_t* slice_p = c.stage_p + (size_t)c.warp * kTcStageSlice;
        if (c.lane < 16) {
#pragma unroll
            for (int j = 0; j < kTcChunksPerWarp; ++j) {
                const int ch = c.h + kTcSlots * j;
                if (ch < n_chunks) {
                    float4* st = reinterpret_cast<float4*>(slice_p + j * 1024);
                    const float4* zz = reinterpret_cast<const float4*>(c.bz_z + Lp.slab_off + (size_t)(c.q * n_chunks + ch) * 256);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float4 v = st[i * 16 + c.lane];
                        const float4 z = __ldg(zz + i * 16 + c.lane);
                        v.x *= z.x; v.y *= z.y; v.z *= z.z; v.w *= z.w;
                        st[i * 16 + c.lane] = v;
                    }
                }
            }
            if (c.h == 0)
                reinterpret_cast<float*>(slice_p + kTcChunksPerWarp * 1024)[c.lane] *= __ldg(c.bz_z + Lp.slab_off + 64 * Lp.in_p + 16 * c.q + c.lane);
        }
        tc_proxy_fence();
        __syncwarp();
        if (c.lane == 0) {
            float* base = c.bz_slab + Lp.slab_off;
            const uint32_t slice = smem_u32(slice_p);
#pragma unroll
            for (int j = 0; j < kTcChunksPerWarp; ++j) {
                const int ch = c.h + kTcSlots * j;
                if (ch < n_chunks)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                                     base + (size_t)(c.q * n_chunks + ch) * 256),
                                 "r"(slice + j * 1024), "n"(1024)
                                 : "memory");
            }
            if (c.h == 0)
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                                 base + 64 * Lp.in_p + 16 * c.q),
                             "r"(slice + kTcChunksPerWarp * 1024), "n"(64)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
}
__device__ __forceinline__ void tc_drained(TcCtx& c) {
    tc_fence_before();
    asm volatile("bar.arrive %0, %1;" ::"n"(kTcBarDrained), "n"(kTcBarThreads) : "memory");
}
// image [a2 | a1 | ones] of the INPUT of layer li (a_{li-1}, or x for the first layer) into sweep buffer `buf`
template <int D0>
__device__ __forceinline__ void tc_sweep_image(const TcLayer* Ls, int li, uint32_t buf, TcCtx& c, const float (&x)[D0], bool is_traj) {
    if (li >= 1) {
        const int w = Ls[li].in_p;
        for (int ch = c.h; ch < (w >> 4); ch += kTcSlots) {
            float v[16];
            tc_ld16(c.tmem + c.lane_addr + Ls[li - 1].save_col + 16 * ch, v);
            tc_store_sweep16(c, buf, v, ch, w);
        }
        if (c.h == 0) tc_store_sweep_ones(c, buf, w);
    } else if (is_traj) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = i < D0 ? x[i] : 0.f;
        tc_store_sweep16(c, buf, v, 0, 16);
        tc_store_sweep_ones(c, buf, 16);
    }
}
template <int D0, int NH, bool BAYES>
__device__ __forceinline__ void tc_net_backward(const TcModel& M, const TcLayer* Ls, int n, int chain, TcCtx& c,
                                                const float (&g_head)[NH], const float (&x)[D0], float* slab,
                                                float (&dx)[D0], bool is_traj) {
    const uint32_t tm_x = c.tmem + c.lane_addr + M.tm_dd;      // exchange columns between the warps of a lane quarter
    // sync/sync exchanges of a lane quarter's warps; NOT the forward heads' arrive/sync barriers: a helper warp that
    // arrived there could reach this sync before its trajectory warp consumed the arrival
    const int bar_q = kTcBarSweep + c.q;
    (void)chain;
    uint32_t par = c.sweep_par;
    float pdx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pdx[i] = 0.f;
    // ---------------- step H
    {
        TCP_T0();
        const TcLayer& Lh = Ls[n - 1];
        const TcLayer& Lq = Ls[n - 2];
        float gh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) gh[i] = 0.f;
        if (is_traj) {
#pragma unroll
            for (int i = 0; i < NH; ++i) gh[i] = g_head[i];
        }
        if (is_traj) {
            static_assert(NH <= 8, "head cotangent image: one 8-feature group per split array");
            uint8_t* base = c.ghead_p + c.t * 16;              // [G1 | G2], ONE 8-feature group each
            split2_store8(gh, base, base + kTcGroupBytes);
            if (kTcSlots > 1) {
                tc_st16(tm_x, gh);
                tc_fence_before();
            }
        }
        if (kTcSlots > 1) {
            asm volatile("bar.sync %0, %1;" ::"r"(bar_q), "n"(32 * kTcSlots) : "memory");
            if (!is_traj) {
                tc_fence_after();
                tc_ld16(tm_x, gh);
            }
        }
        // H1: operands of the head's weight-gradient products -> the issuer starts W_{n-1} right away
        for (int ch = c.h; ch < (Lh.in_p >> 4); ch += kTcSlots) {
            float sv[16];
            tc_ld16(c.tmem + c.lane_addr + Lq.save_col + 16 * ch, sv);
            tc_store_sweep16(c, par, sv, ch, Lh.in_p);
        }
        if (c.h == 0) tc_store_sweep_ones(c, par, Lh.in_p);
        // NOT the barrier of H2's hand-off: the arrivals are non-blocking and nothing blocks between H1 and H2, so a warp
        // that runs ahead (the helper warps skip the tail VJP) would otherwise arrive twice in one barrier generation and
        // release the issuer before a slow warp has written its part of the operands
        tc_operands_ready(c, kTcBarOperands2);
        // H2 (behind W_{n-1}): G_{n-2} = (Wh^T g) * ELU'(a_{n-2}) on the FMA pipe, image of a_{n-3}
        const float* hw = c.bias + Lh.hw_off;
        for (int ch = c.h; ch < (Lh.in_p >> 4); ch += kTcSlots) {
            float d[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) d[i] = 0.f;
#pragma unroll
            for (int o = 0; o < NH; ++o) {
                const float4* w4 = reinterpret_cast<const float4*>(hw + o * Lh.in_p + 16 * ch);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 w = w4[i];
                    fma2(d[4 * i + 0], d[4 * i + 1], gh[o], gh[o], w.x, w.y, d[4 * i + 0], d[4 * i + 1]);
                    fma2(d[4 * i + 2], d[4 * i + 3], gh[o], gh[o], w.z, w.w, d[4 * i + 2], d[4 * i + 3]);
                }
            }
            if (Lq.elu) {
                float sv[16];
                tc_ld16(c.tmem + c.lane_addr + Lq.save_col + 16 * ch, sv);
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float e0, e1;
                    elu_grad2(e0, e1, sv[i], sv[i + 1]);
                    mul2(d[i], d[i + 1], d[i], d[i + 1], e0, e1);
                }
            }
            tc_store_delta16(c, d, ch, Lh.in_p);
        }
        tc_sweep_image<D0>(Ls, n - 2, par ^ 1u, c, x, is_traj);
        tc_operands_ready(c);
        TCP_ADD(c, TCP_ISSUE);
        // the head's gradients leave the shared accumulator (W_{n-1} ran behind H2), W_{n-2} may follow
        tc_wait_mma_b(c);
        TCP_T1();
        tc_flush_grad(M, Lh, c, slab);
        tc_drained(c);
        tc_flush_commit<BAYES>(Lh, c, slab);
        TCP_ADD1(c, TCP_TAIL);
    }
    // ---------------- steps l = n-2 .. 1
    for (int l = n - 2; l >= 1; --l) {
        const TcLayer& Lp = Ls[l];
        const TcLayer& Lq = Ls[l - 1];
        par ^= 1u;                              // W_l reads sweep buffer `par`
        uint4 sg[kTcChunksPerWarp][6];
        {
            TCP_T0();
            tc_sweep_image<D0>(Ls, l - 1, par ^ 1u, c, x, is_traj);
            TCP_ADD(c, TCP_ISSUE);
        }
        tc_wait_mma(c);                         // C_l
        {
            TCP_T0();
            const uint32_t dd = c.tmem + c.lane_addr + M.tm_dd;
#pragma unroll
            for (int j = 0; j < kTcChunksPerWarp; ++j) {
                const int ch = c.h + kTcSlots * j;
                if (ch < (Lp.in_p >> 4)) {
                    float d[16];
                    if (Lq.elu) {
                        float sv[16];
                        tc_ld16x2(dd + 16 * ch, c.tmem + c.lane_addr + Lq.save_col + 16 * ch, d, sv);
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            float e0, e1;
                            elu_grad2(e0, e1, sv[i], sv[i + 1]);
                            mul2(d[i], d[i + 1], d[i], d[i + 1], e0, e1);
                        }
                    } else {
                        tc_ld16(dd + 16 * ch, d);
                    }
                    if (l == 1) {               // W0^T G_0 on the FMA pipe (fp32 rows of W0, broadcast loads)
                        const float* w0 = c.bias + Lq.w0_off + (size_t)(16 * ch) * M.w0_stride;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
#pragma unroll
                            for (int k4 = 0; k4 < (D0 + 3) / 4; ++k4) {
                                const float4 w = *reinterpret_cast<const float4*>(w0 + i * M.w0_stride + 4 * k4);
                                fma2(pdx[4 * k4 + 0], pdx[4 * k4 + 1], d[i], d[i], w.x, w.y, pdx[4 * k4 + 0], pdx[4 * k4 + 1]);
                                fma2(pdx[4 * k4 + 2], pdx[4 * k4 + 3], d[i], d[i], w.z, w.w, pdx[4 * k4 + 2], pdx[4 * k4 + 3]);
                            }
                        }
                    }
                    split3_pack16(d, sg[j]);
                }
            }
            TCP_ADD(c, TCP_EPI_BWD);
        }
        tc_wait_mma_b(c);                       // W_l: Dbuf is free again, its gradients are complete
        {
            TCP_T0();
#pragma unroll
            for (int j = 0; j < kTcChunksPerWarp; ++j) {
                const int ch = c.h + kTcSlots * j;
                if (ch < (Lp.in_p >> 4)) tc_store_pack_delta(c, sg[j], ch, Lp.in_p);
            }
            tc_operands_ready(c);
            TCP_ADD(c, TCP_EPI_BWD);
        }
        TCP_T1();
        tc_flush_grad(M, Lp, c, slab);
        tc_drained(c);
        tc_flush_commit<BAYES>(Lp, c, slab);
        TCP_ADD1(c, TCP_TAIL);
    }
    // ---------------- step 0
    par ^= 1u;
    tc_wait_mma_b(c);
    {
        TCP_T1();
        tc_flush_grad(M, Ls[0], c, slab);
        tc_flush_commit<BAYES>(Ls[0], c, slab);
        if (kTcSlots > 1) {
            if (!is_traj) {
                tc_st16(tm_x + 16 * (c.h - 1), pdx);
                tc_fence_before();
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_q), "n"(32 * kTcSlots) : "memory");
            if (is_traj) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k + 1 < kTcSlots; ++k) {
                    float o[16];
                    tc_ld16(tm_x + 16 * k, o);
#pragma unroll
                    for (int i = 0; i < D0; ++i) pdx[i] += o[i];
                }
            }
        }
        if (is_traj) {
#pragma unroll
            for (int i = 0; i < D0; ++i) dx[i] += pdx[i];
        }
        TCP_ADD1(c, TCP_TAIL);
    }
    c.sweep_par = par ^ 1u;
}

template <int R, int L, bool BAYES = false>
__global__ void __launch_bounds__(kTcThreads, 1) rk4_bwd_tc_kernel(const __grid_constant__ TcModel M, const __grid_constant__ TcTable tab, const TcBwdArgs a_in) {
    const TcBwdArgs& a = a_in;
    constexpr int D0 = R * L, DN = 3 * R;
    static_assert(D0 <= 16, "tensor path: R*L <= 16");
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    uint8_t* const smem = tc_smem;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(8) uint64_t mbar_b;
    __shared__ __align__(8) uint64_t mbar2;
    __shared__ __align__(16) TcHeadScratch head_scratch;
    __shared__ uint32_t tmem_slot;
    __shared__ float dec_acc[4][4];         // decoded mode: per trajectory warp (dWd0, dWd1, dWd2, dbd)

    if (threadIdx.x == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mbar, 1);
        mbar_init(&mbar_b, 1);
        mbar_init(&mbar2, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&head_scratch.bar[i >> 2][i & 3], kTcSlots > 1 ? kTcSlots - 1 : 1);
    }
    if (threadIdx.x < 16) dec_acc[threadIdx.x >> 2][threadIdx.x & 3] = 0.f;
    if (threadIdx.x < 32) tc_tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t wphase = 0;
    const bool grouped = a.n_groups > 1;
    const long long tiles_per_group = (a.B + kTcTile - 1) / kTcTile;
    long long g_cur = grouped ? (long long)blockIdx.x / tiles_per_group : 0;       // group of this CTA's first tile
    tc_stage_blob(smem + M.off_w, a.blob + g_cur * a.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);

    TcCtx c;
    c.abuf_p = smem + M.off_abuf;
    c.dbuf_p = smem + M.off_dbuf;
    c.ghead_p = smem + M.off_ghead;
    c.abuf = smem_u32(smem + M.off_abuf);
    c.dbuf = smem_u32(smem + M.off_dbuf);
    c.wimg = smem_u32(smem + M.off_w);
    c.bias = reinterpret_cast<const float*>(smem + M.off_bias);
    c.tmem = tmem_slot;
    c.bar = &mbar;
    c.bar_b = &mbar_b;
    c.bar2 = &mbar2;
    c.abuf2_p = smem + M.off_abuf2;
    c.phase_b = 0;
    c.phase2 = 0;
    c.phase_h[0] = c.phase_h[1] = 0;
    c.stage_p = smem + M.off_stage;
    c.bz_z = nullptr;
    c.bz_slab = nullptr;
    c.sweep_par = 0;
    c.sweep_buf_bytes = (uint32_t)M.sweep_buf_bytes;
    c.smem_base16 = smem_u32(smem) >> 4;
    c.phase = 0;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform
    c.warp = warp;
    c.q = warp & 3; c.h = warp >> 2; c.lane = threadIdx.x & 31;
    c.t = c.q * 32 + c.lane;
    c.lane_addr = (uint32_t)(c.q * 32) << 16;
    const bool is_traj = warp < 4;
    tc_prof_init(c);
    const TcChain chain_r{c.abuf_p, c.bar, (uint32_t)M.tm_fwd, kTcBarOperands};
    const TcChain chain_a{c.abuf2_p, c.bar2, (uint32_t)(M.tm_fwd + M.tm_acc_stride), kTcBarOperands2};

    // the operand buffers feed MN-major (K = trajectory) products: start from finite contents
    for (int i = threadIdx.x; i < (M.off_abuf - M.off_stage + M.abuf_bytes + M.dbuf_bytes) / 16; i += kTcThreads)
        reinterpret_cast<uint4*>(smem + M.off_stage)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    float gs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.g_side) {
#pragma unroll
        for (int i = 0; i < 5; ++i) gs[i] = __ldg(a.g_side + i);
    }
    float* slab = a.slabs + (long long)blockIdx.x * M.slab_floats;
    if (BAYES) c.bz_slab = a.bz_slabs + (long long)blockIdx.x * M.slab_floats;
    const bool decoded = a.dec.n_dec > 0;
    const float g_reg = (decoded && a.g_side) ? __ldg(a.g_side + 5) : 0.f;

    const long long B = a.B;
    const long long n_tiles = tiles_per_group * a.n_groups;
    const long long tape_stage = (long long)DN * B;
    const int n_steps = a.g.n_steps;

    if (warp >= kTcEpiWarps) {
        if (kTcRegHandOver) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTcIssuerRegs));      // issuer + 3 idle warps of its warpgroup
        if (warp == kTcEpiWarps) {
        const long long n_my = n_tiles > (long long)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        tc_issuer_loop_bwd(M, tab, n_my, n_steps, c.smem_base16, c.tmem, c.bar, c.bar2, c.bar_b);
        }
    } else {
    if (kTcRegHandOver) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTcEpiRegs));
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long grp = grouped ? tile / tiles_per_group : 0;
        const long long b0 = (tile - grp * tiles_per_group) * kTcTile;
        // this group's views of the caller's arrays; per-tile gradient slab (summed per group, in tile order, afterwards)
        TcBwdArgs ag = a_in;
        if (grouped) {
            ag.y0 += grp * B * D0;
            ag.g_y0 += grp * B * D0;
            ag.g_out += grp * (long long)a_in.g.n_out * B * D0;
            ag.tape += grp * a_in.tape_group_floats;
            slab = a_in.slabs + tile * M.slab_floats;
            if (a_in.g_side) {
#pragma unroll
                for (int i = 0; i < 5; ++i) gs[i] = __ldg(a_in.g_side + grp * 5 + i);
            }
            if (grp != g_cur) {     // the issuer waits for this tile's first operands: no product reads the weights now
                asm volatile("bar.sync %0, %1;" ::"n"(kTcBarTile), "n"(32 * kTcEpiWarps) : "memory");
                tc_stage_blob(smem + M.off_w, a_in.blob + grp * a_in.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);
                g_cur = grp;
            }
        }
        const TcBwdArgs& a = ag;
        const bool valid = is_traj && (b0 + c.t < B);
        float x[D0], lam[DN], cb[D0], kb[4][DN];
#pragma unroll
        for (int i = 0; i < D0; ++i) { x[i] = 0.f; cb[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < DN; ++i) lam[i] = 0.f;
        if (valid) {
            const float* src = a.y0 + (b0 + c.t) * (long long)D0;      // constants of x; dynamic comps come from the tape
#pragma unroll
            for (int i = 0; i < D0; ++i) x[i] = __ldg(src + i);
            if (decoded) {      // live part := the final state (the tape's extra record): the outputs of the last step need it
#pragma unroll
                for (int i = 0; i < DN; ++i) x[i] = __ldg(a.tape + (long long)(4 * n_steps) * tape_stage + (long long)i * B + b0 + c.t);
            }
        }
        float xt[DN];       // tape prefetch: stage input of the NEXT evaluation to be reversed
#pragma unroll
        for (int i = 0; i < DN; ++i)
            xt[i] = (valid && n_steps > 0) ? __ldg(a.tape + (long long)(n_steps * 4 - 1) * tape_stage + (long long)i * B + b0 + c.t) : 0.f;
        int j = a.g.n_out - 1;
        for (int n = n_steps - 1; n >= 0; --n) {
            const float dt = __ldg(a.g.dt + n);
            // ---- outputs tied to step n feed lambda_{n+1} (and, for copies of y_n / lerps, lambda_n)
            {
                float pend[DN];
#pragma unroll
                for (int i = 0; i < DN; ++i) pend[i] = 0.f;
                int jj = j;
                while (jj >= 1 && __ldg(a.g.out_step + jj) == n) {
                    if (decoded) {
                        if (is_traj) {      // whole warps (shuffle sums inside); x = state after the step, y_n from the tape
                            const int kind = __ldg(a.g.out_kind + jj);
                            const float w = __ldg(a.g.out_w + jj);
                            float v[3] = {x[0], x[1], x[2]}, gl[3];
                            if (kind != 1 && valid) {
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    const float yy = __ldg(a.tape + (long long)(n * 4) * tape_stage + (long long)i * B + b0 + c.t);
                                    v[i] = kind == 0 ? yy : __fadd_rn(yy, __fmul_rn(w, __fsub_rn(v[i], yy)));
                                }
                            }
                            tc_decode_vjp(a.dec, g_reg, jj, B, b0 + c.t, valid, v[0], v[1], v[2], gl, dec_acc[warp], c.lane);
#pragma unroll
                            for (int l = 0; l < 3; ++l) {
                                const float g = gl[l];
                                if (kind == 1) lam[l] += g;
                                else if (kind == 0) pend[l] += g;
                                else { const float sg = w * g; lam[l] += sg; pend[l] += g - sg; }
                            }
                        }
                    } else if (valid) {
                        const int kind = __ldg(a.g.out_kind + jj);
                        const float w = __ldg(a.g.out_w + jj);
                        const float* gp = a.g_out + ((long long)jj * B + b0 + c.t) * (long long)D0;
                        float gr[D0];
                        if ((D0 & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < D0; i += 4) {
                                const float4 v = __ldg(reinterpret_cast<const float4*>(gp + i));
                                gr[i] = v.x; gr[i + 1] = v.y; gr[i + 2] = v.z; gr[i + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < D0; ++i) gr[i] = __ldg(gp + i);
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r)
#pragma unroll
                            for (int l = 0; l < L; ++l) {
                                const float g = gr[r * L + l];
                                if (l < 3) {
                                    if (kind == 1) lam[3 * r + l] += g;
                                    else if (kind == 0) pend[3 * r + l] += g;
                                    else { const float sg = w * g; lam[3 * r + l] += sg; pend[3 * r + l] += g - sg; }
                                } else {
                                    cb[r * L + l] += g;
                                }
                            }
                    }
                    --jj;
                }
                j = jj;
                const float c8 = dt * 0.125f;
#pragma unroll
                for (int i = 0; i < DN; ++i) {
                    const float base = lam[i] * c8;
                    kb[0][i] = base; kb[1][i] = 3.f * base; kb[2][i] = 3.f * base; kb[3][i] = base;
                    lam[i] = lam[i] + pend[i];        // becomes lambda_n once the four stage VJPs are added
                }
            }
#pragma unroll 1
            for (int stage = 3; stage >= 0; --stage) {
                if (BAYES) {        // this evaluation's weight set and noise: every product of the previous one is consumed
                    asm volatile("bar.sync %0, %1;" ::"n"(kTcBarTile), "n"(32 * kTcEpiWarps) : "memory");
                    tc_stage_blob(smem + M.off_w, a_in.blob + (long long)(n * 4 + stage) * a_in.blob_stride, (uint32_t)M.blob_bytes, &wbar, wphase);
                    c.bz_z = a.bayes_z + (size_t)(n * 4 + stage) * M.slab_floats;
                }
                // ---- stage input from the tape (loaded one evaluation ahead: the latency hides behind the sweep)
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) x[r * L + cc] = xt[3 * r + cc];
                {
                    const int e_next = n * 4 + stage - 1;
                    if (valid && e_next >= 0) {
#pragma unroll
                        for (int i = 0; i < DN; ++i) xt[i] = __ldg(a.tape + (long long)e_next * tape_stage + (long long)i * B + b0 + c.t);
                    }
                }
                // ---- recompute the stage activations
                float o_r[2 * R], o_a[3 * R];
#pragma unroll
                for (int i = 0; i < 2 * R; ++i) o_r[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 3 * R; ++i) o_a[i] = 0.f;
                tc_nets_forward<true, R, L>(M, chain_r, chain_a, c, &head_scratch, is_traj, x, o_r, o_a);
                // ---- tail VJP
                float d_r[2 * R], d_a[3 * R], xb[DN], dx[D0];
#pragma unroll
                for (int i = 0; i < 2 * R; ++i) d_r[i] = 0.f;
#pragma unroll
                for (int i = 0; i < 3 * R; ++i) d_a[i] = 0.f;
#pragma unroll
                for (int i = 0; i < DN; ++i) xb[i] = 0.f;
#pragma unroll
                for (int i = 0; i < D0; ++i) dx[i] = 0.f;
                if (is_traj) {
                    float g[DN];
#pragma unroll
                    for (int i = 0; i < DN; ++i) g[i] = stage == 3 ? kb[3][i] : (stage == 2 ? kb[2][i] : (stage == 1 ? kb[1][i] : kb[0][i]));
                    tc_rhs_tail_vjp<R, L>(M, x, o_r, o_a, valid, gs, g, d_r, d_a, xb);
                }
                if (M.n_rate > 0) tc_net_backward<D0, 2 * R, BAYES>(M, M.rate, M.n_rate, 0, c, d_r, x, slab, dx, is_traj);
                if (M.n_aug > 0) tc_net_backward<D0, 3 * R, BAYES>(M, M.aug, M.n_aug, 1, c, d_a, x, slab, dx, is_traj);
                // ---- combine: xbar = direct + W0^T delta0 (both nets); propagate through the tableau
                if (is_traj) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            const int i = 3 * r + cc;
                            const float xbar = xb[i] + dx[r * L + cc];
                            const float u = xbar * dt;
                            if (stage == 3) {            // x4 = y + dt*((k1 - k2) + k3)
                                kb[2][i] += u; kb[0][i] += u; kb[1][i] -= u;
                            } else if (stage == 2) {     // x3 = y + dt*(k2 - k1/3)
                                kb[1][i] += u; kb[0][i] -= u * (1.0f / 3.0f);
                            } else if (stage == 1) {     // x2 = y + (dt*k1)/3
                                kb[0][i] += (xbar * (1.0f / 3.0f)) * dt;
                            }
                            lam[i] += xbar;
                        }
#pragma unroll
                        for (int l = 3; l < L; ++l) cb[r * L + l] += dx[r * L + l];
                    }
                }
            }
        }
        // ---- output 0 (= y0 itself) and the final store of dL/dy0
        float g0dec[3] = {0.f, 0.f, 0.f};
        if (decoded && is_traj) tc_decode_vjp(a.dec, g_reg, 0, B, b0 + c.t, valid, x[0], x[1], x[2], g0dec, dec_acc[warp], c.lane);   // x = y0
        if (valid) {
            const float* gp = a.g_out + (b0 + c.t) * (long long)D0;
            float* dst = a.g_y0 + (b0 + c.t) * (long long)D0;
            float gy[D0];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int l = 0; l < L; ++l) {
                    const float g = decoded ? (l < 3 ? g0dec[l] : 0.f) : __ldg(gp + r * L + l);
                    gy[r * L + l] = (l < 3 ? lam[3 * r + l] : cb[r * L + l]) + g;
                    dst[r * L + l] = gy[r * L + l];
                }
            if (a.rp.mean) {    // fused reparam: dL/dz row
                const long long traj = b0 + c.t;
                float* gz = a.rp.g_z + traj * (long long)(L - 1);
                const float z0 = reparam_z(a.rp, traj, 0, 1, L - 1, 0), z1 = reparam_z(a.rp, traj, 0, 1, L - 1, 1);
#pragma unroll
                for (int k = 0; k < L - 1; ++k) gz[k] = reparam_gz(z0, z1, gy[0], gy[1], gy[2], k >= 2 ? gy[k + 1] : 0.f, k);
            }
        }
    }
    }
    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // gradient reductions have landed
    tc_prof_dump(c);
    tc_fence_before();
    __syncthreads();
    if (decoded && threadIdx.x < 4)
        a.dec.partial[(long long)blockIdx.x * 4 + threadIdx.x] =
            ((dec_acc[0][threadIdx.x] + dec_acc[1][threadIdx.x]) + dec_acc[2][threadIdx.x]) + dec_acc[3][threadIdx.x];
    if (threadIdx.x < 32) tc_tmem_free<512>(c.tmem);
}

// ===================================================================================
// utility kernels of the tensor path
// ===================================================================================
// theta (PyTorch layout) -> weight images (bf16 x3, zero padded) + bias block.  One block per layer.
__global__ void tc_pack_kernel(const __grid_constant__ TcModel M, const float* __restrict__ theta_all, uint8_t* __restrict__ blob_all,
                               long long blob_stride) {
    const float* __restrict__ theta = theta_all + (size_t)blockIdx.y * M.n_theta;       // weight set blockIdx.y
    uint8_t* __restrict__ blob = blob_all + (size_t)blockIdx.y * blob_stride;
    const int li = blockIdx.x;
    const TcLayer& Lp = li < M.n_rate ? M.rate[li] : M.aug[li - M.n_rate];
    uint16_t* img = reinterpret_cast<uint16_t*>(blob + Lp.w_off);
    const int total = Lp.in_p * Lp.out_p;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int n = idx / Lp.in_p, k = idx - n * Lp.in_p;
        const float w = (n < Lp.out && k < Lp.in) ? theta[Lp.th_w + n * Lp.in + k] : 0.f;
        float r = w;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(r);
            r -= __bfloat162float(hb);
            // [k/8][s][n][k%8]
            img[((size_t)(k >> 3) * 3 + s) * Lp.out_p * 8 + (size_t)n * 8 + (k & 7)] = *reinterpret_cast<const uint16_t*>(&hb);
        }
    }
    float* bias = reinterpret_cast<float*>(blob + M.w_bytes) + Lp.b_off;
    for (int n = threadIdx.x; n < Lp.out_p; n += blockDim.x) bias[n] = n < Lp.out ? theta[Lp.th_b + n] : 0.f;
    if (Lp.w0_off >= 0) {
        float* w0 = reinterpret_cast<float*>(blob + M.w_bytes) + Lp.w0_off;
        for (int idx = threadIdx.x; idx < Lp.out_p * M.w0_stride; idx += blockDim.x) {
            const int n = idx / M.w0_stride, k = idx - n * M.w0_stride;
            w0[idx] = (n < Lp.out && k < Lp.in) ? theta[Lp.th_w + n * Lp.in + k] : 0.f;
        }
    }
    if (Lp.hw_off >= 0) {
        float* hw = reinterpret_cast<float*>(blob + M.w_bytes) + Lp.hw_off;
        for (int idx = threadIdx.x; idx < Lp.out * Lp.in_p; idx += blockDim.x) {
            const int n = idx / Lp.in_p, k = idx - n * Lp.in_p;
            hw[idx] = k < Lp.in ? theta[Lp.th_w + n * Lp.in + k] : 0.f;
        }
    }
}

// side[0..7] = sum over CTA partials (fixed order)
__global__ void tc_reduce_side_kernel(const double* __restrict__ partial_all, int n_parts, double* __restrict__ side_all) {
    const double* __restrict__ partial = partial_all + (size_t)blockIdx.x * n_parts * 8;      // weight set blockIdx.x
    double* __restrict__ side = side_all + (size_t)blockIdx.x * 8;
    const int i = threadIdx.x;
    if (i >= 8) return;
    double s = 0.0;
    for (int cidx = 0; cidx < n_parts; ++cidx) s += partial[cidx * 8 + i];
    if (i == 6) s = partial[6];   // RHS evaluations per trajectory: identical in every CTA
    side[i] = s;
}

// Bayes funcs: the noise of every evaluation from theta layout into SLAB layout (same index map as tc_reduce_dtheta_kernel);
// z_slab must be zero-filled beforehand (padding entries).  grid (ceil(n_theta / 256), n_evals)
__global__ void tc_pack_z_kernel(const __grid_constant__ TcModel M, const float* __restrict__ z_all, float* __restrict__ z_slab_all) {
    const float* __restrict__ z = z_all + (size_t)blockIdx.y * M.n_theta;
    float* __restrict__ zs = z_slab_all + (size_t)blockIdx.y * M.slab_floats;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M.n_theta) return;
    const int nl = M.n_rate + M.n_aug;
    for (int li = 0; li < nl; ++li) {
        const TcLayer& Lp = li < M.n_rate ? M.rate[li] : M.aug[li - M.n_rate];
        if (e >= Lp.th_w && e < Lp.th_w + Lp.in * Lp.out) {
            const int k = e - Lp.th_w;
            const int row = k / Lp.in, col = k - row * Lp.in;
            zs[Lp.slab_off + (((((row >> 4) * (Lp.in_p >> 4) + (col >> 4)) * 4 + ((col & 15) >> 2)) * 16 + (row & 15)) * 4 + (col & 3))] = z[e];
            return;
        }
        if (e >= Lp.th_b && e < Lp.th_b + Lp.out) {
            zs[Lp.slab_off + 64 * Lp.in_p + (e - Lp.th_b)] = z[e];
            return;
        }
    }
}

// g_theta[e] = sum over CTA slabs (fixed order => deterministic)
__global__ void tc_reduce_dtheta_kernel(const __grid_constant__ TcModel M, const float* __restrict__ slabs_all, int n_slabs,
                                        float* __restrict__ g_theta_all) {
    // weight set blockIdx.y owns slabs [y * n_slabs, (y + 1) * n_slabs)
    const float* __restrict__ slabs = slabs_all + (size_t)blockIdx.y * n_slabs * M.slab_floats;
    float* __restrict__ g_theta = g_theta_all + (size_t)blockIdx.y * M.n_theta;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M.n_theta) return;
    int sidx = -1;
    const int nl = M.n_rate + M.n_aug;
    for (int li = 0; li < nl; ++li) {
        const TcLayer& Lp = li < M.n_rate ? M.rate[li] : M.aug[li - M.n_rate];
        if (e >= Lp.th_w && e < Lp.th_w + Lp.in * Lp.out) {
            const int k = e - Lp.th_w;
            const int row = k / Lp.in, col = k - row * Lp.in;
            sidx = Lp.slab_off + (((((row >> 4) * (Lp.in_p >> 4) + (col >> 4)) * 4 + ((col & 15) >> 2)) * 16 + (row & 15)) * 4 + (col & 3));
            break;
        }
        if (e >= Lp.th_b && e < Lp.th_b + Lp.out) {
            sidx = Lp.slab_off + 64 * Lp.in_p + (e - Lp.th_b);
            break;
        }
    }
    float s = 0.f;
    if (sidx >= 0)
        for (int cidx = 0; cidx < n_slabs; ++cidx) s += slabs[(size_t)cidx * M.slab_floats + sidx];
    g_theta[e] = s;
}

}  // namespace ude
