// ude_rk4.cuh — fused fixed-step RK4 (torchdiffeq 'rk4' = 3/8 rule) forward and
// discrete-adjoint kernels.  One persistent CTA per SM walks over tiles of TB
// trajectories; a tile runs the WHOLE time grid before the CTA moves on.
//
// Replaces the Python loop of torchdiffeq FixedGridODESolver.integrate +
// rk4_alt_step_func that the reference drives from /root/reference/lib/VAE.py:137,
// and the autograd replay behind /root/reference/lib/VAE.py:203.
#pragma once
#include "ude_engine.cuh"
#include "ude_ring.cuh"

namespace ude {

struct Rk4FwdArgs {
    const float* wpacked;   // [n_packed]
    const float* y0;        // [B,R,L]
    GridTables g;
    float* y_out;           // [n_out,B,R,L]
    float* tape;            // [n_steps,4,Dn,B] or null
    double* side_partial;   // [gridDim.x, 8]
    float* cold;            // [gridDim.x, cold_floats] when the plan spills the state rows
    long long B;            // trajectories of ONE weight set
    // Weight-set groups (UdeDesc::n_groups = G > 1, BASELINE configs[4]): ONE launch walks the tiles of all G
    // group-major sub-problems; tile -> group = tile / tiles_per_group; the packed weights are re-staged when a CTA
    // moves on to another group; side partials are kept per TILE and summed per group in tile order, which is the
    // summation order of G single-set launches (bit-identical results).
    int n_groups;
    long long tape_group_floats;    // distance between the tapes of two groups
    ReparamArgs rp;                 // rp.mean != nullptr: y0 is PRODUCED here (ude_engine.cuh) and lives in y_out[0]
    DecArgs dec;                    // dec.n_dec > 0: decoded outputs (ude_engine.cuh DecArgs); y_out then holds row 0 only,
                                    // and the tape gets one extra record (index 4 * n_steps): the final state
    BayesArgs bayes;                // bayes.w != nullptr: one weight set per RHS evaluation (SMEM_W kernels only)
};

struct Rk4BwdArgs {
    const float* wpacked;
    const float* y0;
    GridTables g;
    const float* tape;
    const float* g_out;     // [n_out,B,R,L]
    const float* g_side;    // [5] or null
    float* g_y0;            // [B,R,L]
    float* slabs;           // [gridDim.x, n_packed] per-CTA dtheta accumulators ([n_tiles, n_packed] per TILE with groups)
    float* cold;
    long long B;
    int n_groups;           // see Rk4FwdArgs
    long long tape_group_floats;
    ReparamArgs rp;         // rp.mean != nullptr: also write dL/dz (rp.g_z) from the finished dL/dy0 rows
    DecArgs dec;            // dec.n_dec > 0: cotangents arrive as dL/dy_pred [n_out,B,n_dec] + g_side[5] = dL/dreg (g_out unused)
    BayesArgs bayes;        // bayes.w != nullptr: one weight set per RHS evaluation (SMEM_W + SMEM_G kernels only)
};

// Ring-mode prologue shared by the four FFMA-engine kernels (COLD variants): barriers, descriptor, first two fetches.
__device__ __forceinline__ void ring_setup(WRing& wr, const ModelP& P, float* smem, uint64_t* rbar, const float* wpacked) {
    if (threadIdx.x == 0) {
        mbar_init(&rbar[0], 1);
        mbar_init(&rbar[1], 1);
    }
    __syncthreads();
    wr.slots = smem + P.off_ring;
    wr.bar = rbar;
    wr.src = wpacked;
    wr.seq = P.wseq;
    wr.blk_off = P.wblk_off;
    wr.blk_n = P.wblk_n;
    wr.slot_floats = P.ring_slot_floats;
    wr.nseq = P.n_wseq;
    wr_start(wr);
}

// ---- state buffer indices ---------------------------------------------------------
enum { F_Y = 0, F_K1 = 1, F_K2 = 2, F_K3 = 3, F_NSTATE = 4 };
enum { B_LAM = 0, B_KB1 = 1, B_KB2 = 2, B_KB3 = 3, B_KB4 = 4, B_NSTATE = 5 };

__device__ __forceinline__ void copy_model_to_smem(ModelP* dst, const ModelP& src) {
    const int* s = reinterpret_cast<const int*>(&src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(ModelP) / sizeof(int)); i += kThreads) d[i] = s[i];
}

// Load the y0 tile: live components -> X rows (3r + c); `consts`: constant components -> XC rows (r*(L-3) + l-3,
// padding rows zeroed: XC aliases the ACT rows and is rebuilt whenever it is needed).  Invalid columns get zeros.
__device__ __forceinline__ void load_y0_tile(const ModelP& P, float* smem, const float* __restrict__ y0, long long b0,
                                             int nvalid, bool live, bool consts) {
    const int TB = P.TB, TBP = P.TBP, L = P.L, Lc = P.L - 3;
    for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
        const int r = idx >> P.tb_shift, b = idx & (TB - 1);
        float* X = smem + P.off_X + (3 * r) * TBP + b;
        float* XC = smem + P.off_XC + (r * Lc) * TBP + b;
        if (b < nvalid) {
            const float* src = y0 + ((b0 + b) * P.R + r) * (long long)L;
            if ((L & 3) == 0) {
                for (int l = 0; l < L; l += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + l));
                    if (l == 0) {
                        if (live) { X[0] = v.x; X[TBP] = v.y; X[2 * TBP] = v.z; }
                        if (consts) XC[0] = v.w;
                    } else if (consts) {
                        XC[(l - 3) * TBP] = v.x; XC[(l - 2) * TBP] = v.y; XC[(l - 1) * TBP] = v.z; XC[l * TBP] = v.w;
                    }
                }
            } else {
                if (live) { X[0] = __ldg(src); X[TBP] = __ldg(src + 1); X[2 * TBP] = __ldg(src + 2); }
                if (consts)
                    for (int l = 3; l < L; ++l) XC[(l - 3) * TBP] = __ldg(src + l);
            }
        } else {
            if (live) { X[0] = 0.f; X[TBP] = 0.f; X[2 * TBP] = 0.f; }
            if (consts)
                for (int l = 0; l < Lc; ++l) XC[l * TBP] = 0.f;
        }
    }
    if (consts)
        for (int i = threadIdx.x; i < (P.Dc_p - P.Dc) * TBP; i += kThreads) smem[P.off_XC + P.Dc * TBP + i] = 0.f;
}

// y0 tile PRODUCED from (eps, mean, std): the rows go to y_out[0] (their home from now on), the live components to X,
// the constant ones to XC -- same destinations as load_y0_tile.
__device__ __forceinline__ void reparam_y0_tile(const ModelP& P, float* smem, const ReparamArgs& rp, float* __restrict__ y_out0,
                                                long long b0, int nvalid, bool consts) {
    const int TB = P.TB, TBP = P.TBP, L = P.L, Lc = P.L - 3;
    for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
        const int r = idx >> P.tb_shift, b = idx & (TB - 1);
        float* X = smem + P.off_X + (3 * r) * TBP + b;
        float* XC = smem + P.off_XC + (r * Lc) * TBP + b;
        if (b < nvalid) {
            const long long traj = b0 + b;
            const float z0 = reparam_z(rp, traj, r, P.R, L - 1, 0), z1 = reparam_z(rp, traj, r, P.R, L - 1, 1);
            float* row = y_out0 + (traj * P.R + r) * (long long)L;
            for (int l = 0; l < L; ++l) {
                const float y = reparam_y(z0, z1, l >= 3 ? reparam_z(rp, traj, r, P.R, L - 1, l - 1) : 0.f, l);
                row[l] = y;
                if (l < 3) X[l * TBP] = y;
                else if (consts) XC[(l - 3) * TBP] = y;
            }
        } else {
            X[0] = 0.f; X[TBP] = 0.f; X[2 * TBP] = 0.f;
            if (consts)
                for (int l = 0; l < Lc; ++l) XC[l * TBP] = 0.f;
        }
    }
    if (consts)
        for (int i = threadIdx.x; i < (P.Dc_p - P.Dc) * TBP; i += kThreads) smem[P.off_XC + P.Dc * TBP + i] = 0.f;
}

// Store one output row (r, b): dynamic comps from v[], constants straight from the caller's y0 row (L2-resident).
__device__ __forceinline__ void store_out_row(const ModelP& P, const float* y0row, float* dst, const float (&v)[3]) {
    // L2 (coherent) loads: with the fused reparam the y0 rows were written by THIS kernel (same thread, earlier)
    const int L = P.L;
    if ((L & 3) == 0) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(y0row));
        st4(dst, make_float4(v[0], v[1], v[2], a.w));
        for (int l = 4; l < L; l += 4) {
            st4(dst + l, __ldcg(reinterpret_cast<const float4*>(y0row + l)));
        }
    } else {
        dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
        for (int l = 3; l < L; ++l) dst[l] = __ldcg(y0row + l);
    }
}

// ---- decoded outputs (ude_engine.cuh DecArgs) ----------------------------------------------------------------------
// Live value (row k, trajectory b) of output jj: the grid state after the step (kind 1, rows X), before it (kind 0,
// rows SY) or torchdiffeq's linear interpolation between the two (kind 2) -- the arithmetic of the plain output store.
__device__ __forceinline__ float out_value(const float* __restrict__ XN, const float* __restrict__ SY, int at, int kind, float w) {
    if (kind == 1) return XN[at];
    const float y = SY[at];
    return kind == 0 ? y : __fadd_rn(y, __fmul_rn(w, __fsub_rn(XN[at], y)));
}
// y_pred[jj] = Wd . v + bd for the tile, the reg sum and the sign mask.  Pre: XN (and SY) rows synced.  No sync inside.
__device__ __forceinline__ void decode_output(const ModelP& P, const DecArgs& d, const float* __restrict__ XN, const float* __restrict__ SY,
                                              int jj, int kind, float w, long long B, long long b0, int nvalid, SideAcc& side) {
    const int TB = P.TB, TBP = P.TBP, K = P.Dn;
    for (int idx = threadIdx.x; idx < d.n_dec * TB; idx += kThreads) {
        const int o = idx >> P.tb_shift, b = idx & (TB - 1);
        if (b >= nvalid) continue;
        const float* __restrict__ wrow = d.W + (long long)o * K;
        float acc = __ldg(d.b + o);
        for (int k = 0; k < K; ++k) acc = fmaf(__ldg(wrow + k), out_value(XN, SY, k * TBP + b, kind, w), acc);
        d.y_pred[((long long)jj * B + b0 + b) * d.n_dec + o] = acc;
    }
    for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
        const int r = idx >> P.tb_shift, b = idx & (TB - 1);
        if (b >= nvalid) continue;
        const float v0 = out_value(XN, SY, (3 * r) * TBP + b, kind, w), v1 = out_value(XN, SY, (3 * r + 1) * TBP + b, kind, w),
                    v2 = out_value(XN, SY, (3 * r + 2) * TBP + b, kind, w);
        d.mask[((long long)jj * B + b0 + b) * P.R + r] = (unsigned char)reg_mask_bits(v0, v1, v2);
        side.reg += (double)reg_term(v0) + (double)reg_term(v1) + (double)reg_term(v2);
    }
}

// Adjoint of decode_output, cotangent side: dL/dv_c (c < 3) of region r of one trajectory for output jj.
__device__ __forceinline__ void dec_cotangent(const ModelP& P, const DecArgs& d, float g_reg, int jj, long long B, long long traj, int r,
                                              float (&gl)[3]) {
    const float* __restrict__ gp = d.g_pred + ((long long)jj * B + traj) * d.n_dec;
    const unsigned m = d.mask[((long long)jj * B + traj) * P.R + r];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    for (int o = 0; o < d.n_dec; ++o) {
        const float g = __ldg(gp + o);
        const float* __restrict__ wv = d.W + (long long)o * P.Dn + 3 * r;
        g0 = fmaf(__ldg(wv), g, g0);
        g1 = fmaf(__ldg(wv + 1), g, g1);
        g2 = fmaf(__ldg(wv + 2), g, g2);
    }
    gl[0] = fmaf(g_reg, reg_sign(m, 0), g0);
    gl[1] = fmaf(g_reg, reg_sign(m, 1), g1);
    gl[2] = fmaf(g_reg, reg_sign(m, 2), g2);
}
// Adjoint of decode_output, parameter side: partial[o * 3R + k] += sum_b g_pred[jj][b][o] * v_k(b), partial[n_dec * 3R + o] +=
// sum_b g_pred[jj][b][o].  Every entry has ONE owning thread per CTA (fixed order: bit-reproducible).  v: rows XN hold the
// state after the step; the state before it (kinds 0 / 2) is the step's first stage input on the tape (yn, global).
__device__ __forceinline__ void dec_param_grad(const ModelP& P, const DecArgs& d, float* __restrict__ partial, const float* __restrict__ XN,
                                               const float* __restrict__ yn, int jj, int kind, float w, long long B, long long b0, int nvalid) {
    const int K = P.Dn, TBP = P.TBP, nw = d.n_dec * K;
    const float* __restrict__ gp = d.g_pred + ((long long)jj * B + b0) * d.n_dec;
    for (int idx = threadIdx.x; idx < nw + d.n_dec; idx += kThreads) {
        float acc = 0.f;
        if (idx < nw) {
            const int o = idx / K, k = idx - o * K;
            for (int b = 0; b < nvalid; ++b) {
                float v = XN[k * TBP + b];
                if (kind != 1) {
                    const float y = __ldg(yn + (long long)k * B + b0 + b);
                    v = kind == 0 ? y : __fadd_rn(y, __fmul_rn(w, __fsub_rn(v, y)));
                }
                acc = fmaf(__ldg(gp + (long long)b * d.n_dec + o), v, acc);
            }
        } else {
            const int o = idx - nw;
            for (int b = 0; b < nvalid; ++b) acc += __ldg(gp + (long long)b * d.n_dec + o);
        }
        partial[idx] += acc;
    }
}

// ---- hoisted constants (ModelP: "Hoisting") -----------------------------------------------------------------------
// c0 rows of both first layers: W0[:, const] . x_const + b0.  Pre: XC loaded and synced.  Post: synced.
// The constant blocks are read from the packed blob in GLOBAL memory (used twice per tile and solve).
__device__ __forceinline__ void hoist_c0(const ModelP& P, const float* __restrict__ Wg, float* smem, int warp, int lane) {
    if (!P.hoist) return;
    const int u0 = P.n_rate > 0 ? fwd_units(P.rate_c, P.mt_shift) : 0;
    const int u1 = P.n_aug > 0 ? fwd_units(P.aug_c, P.mt_shift) : 0;
    for (int u = warp; u < u0 + u1; u += kWarps) {
        if (u < u0) fwd_run(P.rate_c, Wg, smem, P.TBP, P.mt_shift, u, lane);
        else fwd_run(P.aug_c, Wg, smem, P.TBP, P.mt_shift, u - u0, lane);
    }
    __syncthreads();
}

// Adjoint of the hoisted product, once per tile: with D = sum over evaluations of delta_0 (off_DL rows)
//   dW0[:, const] += D (x) x_const   (global CTA slab)      dL/dx_const += W0[:, const]^T D   (into g_y0)
// Pre: all warps are done with the ACT rows (synced).  Post: synced.
__device__ __forceinline__ void hoist_backward_epilogue(const ModelP& P, const float* __restrict__ Wg, float* smem,
                                                        float* __restrict__ Gg, const float* __restrict__ y0,
                                                        float* __restrict__ g_y0, long long b0, int nvalid, int warp, int lane) {
    if (!P.hoist) return;
    const int TB = P.TB, TBP = P.TBP, L = P.L, Lc = P.L - 3;
    load_y0_tile(P, smem, y0, b0, nvalid, false, true);
    __syncthreads();
    LayerP Lr = P.rate_c, La = P.aug_c;
    Lr.act_off = P.off_DL;
    La.act_off = P.off_DL + (P.n_rate > 0 ? P.rate[0].n_out_p * TBP : 0);
    const int w0 = P.n_rate > 0 ? dw_units(Lr) : 0, w1 = P.n_aug > 0 ? dw_units(La) : 0;
    for (int u = warp; u < w0 + w1; u += kWarps) {
        if (u < w0) dw_run(Lr, smem, Gg, TB, TBP, u, lane);
        else dw_run(La, smem, Gg, TB, TBP, u - w0, lane);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P.Dc_p * TBP; i += kThreads) smem[P.off_XC + i] = 0.f;
    __syncthreads();
    if (P.n_rate > 0) {
        for (int u = warp; u < bwd_units(Lr, P.mt_shift); u += kWarps) bwd_run(Lr, nullptr, Wg, smem, TBP, P.off_XC, P.mt_shift, u, lane);
        __syncthreads();
    }
    if (P.n_aug > 0) {
        for (int u = warp; u < bwd_units(La, P.mt_shift); u += kWarps) bwd_run(La, nullptr, Wg, smem, TBP, P.off_XC, P.mt_shift, u, lane);
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
        const int r = idx >> P.tb_shift, b = idx & (TB - 1);
        if (b < nvalid) {
            float* dst = g_y0 + ((b0 + b) * P.R + r) * (long long)L;
            const float* src = smem + P.off_XC + (r * Lc) * TBP + b;
            for (int l = 3; l < L; ++l) dst[l] += src[(l - 3) * TBP];
        }
    }
    __syncthreads();
}

// Bayes mode, after the sweep of ONE evaluation (weights We, noise ze): the hoisted product's adjoint for this evaluation
//   Gslab[const blocks] = delta_0 (x) x_const,   GXC += We[:, const]^T delta_0        (DL rows hold delta_0 of this evaluation only)
// then the fold  A += G_e,  Bz += G_e * z_e  over the whole packed vector (hot part in shared memory, constant blocks in the
// CTA's global slab), leaving G, the slab's constant blocks and DL zeroed for the next evaluation.
// Pre: synced after rhs_backward_layers.  Post: synced.
__device__ __forceinline__ void bayes_eval_epilogue(const ModelP& P, const float* __restrict__ We, const float* __restrict__ ze,
                                                    float* smem, float* __restrict__ G, float* __restrict__ Gslab,
                                                    float* __restrict__ A, float* __restrict__ Bz, int warp, int lane) {
    const int TB = P.TB, TBP = P.TBP;
    if (P.hoist) {
        LayerP Lr = P.rate_c, La = P.aug_c;
        Lr.act_off = P.off_DL;
        La.act_off = P.off_DL + (P.n_rate > 0 ? P.rate[0].n_out_p * TBP : 0);
        const int w0 = P.n_rate > 0 ? dw_units(Lr) : 0, w1 = P.n_aug > 0 ? dw_units(La) : 0;
        for (int u = warp; u < w0 + w1; u += kWarps) {
            if (u < w0) dw_run(Lr, smem, Gslab, TB, TBP, u, lane);
            else dw_run(La, smem, Gslab, TB, TBP, u - w0, lane);
        }
        if (P.n_rate > 0) {
            for (int u = warp; u < bwd_units(Lr, P.mt_shift); u += kWarps) bwd_run(Lr, nullptr, We, smem, TBP, P.off_GXC, P.mt_shift, u, lane);
            __syncthreads();
        }
        if (P.n_aug > 0) {
            for (int u = warp; u < bwd_units(La, P.mt_shift); u += kWarps) bwd_run(La, nullptr, We, smem, TBP, P.off_GXC, P.mt_shift, u, lane);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < P.dl_rows * TBP; i += kThreads) smem[P.off_DL + i] = 0.f;
    }
    for (int i = threadIdx.x; i < P.n_packed; i += kThreads) {       // every entry has one owning thread: plain read-modify-writes
        float g;
        if (i < P.n_packed_hot) { g = G[i]; G[i] = 0.f; }
        else { g = Gslab[i]; Gslab[i] = 0.f; }
        A[i] += g;
        Bz[i] = fmaf(g, __ldg(ze + i), Bz[i]);
    }
    __syncthreads();
}

// One output row's cotangent: the live components come back in gl[], the constant ones are added straight into the
// caller's g_y0 row (each row is owned by one thread; the lines stay in L2) -- no shared-memory rows for that running
// sum.  128-bit accesses when the row allows it (the live lanes of the first quad are rewritten with their own values).
__device__ __forceinline__ void add_const_cotangent(float* __restrict__ dst, const float* __restrict__ gp, int L, float (&gl)[3]) {
    if ((L & 3) == 0) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp));
        float4 d0 = ld4(dst);
        gl[0] = g0.x; gl[1] = g0.y; gl[2] = g0.z;
        d0.w += g0.w;
        st4(dst, d0);
        for (int l = 4; l < L; l += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gp + l));
            float4 d = ld4(dst + l);
            d.x += g.x; d.y += g.y; d.z += g.z; d.w += g.w;
            st4(dst + l, d);
        }
    } else {
        gl[0] = __ldg(gp); gl[1] = __ldg(gp + 1); gl[2] = __ldg(gp + 2);
        for (int l = 3; l < L; ++l) dst[l] += __ldg(gp + l);
    }
}

// ===================================================================================
// forward
// ===================================================================================
// MINB = 2: register-capped build (<= 128 per thread) for the SMALL models (hyper-parameter sweeps, Fp with 32-wide nets):
// their tiles are a chain of short phases separated by CTA barriers, so two resident CTAs per SM hide each other's
// barrier latency.  Same code, same results.
template <bool SMEM_W, bool COLD, int MINB = 1>
__global__ void __launch_bounds__(kThreads, MINB) rk4_fwd_kernel(const __grid_constant__ ModelP Pg, const Rk4FwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ ModelP P;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ __align__(8) uint64_t rbar[2];
    __shared__ double red[kWarps * 8];

    copy_model_to_smem(&P, Pg);
    if (threadIdx.x == 0 && SMEM_W) mbar_init(&wbar, 1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TB = P.TB, TBP = P.TBP, L = P.L, Dn = P.Dn;
    const bool ring = COLD && P.w_ring;          // wide models: weights through the TMA ring (ude_ring.cuh)
    WRing wr;
    if (ring) ring_setup(wr, P, smem, rbar, a.wpacked);
    uint32_t wphase = 0;
    const float* Wp = a.wpacked;
    if (SMEM_W) {
        stage_weights_tma(smem + P.off_W, a.wpacked, P.n_packed_hot, &wbar, wphase);
        Wp = smem + P.off_W;
    }
    // zero the padding rows of X once (they feed layer 0 with K = Dn_p)
    for (int i = threadIdx.x; i < (P.Dn_p - P.Dn) * TBP; i += kThreads) smem[P.off_X + P.Dn * TBP + i] = 0.f;

    float* SB = COLD ? a.cold + (long long)blockIdx.x * P.cold_floats : smem;    // base of the state rows
    float* SY = SB + P.off_S + F_Y * P.Dn_p * TBP;
    float* SK1 = SB + P.off_S + F_K1 * P.Dn_p * TBP;
    float* SK2 = SB + P.off_S + F_K2 * P.Dn_p * TBP;
    float* SK3 = SB + P.off_S + F_K3 * P.Dn_p * TBP;

    SideAcc side;
    side.clear();
    const double n_evals = 4.0 * a.g.n_steps;      // RHS evaluations per trajectory
    const long long B = a.B;
    const long long tiles_per_group = (B + TB - 1) / TB;
    const long long n_tiles = tiles_per_group * a.n_groups;
    const long long tape_stage = (long long)Dn * B;           // floats per (step, stage)
    const bool grouped = a.n_groups > 1;
    long long g_cur = 0;
    const long long out_group = (long long)a.g.n_out * B * P.R * L, y0_group = B * P.R * L;
    const bool decoded = a.dec.n_dec > 0;      // outputs leave as y_pred / reg / mask (single weight set)
    const bool bayes = SMEM_W && a.bayes.w != nullptr;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long grp = grouped ? tile / tiles_per_group : 0;
        const long long b0 = (tile - grp * tiles_per_group) * TB;
        const int nvalid = (int)((B - b0) < TB ? (B - b0) : TB);
        // this group's views of the caller's arrays (group-major layouts, include/ude_b200.h)
        float* const a_y_out = a.y_out + grp * out_group;
        const bool reparam = a.rp.mean != nullptr;             // y0 = y_out[0], produced below
        const float* const a_y0 = reparam ? a_y_out : a.y0 + grp * y0_group;
        float* const a_tape = a.tape ? a.tape + grp * a.tape_group_floats : nullptr;
        const float* const a_wpacked = a.wpacked + grp * P.n_packed;
        __syncthreads();                                       // previous tile done with X, the ACT rows and the weights
        if (SMEM_W && grp != g_cur) {
            stage_weights_tma(smem + P.off_W, a_wpacked, P.n_packed_hot, &wbar, wphase);
            g_cur = grp;
        } else if (!SMEM_W) {
            Wp = a_wpacked;
        }
        if (reparam) reparam_y0_tile(P, smem, a.rp, a_y_out, b0, nvalid, P.hoist != 0);
        else load_y0_tile(P, smem, a_y0, b0, nvalid, true, P.hoist != 0);
        __syncthreads();
        if (ring) hoist_c0_ring(P, a_wpacked, smem, warp, lane, 0u);
        else if (!bayes) hoist_c0(P, a_wpacked, smem, warp, lane);          // constant part of both first layers, once per tile
        // output 0 = y0 (torchdiffeq: solution[0] = y0) and Y = dynamic part of y0
        for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            const float* X = smem + P.off_X + (3 * r) * TBP + b;
            float v[3] = {X[0], X[TBP], X[2 * TBP]};
            SY[(3 * r + 0) * TBP + b] = v[0];
            SY[(3 * r + 1) * TBP + b] = v[1];
            SY[(3 * r + 2) * TBP + b] = v[2];
            if (b < nvalid && !decoded) {
                const long long row = ((b0 + b) * P.R + r) * (long long)L;
                store_out_row(P, a_y0 + row, a_y_out + row, v);
            }
        }
        if (decoded) decode_output(P, a.dec, smem + P.off_X, SY, 0, 1, 0.f, B, b0, nvalid, side);     // output 0 = y0 (X rows, synced above)
        // (each thread only re-reads rows it wrote itself below, no sync needed here)
        int j = 1;
        for (int n = 0; n < a.g.n_steps; ++n) {
            const float dt = __ldg(a.g.dt + n);
            int j_next = j;
#pragma unroll 1
            for (int stage = 0; stage < 4; ++stage) {
                if (SMEM_W && bayes) {      // this evaluation's weight set (the previous one's readers are behind a CTA barrier)
                    const float* We = a.bayes.w + (long long)(n * 4 + stage) * P.n_packed;
                    stage_weights_tma(smem + P.off_W, We, P.n_packed_hot, &wbar, wphase);
                    hoist_c0(P, We, smem, warp, lane);
                }
                if (ring) rhs_forward_layers_ring(P, wr, smem, warp, lane, 0u);
                else rhs_forward_layers(P, Wp, smem, warp, lane);
                UDE_PROF_DECL;
                for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    const bool valid = b < nvalid;
                    float f[3];
                    rhs_tail(P, smem, r, b, valid, side, f);
                    float* X = smem + P.off_X + (3 * r) * TBP + b;
                    float* tp = a_tape ? a_tape + ((long long)(n * 4 + stage) * tape_stage + (long long)(3 * r) * B + b0 + b) : nullptr;
                    // All state loads of the task first, then the stores: with the state rows in the CTA's global scratch
                    // (wide models) a store followed by a load the compiler cannot prove disjoint costs one L2 round trip each.
                    float xn[3], yv[3], k1v[3], k2v[3], k3v[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        yv[c] = SY[row];
                        k1v[c] = stage >= 1 ? SK1[row] : 0.f;
                        k2v[c] = stage >= 2 ? SK2[row] : 0.f;
                        k3v[c] = stage >= 3 ? SK3[row] : 0.f;
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        const float y = yv[c];
                        if (tp && valid) tp[(long long)c * B] = X[c * TBP];     // stage input x_{stage+1}
                        if (stage == 0) {
                            SK1[row] = f[c];
                            xn[c] = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, f[c]), 1.0f / 3.0f));
                        } else if (stage == 1) {
                            SK2[row] = f[c];
                            xn[c] = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(f[c], __fmul_rn(k1v[c], 1.0f / 3.0f))));
                        } else if (stage == 2) {
                            SK3[row] = f[c];
                            xn[c] = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1v[c], k2v[c]), f[c])));
                        } else {
                            const float s = __fadd_rn(__fadd_rn(k1v[c], __fmul_rn(3.0f, __fadd_rn(k2v[c], k3v[c]))), f[c]);
                            xn[c] = __fadd_rn(y, __fmul_rn(__fmul_rn(s, dt), 0.125f));
                        }
                    }
                    if (stage == 3) {
                        // outputs that fall into (t_n, t_{n+1}]  (FixedGridODESolver.integrate)
                        int jj = j;
                        while (jj < a.g.n_out && __ldg(a.g.out_step + jj) == n) {
                            if (valid && !decoded) {
                                const int kind = __ldg(a.g.out_kind + jj);
                                float v[3];
                                if (kind == 1) {
                                    v[0] = xn[0]; v[1] = xn[1]; v[2] = xn[2];
                                } else {
                                    const float w = __ldg(a.g.out_w + jj);
#pragma unroll
                                    for (int c = 0; c < 3; ++c) {
                                        const float y = yv[c];
                                        v[c] = kind == 0 ? y : __fadd_rn(y, __fmul_rn(w, __fsub_rn(xn[c], y)));
                                    }
                                }
                                store_out_row(P, a_y0 + ((b0 + b) * P.R + r) * (long long)L,
                                              a_y_out + (((long long)jj * B + b0 + b) * P.R + r) * (long long)L, v);
                            }
                            ++jj;
                        }
                        j_next = jj;
                        if (!decoded) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) SY[(3 * r + c) * TBP + b] = xn[c];
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) X[c * TBP] = xn[c];
                }
                if (stage == 3) {
                    // threads with no (r,b) task still need the advanced output cursor
                    int jj = j;
                    while (jj < a.g.n_out && __ldg(a.g.out_step + jj) == n) ++jj;
                    j_next = jj;
                }
                UDE_PROF_WORK_DONE();
                __syncthreads();
                UDE_PROF_PHASE_END(6);
                if (stage == 3 && decoded) {
                    // X = y_{n+1}, SY still = y_n: decode the outputs of (t_n, t_{n+1}], then advance SY
                    for (int jj = j; jj < j_next; ++jj)
                        decode_output(P, a.dec, smem + P.off_X, SY, jj, __ldg(a.g.out_kind + jj), __ldg(a.g.out_w + jj), B, b0, nvalid, side);
                    if (a_tape && n == a.g.n_steps - 1) {      // the extra tape record: the final state
                        for (int idx = threadIdx.x; idx < Dn * TB; idx += kThreads) {
                            const int k = idx >> P.tb_shift, b = idx & (TB - 1);
                            if (b < nvalid) a_tape[(long long)(4 * a.g.n_steps) * tape_stage + (long long)k * B + b0 + b] = smem[P.off_X + k * TBP + b];
                        }
                    }
                    __syncthreads();
                    for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                        const int r = idx >> P.tb_shift, b = idx & (TB - 1);
#pragma unroll
                        for (int c = 0; c < 3; ++c) SY[(3 * r + c) * TBP + b] = smem[P.off_X + (3 * r + c) * TBP + b];
                    }
                }
            }
            j = j_next;
        }
        if (grouped) {                                         // side sums per TILE (summed per group afterwards)
            side_block_reduce(side, n_evals, red, a.side_partial + tile * 8);
            side.clear();
        }
    }
    if (!grouped) side_block_reduce(side, n_evals, red, a.side_partial + (long long)blockIdx.x * 8);
    if (ring) wr_drain(wr);
}

// ===================================================================================
// backward (discrete adjoint)
// ===================================================================================
template <bool SMEM_W, bool SMEM_G, bool COLD, int MINB = 1>
__global__ void __launch_bounds__(kThreads, MINB) rk4_bwd_kernel(const __grid_constant__ ModelP Pg, const Rk4BwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ ModelP P;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ __align__(8) uint64_t rbar[2];
    __shared__ uint32_t tmem_slot;

    copy_model_to_smem(&P, Pg);
    if (threadIdx.x == 0 && SMEM_W) mbar_init(&wbar, 1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TB = P.TB, TBP = P.TBP, L = P.L, Dn = P.Dn;
    const bool ring = COLD && P.w_ring;          // wide models: weights through the TMA ring, c0 / D in tensor memory
    WRing wr;
    if (ring) ring_setup(wr, P, smem, rbar, a.wpacked);
    // plans of last resort (wide models): the weight-gradient accumulators live in tensor memory for the whole launch
    const bool tm_acc = COLD && P.tm_acc;
    uint32_t tmem = 0u;
    if (tm_acc) {
        if (warp == 0) tm_alloc512(&tmem_slot);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem = tmem_slot;
        dw_tm_clear(P, smem + P.off_GB, tmem, warp);
    }
    uint32_t wphase = 0;
    const float* Wp = a.wpacked;
    if (SMEM_W) {
        stage_weights_tma(smem + P.off_W, a.wpacked, P.n_packed_hot, &wbar, wphase);
        Wp = smem + P.off_W;
    }
    // dtheta accumulators: the hot part (every evaluation) in shared memory when it fits, else in the CTA's global slab;
    // the constant blocks of the first layers (once per tile) always in the global slab
    const bool grouped = a.n_groups > 1;      // per-TILE slabs, zeroed and flushed inside the tile loop
    float* Gslab = a.slabs + (long long)blockIdx.x * P.n_packed;
    float* G = SMEM_G ? smem + P.off_G : Gslab;
    if (!grouped) {
        for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) G[i] = 0.f;
        for (int i = P.n_packed_hot + threadIdx.x; i < P.n_packed; i += kThreads) Gslab[i] = 0.f;
    }
    for (int i = threadIdx.x; i < (P.Dn_p - P.Dn) * TBP; i += kThreads) smem[P.off_X + P.Dn * TBP + i] = 0.f;

    float gs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.g_side && !grouped) {
#pragma unroll
        for (int i = 0; i < 5; ++i) gs[i] = __ldg(a.g_side + i);
    }
    const bool bayes = SMEM_W && SMEM_G && a.bayes.w != nullptr;     // one weight set per RHS evaluation (ude_engine.cuh BayesArgs)
    float* const bayesA = bayes ? a.bayes.fold + (long long)blockIdx.x * P.n_packed : nullptr;
    float* const bayesBz = bayes ? a.bayes.fold + ((long long)gridDim.x + blockIdx.x) * P.n_packed : nullptr;
    if (bayes)
        for (int i = threadIdx.x; i < P.n_packed; i += kThreads) { bayesA[i] = 0.f; bayesBz[i] = 0.f; }
    const bool decoded = a.dec.n_dec > 0;      // cotangents arrive as dL/dy_pred + dL/dreg (ude_engine.cuh DecArgs)
    const float g_reg = (decoded && a.g_side) ? __ldg(a.g_side + 5) : 0.f;
    float* const dpart = decoded ? a.dec.partial + (long long)blockIdx.x * (a.dec.n_dec * (P.Dn + 1)) : nullptr;
    if (decoded)
        for (int i = threadIdx.x; i < a.dec.n_dec * (P.Dn + 1); i += kThreads) dpart[i] = 0.f;      // own entries only, see dec_param_grad

    const int SS = P.Dn_p * TBP;
    float* SB = COLD ? a.cold + (long long)blockIdx.x * P.cold_floats : smem;
    float* LAM = SB + P.off_S + B_LAM * SS;
    float* KB = SB + P.off_S + B_KB1 * SS;          // KB[s] = KB + s*SS, s = 0..3
    float* GX = smem + P.off_GX;
    float* DL = smem + P.off_DL;

    const long long B = a.B;
    const long long tiles_per_group = (B + TB - 1) / TB;
    const long long n_tiles = tiles_per_group * a.n_groups;
    const long long tape_stage = (long long)Dn * B;
    const int n_steps = a.g.n_steps;
    long long g_cur = 0;
    const long long out_group = (long long)a.g.n_out * B * P.R * L, y0_group = B * P.R * L;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long grp = grouped ? tile / tiles_per_group : 0;
        const long long b0 = (tile - grp * tiles_per_group) * TB;
        const int nvalid = (int)((B - b0) < TB ? (B - b0) : TB);
        const float* const a_y0 = a.y0 + grp * y0_group;
        float* const a_g_y0 = a.g_y0 + grp * y0_group;
        const float* const a_g_out = a.g_out + grp * out_group;
        const float* const a_tape = a.tape + grp * a.tape_group_floats;
        const float* const a_wpacked = a.wpacked + grp * P.n_packed;
        __syncthreads();
        if (SMEM_W && grp != g_cur) {
            stage_weights_tma(smem + P.off_W, a_wpacked, P.n_packed_hot, &wbar, wphase);
            g_cur = grp;
        } else if (!SMEM_W) {
            Wp = a_wpacked;
        }
        if (grouped) {
            Gslab = a.slabs + tile * P.n_packed;
            if (!SMEM_G) G = Gslab;
            for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) G[i] = 0.f;
            for (int i = P.n_packed_hot + threadIdx.x; i < P.n_packed; i += kThreads) Gslab[i] = 0.f;
            if (a.g_side) {
#pragma unroll
                for (int i = 0; i < 5; ++i) gs[i] = __ldg(a.g_side + grp * 5 + i);
            }
        }
        load_y0_tile(P, smem, a_y0, b0, nvalid, false, P.hoist != 0);     // constants only; the dynamic rows come from the tape
        for (int i = threadIdx.x; i < SS; i += kThreads) LAM[i] = 0.f;
        for (int i = threadIdx.x; i < P.Dn_p * TBP; i += kThreads) GX[i] = 0.f;
        if (ring && P.c0_tm) dl_tm_clear(P, tmem, warp);
        else
            for (int i = threadIdx.x; i < P.dl_rows * TBP; i += kThreads) DL[i] = 0.f;
        // dL/dy0 starts from the cotangent of output 0 (= y0 itself); the constant components then collect the other
        // outputs' cotangents as they are visited, the live ones are overwritten at the end
        for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            if (b < nvalid) {
                const long long row = ((b0 + b) * P.R + r) * (long long)L;
                for (int l = 0; l < L; ++l) a_g_y0[row + l] = decoded ? 0.f : __ldg(a_g_out + row + l);
            }
        }
        if (decoded) {      // X := the final state (the tape's extra record): the outputs of the last step need it
            for (int idx = threadIdx.x; idx < Dn * TB; idx += kThreads) {
                const int k = idx >> P.tb_shift, b = idx & (TB - 1);
                smem[P.off_X + k * TBP + b] = b < nvalid ? __ldg(a_tape + (long long)(4 * n_steps) * tape_stage + (long long)k * B + b0 + b) : 0.f;
            }
        }
        __syncthreads();
        if (bayes)
            for (int i = threadIdx.x; i < P.Dc_p * TBP; i += kThreads) smem[P.off_GXC + i] = 0.f;
        if (ring) hoist_c0_ring(P, a_wpacked, smem, warp, lane, tmem);
        else if (!bayes) hoist_c0(P, a_wpacked, smem, warp, lane);

        int j = a.g.n_out - 1;
        for (int n = n_steps - 1; n >= 0; --n) {
            const float dt = __ldg(a.g.dt + n);
            // ---- outputs tied to step n feed lambda_{n+1} (and, for copies of y_n / lerps, lambda_n)
            int j_next = j;
            {
                int jj = j;
                while (jj >= 1 && __ldg(a.g.out_step + jj) == n) --jj;
                j_next = jj;
            }
            if (decoded) {      // X still holds y_{n+1} (stage-1 input of step n+1, or the final record): decoder gradients first
                for (int jj = j; jj > j_next; --jj)
                    dec_param_grad(P, a.dec, dpart, smem + P.off_X, a_tape + (long long)(n * 4) * tape_stage, jj, __ldg(a.g.out_kind + jj),
                                   __ldg(a.g.out_w + jj), B, b0, nvalid);
                __syncthreads();
            }
            for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                const bool valid = b < nvalid;
                float lam[3], pend[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 3; ++c) lam[c] = LAM[(3 * r + c) * TBP + b];
                if (valid) {
                    for (int jj = j; jj > j_next; --jj) {
                        const int kind = __ldg(a.g.out_kind + jj);
                        const float w = __ldg(a.g.out_w + jj);
                        float gl[3];
                        if (decoded) {
                            dec_cotangent(P, a.dec, g_reg, jj, B, b0 + b, r, gl);
                        } else {
                            const float* gp = a_g_out + (((long long)jj * B + b0 + b) * P.R + r) * (long long)L;
                            add_const_cotangent(a_g_y0 + ((b0 + b) * P.R + r) * (long long)L, gp, L, gl);
                        }
#pragma unroll
                        for (int l = 0; l < 3; ++l) {
                            const float g = gl[l];
                            if (kind == 1) lam[l] += g;
                            else if (kind == 0) pend[l] += g;
                            else { const float sg = w * g; lam[l] += sg; pend[l] += g - sg; }
                        }
                    }
                }
                const float c8 = dt * 0.125f;
                float x4[3];          // stage input of stage 4 (loads before the state stores, see the forward kernel)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    x4[c] = valid ? __ldg(a_tape + (long long)(n * 4 + 3) * tape_stage + (long long)(3 * r + c) * B + b0 + b) : 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int row = (3 * r + c) * TBP + b;
                    const float base = lam[c] * c8;
                    KB[0 * SS + row] = base;
                    KB[1 * SS + row] = 3.f * base;
                    KB[2 * SS + row] = 3.f * base;
                    KB[3 * SS + row] = base;
                    LAM[row] = lam[c] + pend[c];    // becomes lambda_n once the four stage VJPs are added
                    smem[P.off_X + row] = x4[c];
                }
            }
            j = j_next;
            __syncthreads();
#pragma unroll 1
            for (int stage = 3; stage >= 0; --stage) {
                const float* We = nullptr;
                if (bayes) {        // this evaluation's weight set (the previous one's readers are behind a CTA barrier)
                    We = a.bayes.w + (long long)(n * 4 + stage) * P.n_packed;
                    stage_weights_tma(smem + P.off_W, We, P.n_packed_hot, &wbar, wphase);
                    hoist_c0(P, We, smem, warp, lane);
                }
                if (ring) rhs_forward_layers_ring(P, wr, smem, warp, lane, tmem);
                else rhs_forward_layers(P, Wp, smem, warp, lane);
                UDE_PROF_DECL;
                // ---- tail VJP: cotangent KB[stage] -> head cotangents (in place) + direct dx
                for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    const int row = (3 * r) * TBP + b;
                    float xb[3];
                    rhs_tail_vjp(P, smem, r, b, b < nvalid, gs, KB[stage * SS + row], KB[stage * SS + row + TBP],
                                 KB[stage * SS + row + 2 * TBP], xb);
                    float* gx = GX + (3 * r) * TBP + b;     // GX := direct term; the layer-0 deltas accumulate on top
                    gx[0] = xb[0]; gx[TBP] = xb[1]; gx[2 * TBP] = xb[2];
                }
                UDE_PROF_WORK_DONE();
                __syncthreads();
                UDE_PROF_PHASE_END(20);
                if (ring) rhs_backward_layers_ring(P, wr, smem, smem + P.off_GB, warp, lane, tmem);
                else if (tm_acc) rhs_backward_layers<true>(P, Wp, smem, smem + P.off_GB, warp, lane, tmem);
                else rhs_backward_layers<false>(P, Wp, smem, G, warp, lane);
                if (bayes)
                    bayes_eval_epilogue(P, We, a.bayes.z + (long long)(n * 4 + stage) * P.n_packed, smem, G, Gslab, bayesA, bayesBz, warp, lane);
                UDE_PROF_RESTART();
                // ---- combine: xbar = direct + W0^T delta0 (both nets); propagate through the tableau
                for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    const bool valid = b < nvalid;
                    // loads first (state rows may live in the CTA's global scratch), then the arithmetic, then the stores
                    float xbv[3], k0[3], k1[3], k2[3], lm[3], xnext[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        xbv[c] = GX[row];
                        k0[c] = stage >= 1 ? KB[0 * SS + row] : 0.f;
                        k1[c] = stage >= 2 ? KB[1 * SS + row] : 0.f;
                        k2[c] = stage >= 3 ? KB[2 * SS + row] : 0.f;
                        lm[c] = LAM[row];
                        xnext[c] = (stage > 0 && valid)
                                       ? __ldg(a_tape + (long long)(n * 4 + stage - 1) * tape_stage + (long long)(3 * r + c) * B + b0 + b) : 0.f;
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        const float xb = xbv[c];
                        const float u = xb * dt;
                        if (stage == 3) {            // x4 = y + dt*((k1 - k2) + k3)
                            KB[2 * SS + row] = k2[c] + u; KB[0 * SS + row] = k0[c] + u; KB[1 * SS + row] = k1[c] - u;
                        } else if (stage == 2) {     // x3 = y + dt*(k2 - k1/3)
                            KB[1 * SS + row] = k1[c] + u; KB[0 * SS + row] = k0[c] - u * (1.0f / 3.0f);
                        } else if (stage == 1) {     // x2 = y + (dt*k1)/3
                            KB[0 * SS + row] = k0[c] + (xb * (1.0f / 3.0f)) * dt;
                        }
                        LAM[row] = lm[c] + xb;
                        if (stage > 0) smem[P.off_X + row] = xnext[c];     // next stage input (stage-1), from the tape
                    }
                }
                UDE_PROF_WORK_DONE();
                __syncthreads();
                UDE_PROF_PHASE_END(21);
            }
        }
        // ---- output 0 (= y0 itself, already in g_y0) + lambda_0; then the hoisted constants' share
        if (decoded) dec_param_grad(P, a.dec, dpart, smem + P.off_X, nullptr, 0, 1, 0.f, B, b0, nvalid);      // X = stage-1 input of step 0 = y0
        for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            if (b < nvalid) {
                float* dst = a_g_y0 + ((b0 + b) * P.R + r) * (long long)L;
                if (decoded) {
                    float gl[3];
                    dec_cotangent(P, a.dec, g_reg, 0, B, b0 + b, r, gl);
#pragma unroll
                    for (int l = 0; l < 3; ++l) dst[l] = LAM[(3 * r + l) * TBP + b] + gl[l];
                } else {
#pragma unroll
                    for (int l = 0; l < 3; ++l) dst[l] = LAM[(3 * r + l) * TBP + b] + dst[l];
                }
            }
        }
        __syncthreads();
        if (ring && P.c0_tm) dl_tm_dump(P, smem, P.off_DL, tmem, warp, lane);      // D: tensor memory -> rows behind the constants
        if (bayes) {        // the hoisted adjoint ran per evaluation: only the summed dL/dx_const is left to hand over
            if (P.hoist) {
                const int Lc = L - 3;
                for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    if (b < nvalid) {
                        float* dst = a_g_y0 + ((b0 + b) * P.R + r) * (long long)L;
                        const float* src = smem + P.off_GXC + (r * Lc) * TBP + b;
                        for (int l = 3; l < L; ++l) dst[l] += src[(l - 3) * TBP];
                    }
                }
                __syncthreads();
            }
        } else
        hoist_backward_epilogue(P, a_wpacked, smem, Gslab, a_y0, a_g_y0, b0, nvalid, warp, lane);
        if (a.rp.mean) {        // fused reparam: dL/dz rows from the finished dL/dy0 rows (same thread wrote them)
            const int Lm1 = L - 1;
            for (int idx = threadIdx.x; idx < P.R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                if (b < nvalid) {
                    const long long traj = b0 + b;
                    const float* g = a_g_y0 + (traj * P.R + r) * (long long)L;
                    float* gz = a.rp.g_z + (traj * P.R + r) * (long long)Lm1;
                    const float z0 = reparam_z(a.rp, traj, r, P.R, Lm1, 0), z1 = reparam_z(a.rp, traj, r, P.R, Lm1, 1);
                    const float g0 = g[0], g1 = g[1], g2 = g[2];
                    for (int k = 0; k < Lm1; ++k) gz[k] = reparam_gz(z0, z1, g0, g1, g2, k >= 2 ? g[k + 1] : 0.f, k);
                }
            }
        }
        if (grouped && SMEM_G) {
            for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) Gslab[i] = G[i];
        }
    }
    __syncthreads();
    if (SMEM_G && !grouped)
        for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) Gslab[i] = G[i];
    if (tm_acc) {
        dw_tm_flush(P, smem + P.off_GB, Gslab, tmem, warp, lane);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) tm_free512(tmem);
    }
    if (ring) wr_drain(wr);
}

// ===================================================================================
// small utility kernels
// ===================================================================================
// Column of theta's W0 row that feeds packed column `col` of the live (c = 0) / constant (c = 1) block of a first layer:
// live d = 3r + c  <-  u = r*L + c;   constant e = r*(L-3) + k  <-  u = r*L + 3 + k   (u = index into vec(x), lib/models.py:119)
__device__ __forceinline__ int first_layer_src_col(const ModelP& P, int col, bool constant) {
    if (!constant) { const int r = col / 3; return r * P.L + (col - 3 * r); }
    const int Lc = P.L - 3, r = col / Lc;
    return r * P.L + 3 + (col - Lc * r);
}

// theta (PyTorch layout) -> packed blob (padded rows, zero fill).  One block per layer, then one per constant
// block of the first layers (ModelP: "Hoisting"), then the tail padding.
__global__ void pack_theta_kernel(const __grid_constant__ ModelP P, const float* __restrict__ theta_all, float* __restrict__ packed_all) {
    const float* __restrict__ theta = theta_all + (long long)blockIdx.y * P.n_theta;          // weight set blockIdx.y
    float* __restrict__ packed = packed_all + (long long)blockIdx.y * P.n_packed;
    const int li = blockIdx.x, nl = P.n_rate + P.n_aug;
    if (li == nl + 2) {       // tail padding
        for (int i = threadIdx.x; i < kPadTail; i += blockDim.x) packed[P.n_packed - kPadTail + i] = 0.f;
        return;
    }
    const bool constant = li >= nl;
    if (constant && (!P.hoist || (li == nl ? P.n_rate == 0 : P.n_aug == 0))) return;
    const LayerP& Lp = constant ? (li == nl ? P.rate_c : P.aug_c) : (li < P.n_rate ? P.rate[li] : P.aug[li - P.n_rate]);
    const bool first = constant || li == 0 || li == P.n_rate;
    const int nW = Lp.n_out_p * Lp.ldw;
    for (int i = threadIdx.x; i < nW; i += blockDim.x) {
        const int row = i / Lp.ldw, col = i - row * Lp.ldw;
        float v = 0.f;
        if (row < Lp.n_out && col < Lp.n_in)
            v = theta[Lp.th_w + row * Lp.th_ld + (first && P.hoist ? first_layer_src_col(P, col, constant) : col)];
        packed[Lp.w_off + i] = v;
    }
    if (!constant)
        for (int i = threadIdx.x; i < Lp.n_out_p; i += blockDim.x)
            packed[Lp.b_off + i] = i < Lp.n_out ? theta[Lp.th_b + i] : 0.f;
}

// g_theta[e] = sum over CTA slabs (fixed order => deterministic) of the packed accumulators.
__global__ void reduce_dtheta_kernel(const __grid_constant__ ModelP P, const float* __restrict__ slabs_all, int n_slabs,
                                     float* __restrict__ g_theta_all) {
    // weight set blockIdx.y owns slabs [y * n_slabs, (y + 1) * n_slabs)
    const float* __restrict__ slabs = slabs_all + (long long)blockIdx.y * n_slabs * P.n_packed;
    float* __restrict__ g_theta = g_theta_all + (long long)blockIdx.y * P.n_theta;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.n_theta) return;
    int pidx = -1;
    const int nl = P.n_rate + P.n_aug;
    for (int li = 0; li < nl; ++li) {
        const LayerP& Lp = li < P.n_rate ? P.rate[li] : P.aug[li - P.n_rate];
        const bool first = li == 0 || li == P.n_rate;
        if (e >= Lp.th_w && e < Lp.th_w + Lp.th_ld * Lp.n_out) {
            const int k = e - Lp.th_w;
            const int row = k / Lp.th_ld, u = k - row * Lp.th_ld;
            if (first && P.hoist) {
                const int r = u / P.L, l = u - r * P.L;
                const LayerP& Lc = li == 0 && P.n_rate > 0 ? P.rate_c : P.aug_c;
                pidx = l < 3 ? Lp.w_off + row * Lp.ldw + 3 * r + l : Lc.w_off + row * Lc.ldw + r * (P.L - 3) + (l - 3);
            } else {
                pidx = Lp.w_off + row * Lp.ldw + u;
            }
            break;
        }
        if (e >= Lp.th_b && e < Lp.th_b + Lp.n_out) {
            pidx = Lp.b_off + (e - Lp.th_b);
            break;
        }
    }
    float s = 0.f;
    if (pidx >= 0)
        for (int c = 0; c < n_slabs; ++c) s += slabs[(long long)c * P.n_packed + pidx];
    g_theta[e] = s;
}

// side[0..7] = sum over CTA partials (fixed order).
__global__ void reduce_side_kernel(const double* __restrict__ partial_all, int n_parts, double* __restrict__ side_all) {
    const double* __restrict__ partial = partial_all + (long long)blockIdx.x * n_parts * 8;    // weight set blockIdx.x
    double* __restrict__ side = side_all + (long long)blockIdx.x * 8;
    const int i = threadIdx.x;
    if (i >= 8) return;
    double s = 0.0;
    for (int c = 0; c < n_parts; ++c) s += partial[c * 8 + i];
    if (i == 6) s = partial[6];   // RHS evaluations per trajectory: identical in every CTA
    side[i] = s;
}

// decoder gradients: g_dec[e] = sum over the CTAs' partial blocks (fixed order)
__global__ void reduce_dec_kernel(const float* __restrict__ partial, int n_parts, int n, float* __restrict__ g_dec) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int c = 0; c < n_parts; ++c) s += partial[(long long)c * n + e];
    g_dec[e] = s;
}

// Register-resident FFMA chains: the measured fp32 roofline denominator.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i) * 1e-3f;
    const float m = 0.999f + seed * 1e-6f, c = 1e-4f * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;       // never true; keeps the chain alive
}

}  // namespace ude
