// ude_dopri5.cuh — fused adaptive Dormand-Prince 5(4) forward and discrete-adjoint kernels.
//
// Restates torchdiffeq's RKAdaptiveStepsizeODESolver / Dopri5Solver (rk_common.py, dopri5.py,
// interp.py, misc.py; the solver the reference reaches through odeint's default method,
// /root/reference/lib/train_functions.py:190) with ONE controller PER TRAJECTORY: every
// trajectory of a tile carries its own (t, dt), error norm and accept/reject decision; the
// tile advances in lock step until all of its trajectories have produced all outputs.
// Time keeping is fp64, the state fp32 — as in torchdiffeq.
//
// Tape (for the adjoint): per accepted step s and trajectory b
//   tx[s][i][d][b]  stage inputs x_1..x_6 (x_1 = y_s); step index n_accept holds the final state
//   tt[s][b]        (t0, dt) as two fp64
#pragma once
#include <cooperative_groups.h>

#include "ude_engine.cuh"
#include "ude_rk4.cuh"

namespace ude {
namespace cg = cooperative_groups;

constexpr int kDopriStages = 6;          // stage inputs stored per accepted step
constexpr int kMaxTB = 64;

// Butcher tableau (fp32, like torchdiffeq casts it to the state dtype)
__constant__ float c_beta[7][6] = {
    {0.f, 0.f, 0.f, 0.f, 0.f, 0.f},
    {(float)(1.0 / 5), 0.f, 0.f, 0.f, 0.f, 0.f},
    {(float)(3.0 / 40), (float)(9.0 / 40), 0.f, 0.f, 0.f, 0.f},
    {(float)(44.0 / 45), (float)(-56.0 / 15), (float)(32.0 / 9), 0.f, 0.f, 0.f},
    {(float)(19372.0 / 6561), (float)(-25360.0 / 2187), (float)(64448.0 / 6561), (float)(-212.0 / 729), 0.f, 0.f},
    {(float)(9017.0 / 3168), (float)(-355.0 / 33), (float)(46732.0 / 5247), (float)(49.0 / 176), (float)(-5103.0 / 18656), 0.f},
    {(float)(35.0 / 384), 0.f, (float)(500.0 / 1113), (float)(125.0 / 192), (float)(-2187.0 / 6784), (float)(11.0 / 84)},
};
__constant__ float c_err[7] = {
    (float)(35.0 / 384 - 1951.0 / 21600), 0.f, (float)(500.0 / 1113 - 22642.0 / 50085), (float)(125.0 / 192 - 451.0 / 720),
    (float)(-2187.0 / 6784 - -12231.0 / 42400), (float)(11.0 / 84 - 649.0 / 6300), (float)(-1.0 / 60)};
__constant__ float c_mid[7] = {
    (float)(6025192743.0 / 30085553152.0 / 2), 0.f, (float)(51252292925.0 / 65400821598.0 / 2),
    (float)(-2691868925.0 / 45128329728.0 / 2), (float)(187940372067.0 / 1594534317056.0 / 2),
    (float)(-1776094331.0 / 19743644256.0 / 2), (float)(11237099.0 / 235043384.0 / 2)};

struct DopriOpts {
    float rtol, atol;
    double first_step, safety, ifactor, dfactor;
    int max_steps, max_attempts;
    int shared_step;        // torchdiffeq's batch semantics: ONE error norm (RMS over the whole [B,R,L] tensor), one dt and one
                            // accept / reject decision for all trajectories; the grid is launched cooperatively (one tile per CTA)
                            // and the norms are summed across CTAs (warp shuffles -> CTA -> global partials -> grid barrier)
};

struct DopriFwdArgs {
    const float* wpacked;
    const float* y0;
    int n_out;
    const double* t_out;
    float* y_out;
    float* tape_x;          // [(max_steps+1), 6, Dn, B] or null
    double* tape_t;         // [max_steps, B, 2]
    int* counts;            // [B,4]
    double* side_partial;
    float* cold;
    long long B;
    DopriOpts o;
    double* gpart;          // shared_step: [2][gridDim.x][2] partial sums of the batch-wide norms
};

struct DopriBwdArgs {
    const float* wpacked;
    const float* y0;
    int n_out;
    const double* t_out;
    const float* tape_x;
    const double* tape_t;
    const int* counts;
    const float* g_out;
    const float* g_side;
    float* g_y0;
    float* slabs;
    float* cold;
    long long B;
    DopriOpts o;
};

// state buffers ([Dn_p][TBP] each)
enum { DF_Y = 0, DF_K1 = 1, DF_COEF = 8, DF_NSTATE = 13 };            // K1..K7 = 1..7, COEF e,d,c,b,a = 8..12
enum { DF_RB_PART = 0, DF_RB_SA = 1, DF_NRB = 6 };                    // [R_p][TBP]: error partials, 5 side accumulators
enum { DB_LAM1 = 0, DB_Y0B = 1, DB_KB1 = 2, DB_NSTATE = 9 };          // KB1..KB7 = 2..8

// y + sum_j k_j * (dt * w_j), accumulated left to right, zero weights skipped (oracle _weighted_sum).
// The k rows may live in the CTA's global scratch (wide models): ALL loads of the sum are issued before the first use, so one
// L2 round trip is paid per call instead of one per term (the loads were 30 % of the forward kernel's warp samples).
template <int N>
__device__ __forceinline__ float rk_combine_n(const float* __restrict__ K, int SS, int row, const float* w, float dt, float y) {
    float kv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) kv[j] = K[j * SS + row];
    float acc = 0.f;
    bool first = true;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const float wj = w[j];
        if (wj == 0.f) continue;
        const float term = __fmul_rn(kv[j], __fmul_rn(dt, wj));
        acc = first ? term : __fadd_rn(acc, term);
        first = false;
    }
    return __fadd_rn(y, acc);
}
__device__ __forceinline__ float rk_combine(const float* __restrict__ K, int SS, int row, const float* w, int n, float dt,
                                            float y) {
    switch (n) {
        case 1: return rk_combine_n<1>(K, SS, row, w, dt, y);
        case 2: return rk_combine_n<2>(K, SS, row, w, dt, y);
        case 3: return rk_combine_n<3>(K, SS, row, w, dt, y);
        case 4: return rk_combine_n<4>(K, SS, row, w, dt, y);
        case 5: return rk_combine_n<5>(K, SS, row, w, dt, y);
        case 6: return rk_combine_n<6>(K, SS, row, w, dt, y);
        default: return rk_combine_n<7>(K, SS, row, w, dt, y);
    }
}

// ===================================================================================
// forward
// ===================================================================================
template <bool SMEM_W, bool COLD>
__global__ void __launch_bounds__(kThreads, 1) dopri5_fwd_kernel(const __grid_constant__ ModelP Pg, const DopriFwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ ModelP P;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ double red[kWarps * 8];
    __shared__ double s_t0[kMaxTB], s_t1[kMaxTB], s_dt[kMaxTB];
    __shared__ float s_dtold[kMaxTB], s_h0[kMaxTB];
    __shared__ int s_j[kMaxTB], s_nacc[kMaxTB], s_nrej[kMaxTB], s_natt[kMaxTB], s_status[kMaxTB], s_acc[kMaxTB];
    __shared__ int s_nint[kMaxTB];          // attempts since the last emitted output: torchdiffeq's max_num_steps is per output interval
    __shared__ __align__(8) uint64_t rbar[2];

    copy_model_to_smem(&P, Pg);
    if (threadIdx.x == 0 && SMEM_W) mbar_init(&wbar, 1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TB = P.TB, TBP = P.TBP, L = P.L, Dn = P.Dn, R = P.R;
    const bool ring = COLD && P.w_ring;          // see rk4_fwd_kernel
    WRing wr;
    if (ring) ring_setup(wr, P, smem, rbar, a.wpacked);
    uint32_t wphase = 0;
    const float* Wp = a.wpacked;
    if (SMEM_W) {
        stage_weights_tma(smem + P.off_W, a.wpacked, P.n_packed_hot, &wbar, wphase);
        Wp = smem + P.off_W;
    }
    for (int i = threadIdx.x; i < (P.Dn_p - P.Dn) * TBP; i += kThreads) smem[P.off_X + P.Dn * TBP + i] = 0.f;

    const int SS = P.Dn_p * TBP;
    float* SB = COLD ? a.cold + (long long)blockIdx.x * P.cold_floats : smem;
    float* SY = SB + P.off_S + DF_Y * SS;
    float* SK = SB + P.off_S + DF_K1 * SS;          // SK + i*SS = k_{i+1}
    float* CO = SB + P.off_S + DF_COEF * SS;        // e, d, c, b, a
    const int RS = P.R_p * TBP;
    float* PART = SB + P.off_RB + DF_RB_PART * RS;
    float* SA = SB + P.off_RB + DF_RB_SA * RS;       // 5 x [R_p][TBP]

    SideAcc side;
    side.clear();
    const long long B = a.B;
    const long long n_tiles = (B + TB - 1) / TB;
    const float rtol = a.o.rtol, atol = a.o.atol;
    const float inv_n = 1.0f / (float)(R * L);
    const long long stage_sz = (long long)Dn * B;
    // ---- shared-step mode: batch-wide sums of two per-trajectory quantities (threads >= TB contribute 0).  Fixed summation
    // order everywhere (shuffle tree, warps in order, CTAs in order): every CTA computes the SAME bits and takes the same decision.
    const bool shared = a.o.shared_step != 0;
    const double inv_nB = 1.0 / ((double)B * R * L);
    __shared__ double s_tot[2];
    int gphase = 0;
    auto grid_sum2 = [&](double v0, double v1, double& t0, double& t1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        if (lane == 0) { red[2 * warp] = v0; red[2 * warp + 1] = v1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double c0 = 0.0, c1 = 0.0;
            for (int w = 0; w < kWarps; ++w) { c0 += red[2 * w]; c1 += red[2 * w + 1]; }
            double* dst = a.gpart + ((size_t)gphase * gridDim.x + blockIdx.x) * 2;
            dst[0] = c0; dst[1] = c1;
            __threadfence();
        }
        cg::this_grid().sync();
        if (threadIdx.x == 0) {
            const volatile double* src = a.gpart + (size_t)gphase * gridDim.x * 2;
            double c0 = 0.0, c1 = 0.0;
            for (unsigned c = 0; c < gridDim.x; ++c) { c0 += src[2 * c]; c1 += src[2 * c + 1]; }
            s_tot[0] = c0; s_tot[1] = c1;
        }
        __syncthreads();
        t0 = s_tot[0]; t1 = s_tot[1];
        gphase ^= 1;
        __syncthreads();
    };

    // one RHS evaluation at X; `record`: 0 none, 1 straight into the thread's accumulators, 2 into SA (attempt)
    auto eval_into = [&](float* Kdst, int record, int nvalid) {
        if (shared) record = 1;     // the reference's lists hold EVERY evaluation: probe and rejected attempts included
        if (ring) rhs_forward_layers_ring(P, wr, smem, warp, lane, 0u);
        else rhs_forward_layers(P, Wp, smem, warp, lane);
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            const bool valid = b < nvalid;
            float f[3];
            SideAcc tmp;
            tmp.clear();
            rhs_tail(P, smem, r, b, valid, tmp, f);
            Kdst[(3 * r + 0) * TBP + b] = f[0];
            Kdst[(3 * r + 1) * TBP + b] = f[1];
            Kdst[(3 * r + 2) * TBP + b] = f[2];
            if (record == 1) {
                side.sb += tmp.sb; side.sg += tmp.sg; side.sb2 += tmp.sb2; side.sg2 += tmp.sg2; side.sf2 += tmp.sf2;
                side.rows += tmp.rows;
            } else if (record == 2) {
                const int o = r * TBP + b;
                SA[0 * RS + o] += (float)tmp.sb; SA[1 * RS + o] += (float)tmp.sg;
                SA[2 * RS + o] += (float)tmp.sb2; SA[3 * RS + o] += (float)tmp.sg2; SA[4 * RS + o] += (float)tmp.sf2;
            }
        }
        __syncthreads();
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * TB;
        const int nvalid = (int)((B - b0) < TB ? (B - b0) : TB);
        __syncthreads();
        load_y0_tile(P, smem, a.y0, b0, nvalid, true, P.hoist != 0);
        for (int i = threadIdx.x; i < 5 * RS; i += kThreads) SA[i] = 0.f;
        if (threadIdx.x < TB) {
            const int b = threadIdx.x;
            s_t0[b] = s_t1[b] = a.t_out[0];
            s_dt[b] = a.o.first_step;
            s_j[b] = b < nvalid ? 1 : a.n_out;
            s_nacc[b] = s_nrej[b] = s_natt[b] = s_status[b] = 0;
            s_acc[b] = 0;
            s_nint[b] = 0;
        }
        __syncthreads();
        if (ring) hoist_c0_ring(P, a.wpacked, smem, warp, lane, 0u);
        else hoist_c0(P, a.wpacked, smem, warp, lane);
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            const float* X = smem + P.off_X + (3 * r) * TBP + b;
            float v[3] = {X[0], X[TBP], X[2 * TBP]};
            for (int c = 0; c < 3; ++c) SY[(3 * r + c) * TBP + b] = v[c];
            if (b < nvalid) {
                const long long row = ((b0 + b) * R + r) * (long long)L;
                store_out_row(P, a.y0 + row, a.y_out + row, v);
            }
        }
        __syncthreads();
        eval_into(SK, 1, nvalid);                                   // f0 = func(t0, y0): first stage of the first step

        if (!(a.o.first_step > 0.0)) {
            // ---- torchdiffeq _select_initial_step (order p = 4), per trajectory
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                const float* X = smem + P.off_X + (3 * r) * TBP + b;
                const float* y0row = a.y0 + ((b0 + b) * R + r) * (long long)L;      // constants: straight from the caller's y0
                float s0 = 0.f, s1 = 0.f;
                for (int l = 0; l < L; ++l) {
                    const float y = l < 3 ? X[l * TBP] : (b < nvalid ? __ldg(y0row + l) : 0.f);
                    const float scale = __fadd_rn(atol, __fmul_rn(fabsf(y), rtol));
                    const float q0 = y / scale;
                    s0 += q0 * q0;
                    if (l < 3) {
                        const float q1 = SK[(3 * r + l) * TBP + b] / scale;
                        s1 += q1 * q1;
                    }
                }
                PART[r * TBP + b] = s0;
                SA[0 * RS + r * TBP + b] = s1;          // SA is free until the first attempt
            }
            __syncthreads();
            float s0 = 0.f, s1 = 0.f;
            if (threadIdx.x < TB)
                for (int r = 0; r < R; ++r) { s0 += PART[r * TBP + threadIdx.x]; s1 += SA[r * TBP + threadIdx.x]; }
            double g0 = 0.0, g1 = 0.0;
            if (shared) grid_sum2(threadIdx.x < nvalid ? (double)s0 : 0.0, threadIdx.x < nvalid ? (double)s1 : 0.0, g0, g1);
            if (threadIdx.x < TB) {
                const int b = threadIdx.x;
                const float d0 = shared ? sqrtf((float)(g0 * inv_nB)) : sqrtf(s0 * inv_n);
                const float d1 = shared ? sqrtf((float)(g1 * inv_nB)) : sqrtf(s1 * inv_n);
                const float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : fabsf(0.01f * d0 / d1);
                s_h0[b] = h0;
                s_dtold[b] = d1;
            }
            __syncthreads();
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                for (int c = 0; c < 3; ++c)
                    smem[P.off_X + (3 * r + c) * TBP + b] =
                        __fadd_rn(SY[(3 * r + c) * TBP + b], __fmul_rn(s_h0[b], SK[(3 * r + c) * TBP + b]));
            }
            __syncthreads();
            eval_into(SK + 1 * SS, 0, nvalid);                      // f1 = func(t0 + h0, y0 + h0 f0): probe, not tracked
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                float s2 = 0.f;
                for (int c = 0; c < 3; ++c) {
                    const float y = SY[(3 * r + c) * TBP + b];
                    const float scale = __fadd_rn(atol, __fmul_rn(fabsf(y), rtol));
                    const float q = __fsub_rn(SK[1 * SS + (3 * r + c) * TBP + b], SK[(3 * r + c) * TBP + b]) / scale;
                    s2 += q * q;
                }
                PART[r * TBP + b] = s2;
                SA[0 * RS + r * TBP + b] = 0.f;
                for (int c = 0; c < 3; ++c) smem[P.off_X + (3 * r + c) * TBP + b] = SY[(3 * r + c) * TBP + b];
            }
            __syncthreads();
            float s2 = 0.f;
            if (threadIdx.x < TB)
                for (int r = 0; r < R; ++r) s2 += PART[r * TBP + threadIdx.x];
            double g2 = 0.0, g2b = 0.0;
            if (shared) grid_sum2(threadIdx.x < nvalid ? (double)s2 : 0.0, 0.0, g2, g2b);
            if (threadIdx.x < TB) {
                const int b = threadIdx.x;
                const float h0 = s_h0[b], d1 = s_dtold[b];
                const float d2 = fabsf((shared ? sqrtf((float)(g2 * inv_nB)) : sqrtf(s2 * inv_n)) / h0);
                float h1;
                if (d1 <= 1e-15f && d2 <= 1e-15f) h1 = fmaxf(1e-6f, h0 * 1e-3f);
                else h1 = powf(0.01f / fmaxf(d1, d2), 1.0f / 5.0f);
                s_dt[b] = (double)fminf(100.f * h0, fabsf(h1));
            }
            __syncthreads();
        }

        // ---- main loop: emit what the current interpolant covers, then attempt one step
        while (true) {
            // (1) outputs t_out[j] <= t1  (torchdiffeq _advance: step only while next_t > t1)
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                if (b >= nvalid) continue;
                int jj = s_j[b];
                const double t0 = s_t0[b], t1 = s_t1[b];
                while (jj < a.n_out && !(a.t_out[jj] > t1)) {
                    const float x = (float)((a.t_out[jj] - t0) / (t1 - t0));
                    float v[3];
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        float total = __fadd_rn(CO[0 * SS + row], __fmul_rn(x, CO[1 * SS + row]));
                        float xp = x;
                        for (int p = 2; p < 5; ++p) {
                            xp = __fmul_rn(xp, x);
                            total = __fadd_rn(total, __fmul_rn(xp, CO[p * SS + row]));
                        }
                        v[c] = s_status[b] == 0 ? total : __int_as_float(0x7fc00000);
                    }
                    store_out_row(P, a.y0 + ((b0 + b) * R + r) * (long long)L,
                                  a.y_out + (((long long)jj * B + b0 + b) * R + r) * (long long)L, v);
                    ++jj;
                }
            }
            __syncthreads();
            int pending = 0;
            if (threadIdx.x < TB) {
                const int b = threadIdx.x;
                int jj = s_j[b];
                while (jj < a.n_out && !(a.t_out[jj] > s_t1[b])) ++jj;
                if (jj != s_j[b]) s_nint[b] = 0;          // an output was reached: the step budget of the next interval starts over
                if (jj < a.n_out) {
                    if (s_nint[b] >= a.o.max_attempts) { s_status[b] = 2; }
                    else if (a.tape_x && s_nacc[b] >= a.o.max_steps) { s_status[b] = 1; }
                    else if (!(s_t1[b] + s_dt[b] > s_t1[b]) || !isfinite(s_dt[b])) { s_status[b] = 3; }
                    if (s_status[b] != 0) {
                        // failed trajectory: flush the remaining outputs as NaN through the emit path
                        s_t1[b] = 1e300;
                        pending = 1;
                    } else {
                        pending = 1;
                        s_natt[b] += 1;
                        s_nint[b] += 1;
                    }
                }
                s_j[b] = jj;
            }
            if (!__syncthreads_or(pending)) break;
            // (2) one attempt for every trajectory of the tile (finished ones are masked)
            for (int st = 1; st <= 6; ++st) {
                for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    const float dt = (float)s_dt[b];
                    float ys[3], xs[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) ys[c] = SY[(3 * r + c) * TBP + b];
#pragma unroll
                    for (int c = 0; c < 3; ++c) xs[c] = rk_combine(SK, SS, (3 * r + c) * TBP + b, c_beta[st], st, dt, ys[c]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) smem[P.off_X + (3 * r + c) * TBP + b] = xs[c];
                }
                __syncthreads();
                eval_into(SK + st * SS, 2, nvalid);
            }
            // (3) error norm, per trajectory
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                const float dt = (float)s_dt[b];
                float s = 0.f;
                for (int c = 0; c < 3; ++c) {
                    const int row = (3 * r + c) * TBP + b;
                    const float y0v = SY[row], y1v = smem[P.off_X + row];
                    const float err = __fsub_rn(rk_combine(SK, SS, row, c_err, 7, dt, 0.f), 0.f);
                    const float tol = __fadd_rn(atol, __fmul_rn(rtol, fmaxf(fabsf(y0v), fabsf(y1v))));
                    const float q = err / tol;
                    s += q * q;
                }
                PART[r * TBP + b] = s;
            }
            __syncthreads();
            float s_err = 0.f;
            if (threadIdx.x < TB)
                for (int r = 0; r < R; ++r) s_err += PART[r * TBP + threadIdx.x];
            double ge = 0.0, geb = 0.0;
            if (shared) grid_sum2(threadIdx.x < nvalid ? (double)s_err : 0.0, 0.0, ge, geb);
            if (threadIdx.x < TB) {
                const int b = threadIdx.x;
                const bool active = s_j[b] < a.n_out && s_status[b] == 0;
                const float ratio = shared ? sqrtf((float)(ge * inv_nB)) : sqrtf(s_err * inv_n);
                const bool accept = active && (ratio <= 1.0f);
                const double dt = s_dt[b];
                double dt_next;
                if (ratio == 0.f) dt_next = dt * a.o.ifactor;
                else {
                    const double dfac = ratio < 1.0f ? 1.0 : a.o.dfactor;
                    const double fac = fmin(a.o.ifactor, fmax(a.o.safety / pow((double)ratio, 0.2), dfac));
                    dt_next = dt * fac;
                }
                s_acc[b] = accept ? 1 : 0;
                s_dtold[b] = (float)dt;
                if (accept) {
                    if (a.tape_t) {
                        a.tape_t[((long long)s_nacc[b] * B + b0 + b) * 2 + 0] = s_t1[b];
                        a.tape_t[((long long)s_nacc[b] * B + b0 + b) * 2 + 1] = dt;
                    }
                    s_t0[b] = s_t1[b];
                    s_t1[b] = s_t1[b] + dt;
                } else if (active) {
                    s_nrej[b] += 1;
                }
                if (active) s_dt[b] = dt_next;
            }
            __syncthreads();
            // (4) commit accepted steps: interpolant, tape, FSAL, side channels
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                const int o = r * TBP + b;
                if (s_acc[b]) {
                    const float dt = s_dtold[b];
                    const int step = s_nacc[b];
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        const float y0v = SY[row], y1v = smem[P.off_X + row];
                        const float f0 = SK[row], f1 = SK[6 * SS + row];
                        if (a.tape_x) {
                            float* tp = a.tape_x + ((long long)step * kDopriStages) * stage_sz + (long long)(3 * r + c) * B + b0 + b;
                            tp[0] = y0v;
                            for (int st = 1; st < 6; ++st)
                                tp[(long long)st * stage_sz] = rk_combine(SK, SS, row, c_beta[st], st, dt, y0v);
                        }
                        const float ymid = rk_combine(SK, SS, row, c_mid, 7, dt, y0v);
                        // interp._interp_fit, operation order as written there
                        const float ca = __fadd_rn(__fsub_rn(__fmul_rn(__fmul_rn(2.f, dt), __fsub_rn(f1, f0)),
                                                             __fmul_rn(8.f, __fadd_rn(y1v, y0v))), __fmul_rn(16.f, ymid));
                        const float cb = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.f, f0), __fmul_rn(3.f, f1))),
                                                                       __fmul_rn(18.f, y0v)), __fmul_rn(14.f, y1v)), __fmul_rn(32.f, ymid));
                        const float cc = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(f1, __fmul_rn(4.f, f0))),
                                                                       __fmul_rn(11.f, y0v)), __fmul_rn(5.f, y1v)), __fmul_rn(16.f, ymid));
                        CO[0 * SS + row] = y0v;
                        CO[1 * SS + row] = __fmul_rn(dt, f0);
                        CO[2 * SS + row] = cc;
                        CO[3 * SS + row] = cb;
                        CO[4 * SS + row] = ca;
                        SY[row] = y1v;
                        SK[row] = f1;                               // FSAL
                    }
                    if (b < nvalid && !shared) {
                        side.sb += SA[0 * RS + o]; side.sg += SA[1 * RS + o]; side.sb2 += SA[2 * RS + o];
                        side.sg2 += SA[3 * RS + o]; side.sf2 += SA[4 * RS + o];
                        side.rows += 6;
                    }
                }
                for (int q = 0; q < 5; ++q) SA[q * RS + o] = 0.f;
            }
            __syncthreads();
            if (threadIdx.x < TB && s_acc[threadIdx.x]) s_nacc[threadIdx.x] += 1;
            __syncthreads();
        }
        // ---- per-trajectory bookkeeping + final state on the tape
        if (threadIdx.x < nvalid) {
            const int b = threadIdx.x;
            int* cnt = a.counts + (b0 + b) * 4;
            cnt[0] = s_nacc[b]; cnt[1] = s_nrej[b]; cnt[2] = s_status[b]; cnt[3] = s_natt[b];
        }
        if (a.tape_x) {
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                if (b >= nvalid) continue;
                for (int c = 0; c < 3; ++c)
                    a.tape_x[((long long)s_nacc[b] * kDopriStages) * stage_sz + (long long)(3 * r + c) * B + b0 + b] =
                        SY[(3 * r + c) * TBP + b];
            }
        }
    }
    side_block_reduce(side, 0.0, red, a.side_partial + (long long)blockIdx.x * 8);
    if (ring) wr_drain(wr);
}

// ===================================================================================
// backward: discrete adjoint over each trajectory's accepted-step tape
// ===================================================================================
template <bool SMEM_W, bool SMEM_G, bool COLD>
__global__ void __launch_bounds__(kThreads, 1) dopri5_bwd_kernel(const __grid_constant__ ModelP Pg, const DopriBwdArgs a) {
    extern __shared__ __align__(128) float smem[];
    __shared__ ModelP P;
    __shared__ __align__(8) uint64_t wbar;
    __shared__ double s_t0[kMaxTB], s_t1[kMaxTB];
    __shared__ float s_dt[kMaxTB];
    __shared__ int s_j[kMaxTB], s_nacc[kMaxTB], s_act[kMaxTB], s_max;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t rbar[2];

    copy_model_to_smem(&P, Pg);
    if (threadIdx.x == 0 && SMEM_W) mbar_init(&wbar, 1);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TB = P.TB, TBP = P.TBP, L = P.L, Dn = P.Dn, R = P.R;
    const bool ring = COLD && P.w_ring;            // see rk4_bwd_kernel
    WRing wr;
    if (ring) ring_setup(wr, P, smem, rbar, a.wpacked);
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
    float* Gslab = a.slabs + (long long)blockIdx.x * P.n_packed;      // see rk4_bwd_kernel
    float* G = SMEM_G ? smem + P.off_G : Gslab;
    for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) G[i] = 0.f;
    for (int i = P.n_packed_hot + threadIdx.x; i < P.n_packed; i += kThreads) Gslab[i] = 0.f;
    for (int i = threadIdx.x; i < (P.Dn_p - P.Dn) * TBP; i += kThreads) smem[P.off_X + P.Dn * TBP + i] = 0.f;
    float gs[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (a.g_side) {
#pragma unroll
        for (int i = 0; i < 5; ++i) gs[i] = __ldg(a.g_side + i);
    }
    const int SS = P.Dn_p * TBP;
    float* SB = COLD ? a.cold + (long long)blockIdx.x * P.cold_floats : smem;
    float* LAM1 = SB + P.off_S + DB_LAM1 * SS;
    float* Y0B = SB + P.off_S + DB_Y0B * SS;
    float* KB = SB + P.off_S + DB_KB1 * SS;          // KB + i*SS = kbar_{i+1}
    float* GX = smem + P.off_GX;
    float* DL = smem + P.off_DL;
    const long long B = a.B;
    const long long n_tiles = (B + TB - 1) / TB;
    const long long stage_sz = (long long)Dn * B;

    // VJP of one RHS evaluation at X with cotangent KB[ki]; afterwards GX holds dL/dx (all rows)
    auto vjp_at_x = [&](int ki, int nvalid) {
        if (ring) rhs_forward_layers_ring(P, wr, smem, warp, lane, tmem);
        else rhs_forward_layers(P, Wp, smem, warp, lane);
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            const int row = (3 * r) * TBP + b;
            float xb[3];
            rhs_tail_vjp(P, smem, r, b, (b < nvalid) && s_act[b], gs, KB[ki * SS + row], KB[ki * SS + row + TBP],
                         KB[ki * SS + row + 2 * TBP], xb);
            float* gx = GX + (3 * r) * TBP + b;
            gx[0] = xb[0]; gx[TBP] = xb[1]; gx[2 * TBP] = xb[2];
        }
        __syncthreads();
        if (ring) rhs_backward_layers_ring(P, wr, smem, smem + P.off_GB, warp, lane, tmem);
        else if (tm_acc) rhs_backward_layers<true>(P, Wp, smem, smem + P.off_GB, warp, lane, tmem);
        else rhs_backward_layers<false>(P, Wp, smem, G, warp, lane);
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long b0 = tile * TB;
        const int nvalid = (int)((B - b0) < TB ? (B - b0) : TB);
        __syncthreads();
        // live rows too: trajectories that are not active in a step (fewer accepted steps than the tile's maximum) and the
        // invalid columns of a ragged tile still run through the MLPs and must see FINITE inputs (0 * NaN = NaN)
        load_y0_tile(P, smem, a.y0, b0, nvalid, true, P.hoist != 0);
        for (int i = threadIdx.x; i < DB_NSTATE * SS; i += kThreads) SB[P.off_S + i] = 0.f;
        for (int i = threadIdx.x; i < P.Dn_p * TBP; i += kThreads) GX[i] = 0.f;
        if (ring && P.c0_tm) dl_tm_clear(P, tmem, warp);
        else
            for (int i = threadIdx.x; i < P.dl_rows * TBP; i += kThreads) DL[i] = 0.f;
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {      // dL/dy0 starts from the cotangent of output 0
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            if (b < nvalid) {
                const long long row = ((b0 + b) * R + r) * (long long)L;
                for (int l = 0; l < L; ++l) a.g_y0[row + l] = __ldg(a.g_out + row + l);
            }
        }
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
        if (ring) hoist_c0_ring(P, a.wpacked, smem, warp, lane, tmem);
        else hoist_c0(P, a.wpacked, smem, warp, lane);
        if (threadIdx.x < TB) {
            const int b = threadIdx.x;
            s_nacc[b] = b < nvalid ? a.counts[(b0 + b) * 4 + 0] : 0;
            s_j[b] = a.n_out - 1;
            atomicMax(&s_max, s_nacc[b]);
        }
        __syncthreads();
        const int smax = s_max;
        for (int s = smax - 1; s >= 0; --s) {
            if (threadIdx.x < TB) {
                const int b = threadIdx.x;
                const int act = s < s_nacc[b];
                s_act[b] = act;
                if (act) {
                    const double t0 = a.tape_t[((long long)s * B + b0 + b) * 2 + 0];
                    const double dt = a.tape_t[((long long)s * B + b0 + b) * 2 + 1];
                    s_t0[b] = t0; s_t1[b] = t0 + dt; s_dt[b] = (float)dt;
                }
            }
            __syncthreads();
            // ---- outputs interpolated inside step s: cotangents of the five coefficients
            for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                const bool act = s_act[b] != 0;
                float ce[3] = {0, 0, 0}, cd[3] = {0, 0, 0}, cc[3] = {0, 0, 0}, cbv[3] = {0, 0, 0}, ca[3] = {0, 0, 0};
                const float dt = act ? s_dt[b] : 0.f;
                if (act) {
                    const double t0 = s_t0[b], t1 = s_t1[b];
                    for (int jj = s_j[b]; jj >= 1 && a.t_out[jj] > t0; --jj) {
                        const float x = (float)((a.t_out[jj] - t0) / (t1 - t0));
                        const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                        const float* gp = a.g_out + (((long long)jj * B + b0 + b) * R + r) * (long long)L;
                        float gl[3];
                        add_const_cotangent(a.g_y0 + ((b0 + b) * R + r) * (long long)L, gp, L, gl);
                        for (int l = 0; l < 3; ++l) {
                            const float g = gl[l];
                            ce[l] += g; cd[l] += x * g; cc[l] += x2 * g; cbv[l] += x3 * g; ca[l] += x4 * g;
                        }
                    }
                }
                for (int c = 0; c < 3; ++c) {
                    const int row = (3 * r + c) * TBP + b;
                    if (!act) {
                        for (int i = 0; i < 7; ++i) KB[i * SS + row] = 0.f;
                        continue;
                    }
                    const float ymid_b = 16.f * cc[c] - 32.f * cbv[c] + 16.f * ca[c];
                    const float y0_b = ce[c] - 11.f * cc[c] + 18.f * cbv[c] - 8.f * ca[c] + ymid_b;
                    const float y1_b = -5.f * cc[c] + 14.f * cbv[c] - 8.f * ca[c];
                    const float f0_b = dt * (cd[c] - 4.f * cc[c] + 5.f * cbv[c] - 2.f * ca[c]);
                    const float f1_b = dt * (cc[c] - 3.f * cbv[c] + 2.f * ca[c]);
                    const float carry = KB[0 * SS + row];           // kbar_1 of step s+1 == extra cotangent of k_7 (FSAL)
                    for (int i = 0; i < 7; ++i) KB[i * SS + row] = dt * c_mid[i] * ymid_b;
                    KB[0 * SS + row] += f0_b;
                    KB[6 * SS + row] += f1_b + carry;
                    LAM1[row] += y1_b;
                    Y0B[row] = y0_b;
                    // stage input of the k_7 evaluation: y_{s+1} = x_1 of step s+1 (or the final state)
                    smem[P.off_X + row] =
                        __ldg(a.tape_x + ((long long)(s + 1) * kDopriStages) * stage_sz + (long long)(3 * r + c) * B + b0 + b);
                }
            }
            __syncthreads();
            if (threadIdx.x < TB && s_act[threadIdx.x]) {
                const int b = threadIdx.x;
                int jj = s_j[b];
                while (jj >= 1 && a.t_out[jj] > s_t0[b]) --jj;
                s_j[b] = jj;
            }
            // ---- stage VJPs 7 .. 2
            for (int st = 6; st >= 1; --st) {
                vjp_at_x(st, nvalid);
                for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
                    const int r = idx >> P.tb_shift, b = idx & (TB - 1);
                    const bool act = s_act[b] != 0;
                    const float dt = act ? s_dt[b] : 0.f;
                    for (int c = 0; c < 3; ++c) {
                        const int row = (3 * r + c) * TBP + b;
                        const int xrow = row;
                        if (act) {
                            float xb = GX[xrow];
                            if (st == 6) {                       // x_7 = y_1: close lambda_1, then y_1 = y_0 + dt sum c_i k_i
                                xb += LAM1[row];
                                LAM1[row] = xb;
                            }
                            Y0B[row] += xb;
                            for (int jst = 0; jst < st; ++jst) {
                                const float w = c_beta[st][jst];
                                if (w != 0.f) KB[jst * SS + row] += dt * w * xb;
                            }
                            if (st == 1) LAM1[row] = Y0B[row];   // becomes lambda_1 of step s-1 (x_1 VJP is deferred: FSAL)
                            if (st > 1)
                                smem[P.off_X + xrow] = __ldg(a.tape_x + ((long long)s * kDopriStages + (st - 1)) * stage_sz +
                                                             (long long)(3 * r + c) * B + b0 + b);
                        }
                    }
                }
                __syncthreads();
            }
        }
        // ---- the very first evaluation f0 = f(y0): VJP with the carried kbar_1 of step 0
        if (threadIdx.x < TB) s_act[threadIdx.x] = s_nacc[threadIdx.x] > 0;
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            for (int c = 0; c < 3; ++c)
                smem[P.off_X + (3 * r + c) * TBP + b] =
                    (b < nvalid) ? __ldg(a.y0 + ((b0 + b) * R + r) * (long long)L + c) : 0.f;
        }
        __syncthreads();
        vjp_at_x(0, nvalid);
        for (int idx = threadIdx.x; idx < R * TB; idx += kThreads) {
            const int r = idx >> P.tb_shift, b = idx & (TB - 1);
            if (b < nvalid) {
                float* dst = a.g_y0 + ((b0 + b) * R + r) * (long long)L;      // holds the cotangent of output 0 (+ constants' sums)
                for (int l = 0; l < 3; ++l) {
                    const int row = (3 * r + l) * TBP + b;
                    dst[l] = (LAM1[row] + GX[row]) + dst[l];
                }
            }
        }
        __syncthreads();
        if (ring && P.c0_tm) dl_tm_dump(P, smem, P.off_DL, tmem, warp, lane);
        hoist_backward_epilogue(P, a.wpacked, smem, Gslab, a.y0, a.g_y0, b0, nvalid, warp, lane);
    }
    __syncthreads();
    if (SMEM_G)
        for (int i = threadIdx.x; i < P.n_packed_hot; i += kThreads) Gslab[i] = G[i];
    if (tm_acc) {
        dw_tm_flush(P, smem + P.off_GB, Gslab, tmem, warp, lane);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) tm_free512(tmem);
    }
    if (ring) wr_drain(wr);
}

}  // namespace ude
