// ude_ring.cuh — the FFMA engine's plan for WIDE models (R = 49: 168 KB of live weights, 166 KB of dtheta and 157 KB of
// activations cannot all sit in 227 KB of shared memory).  What streams, and what stays:
//   * weights   : every layer block [W | b] passes through a 2-slot shared-memory ring, fetched from L2 with TMA bulk
//                 copies (cp.async.bulk + mbarrier) ONE PHASE AHEAD of its use -- the inner loops read shared memory
//                 only; the layers of the two nets run as separate phases (one block per phase);
//   * dtheta    : 4x4 accumulators per thread in tensor memory for the whole launch (ude_engine.cuh: TmIo);
//   * c0 / D    : the hoisted first-layer constants and the summed first-layer deltas, also in the owning threads'
//                 tensor-memory columns (adjoint kernel), which frees 37 KB of shared memory for the ring;
//   * solver state rows: the CTA's scratch in global memory (L2-resident), touched only in the short tail phases.
// Same micro-kernels, same arithmetic and summation order as the lockstep path of ude_engine.cuh.
#pragma once
#include "ude_engine.cuh"

namespace ude {

// ----------------------------------------------------------------------------------
// weight ring
// ----------------------------------------------------------------------------------
struct WRing {
    float* slots;           // 2 x slot_floats
    uint64_t* bar;          // 2 mbarriers
    const float* src;       // packed blob of this weight set (global)
    const short* seq;       // layer ids in acquisition order (one RHS evaluation), in shared memory (ModelP)
    const int* blk_off;     // per layer id: float offset / float count of its [W | b] block
    const int* blk_n;
    int slot_floats, nseq;
    uint32_t k;             // acquisitions so far
    int pos;                // k mod nseq
};

// mbarrier wait that traps instead of hanging the GPU if a fetch never lands (a scheduling bug would otherwise spin forever)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 24) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done) __trap();
}

__device__ __forceinline__ void wr_issue(const WRing& r, int slot, int pos) {      // one thread
    const int id = r.seq[pos];
    const uint32_t bytes = (uint32_t)r.blk_n[id] * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the slot precede the async writes
    mbar_expect_tx(&r.bar[slot], bytes);
    uint32_t done = 0;
    char* dst = reinterpret_cast<char*>(r.slots + (size_t)slot * r.slot_floats);
    const char* src = reinterpret_cast<const char*>(r.src + r.blk_off[id]);
    while (done < bytes) {
        uint32_t chunk = bytes - done;
        if (chunk > 32768u) chunk = 32768u;
        tma_bulk_g2s(dst + done, src + done, chunk, &r.bar[slot]);
        done += chunk;
    }
}
// Called by ALL threads once, after the barriers are initialised and visible.
__device__ __forceinline__ void wr_start(WRing& r) {
    r.k = 0;
    r.pos = 0;
    if (threadIdx.x == 0) {
        wr_issue(r, 0, 0);
        wr_issue(r, 1, 1 % r.nseq);
    }
}
// All threads: wait for the block of the current phase.  Returns Wp such that Wp + Lp.w_off / Lp.b_off address the slot copy.
__device__ __forceinline__ const float* wr_acquire(const WRing& r, const LayerP& Lp) {
    const int slot = (int)(r.k & 1u);
    mbar_wait_bounded(&r.bar[slot], (r.k >> 1) & 1u);
    return r.slots + (size_t)slot * r.slot_floats - Lp.w_off;
}
// All threads, AFTER the __syncthreads that ends the phase: the slot is free, fetch the block two phases ahead.
__device__ __forceinline__ void wr_release(WRing& r) {
    int p2 = r.pos + 2;
    if (p2 >= r.nseq) p2 -= r.nseq;
    if (p2 >= r.nseq) p2 -= r.nseq;
    if (threadIdx.x == 0) wr_issue(r, (int)(r.k & 1u), p2);
    r.k += 1u;
    r.pos = r.pos + 1 == r.nseq ? 0 : r.pos + 1;
}
// Kernel end: two fetches are in flight; they must land before the CTA gives its shared memory back.
__device__ __forceinline__ void wr_drain(WRing& r) {
    mbar_wait_bounded(&r.bar[r.k & 1u], (r.k >> 1) & 1u);
    r.k += 1u;
    mbar_wait_bounded(&r.bar[r.k & 1u], (r.k >> 1) & 1u);
}

// ----------------------------------------------------------------------------------
// micro-kernel variants
// ----------------------------------------------------------------------------------
// forward unit with any TN (n-tile = 4 TN outputs) and the hoisted constants in tensor memory:
//   C0 = 0: bias / c0 rows in shared memory (fwd_unit)     C0 = 1: c0 from the thread's TMEM slot `tm`
//   C0 = 2: the product IS c0 (hoist_c0): + bias, written to the TMEM slot instead of shared memory
template <int TN, int C0>
__device__ __forceinline__ void fwd_unit_r(const LayerP& Lp, const float* __restrict__ Wp, float* __restrict__ smem,
                                           const int TBP, const int mt, const int nt, const int lane, const uint32_t tm) {
    static_assert(C0 == 0 || TN <= 4, "c0 slots hold 16 values per thread");
    const int tx = lane & 7, ty = lane >> 3;
    const int m_off = mt * 32 + tx * 4;
    const int n0 = nt * (4 * TN);
    const float* __restrict__ A = smem + Lp.in_off + m_off;
    const float* __restrict__ W = Wp + Lp.w_off;
    int wofs[TN];
#pragma unroll
    for (int q = 0; q < TN; ++q) {
        int n = n0 + ty + 4 * q;
        n = n < Lp.n_out_p ? n : Lp.n_out_p - 1;
        wofs[q] = n * Lp.ldw;
    }
    float acc[TN][4];
#pragma unroll
    for (int q = 0; q < TN; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    const int K = Lp.n_in_p;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float4 a0 = ld4(A + (k + 0) * TBP);
        const float4 a1 = ld4(A + (k + 1) * TBP);
        const float4 a2 = ld4(A + (k + 2) * TBP);
        const float4 a3 = ld4(A + (k + 3) * TBP);
#pragma unroll
        for (int q = 0; q < TN; ++q) {
            const float4 w = ld4(W + wofs[q] + k);
            ffma2(acc[q][0], acc[q][1], a0.x, a0.y, w.x, w.x);
            ffma2(acc[q][2], acc[q][3], a0.z, a0.w, w.x, w.x);
            ffma2(acc[q][0], acc[q][1], a1.x, a1.y, w.y, w.y);
            ffma2(acc[q][2], acc[q][3], a1.z, a1.w, w.y, w.y);
            ffma2(acc[q][0], acc[q][1], a2.x, a2.y, w.z, w.z);
            ffma2(acc[q][2], acc[q][3], a2.z, a2.w, w.z, w.z);
            ffma2(acc[q][0], acc[q][1], a3.x, a3.y, w.w, w.w);
            ffma2(acc[q][2], acc[q][3], a3.z, a3.w, w.w, w.w);
        }
    }
    const float* __restrict__ bias = Wp + Lp.b_off;
    float* __restrict__ out = smem + Lp.act_off + m_off;
    const bool elu = Lp.elu != 0;
    float c0v[16];
    if (C0 == 1) TmIo<16>::ld(tm, c0v);
#pragma unroll
    for (int q = 0; q < TN; ++q) {
        const int n = n0 + ty + 4 * q;
        const int nc = n < Lp.n_out_p ? n : Lp.n_out_p - 1;
        float4 v;
        if (C0 == 1) {
            v = make_float4(acc[q][0] + c0v[(4 * q + 0) & 15], acc[q][1] + c0v[(4 * q + 1) & 15], acc[q][2] + c0v[(4 * q + 2) & 15],
                            acc[q][3] + c0v[(4 * q + 3) & 15]);
        } else if (C0 == 0 && Lp.c0_off >= 0) {
            const float4 c = ld4(smem + Lp.c0_off + nc * TBP + m_off);
            v = make_float4(acc[q][0] + c.x, acc[q][1] + c.y, acc[q][2] + c.z, acc[q][3] + c.w);
        } else {
            const float b = bias[nc];
            v = make_float4(acc[q][0] + b, acc[q][1] + b, acc[q][2] + b, acc[q][3] + b);
        }
        if (C0 == 2) {
            c0v[(4 * q + 0) & 15] = v.x; c0v[(4 * q + 1) & 15] = v.y; c0v[(4 * q + 2) & 15] = v.z; c0v[(4 * q + 3) & 15] = v.w;
        } else if (n < Lp.n_out_p) {
            if (elu) {
                v.x = elu_f(v.x); v.y = elu_f(v.y); v.z = elu_f(v.z); v.w = elu_f(v.w);
            }
            st4(out + n * TBP, v);
        }
    }
    if (C0 == 2) {
#pragma unroll
        for (int i = 4 * TN; i < 16; ++i) c0v[i] = 0.f;
        TmIo<16>::st(tm, c0v);
    }
}

template <int C0>
__device__ __forceinline__ void fwd_run_r(const LayerP& Lp, const float* Wp, float* smem, int TBP, int mt_shift, int u, int lane,
                                          uint32_t tm) {
    const int mt = u & ((1 << mt_shift) - 1), nt = u >> mt_shift;
    switch (Lp.tn_fwd) {
        case 1: fwd_unit_r<1, C0>(Lp, Wp, smem, TBP, mt, nt, lane, tm); break;
        case 2: fwd_unit_r<2, C0>(Lp, Wp, smem, TBP, mt, nt, lane, tm); break;
        case 3: fwd_unit_r<3, C0>(Lp, Wp, smem, TBP, mt, nt, lane, tm); break;
        case 4: fwd_unit_r<4, C0>(Lp, Wp, smem, TBP, mt, nt, lane, tm); break;
        default:
            if (C0 == 0) {
                if (Lp.tn_fwd == 5) fwd_unit_r<5, 0>(Lp, Wp, smem, TBP, mt, nt, lane, tm);
                else if (Lp.tn_fwd == 6) fwd_unit_r<6, 0>(Lp, Wp, smem, TBP, mt, nt, lane, tm);
                else fwd_unit_r<8, 0>(Lp, Wp, smem, TBP, mt, nt, lane, tm);
            }
            break;
    }
}

// delta unit with CW consecutive inputs per thread and NG column groups: n-tile = 4 CW NG inputs (8, 16 or 32).
//   DLM = 1: (prev = first layer, hoisting) the result, i.e. delta_0, is ALSO added to the thread's TMEM slot `tm`
//            (D = sum over evaluations of delta_0, ModelP "Hoisting")
template <int CW, int NG, int DLM>
__device__ __forceinline__ void bwd_unit_g(const LayerP& Lp, const LayerP* prev, const float* __restrict__ Wp,
                                           float* __restrict__ smem, const int TBP, const int gx_off, const int mt,
                                           const int nt, const int lane, const uint32_t tm) {
    static_assert(CW == 2 || CW == 4, "CW");
    static_assert(!DLM || CW * NG <= 4, "D slots hold 16 values per thread");
    constexpr int TN = CW * NG;
    const int tx = lane & 7, ty = lane >> 3;
    const int m_off = mt * 32 + tx * 4;
    const int n0 = nt * (4 * TN);
    const float* __restrict__ Dl = smem + Lp.act_off + m_off;
    const float* __restrict__ W = Wp + Lp.w_off + n0 + ty * CW;
    const int ldw = Lp.ldw;
    float acc[TN][4];
#pragma unroll
    for (int q = 0; q < TN; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    const int K = Lp.n_out;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a = ld4(Dl + k * TBP);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float w[CW];
            if (CW == 4) {
                const float4 w4 = ld4(W + k * ldw + g * 16);
                w[0] = w4.x; w[1] = w4.y; w[CW - 2] = w4.z; w[CW - 1] = w4.w;
            } else {
                const float2 w2 = *reinterpret_cast<const float2*>(W + k * ldw + g * 8);
                w[0] = w2.x; w[1] = w2.y;
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                ffma2(acc[g * CW + j][0], acc[g * CW + j][1], a.x, a.y, w[j], w[j]);
                ffma2(acc[g * CW + j][2], acc[g * CW + j][3], a.z, a.w, w[j], w[j]);
            }
        }
    }
    float* __restrict__ dst = smem + (prev ? prev->act_off : gx_off) + m_off;
    const bool elu = prev && prev->elu;
    const bool accumulate = prev == nullptr;
    float dl[16];
    if (DLM) TmIo<16>::ld(tm, dl);
#pragma unroll
    for (int q = 0; q < TN; ++q) {
        const int n = n0 + (q / CW) * (4 * CW) + ty * CW + (q % CW);
        float4 v = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        if (n < Lp.n_in_p) {
            if (elu) {
                const float4 e = ld4(dst + n * TBP);
                v.x *= elu_grad_from_out(e.x); v.y *= elu_grad_from_out(e.y);
                v.z *= elu_grad_from_out(e.z); v.w *= elu_grad_from_out(e.w);
            } else if (accumulate) {
                const float4 e = ld4(dst + n * TBP);
                v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
            }
            st4(dst + n * TBP, v);
        } else {
            v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (DLM) {
            dl[(4 * q + 0) & 15] += v.x; dl[(4 * q + 1) & 15] += v.y; dl[(4 * q + 2) & 15] += v.z; dl[(4 * q + 3) & 15] += v.w;
        }
    }
    if (DLM) TmIo<16>::st(tm, dl);
}
template <int DLM>
__device__ __forceinline__ void bwd_run_r(const LayerP& Lp, const LayerP* prev, const float* Wp, float* smem, int TBP, int gx_off,
                                          int mt_shift, int u, int lane, uint32_t tm) {
    const int mt = u & ((1 << mt_shift) - 1), nt = u >> mt_shift;
    if (Lp.tn_bwd == 2) bwd_unit_g<2, 1, DLM>(Lp, prev, Wp, smem, TBP, gx_off, mt, nt, lane, tm);
    else if (Lp.tn_bwd == 4) bwd_unit_g<4, 1, DLM>(Lp, prev, Wp, smem, TBP, gx_off, mt, nt, lane, tm);
    else if (!DLM) bwd_unit_g<4, 2, 0>(Lp, prev, Wp, smem, TBP, gx_off, mt, nt, lane, tm);
}

// TMEM slot of the `it`-th unit this warp runs in the first-layer phase of net `net` (c0) / in the delta phase that
// produces that net's delta_0 (D): 16 columns, warps w and w + 4 interleaved
__device__ __forceinline__ uint32_t aux_tm_addr(uint32_t tmem, int base, int it, int warp) {
    return tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(base + (it * (kWarps / 4) + (warp >> 2)) * kTmSlot);
}

// ----------------------------------------------------------------------------------
// phases (ring mode: one layer per phase, its weights from the ring)
// ----------------------------------------------------------------------------------
// c0 of both first layers (constant blocks straight from global memory: twice per tile and solve)
__device__ __forceinline__ void hoist_c0_ring(const ModelP& P, const float* __restrict__ Wg, float* smem, int warp, int lane,
                                              uint32_t tmem) {
    if (!P.hoist) return;
    for (int net = 0; net < 2; ++net) {
        if ((net == 0 ? P.n_rate : P.n_aug) == 0) continue;
        const LayerP& Lc = net == 0 ? P.rate_c : P.aug_c;
        const int units = fwd_units(Lc, P.mt_shift);
        int it = 0;
        for (int u = warp; u < units; u += kWarps, ++it) {
            if (P.c0_tm) fwd_run_r<2>(Lc, Wg, smem, P.TBP, P.mt_shift, u, lane, aux_tm_addr(tmem, P.tm_c0_base[net], it, warp));
            else fwd_run_r<0>(Lc, Wg, smem, P.TBP, P.mt_shift, u, lane, 0u);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void rhs_forward_layers_ring(const ModelP& P, WRing& wr, float* smem, int warp, int lane, uint32_t tmem) {
    const int nl = P.n_rate > P.n_aug ? P.n_rate : P.n_aug;
    for (int li = 0; li < nl; ++li) {
        for (int net = 0; net < 2; ++net) {
            if (li >= (net == 0 ? P.n_rate : P.n_aug)) continue;
            const LayerP& Lp = net == 0 ? P.rate[li] : P.aug[li];
            const float* Wp = wr_acquire(wr, Lp);
            const int units = fwd_units(Lp, P.mt_shift);
            const bool c0_tm = li == 0 && P.hoist && P.c0_tm;
            int it = 0;
            for (int u = warp; u < units; u += kWarps, ++it) {
                if (c0_tm) fwd_run_r<1>(Lp, Wp, smem, P.TBP, P.mt_shift, u, lane, aux_tm_addr(tmem, P.tm_c0_base[net], it, warp));
                else fwd_run_r<0>(Lp, Wp, smem, P.TBP, P.mt_shift, u, lane, 0u);
            }
            __syncthreads();
            wr_release(wr);
        }
    }
}

// Backward through all layers; the weight-gradient phases run both nets side by side (no weights needed), the delta
// phases one layer at a time.  Pre / post as rhs_backward_layers.  G = compact bias-sum array (TM_ACC mode).
__device__ __forceinline__ void rhs_backward_layers_ring(const ModelP& P, WRing& wr, float* smem, float* Gb, int warp, int lane,
                                                         uint32_t tmem) {
    const int nl = P.n_rate > P.n_aug ? P.n_rate : P.n_aug;
    for (int s = 0; s < nl; ++s) {
        const int lr = P.n_rate - 1 - s, la = P.n_aug - 1 - s;
        const LayerP* Lr = lr >= 0 ? &P.rate[lr] : nullptr;
        const LayerP* La = la >= 0 ? &P.aug[la] : nullptr;
        {   // ---- weight and bias gradients of this depth (TMEM accumulators)
            const int w0 = Lr ? dw_units(*Lr) : 0, w1 = La ? dw_units(*La) : 0;
            for (int u = dw_first_unit(P, s, warp); u < w0 + w1; u += kWarps) {
                const uint32_t tm = dw_tm_addr(P, tmem, s, u, warp);
                if (u < w0) dw_run<true>(*Lr, smem, Gb, P.TB, P.TBP, u, lane, tm);
                else dw_run<true>(*La, smem, Gb, P.TB, P.TBP, u - w0, lane, tm);
            }
        }
        __syncthreads();
        // ---- cotangents of this depth's inputs, one layer (= one ring block) per phase
        for (int net = 0; net < 2; ++net) {
            const LayerP* Lp = net == 0 ? Lr : La;
            if (!Lp) continue;
            const int l = net == 0 ? lr : la;
            const LayerP* prev = l > 0 ? (net == 0 ? &P.rate[l - 1] : &P.aug[l - 1]) : nullptr;
            const float* Wp = wr_acquire(wr, *Lp);
            const int units = bwd_units(*Lp, P.mt_shift);
            const bool dl = l == 1 && P.hoist;       // this phase produces delta_0 of the net
            int it = 0;
            for (int u = warp; u < units; u += kWarps, ++it) {
                if (dl) bwd_run_r<1>(*Lp, prev, Wp, smem, P.TBP, P.off_GX, P.mt_shift, u, lane, aux_tm_addr(tmem, P.tm_dl_base[net], it, warp));
                else bwd_run_r<0>(*Lp, prev, Wp, smem, P.TBP, P.off_GX, P.mt_shift, u, lane, 0u);
            }
            __syncthreads();
            wr_release(wr);
        }
    }
}

// D slots <- 0 (tile start) / D slots -> the rows at `row_off` (tile end; same walk as the delta phases of layer 1)
__device__ __forceinline__ void dl_tm_clear(const ModelP& P, uint32_t tmem, int warp) {
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0.f;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = P.tm_dl_base[0] + (warp >> 2) * kTmSlot; c < P.tm_dl_end; c += (kWarps / 4) * kTmSlot) TmIo<16>::st(lane_base + c, z);
}
__device__ __forceinline__ void dl_tm_dump(const ModelP& P, float* smem, int row_off, uint32_t tmem, int warp, int lane) {
    const int tx = lane & 7, ty = lane >> 3;
    int rows_before = 0;
    for (int net = 0; net < 2; ++net) {
        if ((net == 0 ? P.n_rate : P.n_aug) == 0) continue;
        const LayerP& L1 = net == 0 ? P.rate[1] : P.aug[1];          // its delta phase produced delta_0 (n = its inputs)
        const int TN = L1.tn_bwd;                                    // 2 | 4 (host guarantees <= 4)
        const int CW = TN == 2 ? 2 : 4;
        const int units = bwd_units(L1, P.mt_shift);
        int it = 0;
        for (int u = warp; u < units; u += kWarps, ++it) {
            const int mt = u & ((1 << P.mt_shift) - 1), nt = u >> P.mt_shift;
            float dl[16];
            TmIo<16>::ld(aux_tm_addr(tmem, P.tm_dl_base[net], it, warp), dl);
            for (int q = 0; q < TN; ++q) {
                const int n = nt * (4 * TN) + (q / CW) * (4 * CW) + ty * CW + (q % CW);
                if (n < L1.n_in_p)
                    st4(smem + row_off + (rows_before + n) * P.TBP + mt * 32 + tx * 4,
                        make_float4(dl[(4 * q) & 15], dl[(4 * q + 1) & 15], dl[(4 * q + 2) & 15], dl[(4 * q + 3) & 15]));
            }
        }
        rows_before += L1.n_in_p;
    }
}

}  // namespace ude
