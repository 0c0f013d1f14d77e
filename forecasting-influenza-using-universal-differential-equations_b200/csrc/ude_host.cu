// ude_host.cu — C-ABI entry points (include/ude_b200.h) and host-side planning.
//
// Everything here is plain C++/CUDA runtime: no torch, no allocation of persistent
// memory, no exceptions across the boundary.
#include <stdlib.h>

#include "ude_host_common.h"
#include "ude_dopri5.cuh"
#include "ude_rk4.cuh"
#include "ude_tc_model.h"

namespace ude {

// ---------------------------------------------------------------------------------
// descriptor -> ModelP (layer tables, packed layout).  No GPU needed.
// ---------------------------------------------------------------------------------
static int check_desc(const UdeDesc* d) {
    if (!d) return UDE_ERR_INVALID;
    if (d->kind < UDE_KIND_FP || d->kind > UDE_KIND_FAFP) return UDE_ERR_INVALID;
    if (d->n_regions < 1 || d->latent_dim < 3) return UDE_ERR_INVALID;
    const bool rate = d->kind != UDE_KIND_FA, aug = d->kind != UDE_KIND_FP;
    if (rate) {
        if (d->n_rate_layers < 2 || d->n_rate_layers > UDE_MAX_LAYERS) return UDE_ERR_INVALID;
        if (d->rate_sizes[d->n_rate_layers - 1] != 2 * d->n_regions) return UDE_ERR_INVALID;
        for (int i = 0; i < d->n_rate_layers; ++i)
            if (d->rate_sizes[i] < 1) return UDE_ERR_INVALID;
    } else if (d->n_rate_layers != 0) return UDE_ERR_INVALID;
    if (aug) {
        if (d->n_aug_layers < 2 || d->n_aug_layers > UDE_MAX_LAYERS) return UDE_ERR_INVALID;
        if (d->aug_sizes[d->n_aug_layers - 1] != 3 * d->n_regions) return UDE_ERR_INVALID;
        for (int i = 0; i < d->n_aug_layers; ++i)
            if (d->aug_sizes[i] < 1) return UDE_ERR_INVALID;
    } else if (d->n_aug_layers != 0) return UDE_ERR_INVALID;
    if (d->n_groups < 1 || d->batch < 1 || d->batch % d->n_groups != 0) return UDE_ERR_INVALID;
    return UDE_OK;
}

static void fill_layers(ModelP& P, LayerP* Ls, int n, const int32_t* sizes, int head_elu, int& th, int& pk) {
    int n_in = P.D;
    for (int i = 0; i < n; ++i) {
        LayerP& Lp = Ls[i];
        memset(&Lp, 0, sizeof(Lp));
        const bool split = i == 0 && P.hoist;        // first layer: the packed block holds the LIVE columns only
        Lp.n_in = split ? P.Dn : n_in;
        Lp.th_ld = n_in;
        Lp.n_out = sizes[i];
        Lp.n_in_p = round4(Lp.n_in);
        Lp.n_out_p = round4(Lp.n_out);
        Lp.ldw = Lp.n_in_p;
        if (Lp.ldw % 32 == 0) Lp.ldw += 4;     // keep the 4 rows a warp reads at once in distinct banks
        Lp.th_w = th;
        th += n_in * Lp.n_out;
        Lp.th_b = th;
        th += Lp.n_out;
        Lp.w_off = pk;
        pk += Lp.n_out_p * Lp.ldw;
        Lp.b_off = pk;
        pk += Lp.n_out_p;
        Lp.elu = (i <= n - 3 + (head_elu ? 1 : 0)) ? 1 : 0;   // lib/models.py: no ELU before the head
        Lp.c0_off = -1;
        n_in = Lp.n_out;
    }
}

// constant block of a first layer (ModelP: "Hoisting"): same rows, the (L-3)R constant columns
static void fill_const_block(const ModelP& P, const LayerP& L0, LayerP& Lc, int& pk) {
    memset(&Lc, 0, sizeof(Lc));
    Lc.n_in = P.Dc;
    Lc.th_ld = L0.th_ld;
    Lc.n_out = L0.n_out;
    Lc.n_in_p = P.Dc_p;
    Lc.n_out_p = L0.n_out_p;
    Lc.ldw = Lc.n_in_p;
    if (Lc.ldw % 32 == 0) Lc.ldw += 4;
    Lc.th_w = L0.th_w;
    Lc.th_b = L0.th_b;
    Lc.w_off = pk;
    pk += Lc.n_out_p * Lc.ldw;
    Lc.b_off = L0.b_off;          // c0 = W0[:, const] . x_const + b0
    Lc.elu = 0;
    Lc.c0_off = -1;
    Lc.no_db = 1;                 // the live block's units own the bias gradient
}

static void base_model(const UdeDesc* d, ModelP& P) {
    memset(&P, 0, sizeof(P));
    P.kind = d->kind;
    P.R = d->n_regions;
    P.L = d->latent_dim;
    P.D = P.R * P.L;
    P.D_p = round4(P.D);
    P.Dn = 3 * P.R;
    P.Dn_p = round4(P.Dn);
    P.Dc = (P.L - 3) * P.R;
    P.Dc_p = round4(P.Dc);
    P.hoist = P.Dc > 0;
    P.bayes = (d->flags & UDE_FLAG_BAYES) != 0;
    P.R_p = round4(P.R);
    P.n_rate = d->n_rate_layers;
    P.n_aug = d->n_aug_layers;
    P.fa_w = d->fa_w;
    int th = 0, pk = 0;
    fill_layers(P, P.rate, P.n_rate, d->rate_sizes, d->flags & UDE_FLAG_RATE_HEAD_ELU, th, pk);
    fill_layers(P, P.aug, P.n_aug, d->aug_sizes, d->flags & UDE_FLAG_AUG_HEAD_ELU, th, pk);
    P.n_theta = th;
    P.n_packed_hot = pk;
    if (P.hoist) {
        if (P.n_rate > 0) fill_const_block(P, P.rate[0], P.rate_c, pk);
        if (P.n_aug > 0) fill_const_block(P, P.aug[0], P.aug_c, pk);
    }
    P.n_packed = pk + kPadTail;
    if (!P.hoist) P.n_packed_hot = P.n_packed;
}

static void tile_shapes(LayerP& Lp, int m_tiles, int want) {
    if (Lp.n_out_p <= 4) Lp.tn_fwd = 1;
    else Lp.tn_fwd = (m_tiles * ((Lp.n_out_p + 31) / 32) >= want) ? 8 : 4;
    Lp.tn_bwd = (m_tiles * ((Lp.n_in_p + 31) / 32) >= want) ? 8 : 4;
    Lp.tm_dw = Lp.n_in_p > 8 ? 4 : 1;
    Lp.tn_dw = Lp.n_out_p > 4 ? 4 : 1;
    Lp.fwd_nt = (Lp.n_out_p + 4 * Lp.tn_fwd - 1) / (4 * Lp.tn_fwd);
    Lp.bwd_nt = (Lp.n_in_p + 4 * Lp.tn_bwd - 1) / (4 * Lp.tn_bwd);
    Lp.dw_mt = (Lp.n_in_p + 8 * Lp.tm_dw - 1) / (8 * Lp.tm_dw);
    Lp.dw_nt = (Lp.n_out_p + 4 * Lp.tn_dw - 1) / (4 * Lp.tn_dw);
}

// Shared-memory plan for a given tile size; returns the arena size in floats.
// ring-mode tile shapes: one net per phase, so a phase wants >= kWarps units of its own (ude_ring.cuh)
static void tile_shapes_ring(LayerP& Lp, int m_tiles, bool c0_slot, bool dl_slot) {
    const int groups = (Lp.n_out_p + 3) / 4;
    int tn = (groups * m_tiles + kWarps - 1) / kWarps;
    tn = tn < 1 ? 1 : (tn > 6 ? 8 : tn);
    if (c0_slot && tn > 4) tn = 4;                      // c0 slots hold 16 values per thread
    Lp.tn_fwd = tn;
    Lp.fwd_nt = (Lp.n_out_p + 4 * tn - 1) / (4 * tn);
    int best = 2, best_cost = 1 << 30;
    for (int t = 2; t <= (dl_slot ? 4 : 8); t *= 2) {   // D slots hold 16 values per thread
        const int units = (Lp.n_in_p + 4 * t - 1) / (4 * t) * m_tiles;
        const int cost = (units + kWarps - 1) / kWarps * t;
        if (cost <= best_cost) { best = t; best_cost = cost; }
    }
    Lp.tn_bwd = best;
    Lp.bwd_nt = (Lp.n_in_p + 4 * best - 1) / (4 * best);
}

static int plan_smem(ModelP& P, int TB, bool w_smem, bool g_smem, bool cold, int n_state, int n_rbuf, bool bwd, bool ring = false) {
    P.TB = TB;
    P.TBP = TB + 4;
    P.tb_shift = TB == 64 ? 6 : 5;
    P.mt_shift = TB == 64 ? 1 : 0;
    P.w_in_smem = w_smem;
    P.g_in_smem = g_smem;
    P.n_state = n_state;
    int off = 0;
    P.off_W = off;
    if (w_smem) off += P.n_packed_hot;
    P.off_G = off;
    if (g_smem) off += P.n_packed_hot;
    P.off_X = off;
    off += P.Dn_p * P.TBP;
    const int act0 = off;
    for (int net = 0; net < 2; ++net) {
        LayerP* Ls = net == 0 ? P.rate : P.aug;
        const int n = net == 0 ? P.n_rate : P.n_aug;
        for (int i = 0; i < n; ++i) {
            Ls[i].in_off = i == 0 ? P.off_X : Ls[i - 1].act_off;
            Ls[i].act_off = off;
            off += Ls[i].n_out_p * P.TBP;
        }
    }
    // constants of the tile: alias the ACT rows (only alive in the tile prologue / epilogue) -- unless the weights change per
    // evaluation (bayes): then every evaluation redoes the constant product and the rows stay
    P.off_XC = act0;
    P.off_GXC = 0;
    if (P.bayes) {
        P.off_XC = off;
        off += P.Dc_p * P.TBP;
        if (bwd) { P.off_GXC = off; off += P.Dc_p * P.TBP; }
    } else if (off - act0 < P.Dc_p * P.TBP) off = act0 + P.Dc_p * P.TBP;
    // hoisted first-layer parts: c0 rows (both directions), D rows (adjoint); in the ring-mode adjoint both live in
    // tensor memory and D is dumped behind the constants' rows for the tile epilogue
    P.w_ring = ring;
    P.c0_tm = ring && bwd && P.hoist;
    P.dl_rows = 0;
    if (P.hoist) {
        for (int net = 0; net < 2; ++net) {
            LayerP* Ls = net == 0 ? P.rate : P.aug;
            LayerP& Lc = net == 0 ? P.rate_c : P.aug_c;
            if ((net == 0 ? P.n_rate : P.n_aug) == 0) continue;
            P.dl_rows += Ls[0].n_out_p;
            Lc.in_off = P.off_XC;
            if (P.c0_tm) { Ls[0].c0_off = -1; Lc.act_off = 0; continue; }
            Ls[0].c0_off = off;
            Lc.act_off = off;
            off += Ls[0].n_out_p * P.TBP;
        }
        if (P.c0_tm) {
            P.off_DL = P.off_XC + P.Dc_p * P.TBP;
            if (off < P.off_DL + P.dl_rows * P.TBP) off = P.off_DL + P.dl_rows * P.TBP;
        } else {
            P.off_DL = off;
            if (bwd) off += P.dl_rows * P.TBP;
        }
    }
    // weight ring: two slots, each as large as the largest [W | b] layer block
    P.off_ring = off;
    P.ring_slot_floats = 0;
    if (ring) {
        int id = 0;
        for (int net = 0; net < 2; ++net) {
            LayerP* Ls = net == 0 ? P.rate : P.aug;
            const int n = net == 0 ? P.n_rate : P.n_aug;
            for (int i = 0; i < n; ++i, ++id) {
                P.wblk_off[id] = Ls[i].w_off;
                P.wblk_n[id] = Ls[i].n_out_p * Ls[i].ldw + Ls[i].n_out_p;
                if (P.wblk_n[id] > P.ring_slot_floats) P.ring_slot_floats = P.wblk_n[id];
            }
        }
        off += 2 * P.ring_slot_floats;
        // acquisition order of one RHS evaluation: forward layer by layer (rate, aug), then -- adjoint kernels -- the
        // delta phases from the heads down (rhs_forward_layers_ring / rhs_backward_layers_ring)
        const int nl = P.n_rate > P.n_aug ? P.n_rate : P.n_aug;
        int q = 0;
        for (int li = 0; li < nl; ++li) {
            if (li < P.n_rate) P.wseq[q++] = (short)li;
            if (li < P.n_aug) P.wseq[q++] = (short)(P.n_rate + li);
        }
        if (bwd)
            for (int sdepth = 0; sdepth < nl; ++sdepth) {
                if (P.n_rate - 1 - sdepth >= 0) P.wseq[q++] = (short)(P.n_rate - 1 - sdepth);
                if (P.n_aug - 1 - sdepth >= 0) P.wseq[q++] = (short)(P.n_rate + P.n_aug - 1 - sdepth);
            }
        P.n_wseq = q;
    }
    // solver-state rows: in shared memory, or (large R) in a per-CTA global scratch ("cold")
    int coff = 0;
    int& soff = cold ? coff : off;
    P.cold_global = cold;
    P.off_S = soff;
    soff += n_state * P.Dn_p * P.TBP;
    P.off_RB = soff;
    P.n_rbuf = n_rbuf;
    soff += n_rbuf * P.R_p * P.TBP;
    P.cold_floats = coff + 8;
    P.off_GX = off;
    if (bwd) off += P.Dn_p * P.TBP;
    off += 8;    // slack
    // tile shapes: aim at >= kWarps units per phase (both nets run in the same phase)
    const int nets = (P.n_rate > 0) + (P.n_aug > 0);
    const int want = kWarps / nets;
    const int m_tiles = TB / 32;
    for (int net = 0; net < 2; ++net) {
        LayerP* Ls = net == 0 ? P.rate : P.aug;
        const int n = net == 0 ? P.n_rate : P.n_aug;
        for (int i = 0; i < n; ++i) tile_shapes(Ls[i], m_tiles, want);
    }
    if (P.hoist) {
        if (P.n_rate > 0) tile_shapes(P.rate_c, m_tiles, want);
        if (P.n_aug > 0) tile_shapes(P.aug_c, m_tiles, want);
    }
    if (ring) {
        for (int net = 0; net < 2; ++net) {
            LayerP* Ls = net == 0 ? P.rate : P.aug;
            const int n = net == 0 ? P.n_rate : P.n_aug;
            for (int i = 0; i < n; ++i) tile_shapes_ring(Ls[i], m_tiles, i == 0 && P.c0_tm, i == 1 && P.c0_tm);
            if (n > 0 && P.hoist) {          // the constant block runs with the live block's unit shape (same c0 slots)
                LayerP& Lc = net == 0 ? P.rate_c : P.aug_c;
                Lc.tn_fwd = Ls[0].tn_fwd;
                Lc.fwd_nt = Ls[0].fwd_nt;
            }
        }
    }
    // weight-gradient accumulators in tensor memory: only on the plans of last resort (nothing else fits on chip),
    // one 16-column slot per unit and warp, the units of all phases (heads first) numbered through
    P.tm_acc = 0;
    P.tm_cols = 0;
    P.n_bias_c = 0;
    P.off_GB = off;
    if (bwd && cold && !g_smem) {
        int bc = 0;
        for (int net = 0; net < 2; ++net) {
            LayerP* Ls = net == 0 ? P.rate : P.aug;
            const int n = net == 0 ? P.n_rate : P.n_aug;
            for (int i = 0; i < n; ++i) { Ls[i].bc_off = bc; bc += Ls[i].n_out_p; }
        }
        const int nl = P.n_rate > P.n_aug ? P.n_rate : P.n_aug;
        int total = 0;
        for (int s = 0; s < nl; ++s) {
            const int lr = P.n_rate - 1 - s, la = P.n_aug - 1 - s;
            P.tm_dw_prefix[s] = total;
            for (int net = 0; net < 2; ++net) {
                const int l = net == 0 ? lr : la;
                if (l < 0) continue;
                const LayerP& Lp = net == 0 ? P.rate[l] : P.aug[l];
                total += Lp.dw_mt * Lp.dw_nt;
            }
        }
        int cols = (total + kWarps - 1) / kWarps * (kWarps / 4) * 16;
        const int dw_cols = cols;
        if (P.c0_tm) {       // c0 and D slots of the first-layer / layer-1-delta phases, per net
            for (int net = 0; net < 2; ++net) {
                P.tm_c0_base[net] = cols;
                if ((net == 0 ? P.n_rate : P.n_aug) == 0) continue;
                const LayerP& L0 = net == 0 ? P.rate[0] : P.aug[0];
                cols += ((L0.fwd_nt << P.mt_shift) + kWarps - 1) / kWarps * (kWarps / 4) * 16;
            }
            for (int net = 0; net < 2; ++net) {
                P.tm_dl_base[net] = cols;
                if ((net == 0 ? P.n_rate : P.n_aug) == 0) continue;
                const LayerP& L1 = net == 0 ? P.rate[1] : P.aug[1];
                cols += ((L1.bwd_nt << P.mt_shift) + kWarps - 1) / kWarps * (kWarps / 4) * 16;
            }
            P.tm_dl_end = cols;
        }
        if (cols <= 512) {
            P.tm_acc = 1;
            P.tm_cols = dw_cols;
            P.n_bias_c = bc;
            off += bc;
        } else if (ring) {
            off = 1 << 30;          // the ring plan needs its tensor-memory slots: does not fit
        }
    }
    P.smem_floats = off;
    return off;
}

static const int kSmemBudgetBytes = 220 * 1024;     // 227 KB per CTA minus the static part (<= 5.4 KB)

// Pick (TB, weights in smem, accumulators in smem) for a direction.
static int choose_plan(const UdeDesc* d, ModelP& P, int n_state, int n_rbuf, bool bwd) {
    base_model(d, P);
    const int budget = kSmemBudgetBytes / 4;
    const long long tiles64 = (d->batch + 63) / 64;
    const DeviceInfo& di = device_info();
    const int n_sm = di.ok ? di.n_sm : 148;
    // small batches: smaller tiles spread over more SMs
    const int tb_order[2] = {tiles64 >= n_sm ? 64 : 32, tiles64 >= n_sm ? 32 : 64};
    struct Cand { int w, g, cold, ring; };
    // weights (+ accumulators) resident; else the ring plan for wide models (ude_ring.cuh: weights streamed through
    // shared memory, accumulators in tensor memory, state rows in the CTA's global scratch); else everything from L2
    const Cand cands_bwd[5] = {{1, 1, 0, 0}, {1, 0, 0, 0}, {0, 0, 1, 1}, {0, 0, 0, 0}, {0, 0, 1, 0}};
    const Cand cands_fwd[4] = {{1, 0, 0, 0}, {0, 0, 1, 1}, {0, 0, 0, 0}, {0, 0, 1, 0}};
    const Cand* cands = bwd ? cands_bwd : cands_fwd;
    const int nc = bwd ? 5 : 4;
    const bool no_ring = getenv("UDE_B200_NO_RING") != nullptr;         // A/B switch for profiling
    for (int c = 0; c < nc; ++c)
        for (int t = 0; t < 2; ++t) {
            if (cands[c].ring && no_ring) continue;
            if (P.bayes && c > 0) continue;      // per-evaluation weights: only the all-resident plans re-stage them
            if (plan_smem(P, tb_order[t], cands[c].w, cands[c].g, cands[c].cold, n_state, n_rbuf, bwd, cands[c].ring) <= budget)
                return UDE_OK;
        }
    return UDE_ERR_UNSUPPORTED;
}

struct Workspace {
    float* wpacked;
    double* side_partial;
    float* slabs;
    float* cold;
    long long bytes;
};
static Workspace carve_workspace(void* base, const ModelP& P, int n_ctas, int cold_floats) {
    Workspace w;
    long long off = 0;
    w.wpacked = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)P.n_packed * 4, 256);
    w.side_partial = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)n_ctas * 8 * 8, 256);
    w.slabs = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)n_ctas * P.n_packed * 4, 256);
    w.cold = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)n_ctas * cold_floats * 4, 256);
    w.bytes = off;
    return w;
}

// Bayes mode: behind the usual carve, the packed weight sets and noise of all 4 * n_steps evaluations and the fold blocks
struct BayesWs {
    float* w;        // [n_evals][n_packed]
    float* z;        // [n_evals][n_packed]
    float* fold;     // [2][n_ctas][n_packed]
    long long bytes;
};
static BayesWs bayes_carve(void* base, long long off, const ModelP& P, int n_steps, int n_ctas) {
    BayesWs b;
    const long long set = (long long)4 * n_steps * P.n_packed * 4;
    off = align_up(off, 256);
    b.w = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + set, 256);
    b.z = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + set, 256);
    b.fold = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)2 * n_ctas * P.n_packed * 4, 256);
    b.bytes = off;
    return b;
}

static int grid_size(const ModelP& P, long long B) {
    const DeviceInfo& di = device_info();
    const long long tiles = (B + P.TB - 1) / P.TB;
    const int n_sm = di.ok ? di.n_sm : 148;
    return (int)(tiles < n_sm ? tiles : n_sm);
}

static long long dopri_tape_x_bytes(const ModelP& P, long long B, int max_steps) {
    return align_up((long long)(max_steps + 1) * kDopriStages * P.Dn * B * 4, 256);
}

static int n_state_for(int method, bool bwd) {
    if (method == UDE_METHOD_RK4) return bwd ? (int)B_NSTATE : (int)F_NSTATE;
    return bwd ? (int)DB_NSTATE : (int)DF_NSTATE;
}
static int n_rbuf_for(int method, bool bwd) { return (method == UDE_METHOD_DOPRI5 && !bwd) ? (int)DF_NRB : 0; }

}  // namespace ude

namespace ude {
ReparamArgs& pending_reparam() {
    static thread_local ReparamArgs rp = {};
    return rp;
}
// the *_decoded entry points park their arguments here for the launch they forward to; the adjoint launch records how many
// per-CTA partial blocks it wrote
BayesArgs& pending_bayes() {
    static thread_local BayesArgs b = {};
    return b;
}
DecArgs& pending_decoded() {
    static thread_local DecArgs d = {};
    return d;
}
int& decoded_ctas() {
    static thread_local int n = 0;
    return n;
}
// tensor-core path, defined in ude_tc_host.cu
int tc_choose(const UdeDesc* d, TcModel& M);
void tc_plan_info(const UdeDesc* d, const TcModel& M, long long* workspace_bytes, int* n_ctas);
long long tc_bayes_workspace(const UdeDesc* d, const TcModel& M, int n_steps, int* n_ctas);
int tc_bayes_forward(const UdeDesc* desc, const TcModel& M, const float* theta_evals, const float* y0, const GridTables& g,
                     float* y_out, float* tape, double* side, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int tc_bayes_backward(const UdeDesc* desc, const TcModel& M, const float* theta_evals, const float* z_evals, const float* y0,
                      const GridTables& g, const float* tape, const float* g_out, const float* g_side, float* g_y0,
                      float* g_mean_zsum, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int tc_forward(const UdeDesc* desc, const TcModel& M, const float* theta, const float* y0, const GridTables& g,
               float* y_out, float* tape, double* side, void* workspace, int64_t workspace_bytes, cudaStream_t s);
int tc_backward(const UdeDesc* desc, const TcModel& M, const float* theta, const float* y0, const GridTables& g,
                const float* tape, const float* g_out, const float* g_side, float* g_y0, float* g_theta,
                void* workspace, int64_t workspace_bytes, cudaStream_t s);
}  // namespace ude

namespace ude {
// ---------------------------------------------------------------------------------
// Weight-set groups in ONE launch (Rk4FwdArgs::n_groups): possible when the plans keep the packed weights (and, in the
// adjoint, the dtheta accumulators) in shared memory -- the small models of the hyper-parameter sweeps
// (tuning/tune_node.py:169-171).  Anything else falls back to G single-set launch sets (for_each_group).
// ---------------------------------------------------------------------------------
// two resident CTAs per SM (the register-capped kernel builds): dynamic + static shared memory of both must fit
static bool two_ctas_fit(const ModelP& P) { return 2 * (P.smem_floats * 4 + 8 * 1024) <= 227 * 1024; }
struct GroupedPlan {
    ModelP Pf, Pb;
    bool two_f, two_b;
    long long tiles_f, tiles_b;      // tiles over all groups
    int ctas_f, ctas_b;
    float* wpacked;                  // [G][n_packed]
    double* side_partial;            // [tiles_f][8]
    float* slabs;                    // [tiles_b][n_packed]
    long long bytes;
    long long tape_group_floats;
};
static bool grouped_single_launch(const UdeDesc* desc, int n_steps, void* base, GroupedPlan& gp) {
    const int G = desc->n_groups;
    if (G <= 1 || (desc->flags & UDE_FLAG_FORCE_TENSOR)) return false;
    UdeDesc d = *desc;
    d.n_groups = 1;                  // tile size from the TOTAL batch: all groups share one persistent grid
    if (choose_plan(&d, gp.Pf, F_NSTATE, 0, false) != UDE_OK || choose_plan(&d, gp.Pb, B_NSTATE, 0, true) != UDE_OK) return false;
    if (!gp.Pf.w_in_smem || gp.Pf.cold_global || !gp.Pb.w_in_smem || !gp.Pb.g_in_smem || gp.Pb.cold_global) return false;
    const long long Bg = desc->batch / G;
    gp.tiles_f = (Bg + gp.Pf.TB - 1) / gp.Pf.TB * G;
    gp.tiles_b = (Bg + gp.Pb.TB - 1) / gp.Pb.TB * G;
    const DeviceInfo& di = device_info();
    const int n_sm = di.ok ? di.n_sm : 148;
    gp.two_f = two_ctas_fit(gp.Pf);
    gp.two_b = two_ctas_fit(gp.Pb);
    const long long cap_f = (long long)n_sm * (gp.two_f ? 2 : 1), cap_b = (long long)n_sm * (gp.two_b ? 2 : 1);
    gp.ctas_f = (int)(gp.tiles_f < cap_f ? gp.tiles_f : cap_f);
    gp.ctas_b = (int)(gp.tiles_b < cap_b ? gp.tiles_b : cap_b);
    long long off = 0;
    gp.wpacked = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)G * gp.Pf.n_packed * 4, 256);
    gp.side_partial = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + gp.tiles_f * 8 * 8, 256);
    gp.slabs = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + gp.tiles_b * gp.Pb.n_packed * 4, 256);
    gp.bytes = off;
    gp.tape_group_floats = align_up((long long)n_steps * 4 * gp.Pf.Dn * Bg * 4, 256) / 4;
    return true;
}
}  // namespace ude

using namespace ude;

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int ude_abi_version(void) { return UDE_ABI_VERSION; }

const char* ude_strerror(int status) {
    switch (status) {
        case UDE_OK: return "ok";
        case UDE_ERR_INVALID: return "invalid argument or descriptor";
        case UDE_ERR_UNSUPPORTED: return "configuration not supported by the fused kernels";
        case UDE_ERR_WORKSPACE: return "workspace or tape too small (see ude_plan)";
        case UDE_ERR_CUDA: return "CUDA runtime error (see ude_last_cuda_error)";
        case UDE_ERR_NO_DEVICE: return "no sm_100 CUDA device visible";
        case UDE_ERR_GRID: return "time points must be strictly increasing and step_size positive";
        case UDE_ERR_CAPACITY: return "caller-provided table too small";
        default: return "unknown status";
    }
}

const char* ude_last_cuda_error(void) { return cuda_err_buf(); }

int64_t ude_theta_size(const UdeDesc* desc) {
    if (check_desc(desc) != UDE_OK) return UDE_ERR_INVALID;
    ModelP P;
    base_model(desc, P);
    return P.n_theta;
}

int64_t ude_macs_per_eval(const UdeDesc* desc) {
    if (check_desc(desc) != UDE_OK) return UDE_ERR_INVALID;
    ModelP P;
    base_model(desc, P);
    int64_t macs = 0;
    // algorithmic count of the reference's layers (the first layers at their full R*L width, hoisting or not)
    for (int i = 0; i < P.n_rate; ++i) macs += (int64_t)P.rate[i].th_ld * P.rate[i].n_out;
    for (int i = 0; i < P.n_aug; ++i) macs += (int64_t)P.aug[i].th_ld * P.aug[i].n_out;
    return macs;
}

int ude_plan(const UdeDesc* desc, int method, int n_steps, UdePlan* plan) {
    int st = check_desc(desc);
    if (st != UDE_OK || !plan || n_steps < 0) return UDE_ERR_INVALID;
    if (method != UDE_METHOD_RK4 && method != UDE_METHOD_DOPRI5) return UDE_ERR_INVALID;
    if (desc->n_groups != 1) {
        // G independent weight sets: G single-group solves side by side (group-major layouts, see ude_rk4_forward)
        if (method != UDE_METHOD_RK4) return UDE_ERR_UNSUPPORTED;
        UdeDesc d1 = *desc;
        d1.n_groups = 1;
        d1.batch = desc->batch / desc->n_groups;
        if ((st = ude_plan(&d1, method, n_steps, plan)) != UDE_OK) return st;
        plan->workspace_bytes = align_up(plan->workspace_bytes, 256) * desc->n_groups;
        plan->tape_bytes = align_up(plan->tape_bytes, 256) * desc->n_groups;
        {
            TcModel Mt;
            const int path = tc_choose(desc, Mt);
            if (path < 0) return path;
            if (path == 1) {          // tcgen05 engine, all groups in one launch
                int tc_ctas = 0;
                long long tc_ws = 0;
                tc_plan_info(desc, Mt, &tc_ws, &tc_ctas);
                if (tc_ws > plan->workspace_bytes) plan->workspace_bytes = tc_ws;
                plan->tile_traj = kTcTile;
                plan->n_ctas = tc_ctas;
                plan->smem_bytes_fwd = Mt.smem_bytes_fwd;
                plan->smem_bytes_bwd = Mt.smem_bytes_bwd;
                plan->weights_in_smem = 1;
                plan->path = 3;
                return UDE_OK;
            }
        }
        GroupedPlan gp;
        if (grouped_single_launch(desc, n_steps, nullptr, gp)) {      // one launch for all groups (same tape layout)
            if (gp.bytes > plan->workspace_bytes) plan->workspace_bytes = gp.bytes;
            plan->tile_traj = gp.Pb.TB;
            plan->n_ctas = gp.ctas_f > gp.ctas_b ? gp.ctas_f : gp.ctas_b;
            plan->smem_bytes_fwd = gp.Pf.smem_floats * 4;
            plan->smem_bytes_bwd = gp.Pb.smem_floats * 4;
            plan->weights_in_smem = 1;
            plan->path = 2;           // grouped single launch
        }
        return UDE_OK;
    }
    ModelP Pf, Pb;      // pure host arithmetic: works without a GPU (148 SMs assumed then)
    if ((st = choose_plan(desc, Pf, n_state_for(method, false), n_rbuf_for(method, false), false)) != UDE_OK) return st;
    if ((st = choose_plan(desc, Pb, n_state_for(method, true), n_rbuf_for(method, true), true)) != UDE_OK) return st;
    memset(plan, 0, sizeof(*plan));
    plan->n_theta = Pf.n_theta;
    const int ctas_f = grid_size(Pf, desc->batch), ctas_b = grid_size(Pb, desc->batch);
    const int n_ctas = ctas_f > ctas_b ? ctas_f : ctas_b;
    const int cold_max = Pf.cold_floats > Pb.cold_floats ? Pf.cold_floats : Pb.cold_floats;
    Workspace w = carve_workspace(nullptr, Pf, n_ctas, cold_max);
    plan->workspace_bytes = w.bytes;
    TcModel Mt;
    const int path = (method == UDE_METHOD_RK4 && !(desc->flags & UDE_FLAG_BAYES)) ? tc_choose(desc, Mt) : 0;
    if (path < 0) return path;
    // rk4: [n_steps][4 stages][Dn][B] stage inputs + ONE record for the final state (read by the decoded adjoint only)
    if (method == UDE_METHOD_RK4) plan->tape_bytes = align_up(((long long)n_steps * 4 + 1) * Pf.Dn * desc->batch * 4, 256);
    else plan->tape_bytes = dopri_tape_x_bytes(Pf, desc->batch, n_steps) + align_up((long long)n_steps * desc->batch * 16, 256);
    if (desc->flags & UDE_FLAG_BAYES) {       // per-evaluation weight sets (ude_rk4_*_bayes)
        if (method != UDE_METHOD_RK4) return UDE_ERR_UNSUPPORTED;
        UdeDesc db = *desc;
        db.flags &= ~UDE_FLAG_BAYES;
        TcModel Mb;
        const int pb = tc_choose(&db, Mb);       // tcgen05 engine where it covers the model, else the FFMA engine's all-resident plans
        if (pb < 0) return pb;
        if (pb == 1) {
            int tc_ctas = 0;
            plan->workspace_bytes = tc_bayes_workspace(desc, Mb, n_steps, &tc_ctas);
            plan->tape_bytes = align_up(((long long)n_steps * 4 + 1) * Pf.Dn * desc->batch * 4, 256);
            plan->tile_traj = kTcTile;
            plan->n_ctas = tc_ctas;
            plan->smem_bytes_fwd = Mb.smem_bytes_fwd;
            plan->smem_bytes_bwd = Mb.smem_bytes_bwd;
            plan->weights_in_smem = 1;
            plan->path = 1;
            return UDE_OK;
        }
        if (!Pf.w_in_smem || Pf.cold_global || !Pb.w_in_smem || !Pb.g_in_smem || Pb.cold_global) return UDE_ERR_UNSUPPORTED;
        plan->workspace_bytes = bayes_carve(nullptr, w.bytes, Pf, n_steps, n_ctas).bytes;
    }
    plan->tile_traj = Pb.TB;
    plan->n_ctas = n_ctas;
    plan->smem_bytes_fwd = Pf.smem_floats * 4;
    plan->smem_bytes_bwd = Pb.smem_floats * 4;
    plan->weights_in_smem = Pb.w_in_smem;
    if (path == 1) {
        int tc_ctas = 0;
        long long tc_ws = 0;
        tc_plan_info(desc, Mt, &tc_ws, &tc_ctas);
        if (tc_ws > plan->workspace_bytes) plan->workspace_bytes = tc_ws;
        plan->tile_traj = kTcTile;
        plan->n_ctas = tc_ctas;
        plan->smem_bytes_fwd = Mt.smem_bytes_fwd;
        plan->smem_bytes_bwd = Mt.smem_bytes_bwd;
        plan->weights_in_smem = 1;
        plan->path = 1;           // tensor path
    }
    return UDE_OK;
}

int ude_rk4_path(const UdeDesc* desc) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    TcModel Mt;
    return tc_choose(desc, Mt);
}

int ude_rk4_grid(const float* t, int32_t n_t, float step_size, int32_t use_step_size, int32_t cap_steps, float* dt,
                 int32_t* n_steps, int32_t* out_step, int32_t* out_kind, float* out_w) {
    if (!t || n_t < 2 || !dt || !n_steps || !out_step || !out_kind || !out_w) return UDE_ERR_INVALID;
    for (int i = 1; i < n_t; ++i)
        if (!(t[i] > t[i - 1])) return UDE_ERR_GRID;
    // grid points are generated on the fly: g(i) = fl32(fl32(i * h) + t0), last one overwritten by t[-1]
    long long niters;
    volatile float h = step_size;
    if (use_step_size) {
        if (!(step_size > 0.f)) return UDE_ERR_GRID;
        volatile float span = t[n_t - 1] - t[0];
        volatile float q = span / h;
        volatile float q1 = q + 1.0f;
        const float c = ceilf(q1);
        if (!(c >= 2.f) || c > 1e9f) return UDE_ERR_GRID;
        niters = (long long)c;
    } else {
        niters = n_t;
    }
    const long long ns = niters - 1;
    if (ns > cap_steps) {
        *n_steps = (int32_t)ns;
        return UDE_ERR_CAPACITY;
    }
    auto grid_at = [&](long long i) -> float {
        if (i == niters - 1) return t[n_t - 1];
        if (!use_step_size) return t[i];
        volatile float fi = (float)i;       // arange(0, niters, dtype=float32)
        volatile float p = fi * h;          // * step_size
        volatile float g = p + t[0];        // + start_time
        return g;
    };
    if (grid_at(0) != t[0]) return UDE_ERR_GRID;
    out_step[0] = -1;
    out_kind[0] = 0;
    out_w[0] = 0.f;
    int j = 1;
    for (long long n = 0; n < ns; ++n) {
        const float t0 = grid_at(n), t1 = grid_at(n + 1);
        volatile float d = t1 - t0;
        dt[n] = d;
        while (j < n_t && t1 >= t[j]) {
            out_step[j] = (int32_t)n;
            if (t[j] == t0) {
                out_kind[j] = 0;
                out_w[j] = 0.f;
            } else if (t[j] == t1) {
                out_kind[j] = 1;
                out_w[j] = 1.f;
            } else {
                volatile float num = t[j] - t0;
                volatile float den = t1 - t0;
                volatile float s = num / den;
                out_kind[j] = 2;
                out_w[j] = s;
            }
            ++j;
        }
    }
    if (j != n_t) return UDE_ERR_GRID;
    *n_steps = (int32_t)ns;
    return UDE_OK;
}

static int launch_pack(const ModelP& P, const float* theta, float* wpacked, cudaStream_t s) {
    pack_theta_kernel<<<P.n_rate + P.n_aug + 3, 256, 0, s>>>(P, theta, wpacked);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

}  // extern "C"

// ----------------------------------------------------------------------------------
// Weight-set groups (BASELINE configs[4]: tuning/tune_node.py-style sweeps, G independent Fp models trained side
// by side).  The reference trains every configuration in its own process with its own odeint call
// (tuning/tune_node.py:204,227); here ONE call carries G of them in group-major layouts
//   theta [G][n_theta], y0 [G][B_g][R][L], y_out [G][n_out][B_g][R][L], tape / workspace [G][per-group bytes of the plan],
//   side [G][8], g_side [G][5], g_theta [G][n_theta]
// and fans the G single-group launch sets out over an internal pool of streams that fork from / join into the
// caller's stream (event edges only: legal inside stream capture).  A group is 32 tiles of 64 trajectories at the
// reference's 64 x 32 samples: eight in flight fill the 148 SMs.
// ----------------------------------------------------------------------------------
constexpr int kFan = 8;
struct GroupFan {
    int dev = -1;
    cudaStream_t s[kFan];
    cudaEvent_t fork, join[kFan];
};
static int group_fan(GroupFan** out) {
    static thread_local GroupFan fans[kMaxDevices];      // one pool per device and thread, created once, never re-created
    int dev = 0;
    UDE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return UDE_ERR_INVALID;
    GroupFan& fan = fans[dev];
    if (fan.dev != dev) {
        for (int i = 0; i < kFan; ++i) {
            UDE_CUDA(cudaStreamCreateWithFlags(&fan.s[i], cudaStreamNonBlocking));
            UDE_CUDA(cudaEventCreateWithFlags(&fan.join[i], cudaEventDisableTiming));
        }
        UDE_CUDA(cudaEventCreateWithFlags(&fan.fork, cudaEventDisableTiming));
        fan.dev = dev;
    }
    *out = &fan;
    return UDE_OK;
}
template <class F>
static int for_each_group(const UdeDesc* desc, int n_steps, cudaStream_t user, F&& launch_one) {
    const int G = desc->n_groups;
    UdeDesc d1 = *desc;
    d1.n_groups = 1;
    d1.batch = desc->batch / G;
    UdePlan p1;
    int st = ude_plan(&d1, UDE_METHOD_RK4, n_steps, &p1);
    if (st != UDE_OK) return st;
    const long long ws_g = align_up(p1.workspace_bytes, 256), tape_g = align_up(p1.tape_bytes, 256);
    GroupFan* fan = nullptr;
    if ((st = group_fan(&fan)) != UDE_OK) return st;
    const int n_fan = G < kFan ? G : kFan;
    UDE_CUDA(cudaEventRecord(fan->fork, user));
    for (int i = 0; i < n_fan; ++i) UDE_CUDA(cudaStreamWaitEvent(fan->s[i], fan->fork, 0));
    for (int g = 0; g < G; ++g)
        if ((st = launch_one(d1, p1, g, ws_g, tape_g, fan->s[g % n_fan])) != UDE_OK) break;
    for (int i = 0; i < n_fan; ++i) {       // always join, also after a failed launch
        cudaEventRecord(fan->join[i], fan->s[i]);
        cudaStreamWaitEvent(user, fan->join[i], 0);
    }
    return st;
}

extern "C" {

int ude_rk4_forward(const UdeDesc* desc, const float* theta, const float* y0, int32_t n_steps, const float* dt,
                    int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w, float* y_out,
                    float* tape, double* side, void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    if (!theta || !y0 || !dt || !out_step || !out_kind || !out_w || !y_out || !side || !workspace) return UDE_ERR_INVALID;
    if (n_steps < 1 || n_out < 1) return UDE_ERR_INVALID;
    if (!device_info().ok) return UDE_ERR_NO_DEVICE;
    if (desc->n_groups != 1) {
        {   // the tensor-core kernels walk the tiles of ALL groups in one launch when they cover the model
            TcModel Mt;
            const int path = tc_choose(desc, Mt);
            if (path < 0) return path;
            if (path == 1)
                return tc_forward(desc, Mt, theta, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, y_out, tape, side,
                                  workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
        }
        GroupedPlan gp;
        if (grouped_single_launch(desc, n_steps, workspace, gp)) {
            if (gp.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
            cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
            const ModelP& P = gp.Pf;
            pack_theta_kernel<<<dim3(P.n_rate + P.n_aug + 3, desc->n_groups), 256, 0, s>>>(P, theta, gp.wpacked);
            UDE_CUDA(cudaGetLastError());
            Rk4FwdArgs a{};
            a.wpacked = gp.wpacked;
            a.y0 = y0;
            a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
            a.y_out = y_out;
            a.tape = tape;
            a.side_partial = gp.side_partial;
            a.cold = nullptr;
            a.B = desc->batch / desc->n_groups;
            a.n_groups = desc->n_groups;
            a.tape_group_floats = gp.tape_group_floats;
            const int smem = P.smem_floats * 4;
            if (gp.two_f) {
                if ((st = set_smem(rk4_fwd_kernel<true, false, 2>, smem)) != UDE_OK) return st;
                rk4_fwd_kernel<true, false, 2><<<gp.ctas_f, kThreads, smem, s>>>(P, a);
            } else {
                if ((st = set_smem(rk4_fwd_kernel<true, false>, smem)) != UDE_OK) return st;
                rk4_fwd_kernel<true, false><<<gp.ctas_f, kThreads, smem, s>>>(P, a);
            }
            UDE_CUDA(cudaGetLastError());
            reduce_side_kernel<<<desc->n_groups, 32, 0, s>>>(gp.side_partial, (int)(gp.tiles_f / desc->n_groups), side);
            UDE_CUDA(cudaGetLastError());
            return UDE_OK;
        }
        const long long n_theta = ude_theta_size(desc), row = (long long)desc->n_regions * desc->latent_dim;
        return for_each_group(desc, n_steps, reinterpret_cast<cudaStream_t>(stream),
                              [&](const UdeDesc& d1, const UdePlan& p1, int g, long long ws_g, long long tape_g, cudaStream_t sg) {
                                  const long long Bg = d1.batch;
                                  return ude_rk4_forward(&d1, theta + g * n_theta, y0 + g * Bg * row, n_steps, dt, n_out, out_step, out_kind,
                                                         out_w, y_out + g * (long long)n_out * Bg * row,
                                                         tape ? reinterpret_cast<float*>(reinterpret_cast<char*>(tape) + g * tape_g) : nullptr,
                                                         side + g * 8, reinterpret_cast<char*>(workspace) + g * ws_g,
                                                         workspace_bytes - g * ws_g < p1.workspace_bytes ? 0 : p1.workspace_bytes, sg);
                              });
    }
    {
        TcModel Mt;
        const int path = tc_choose(desc, Mt);
        if (path < 0) return path;
        if (path == 1)
            return tc_forward(desc, Mt, theta, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, y_out, tape,
                              side, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
    }
    ModelP P;
    if ((st = choose_plan(desc, P, F_NSTATE, 0, false)) != UDE_OK) return st;
    const int n_ctas = grid_size(P, desc->batch);
    Workspace w = carve_workspace(workspace, P, n_ctas, P.cold_global ? P.cold_floats : 0);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if ((st = launch_pack(P, theta, w.wpacked, s)) != UDE_OK) return st;
    Rk4FwdArgs a{};
    a.wpacked = w.wpacked;
    a.y0 = y0;
    a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
    a.y_out = y_out;
    a.tape = tape;
    a.side_partial = w.side_partial;
    a.cold = w.cold;
    a.B = desc->batch;
    a.n_groups = 1;
    a.tape_group_floats = 0;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    const int smem = P.smem_floats * 4;
#define UDE_LAUNCH_FWD(W, C)                                                              \
    do {                                                                                  \
        if ((st = set_smem(rk4_fwd_kernel<W, C>, smem)) != UDE_OK) return st;             \
        rk4_fwd_kernel<W, C><<<n_ctas, kThreads, smem, s>>>(P, a);                        \
    } while (0)
    if (P.cold_global) UDE_LAUNCH_FWD(false, true);
    else if (P.w_in_smem) UDE_LAUNCH_FWD(true, false);
    else UDE_LAUNCH_FWD(false, false);
#undef UDE_LAUNCH_FWD
    UDE_CUDA(cudaGetLastError());
    reduce_side_kernel<<<1, 32, 0, s>>>(w.side_partial, n_ctas, side);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int ude_rk4_backward(const UdeDesc* desc, const float* theta, const float* y0, int32_t n_steps, const float* dt,
                     int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                     const float* tape, const float* g_out, const float* g_side, float* g_y0, float* g_theta,
                     void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    if (!theta || !y0 || !dt || !out_step || !out_kind || !out_w || !tape || !g_out || !g_y0 || !g_theta || !workspace)
        return UDE_ERR_INVALID;
    if (n_steps < 1 || n_out < 1) return UDE_ERR_INVALID;
    if (!device_info().ok) return UDE_ERR_NO_DEVICE;
    if (desc->n_groups != 1) {
        {   // the tensor-core kernels walk the tiles of ALL groups in one launch when they cover the model
            TcModel Mt;
            const int path = tc_choose(desc, Mt);
            if (path < 0) return path;
            if (path == 1)
                return tc_backward(desc, Mt, theta, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, tape, g_out, g_side,
                                   g_y0, g_theta, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
        }
        GroupedPlan gp;
        if (grouped_single_launch(desc, n_steps, workspace, gp)) {
            if (gp.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
            cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
            const ModelP& P = gp.Pb;
            if (!(desc->flags & UDE_FLAG_WEIGHTS_PACKED)) {
                pack_theta_kernel<<<dim3(P.n_rate + P.n_aug + 3, desc->n_groups), 256, 0, s>>>(P, theta, gp.wpacked);
                UDE_CUDA(cudaGetLastError());
            }
            Rk4BwdArgs a{};
            a.wpacked = gp.wpacked;
            a.y0 = y0;
            a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
            a.tape = tape;
            a.g_out = g_out;
            a.g_side = g_side;
            a.g_y0 = g_y0;
            a.slabs = gp.slabs;
            a.cold = nullptr;
            a.B = desc->batch / desc->n_groups;
            a.n_groups = desc->n_groups;
            a.tape_group_floats = gp.tape_group_floats;
            const int smem = P.smem_floats * 4;
            if (gp.two_b) {
                if ((st = set_smem(rk4_bwd_kernel<true, true, false, 2>, smem)) != UDE_OK) return st;
                rk4_bwd_kernel<true, true, false, 2><<<gp.ctas_b, kThreads, smem, s>>>(P, a);
            } else {
                if ((st = set_smem(rk4_bwd_kernel<true, true, false>, smem)) != UDE_OK) return st;
                rk4_bwd_kernel<true, true, false><<<gp.ctas_b, kThreads, smem, s>>>(P, a);
            }
            UDE_CUDA(cudaGetLastError());
            reduce_dtheta_kernel<<<dim3((P.n_theta + 255) / 256, desc->n_groups), 256, 0, s>>>(
                P, gp.slabs, (int)(gp.tiles_b / desc->n_groups), g_theta);
            UDE_CUDA(cudaGetLastError());
            return UDE_OK;
        }
        const long long n_theta = ude_theta_size(desc), row = (long long)desc->n_regions * desc->latent_dim;
        return for_each_group(desc, n_steps, reinterpret_cast<cudaStream_t>(stream),
                              [&](const UdeDesc& d1, const UdePlan& p1, int g, long long ws_g, long long tape_g, cudaStream_t sg) {
                                  const long long Bg = d1.batch;
                                  return ude_rk4_backward(&d1, theta + g * n_theta, y0 + g * Bg * row, n_steps, dt, n_out, out_step, out_kind,
                                                          out_w, reinterpret_cast<const float*>(reinterpret_cast<const char*>(tape) + g * tape_g),
                                                          g_out + g * (long long)n_out * Bg * row, g_side ? g_side + g * 5 : nullptr,
                                                          g_y0 + g * Bg * row, g_theta + g * n_theta,
                                                          reinterpret_cast<char*>(workspace) + g * ws_g,
                                                          workspace_bytes - g * ws_g < p1.workspace_bytes ? 0 : p1.workspace_bytes, sg);
                              });
    }
    {
        TcModel Mt;
        const int path = tc_choose(desc, Mt);
        if (path < 0) return path;
        if (path == 1)
            return tc_backward(desc, Mt, theta, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, tape, g_out,
                               g_side, g_y0, g_theta, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
    }
    ModelP P;
    if ((st = choose_plan(desc, P, B_NSTATE, 0, true)) != UDE_OK) return st;
    const int n_ctas = grid_size(P, desc->batch);
    Workspace w = carve_workspace(workspace, P, n_ctas, P.cold_global ? P.cold_floats : 0);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!(desc->flags & UDE_FLAG_WEIGHTS_PACKED) && (st = launch_pack(P, theta, w.wpacked, s)) != UDE_OK) return st;
    Rk4BwdArgs a{};
    a.wpacked = w.wpacked;
    a.y0 = y0;
    a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
    a.tape = tape;
    a.g_out = g_out;
    a.g_side = g_side;
    a.g_y0 = g_y0;
    a.slabs = w.slabs;
    a.cold = w.cold;
    a.B = desc->batch;
    a.n_groups = 1;
    a.tape_group_floats = 0;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    decoded_ctas() = n_ctas;
    const int smem = P.smem_floats * 4;
#define UDE_LAUNCH_BWD(W, G, C)                                                           \
    do {                                                                                  \
        if ((st = set_smem(rk4_bwd_kernel<W, G, C>, smem)) != UDE_OK) return st;          \
        rk4_bwd_kernel<W, G, C><<<n_ctas, kThreads, smem, s>>>(P, a);                     \
    } while (0)
    if (P.cold_global) UDE_LAUNCH_BWD(false, false, true);
    else if (P.w_in_smem && P.g_in_smem) UDE_LAUNCH_BWD(true, true, false);
    else if (P.w_in_smem) UDE_LAUNCH_BWD(true, false, false);
    else UDE_LAUNCH_BWD(false, false, false);
#undef UDE_LAUNCH_BWD
    UDE_CUDA(cudaGetLastError());
    reduce_dtheta_kernel<<<(P.n_theta + 255) / 256, 256, 0, s>>>(P, w.slabs, n_ctas, g_theta);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

#ifdef UDE_PROFILE
/* debug builds only: read and clear the per-phase cycle counters of CTA 0 */
int ude_debug_read_prof(unsigned long long* out /*[kWarps*32*2]*/) {
    UDE_CUDA(cudaDeviceSynchronize());
    UDE_CUDA(cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * kWarps * 32 * 2));
    static unsigned long long zeros[kWarps * 32 * 2];
    UDE_CUDA(cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros)));
    return UDE_OK;
}
#endif

static int check_reparam(const UdeDesc* desc, const UdeReparam* rp) {
    if (!desc || !rp || !rp->mean || (!rp->eps && rp->std) || (rp->eps && !rp->std)) return UDE_ERR_INVALID;
    if (rp->n_samples < 1 || rp->batch < 1 || (int64_t)rp->n_samples * rp->batch != desc->batch) return UDE_ERR_INVALID;
    if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
    return UDE_OK;
}

int ude_rk4_forward_reparam(const UdeDesc* desc, const float* theta, const UdeReparam* rp, int32_t n_steps, const float* dt,
                            int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w, float* y_out,
                            float* tape, double* side, void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_reparam(desc, rp);
    if (st != UDE_OK) return st;
    pending_reparam() = ReparamArgs{rp->eps, rp->mean, rp->std, nullptr, rp->batch};
    st = ude_rk4_forward(desc, theta, /*y0 = output 0, produced by the kernel*/ y_out, n_steps, dt, n_out, out_step, out_kind, out_w,
                         y_out, tape, side, workspace, workspace_bytes, stream);
    pending_reparam() = ReparamArgs{};
    return st;
}

int ude_rk4_backward_reparam(const UdeDesc* desc, const float* theta, const UdeReparam* rp, const float* y0, int32_t n_steps,
                             const float* dt, int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                             const float* tape, const float* g_out, const float* g_side, float* g_y0, float* g_z, float* g_theta,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_reparam(desc, rp);
    if (st != UDE_OK) return st;
    if (!g_z) return UDE_ERR_INVALID;
    pending_reparam() = ReparamArgs{rp->eps, rp->mean, rp->std, g_z, rp->batch};
    st = ude_rk4_backward(desc, theta, y0, n_steps, dt, n_out, out_step, out_kind, out_w, tape, g_out, g_side, g_y0, g_theta,
                          workspace, workspace_bytes, stream);
    pending_reparam() = ReparamArgs{};
    return st;
}

// ---- Bayesian ODE funcs: one weight set per RHS evaluation (include/ude_b200.h) -----------------------------------------
static int bayes_common(const UdeDesc* desc, int n_steps, void* workspace, int64_t workspace_bytes, UdeDesc* d, ModelP* Pf, ModelP* Pb,
                        int* n_ctas, Workspace* w, BayesWs* bw) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    if (desc->n_groups != 1 || (desc->flags & UDE_FLAG_FORCE_TENSOR)) return UDE_ERR_UNSUPPORTED;
    if (n_steps < 1 || !workspace) return UDE_ERR_INVALID;
    if (!device_info().ok) return UDE_ERR_NO_DEVICE;
    *d = *desc;
    d->flags |= UDE_FLAG_BAYES | UDE_FLAG_FORCE_FFMA;
    if ((st = choose_plan(d, *Pf, F_NSTATE, 0, false)) != UDE_OK) return st;
    if ((st = choose_plan(d, *Pb, B_NSTATE, 0, true)) != UDE_OK) return st;
    if (!Pf->w_in_smem || Pf->cold_global || !Pb->w_in_smem || !Pb->g_in_smem || Pb->cold_global) return UDE_ERR_UNSUPPORTED;
    const int cf = grid_size(*Pf, d->batch), cb = grid_size(*Pb, d->batch);
    *n_ctas = cf > cb ? cf : cb;
    *w = carve_workspace(workspace, *Pf, *n_ctas, 0);
    *bw = bayes_carve(workspace, w->bytes, *Pf, n_steps, *n_ctas);
    if (bw->bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    return UDE_OK;
}

int ude_rk4_forward_bayes(const UdeDesc* desc, const float* theta_evals, const float* y0, int32_t n_steps, const float* dt,
                          int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w, float* y_out,
                          float* tape, double* side, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!theta_evals || !y0 || !dt || !out_step || !out_kind || !out_w || !y_out || !side || n_out < 1) return UDE_ERR_INVALID;
    {   // the tcgen05 engine where it covers the model (R = 1, nets up to 64 wide), else the FFMA engine's all-resident plans
        int st0 = check_desc(desc);
        if (st0 != UDE_OK) return st0;
        if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
        if (n_steps < 1 || !workspace) return UDE_ERR_INVALID;
        if (!device_info().ok) return UDE_ERR_NO_DEVICE;
        UdeDesc db = *desc;
        db.flags &= ~UDE_FLAG_BAYES;
        TcModel Mt;
        const int path = tc_choose(&db, Mt);
        if (path < 0) return path;
        if (path == 1) {
            return tc_bayes_forward(desc, Mt, theta_evals, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, y_out, tape, side,
                                    workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
        }
    }
    UdeDesc d;
    ModelP Pf, Pb;
    Workspace w;
    BayesWs bw;
    int n_max = 0;
    int st = bayes_common(desc, n_steps, workspace, workspace_bytes, &d, &Pf, &Pb, &n_max, &w, &bw);
    if (st != UDE_OK) return st;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    pack_theta_kernel<<<dim3(Pf.n_rate + Pf.n_aug + 3, 4 * n_steps), 256, 0, s>>>(Pf, theta_evals, bw.w);
    UDE_CUDA(cudaGetLastError());
    const int n_ctas = grid_size(Pf, d.batch);
    Rk4FwdArgs a{};
    a.wpacked = bw.w;
    a.y0 = y0;
    a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
    a.y_out = y_out;
    a.tape = tape;
    a.side_partial = w.side_partial;
    a.B = d.batch;
    a.n_groups = 1;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    a.bayes = BayesArgs{bw.w, nullptr, nullptr};
    const int smem = Pf.smem_floats * 4;
    if ((st = set_smem(rk4_fwd_kernel<true, false>, smem)) != UDE_OK) return st;
    rk4_fwd_kernel<true, false><<<n_ctas, kThreads, smem, s>>>(Pf, a);
    UDE_CUDA(cudaGetLastError());
    reduce_side_kernel<<<1, 32, 0, s>>>(w.side_partial, n_ctas, side);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int ude_rk4_backward_bayes(const UdeDesc* desc, const float* theta_evals, const float* z_evals, const float* y0, int32_t n_steps,
                           const float* dt, int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                           const float* tape, const float* g_out, const float* g_side, float* g_y0, float* g_mean_zsum,
                           void* workspace, int64_t workspace_bytes, void* stream) {
    if (!theta_evals || !z_evals || !y0 || !dt || !out_step || !out_kind || !out_w || !tape || !g_out || !g_y0 || !g_mean_zsum || n_out < 1)
        return UDE_ERR_INVALID;
    {
        int st0 = check_desc(desc);
        if (st0 != UDE_OK) return st0;
        if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
        if (n_steps < 1 || !workspace) return UDE_ERR_INVALID;
        if (!device_info().ok) return UDE_ERR_NO_DEVICE;
        UdeDesc db = *desc;
        db.flags &= ~UDE_FLAG_BAYES;
        TcModel Mt;
        const int path = tc_choose(&db, Mt);
        if (path < 0) return path;
        if (path == 1) {
            return tc_bayes_backward(desc, Mt, theta_evals, z_evals, y0, GridTables{n_steps, dt, n_out, out_step, out_kind, out_w}, tape,
                                     g_out, g_side, g_y0, g_mean_zsum, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
        }
    }
    UdeDesc d;
    ModelP Pf, Pb;
    Workspace w;
    BayesWs bw;
    int n_max = 0;
    int st = bayes_common(desc, n_steps, workspace, workspace_bytes, &d, &Pf, &Pb, &n_max, &w, &bw);
    if (st != UDE_OK) return st;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const dim3 pgrid(Pb.n_rate + Pb.n_aug + 3, 4 * n_steps);
    if (!(desc->flags & UDE_FLAG_WEIGHTS_PACKED)) {
        pack_theta_kernel<<<pgrid, 256, 0, s>>>(Pb, theta_evals, bw.w);
        UDE_CUDA(cudaGetLastError());
    }
    pack_theta_kernel<<<pgrid, 256, 0, s>>>(Pb, z_evals, bw.z);       // the packing is a permutation + zero padding: the same map serves z
    UDE_CUDA(cudaGetLastError());
    const int n_ctas = grid_size(Pb, d.batch);
    // the fold blocks are laid out for the CTAs of THIS launch: [2][n_ctas][n_packed] (inside the plan's [2][n_max][n_packed])
    Rk4BwdArgs a{};
    a.wpacked = bw.w;
    a.y0 = y0;
    a.g = GridTables{n_steps, dt, n_out, out_step, out_kind, out_w};
    a.tape = tape;
    a.g_out = g_out;
    a.g_side = g_side;
    a.g_y0 = g_y0;
    a.slabs = w.slabs;
    a.B = d.batch;
    a.n_groups = 1;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    decoded_ctas() = n_ctas;
    a.bayes = BayesArgs{bw.w, bw.z, bw.fold};
    const int smem = Pb.smem_floats * 4;
    if ((st = set_smem(rk4_bwd_kernel<true, true, false>, smem)) != UDE_OK) return st;
    rk4_bwd_kernel<true, true, false><<<n_ctas, kThreads, smem, s>>>(Pb, a);
    UDE_CUDA(cudaGetLastError());
    reduce_dtheta_kernel<<<dim3((Pb.n_theta + 255) / 256, 2), 256, 0, s>>>(Pb, bw.fold, n_ctas, g_mean_zsum);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

static int check_decoder(const UdeDesc* desc, const UdeDecoder* dec) {
    if (!desc || !dec || !dec->weight || !dec->bias || dec->n_out < 1) return UDE_ERR_INVALID;
    if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
    return UDE_OK;
}
// the tensor-core kernels cover the reference's decoder shape at R = 1 (one output per region); anything wider runs the FFMA engine
static int decoded_desc(const UdeDesc* desc, const UdeDecoder* dec, UdeDesc* d) {
    *d = *desc;
    if (dec->n_out != 1) {
        if (d->flags & UDE_FLAG_FORCE_TENSOR) return UDE_ERR_UNSUPPORTED;
        d->flags |= UDE_FLAG_FORCE_FFMA;
    }
    return UDE_OK;
}

int ude_rk4_forward_decoded(const UdeDesc* desc, const float* theta, const float* y0, const UdeReparam* rp, int32_t n_steps,
                            const float* dt, int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                            const UdeDecoder* dec, float* y_pred, uint8_t* out_mask, float* y0_out, float* tape, double* side,
                            void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_decoder(desc, dec);
    if (st != UDE_OK) return st;
    if (!y_pred || !out_mask || (rp ? !y0_out : !y0)) return UDE_ERR_INVALID;
    UdeDesc d;
    if ((st = decoded_desc(desc, dec, &d)) != UDE_OK) return st;
    pending_decoded() = DecArgs{dec->weight, dec->bias, dec->n_out, y_pred, out_mask, nullptr, nullptr};
    if (rp)
        st = ude_rk4_forward_reparam(&d, theta, rp, n_steps, dt, n_out, out_step, out_kind, out_w, y0_out, tape, side, workspace,
                                     workspace_bytes, stream);
    else    // y_out is not written in decoded mode without the reparam prologue: any non-null pointer passes the checks
        st = ude_rk4_forward(&d, theta, y0, n_steps, dt, n_out, out_step, out_kind, out_w, const_cast<float*>(y0), tape, side,
                             workspace, workspace_bytes, stream);
    pending_decoded() = DecArgs{};
    return st;
}

int ude_rk4_backward_decoded(const UdeDesc* desc, const float* theta, const float* y0, const UdeReparam* rp, int32_t n_steps,
                             const float* dt, int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                             const UdeDecoder* dec, const float* tape, const uint8_t* out_mask, const float* g_pred,
                             const float* g_side, float* g_y0, float* g_z, float* g_theta, float* g_dec, float* dec_partial,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_decoder(desc, dec);
    if (st != UDE_OK) return st;
    if (!out_mask || !g_pred || !g_dec || !dec_partial || !y0) return UDE_ERR_INVALID;
    UdeDesc d;
    if ((st = decoded_desc(desc, dec, &d)) != UDE_OK) return st;
    pending_decoded() = DecArgs{dec->weight, dec->bias, dec->n_out, nullptr, const_cast<uint8_t*>(out_mask), g_pred, dec_partial};
    decoded_ctas() = 0;
    if (rp)
        st = ude_rk4_backward_reparam(&d, theta, rp, y0, n_steps, dt, n_out, out_step, out_kind, out_w, tape, /*g_out unused*/ g_pred,
                                      g_side, g_y0, g_z, g_theta, workspace, workspace_bytes, stream);
    else
        st = ude_rk4_backward(&d, theta, y0, n_steps, dt, n_out, out_step, out_kind, out_w, tape, /*g_out unused*/ g_pred, g_side,
                              g_y0, g_theta, workspace, workspace_bytes, stream);
    pending_decoded() = DecArgs{};
    if (st != UDE_OK) return st;
    const int n = dec->n_out * (3 * desc->n_regions + 1);
    reduce_dec_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dec_partial, decoded_ctas(), n, g_dec);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

// Bayes funcs with the fused consumers (and, optionally, the fused reparam producer): the decoded entry points over
// ude_rk4_*_bayes.  theta_evals / z_evals / g_mean_zsum as there, everything else as in ude_rk4_*_decoded.
int ude_rk4_forward_bayes_decoded(const UdeDesc* desc, const float* theta_evals, const float* y0, const UdeReparam* rp, int32_t n_steps,
                                  const float* dt, int32_t n_out, const int32_t* out_step, const int32_t* out_kind, const float* out_w,
                                  const UdeDecoder* dec, float* y_pred, uint8_t* out_mask, float* y0_out, float* tape, double* side,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_decoder(desc, dec);
    if (st != UDE_OK) return st;
    if (!y_pred || !out_mask || (rp ? !y0_out : !y0)) return UDE_ERR_INVALID;
    if (rp && (st = check_reparam(desc, rp)) != UDE_OK) return st;
    UdeDesc d;
    if ((st = decoded_desc(desc, dec, &d)) != UDE_OK) return st;
    pending_decoded() = DecArgs{dec->weight, dec->bias, dec->n_out, y_pred, out_mask, nullptr, nullptr};
    if (rp) pending_reparam() = ReparamArgs{rp->eps, rp->mean, rp->std, nullptr, rp->batch};
    st = ude_rk4_forward_bayes(&d, theta_evals, rp ? y0_out : y0, n_steps, dt, n_out, out_step, out_kind, out_w,
                               rp ? y0_out : const_cast<float*>(y0), tape, side, workspace, workspace_bytes, stream);
    pending_decoded() = DecArgs{};
    pending_reparam() = ReparamArgs{};
    return st;
}

int ude_rk4_backward_bayes_decoded(const UdeDesc* desc, const float* theta_evals, const float* z_evals, const float* y0,
                                   const UdeReparam* rp, int32_t n_steps, const float* dt, int32_t n_out, const int32_t* out_step,
                                   const int32_t* out_kind, const float* out_w, const UdeDecoder* dec, const float* tape,
                                   const uint8_t* out_mask, const float* g_pred, const float* g_side, float* g_y0, float* g_z,
                                   float* g_mean_zsum, float* g_dec, float* dec_partial, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
    int st = check_decoder(desc, dec);
    if (st != UDE_OK) return st;
    if (!out_mask || !g_pred || !g_dec || !dec_partial || !y0 || (rp && !g_z)) return UDE_ERR_INVALID;
    if (rp && (st = check_reparam(desc, rp)) != UDE_OK) return st;
    UdeDesc d;
    if ((st = decoded_desc(desc, dec, &d)) != UDE_OK) return st;
    pending_decoded() = DecArgs{dec->weight, dec->bias, dec->n_out, nullptr, const_cast<uint8_t*>(out_mask), g_pred, dec_partial};
    if (rp) pending_reparam() = ReparamArgs{rp->eps, rp->mean, rp->std, g_z, rp->batch};
    decoded_ctas() = 0;
    st = ude_rk4_backward_bayes(&d, theta_evals, z_evals, y0, n_steps, dt, n_out, out_step, out_kind, out_w, tape, /*g_out unused*/ g_pred,
                                g_side, g_y0, g_mean_zsum, workspace, workspace_bytes, stream);
    pending_decoded() = DecArgs{};
    pending_reparam() = ReparamArgs{};
    if (st != UDE_OK) return st;
    const int n = dec->n_out * (3 * desc->n_regions + 1);
    reduce_dec_kernel<<<(n + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dec_partial, decoded_ctas(), n, g_dec);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int ude_measure_ffma_tflops(double* tflops, void* stream) {
    if (!tflops) return UDE_ERR_INVALID;
    const DeviceInfo& di = device_info();
    if (!di.ok) return UDE_ERR_NO_DEVICE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    float* out = nullptr;
    UDE_CUDA(cudaMalloc(&out, 4));
    cudaEvent_t e0, e1;
    UDE_CUDA(cudaEventCreate(&e0));
    UDE_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = di.n_sm * 8, threads = 256;
    ffma_peak_kernel<<<blocks, threads, 0, s>>>(out, 256, 1.0f);     // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        UDE_CUDA(cudaEventRecord(e0, s));
        ffma_peak_kernel<<<blocks, threads, 0, s>>>(out, iters, 1.0f);
        UDE_CUDA(cudaEventRecord(e1, s));
        UDE_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        UDE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return UDE_OK;
}

}  // extern "C"

#include "ude_dopri5_host.inl"
