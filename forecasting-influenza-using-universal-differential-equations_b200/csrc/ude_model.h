// ude_model.h — host+device description of one UDE func as the kernels see it.
//
// The reference builds its ODE funcs as ModuleLists
//   Fp   : Flatten, Linear(RL,s0), [ELU, Linear]*, Linear(s_last, 2R)   (/root/reference/lib/models.py:116-124)
//   Fa   : Linear(RL,a0), [ELU, Linear]*, Linear(a_last, 3R)            (:167-172)
//   FaFp : both                                                         (:208-223)
// ModelP is that structure flattened into plain ints plus the shared-memory
// plan of one CTA tile.  Built on the host by ude::build_model (ude_host.cu),
// passed to kernels by value and copied to shared memory there.
#pragma once
#include <stdint.h>

namespace ude {

#ifndef UDE_THREADS
#define UDE_THREADS 256
#endif
constexpr int kThreads = UDE_THREADS;     // threads per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMaxLayers = 8;     // == UDE_MAX_LAYERS
constexpr int kPadTail = 64;      // floats of zero padding after the packed weights

struct LayerP {
    int n_in, n_out;        // true nn.Linear dims
    int n_in_p, n_out_p;    // rounded up to a multiple of 4
    int ldw;                // row stride of the packed W (floats); never == 0 mod 32
    int w_off, b_off;       // offsets into the packed blob (floats)
    int th_w, th_b;         // offsets into theta (PyTorch layout)
    int act_off;            // smem offset (floats) of this layer's output rows [n_out_p][TBP]
    int in_off;             // smem offset of its input rows (X for layer 0)
    int elu;                // ELU on the output (hidden layers except the last one before the head)
    int tn_fwd;             // n-tile shape of the forward micro-kernel: 8 | 4 | 1 outputs per thread
    int tn_bwd;             // 8 | 4 inputs per thread in the delta micro-kernel
    int tm_dw, tn_dw;       // 4 | 1 each, weight-gradient micro-kernel
    int fwd_nt, bwd_nt;     // n-tiles of the forward / delta micro-kernels (host-computed: no device divisions)
    int dw_mt, dw_nt;       // m- / n-tiles of the weight-gradient micro-kernel
    int c0_off;             // >= 0: smem rows [n_out_p][TBP] added INSTEAD of the bias (first layers with hoisted constants)
    int th_ld;              // row stride of W inside theta (== n_in except for the split first layer, where it is R*L)
    int no_db;              // weight-gradient units skip the bias gradient (constant block: the live block owns it)
    int bc_off;             // compact bias offset (sum of n_out_p of the layers before): bias sums of the TMEM-accumulator mode
};

struct ModelP {
    int kind;               // UdeKind
    int R, L;
    int D, D_p;             // R*L and padded to 4  : width of the reference's MLP input u = vec(x) (lib/models.py:119)
    int Dn, Dn_p;           // 3R and padded to 4   : rows of the dynamic state AND of X (row = r*3 + c), see "hoisting"
    int Dc, Dc_p;           // (L-3)*R and padded   : constant inputs (row = r*(L-3) + l-3)
    int R_p;                // R padded to 4        : rows of per-region scratch
    int n_rate, n_aug;
    LayerP rate[kMaxLayers];
    LayerP aug[kMaxLayers];
    // Hoisting (lib/models.py:144,184,249: d/dt of components l >= 3 is identically zero).  Only 3 of the L latent
    // components of a region move, so W0 . u splits into a LIVE product over the 3R dynamic inputs (every RHS
    // evaluation) and a CONSTANT product over the (L-3)R others, done ONCE per tile and solve:
    //     c0 = W0[:, const] . x_const + b0          (rows [n_out_p][TBP] at LayerP::c0_off of the first layers)
    // and in the adjoint  dL/dx_const = W0[:, const]^T . D,  dW0[:, const] = D (x) x_const  with  D = sum over
    // evaluations of delta_0 (off_DL).  rate[0] / aug[0] describe the live block, rate_c / aug_c the constant one.
    LayerP rate_c, aug_c;
    int hoist;              // Dc > 0
    int bayes;              // weights change at EVERY RHS evaluation (Bayes_* funcs, lib/in_development/models_bayes.py:30-48): the
                            // constant product is redone per evaluation, so the tile's constants keep their own rows (off_XC does
                            // not alias the ACT rows) and the adjoint sums dL/dx_const in off_GXC
    int n_theta;            // floats of one weight set (PyTorch layout)
    int n_packed;           // floats of the packed blob incl. kPadTail
    int n_packed_hot;       // leading part used by every RHS evaluation (all but the constant blocks of the first layers)
    int dl_rows;            // rows of off_DL
    float fa_w;
    // ---- shared-memory plan (float offsets into the dynamic smem arena) ----
    int TB, TBP;            // trajectories per tile (32 | 64), padded row stride (TB + 4)
    int tb_shift;           // log2(TB)
    int mt_shift;           // log2(TB / 32): m-tiles per layer are a power of two
    int w_in_smem;          // packed weights resident in smem
    int g_in_smem;          // (bwd) dtheta accumulators resident in smem
    int off_W, off_G;       // valid when the flags above are set
    int off_X;              // [Dn_p][TBP] live inputs
    int off_XC;             // [Dc_p][TBP] constant inputs of the tile (aliases the ACT rows: tile prologue / epilogue only)
    int off_GXC;            // (bwd, bayes) [Dc_p][TBP] sum over evaluations of W0[:, const]^T delta_0
    int off_DL;             // (bwd) [n0_rate_p + n0_aug_p][TBP] sum over evaluations of the first layers' deltas
    int off_S;              // state rows: n_state buffers of [Dn_p][TBP]
    int off_GX;             // (bwd) dL/dx_live of one RHS evaluation, [Dn_p][TBP]
    int n_state;
    int off_RB, n_rbuf;     // per-region scratch: n_rbuf buffers of [R_p][TBP] (dopri5)
    int tm_acc;             // (bwd) weight-gradient accumulators in tensor memory (ude_engine.cuh: TmIo), bias sums at off_GB
    int tm_dw_prefix[kMaxLayers]; // weight-gradient units of all phases before depth s (counted from the heads): unit u of
                                  // depth s is unit g = prefix + u of the launch, runs on warp g % kWarps and owns slot g / kWarps
    int tm_cols;            // accumulator columns in use
    // ---- ring mode (ude_ring.cuh): weights through a 2-slot TMA ring, c0 / D in tensor memory
    int w_ring;
    int off_ring, ring_slot_floats;
    int c0_tm;              // c0 (and, in the adjoint, D) live in tensor memory
    int tm_c0_base[2], tm_dl_base[2], tm_dl_end;      // first column per net
    int wblk_off[2 * kMaxLayers], wblk_n[2 * kMaxLayers];   // [W | b] block of layer id (rate 0.., then aug 0..): floats
    short wseq[4 * kMaxLayers];   // layer ids in acquisition order of ONE RHS evaluation of this kernel
    int n_wseq;
    int n_bias_c, off_GB;   // compact bias array [n_bias_c] in shared memory
    int cold_global;        // solver-state rows (off_S, off_RB, off_CB) live in a per-CTA global scratch
    int cold_floats;        // size of that scratch per CTA
    int smem_floats;        // total arena size
};

}  // namespace ude
