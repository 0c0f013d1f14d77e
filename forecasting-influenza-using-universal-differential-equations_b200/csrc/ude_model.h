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
};

struct ModelP {
    int kind;               // UdeKind
    int R, L;
    int D, D_p;             // R*L and padded to 4  : rows of X (MLP input, row = r*L + l)
    int Dn, Dn_p;           // 3R and padded to 4   : rows of the dynamic state (row = r*3 + c)
    int R_p;                // R padded to 4        : rows of per-region scratch
    int n_rate, n_aug;
    LayerP rate[kMaxLayers];
    LayerP aug[kMaxLayers];
    int n_theta;            // floats of one weight set (PyTorch layout)
    int n_packed;           // floats of the packed blob incl. kPadTail
    float fa_w;
    // ---- shared-memory plan (float offsets into the dynamic smem arena) ----
    int TB, TBP;            // trajectories per tile (32 | 64), padded row stride (TB + 4)
    int tb_shift;           // log2(TB)
    int mt_shift;           // log2(TB / 32): m-tiles per layer are a power of two
    int w_in_smem;          // packed weights resident in smem
    int g_in_smem;          // (bwd) dtheta accumulators resident in smem
    int off_W, off_G;       // valid when the flags above are set
    int off_X;              // [D_p][TBP]
    int off_S;              // state rows: n_state buffers of [Dn_p][TBP]
    int off_GX;             // (bwd) dL/du of one RHS evaluation, [D_p][TBP]
    int off_CB;             // (bwd) running cotangent of the constant inputs, [D_p][TBP]
    int n_state;
    int off_RB, n_rbuf;     // per-region scratch: n_rbuf buffers of [R_p][TBP] (dopri5)
    int cold_global;        // solver-state rows (off_S, off_RB, off_CB) live in a per-CTA global scratch
    int cold_floats;        // size of that scratch per CTA
    int smem_floats;        // total arena size
};

}  // namespace ude
