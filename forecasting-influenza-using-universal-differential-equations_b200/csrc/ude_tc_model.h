// ude_tc_model.h — host+device description of one UDE func for the tcgen05 (tensor-core) path.
//
// The tensor path integrates tiles of 128 trajectories; a trajectory is a ROW of every matrix
// product (M = 128 trajectories for the layer chain, K = 128 trajectories for the weight
// gradients).  fp32 accuracy comes from a 3-way bf16 split of every operand
//     x = x1 + x2 + x3,  x1 = bf16_rn(x), x2 = bf16_rn(x - x1), x3 = bf16_rn(x - x1 - x2)
// and the six products x_i * w_j with i + j <= 4 (error ~2^-24, measured at the fp32 FFMA level:
// tools/umma_probe.cu), issued as kind::f16 tcgen05.mma with fp32 accumulation in TMEM.
//
// Same reference structure as ude_model.h:
//   Fp / Fa / FaFp of /root/reference/lib/models.py:109-265.
#pragma once
#include <stdint.h>

namespace ude {

// Epilogue warps: warp w works on TMEM lane quarter w & 3 and on the 16-column chunks ch = (w >> 2) + kTcSlots * i of a
// layer; warps 0-3 additionally own the per-trajectory solver state.  The next warp is the MMA issuer; three more warps
// only complete its warpgroup (setmaxnreg hands their registers over).
#ifndef UDE_TC_EPI_WARPS
#define UDE_TC_EPI_WARPS 8
#endif
constexpr int kTcEpiWarps = UDE_TC_EPI_WARPS;
constexpr int kTcSlots = kTcEpiWarps / 4;              // column slots per lane quarter
constexpr int kTcChunksPerWarp = (4 + kTcSlots - 1) / kTcSlots;   // layers are at most 64 = 4 x 16 columns wide
constexpr int kTcThreads = 32 * (kTcEpiWarps + 4);
constexpr int kTcBarThreads = 32 * (kTcEpiWarps + 1);   // operand hand-off barrier: epilogue warps + issuer
constexpr int kTcTile = 128;           // trajectories per tile (= UMMA M)
constexpr int kTcMaxLayers = 4;        // nn.Linear layers per net on the tensor path
constexpr int kTcGroupBytes = kTcTile * 16;   // one 8-feature group of an operand array: [128 traj][8 bf16]

struct TcLayer {
    int in, out;            // true nn.Linear dims
    int in_p, out_p;        // padded to a multiple of 16 (UMMA K / N granularity used here)
    int elu;                // ELU on the output
    int w_off;              // byte offset of the weight image inside the packed blob
    int b_off;              // float offset of the bias (padded to out_p) in the bias block
    int hw_off;             // heads only: float offset (bias block) of the fp32 weights [out][in_p] of the FFMA forward
    int save_col;           // TMEM column of the saved OUTPUT activation (backward), -1 for the head
    int th_w, th_b;         // offsets into theta (PyTorch layout)
    int slab_off;           // float offset of this layer's gradients in a CTA slab: [in_p / 4][64 rows] float4 + [64] bias sums
    int mma_fwd, n_fwd;     // forward instructions of this layer in the MMA table
    int mma_c, n_c;         // delta-chain instructions (hidden layers 1 .. n-2 only: the first layer's W0^T delta0 and the
                            // head's Wh^T g run on the FMA pipe)
    int mma_w, n_w;         // weight-gradient instructions
    int w0_off;             // first layer only: float offset (bias block) of the fp32 weights [out_p][w0_stride]
};

// Weight image of one layer (bf16): [k/8][s = 0..2][n (out_p)][k%8], i.e. for every group of 8
// input features the three split arrays are stacked along N.  Viewed as
//   * K-major B operand of the forward product (N-stacked over s: one instruction computes
//     a_i x [w1|w2|w3]):          SBO = 128, LBO = 3*out_p*16
//   * MN-major B operand of the delta product (N' = in, K' = out) for ONE split array s:
//                                  SBO = 3*out_p*16, LBO = 128
struct TcModel {
    int kind;               // UdeKind
    int R, L;
    int D0, D0_p;           // R*L and its padding to 16
    int n_rate, n_aug;
    TcLayer rate[kTcMaxLayers];
    TcLayer aug[kTcMaxLayers];
    float fa_w;
    int n_theta;
    int w_bytes;            // total bytes of all weight images
    int n_bias;             // floats of the bias block
    int blob_bytes;         // w_bytes + 4*n_bias, rounded up to 16
    int slab_floats;        // floats of one CTA gradient slab
    // ---- shared memory plan (byte offsets into the dynamic arena) ----
    int off_w, off_bias, off_abuf, off_dbuf;
    int off_ghead;          // [G1 | G2] image of the head cotangent (16 wide), read by the head's weight-gradient products
    int w0_stride;          // 8 (D0 <= 8) or 16
    int off_stage;          // backward: staging buffer of the gradient bulk reductions (64 x 64 floats + 64 bias sums)
    int off_abuf2;          // forward image of the aug chain (aliases the sweep buffers / Dbuf region in the backward kernel)
    int n_mma;              // entries of the MMA table (kernel parameter)
    int abuf_bytes, dbuf_bytes;
    int sweep_buf_bytes;    // one of the two alternating sweep images [a2|a1|ones] inside Abuf
    int smem_bytes_fwd, smem_bytes_bwd;
    // ---- tensor memory plan (columns) ----
    int tm_save;            // [0, tm_save): saved activations (backward only)
    int tm_q;               // weight-gradient accumulator: 8 + 2*max_in_p columns
    int tm_q_cols;
    int tm_dd;              // delta-chain accumulator: max_in_p columns
    int tm_fwd;             // backward kernel: forward accumulator of the rate chain (max_out_p columns, aliases [tm_q, ...));
                            // the aug chain's follows at + tm_acc_stride
    int tm_acc_fwd;         // forward kernel: accumulator of the rate chain (the aug chain's at + tm_acc_stride)
    int tm_acc_stride;
    int tm_cols_fwd, tm_cols_bwd;   // allocation sizes (power of two >= 32)
};

}  // namespace ude
