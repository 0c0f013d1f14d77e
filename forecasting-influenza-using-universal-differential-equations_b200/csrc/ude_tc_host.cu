// ude_tc_host.cu — host-side planning and launch of the tcgen05 path (own translation unit: the
// two kernel families compile in parallel).
#include <stdlib.h>

#include "ude_host_common.h"
#include "ude_tc.cuh"

namespace ude {

static inline int round16(int x) { return (x + 15) & ~15; }

// Batch from which the covered models (R = 1, nets up to 64 wide) run on the tensor path.  Measured on the B200
// (tools/engine_crossover.py, profiles/r02_engine_crossover.txt; fwd + adjoint of the headline model replayed from a CUDA graph):
// ONE 128-trajectory tile takes 0.80 ms for 7 steps on the tcgen05 engine, one 32/64-trajectory tile 1.05 ms on the FFMA
// engine (its layer phases are CTA-wide barriers around register-tiled products), and the gap only widens with the batch
// (B = 8192: 0.85 vs 2.04 ms, B = 18944: 0.88 vs 3.73 ms).  So the tensor path is the default at EVERY batch size; the FFMA
// engine serves what it does not cover (R > 1, wide / deep nets, Dopri5, Bayes_* funcs, weight-set groups).
static long long tc_min_batch() { return 1; }
static const int kTcSmemLimit = 227 * 1024 - 2048;      // per-CTA opt-in limit minus the static part (~1.2 KB) and alignment slack

static bool tc_fill_layers(TcModel& M, TcLayer* Ls, int n, const int32_t* sizes, int head_elu, int& th) {
    if (n == 0) return true;
    if (n < 3 || n > kTcMaxLayers) return false;      // at least two hidden layers + head (every net of the reference)
    int n_in = M.D0, in_p = M.D0_p;
    for (int i = 0; i < n; ++i) {
        TcLayer& Lp = Ls[i];
        memset(&Lp, 0, sizeof(Lp));
        Lp.in = n_in;
        Lp.out = sizes[i];
        Lp.in_p = in_p;
        Lp.out_p = round16(Lp.out);
        if (i == n - 1 && Lp.out_p != 16) return false;          // heads: 2R / 3R <= 16
        if (Lp.out_p > 64 || Lp.in_p > 64) return false;         // accumulators: 3*out_p columns, M = 64 rows of dW
        Lp.elu = (i <= n - 3 + (head_elu ? 1 : 0)) ? 1 : 0;       // lib/models.py: no ELU before the head
        Lp.th_w = th;
        th += Lp.in * Lp.out;
        Lp.th_b = th;
        th += Lp.out;
        Lp.w_off = M.w_bytes;
        M.w_bytes += Lp.in_p * Lp.out_p * 6;
        Lp.b_off = M.n_bias;
        M.n_bias += Lp.out_p;
        Lp.hw_off = -1;
        if (i == n - 1) {               // head: forward runs on the FMA pipe from fp32 weights (ude_tc.cuh: tc_head_partial)
            Lp.hw_off = M.n_bias;
            M.n_bias += Lp.out * Lp.in_p;
        }
        Lp.save_col = -1;
        if (i < n - 1) {
            Lp.save_col = M.tm_save;
            M.tm_save += Lp.out_p;
        }
        Lp.slab_off = M.slab_floats;
        M.slab_floats += 64 * Lp.in_p + 64;
        Lp.mma_fwd = M.n_mma;
        Lp.n_fwd = (i == n - 1) ? 0 : 6 * (Lp.in_p / 16);
        M.n_mma += Lp.n_fwd;
        Lp.mma_c = M.n_mma;
        Lp.n_c = (i == 0 || i == n - 1) ? 0 : 6 * (Lp.out_p / 16);
        M.n_mma += Lp.n_c;
        Lp.mma_w = M.n_mma;
        Lp.n_w = 2 * (kTcTile / 16);
        M.n_mma += Lp.n_w;
        Lp.w0_off = -1;
        if (i == 0) {                   // first layer: W0^T delta0 runs on the FMA pipe from fp32 weights
            Lp.w0_off = M.n_bias;
            M.n_bias += Lp.out_p * M.w0_stride;
        }
        n_in = Lp.out;
        in_p = Lp.out_p;
    }
    return true;
}

// descriptor -> TcModel; false when the tensor path does not cover the configuration
static bool tc_build_model(const UdeDesc* d, TcModel& M) {
    memset(&M, 0, sizeof(M));
    M.kind = d->kind;
    M.R = d->n_regions;
    M.L = d->latent_dim;
    M.D0 = M.R * M.L;
    M.D0_p = 16;
    M.w0_stride = M.D0 <= 8 ? 8 : 16;
    if (!(M.R == 1 && (M.L == 4 || M.L == 8 || M.L == 16))) return false;      // instantiated (R, L) pairs
    M.n_rate = d->n_rate_layers;
    M.n_aug = d->n_aug_layers;
    M.fa_w = d->fa_w;
    int th = 0;
    if (!tc_fill_layers(M, M.rate, M.n_rate, d->rate_sizes, d->flags & UDE_FLAG_RATE_HEAD_ELU, th)) return false;
    if (!tc_fill_layers(M, M.aug, M.n_aug, d->aug_sizes, d->flags & UDE_FLAG_AUG_HEAD_ELU, th)) return false;
    M.n_theta = th;
    M.blob_bytes = (int)align_up(M.w_bytes + 4 * M.n_bias, 16);
    int max_in = 16, max_out = 16, dcap = 0;
    for (int net = 0; net < 2; ++net) {
        const TcLayer* Ls = net == 0 ? M.rate : M.aug;
        const int n = net == 0 ? M.n_rate : M.n_aug;
        for (int i = 0; i < n; ++i) {
            if (Ls[i].in_p > max_in) max_in = Ls[i].in_p;
            if (Ls[i].out_p > max_out) max_out = Ls[i].out_p;
            const int g = Ls[i].out_p / 8;
            const int need = 2 * g + (g > 8 ? g : 8);        // the M = 64 view of the third split array reads 8 groups
            if (need > dcap) dcap = need;
        }
    }
    // tensor memory (columns)
    M.tm_q = M.tm_save;
    M.tm_q_cols = 2 * max_in + 16;
    M.tm_dd = M.tm_q + M.tm_q_cols;
    M.tm_fwd = M.tm_q;
    M.tm_acc_fwd = 0;
    M.tm_acc_stride = max_out > 16 * kTcSlots ? max_out : 16 * kTcSlots;      // >= one 16-column block per column slot (head partial sums)
    const int bwd_cols = (M.tm_dd + max_in) > (M.tm_fwd + 2 * max_out) ? (M.tm_dd + max_in) : (M.tm_fwd + 2 * max_out);
    if (bwd_cols > 512 || 2 * max_out > 128) return false;
    M.tm_cols_fwd = 128;
    M.tm_cols_bwd = 512;
    // shared memory
    M.off_w = 0;
    M.off_bias = M.w_bytes;
    // [per-warp staging slices of the gradient bulk reductions] [head-cotangent image [G1|G2]]
    M.off_stage = (int)align_up(M.blob_bytes, 128);
    M.off_ghead = (int)align_up(M.off_stage + kTcEpiWarps * kTcStageSlice, 128);
    M.off_abuf = M.off_ghead + 2 * kTcGroupBytes;       // the M = 64 view of the image over-reads into Abuf (finite data)
    M.sweep_buf_bytes = (2 * (max_in / 8) + 1) * kTcGroupBytes;
    M.abuf_bytes = (3 * (max_in / 8) + 1) * kTcGroupBytes;               // forward image [a2|a1|ones|a3]
    if (2 * M.sweep_buf_bytes > M.abuf_bytes) M.abuf_bytes = 2 * M.sweep_buf_bytes;
    M.off_dbuf = M.off_abuf + M.abuf_bytes;
    M.dbuf_bytes = dcap * kTcGroupBytes;
    const int fwd_image_bytes = (3 * (max_in / 8) + 1) * kTcGroupBytes;
    M.off_abuf2 = M.off_abuf + fwd_image_bytes;
    M.smem_bytes_fwd = M.off_abuf2 + fwd_image_bytes;
    M.smem_bytes_bwd = M.off_dbuf + M.dbuf_bytes;
    if (M.smem_bytes_bwd < M.off_abuf2 + fwd_image_bytes) M.smem_bytes_bwd = M.off_abuf2 + fwd_image_bytes;
    if (M.n_mma > kTcMaxMma) return false;
    if (M.smem_bytes_bwd > kTcSmemLimit) return false;
    return true;
}

// 0 = FFMA engine, 1 = tensor path, negative = error
int tc_choose(const UdeDesc* d, TcModel& M) {
    const bool ok = tc_build_model(d, M);
    if (d->flags & UDE_FLAG_FORCE_TENSOR) return ok ? 1 : UDE_ERR_UNSUPPORTED;
    if (d->flags & UDE_FLAG_FORCE_FFMA) return 0;
    return (ok && d->batch >= tc_min_batch()) ? 1 : 0;
}

// MMA instruction lists of every layer (shared-memory addresses relative to the dynamic smem base;
// forward lists relative to the forward accumulator, backward lists to the TMEM allocation base)
static void tc_build_table(const TcModel& M, TcTable& T, bool backward) {
    memset(&T, 0, sizeof(T));
    const int nl = M.n_rate + M.n_aug;
    for (int li = 0; li < nl; ++li) {
        const TcLayer& Lp = li < M.n_rate ? M.rate[li] : M.aug[li - M.n_rate];
        TcRecorder rf{T.e + Lp.mma_fwd, 0};
        tc_issue_forward(Lp, (uint32_t)M.off_abuf, (uint32_t)M.off_w, 0u, rf);
        if (backward) {
            const bool head = (li == M.n_rate - 1) || (li == nl - 1);
            if (Lp.n_c > 0) {
                TcRecorder rc{T.e + Lp.mma_c, 0};
                tc_issue_delta(Lp, (uint32_t)M.off_dbuf, (uint32_t)M.off_w, (uint32_t)M.tm_dd, rc);
            }
            TcRecorder rw{T.e + Lp.mma_w, 0};
            tc_issue_dw(Lp, (uint32_t)(head ? M.off_ghead : M.off_dbuf), (uint32_t)M.off_abuf, (uint32_t)M.tm_q,
                        head ? 1 : Lp.out_p / 8, rw);
        }
    }
}

struct TcWorkspace {
    uint8_t* blob;
    double* side_partial;
    float* slabs;
    long long bytes;
};
// One weight set: a blob, side partials and gradient slabs per CTA.  G weight sets in one launch (UdeDesc::n_groups): G blobs,
// side partials per tile and trajectory warp, gradient slabs per TILE (summed per group in tile order afterwards).
static long long tc_blob_stride(const TcModel& M) { return align_up(M.blob_bytes, 256); }
static long long tc_tiles(long long B, int G) { return (B / G + kTcTile - 1) / kTcTile * G; }
static TcWorkspace tc_carve(void* base, const TcModel& M, int n_ctas, long long B = 0, int G = 1) {
    TcWorkspace w;
    long long off = 0;
    const long long tiles = G > 1 ? tc_tiles(B, G) : 0;
    w.blob = reinterpret_cast<uint8_t*>(base) + off;
    off = align_up(off + tc_blob_stride(M) * G, 256);
    w.side_partial = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (G > 1 ? tiles * 4 : (long long)n_ctas) * 8 * 8, 256);
    w.slabs = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (G > 1 ? tiles : (long long)n_ctas) * M.slab_floats * 4, 256);
    w.bytes = off;
    return w;
}
static int tc_grid(long long B, int G = 1) {
    const DeviceInfo& di = device_info();
    const long long tiles = tc_tiles(B, G);
    const int n_sm = di.ok ? di.n_sm : 148;
    return (int)(tiles < n_sm ? tiles : n_sm);
}

int tc_forward(const UdeDesc* desc, const TcModel& M, const float* theta, const float* y0, const GridTables& g,
                      float* y_out, float* tape, double* side, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    const int G = desc->n_groups;
    const int n_ctas = tc_grid(desc->batch, G);
    TcWorkspace w = tc_carve(workspace, M, n_ctas, desc->batch, G);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    tc_pack_kernel<<<dim3(M.n_rate + M.n_aug, G), 256, 0, s>>>(M, theta, w.blob, tc_blob_stride(M));
    UDE_CUDA(cudaGetLastError());
    static thread_local TcTable tab_storage;
    TcTable* tab = &tab_storage;
    tc_build_table(M, *tab, false);
    TcFwdArgs a{};
    a.blob = w.blob;
    a.y0 = y0;
    a.g = g;
    a.y_out = y_out;
    a.tape = tape;
    a.side_partial = w.side_partial;
    a.B = desc->batch / G;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    a.n_groups = G;
    a.blob_stride = tc_blob_stride(M);
    a.tape_group_floats = align_up(((long long)g.n_steps * 4 + 1) * 3 * M.R * (desc->batch / G) * 4, 256) / 4;
    if (G > 1 && (a.rp.mean || a.dec.n_dec > 0)) return UDE_ERR_UNSUPPORTED;
    int st;
#define UDE_TC_FWD(RR, LL)                                                                     \
    do {                                                                                       \
        if ((st = set_smem(rk4_fwd_tc_kernel<RR, LL>, M.smem_bytes_fwd)) != UDE_OK) return st; \
        rk4_fwd_tc_kernel<RR, LL><<<n_ctas, kTcThreads, M.smem_bytes_fwd, s>>>(M, *tab, a);          \
    } while (0)
    if (M.L == 8) UDE_TC_FWD(1, 8);
    else if (M.L == 4) UDE_TC_FWD(1, 4);
    else UDE_TC_FWD(1, 16);
#undef UDE_TC_FWD
    UDE_CUDA(cudaGetLastError());
    tc_reduce_side_kernel<<<G, 32, 0, s>>>(w.side_partial, G > 1 ? (int)(tc_tiles(desc->batch, G) / G) * 4 : n_ctas, side);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int tc_backward(const UdeDesc* desc, const TcModel& M, const float* theta, const float* y0, const GridTables& g,
                       const float* tape, const float* g_out, const float* g_side, float* g_y0, float* g_theta,
                       void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    const int G = desc->n_groups;
    const int n_ctas = tc_grid(desc->batch, G);
    TcWorkspace w = tc_carve(workspace, M, n_ctas, desc->batch, G);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    if (!(desc->flags & UDE_FLAG_WEIGHTS_PACKED)) {
        tc_pack_kernel<<<dim3(M.n_rate + M.n_aug, G), 256, 0, s>>>(M, theta, w.blob, tc_blob_stride(M));
        UDE_CUDA(cudaGetLastError());
    }
    const long long n_slabs = G > 1 ? tc_tiles(desc->batch, G) : n_ctas;
    UDE_CUDA(cudaMemsetAsync(w.slabs, 0, (size_t)n_slabs * M.slab_floats * 4, s));
    static thread_local TcTable tab_storage;
    TcTable* tab = &tab_storage;
    tc_build_table(M, *tab, true);
    TcBwdArgs a{};
    a.blob = w.blob;
    a.y0 = y0;
    a.g = g;
    a.tape = tape;
    a.g_out = g_out;
    a.g_side = g_side;
    a.g_y0 = g_y0;
    a.slabs = w.slabs;
    a.B = desc->batch / G;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    decoded_ctas() = n_ctas;
    a.n_groups = G;
    a.blob_stride = tc_blob_stride(M);
    a.tape_group_floats = align_up(((long long)g.n_steps * 4 + 1) * 3 * M.R * (desc->batch / G) * 4, 256) / 4;
    if (G > 1 && (a.rp.mean || a.dec.n_dec > 0)) return UDE_ERR_UNSUPPORTED;
    int st;
#define UDE_TC_BWD(RR, LL)                                                                     \
    do {                                                                                       \
        if ((st = set_smem(rk4_bwd_tc_kernel<RR, LL>, M.smem_bytes_bwd)) != UDE_OK) return st; \
        rk4_bwd_tc_kernel<RR, LL><<<n_ctas, kTcThreads, M.smem_bytes_bwd, s>>>(M, *tab, a);          \
    } while (0)
    if (M.L == 8) UDE_TC_BWD(1, 8);
    else if (M.L == 4) UDE_TC_BWD(1, 4);
    else UDE_TC_BWD(1, 16);
#undef UDE_TC_BWD
    UDE_CUDA(cudaGetLastError());
    tc_reduce_dtheta_kernel<<<dim3((M.n_theta + 255) / 256, G), 256, 0, s>>>(M, w.slabs, (int)(n_slabs / G), g_theta);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

// ---- Bayes funcs on the tensor path: one blob per RHS evaluation, gradients folded into two slab sets -------------------------
struct TcBayesWs {
    uint8_t* blobs;         // [n_evals][blob_stride]
    double* side_partial;   // [n_ctas][8]
    float* slabs;           // [2][n_ctas][slab_floats]: sum_e G_e, then sum_e G_e * z_e
    float* z_slab;          // [n_evals][slab_floats]
    long long bytes;
};
static TcBayesWs tc_bayes_carve(void* base, const TcModel& M, int n_ctas, int n_evals) {
    TcBayesWs w;
    long long off = 0;
    w.blobs = reinterpret_cast<uint8_t*>(base) + off;
    off = align_up(off + tc_blob_stride(M) * n_evals, 256);
    w.side_partial = reinterpret_cast<double*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)n_ctas * 8 * 8, 256);
    w.slabs = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)2 * n_ctas * M.slab_floats * 4, 256);
    w.z_slab = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off);
    off = align_up(off + (long long)n_evals * M.slab_floats * 4, 256);
    w.bytes = off;
    return w;
}
long long tc_bayes_workspace(const UdeDesc* d, const TcModel& M, int n_steps, int* n_ctas) {
    *n_ctas = tc_grid(d->batch);
    return tc_bayes_carve(nullptr, M, *n_ctas, 4 * n_steps).bytes;
}

int tc_bayes_forward(const UdeDesc* desc, const TcModel& M, const float* theta_evals, const float* y0, const GridTables& g,
                     float* y_out, float* tape, double* side, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    const int n_ctas = tc_grid(desc->batch), n_evals = 4 * g.n_steps;
    TcBayesWs w = tc_bayes_carve(workspace, M, n_ctas, n_evals);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    tc_pack_kernel<<<dim3(M.n_rate + M.n_aug, n_evals), 256, 0, s>>>(M, theta_evals, w.blobs, tc_blob_stride(M));
    UDE_CUDA(cudaGetLastError());
    static thread_local TcTable tab_storage;
    TcTable* tab = &tab_storage;
    tc_build_table(M, *tab, false);
    TcFwdArgs a{};
    a.blob = w.blobs;
    a.y0 = y0;
    a.g = g;
    a.y_out = y_out;
    a.tape = tape;
    a.side_partial = w.side_partial;
    a.B = desc->batch;
    a.n_groups = 1;
    a.blob_stride = tc_blob_stride(M);
    a.bayes = 1;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    int st;
#define UDE_TC_FWD(RR, LL)                                                                     \
    do {                                                                                       \
        if ((st = set_smem(rk4_fwd_tc_kernel<RR, LL, true>, M.smem_bytes_fwd)) != UDE_OK) return st; \
        rk4_fwd_tc_kernel<RR, LL, true><<<n_ctas, kTcThreads, M.smem_bytes_fwd, s>>>(M, *tab, a);          \
    } while (0)
    if (M.L == 8) UDE_TC_FWD(1, 8);
    else if (M.L == 4) UDE_TC_FWD(1, 4);
    else UDE_TC_FWD(1, 16);
#undef UDE_TC_FWD
    UDE_CUDA(cudaGetLastError());
    tc_reduce_side_kernel<<<1, 32, 0, s>>>(w.side_partial, n_ctas, side);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int tc_bayes_backward(const UdeDesc* desc, const TcModel& M, const float* theta_evals, const float* z_evals, const float* y0,
                      const GridTables& g, const float* tape, const float* g_out, const float* g_side, float* g_y0,
                      float* g_mean_zsum, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    const int n_ctas = tc_grid(desc->batch), n_evals = 4 * g.n_steps;
    TcBayesWs w = tc_bayes_carve(workspace, M, n_ctas, n_evals);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    if (!(desc->flags & UDE_FLAG_WEIGHTS_PACKED)) {
        tc_pack_kernel<<<dim3(M.n_rate + M.n_aug, n_evals), 256, 0, s>>>(M, theta_evals, w.blobs, tc_blob_stride(M));
        UDE_CUDA(cudaGetLastError());
    }
    UDE_CUDA(cudaMemsetAsync(w.z_slab, 0, (size_t)n_evals * M.slab_floats * 4, s));
    tc_pack_z_kernel<<<dim3((M.n_theta + 255) / 256, n_evals), 256, 0, s>>>(M, z_evals, w.z_slab);
    UDE_CUDA(cudaGetLastError());
    UDE_CUDA(cudaMemsetAsync(w.slabs, 0, (size_t)2 * n_ctas * M.slab_floats * 4, s));
    static thread_local TcTable tab_storage;
    TcTable* tab = &tab_storage;
    tc_build_table(M, *tab, true);
    TcBwdArgs a{};
    a.blob = w.blobs;
    a.y0 = y0;
    a.g = g;
    a.tape = tape;
    a.g_out = g_out;
    a.g_side = g_side;
    a.g_y0 = g_y0;
    a.slabs = w.slabs;
    a.B = desc->batch;
    a.n_groups = 1;
    a.blob_stride = tc_blob_stride(M);
    a.bayes_z = w.z_slab;
    a.bz_slabs = w.slabs + (size_t)n_ctas * M.slab_floats;
    a.rp = pending_reparam();
    a.dec = pending_decoded();
    decoded_ctas() = n_ctas;
    int st;
#define UDE_TC_BWD(RR, LL)                                                                     \
    do {                                                                                       \
        if ((st = set_smem(rk4_bwd_tc_kernel<RR, LL, true>, M.smem_bytes_bwd)) != UDE_OK) return st; \
        rk4_bwd_tc_kernel<RR, LL, true><<<n_ctas, kTcThreads, M.smem_bytes_bwd, s>>>(M, *tab, a);          \
    } while (0)
    if (M.L == 8) UDE_TC_BWD(1, 8);
    else if (M.L == 4) UDE_TC_BWD(1, 4);
    else UDE_TC_BWD(1, 16);
#undef UDE_TC_BWD
    UDE_CUDA(cudaGetLastError());
    tc_reduce_dtheta_kernel<<<dim3((M.n_theta + 255) / 256, 2), 256, 0, s>>>(M, w.slabs, n_ctas, g_mean_zsum);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

void tc_plan_info(const UdeDesc* d, const TcModel& M, long long* workspace_bytes, int* n_ctas) {
    *n_ctas = tc_grid(d->batch, d->n_groups);
    *workspace_bytes = tc_carve(nullptr, M, *n_ctas, d->batch, d->n_groups).bytes;
}

}  // namespace ude

#ifdef UDE_TC_PROFILE
/* debug builds only: cycle counters of CTA 0 (thread 0: [0..8), thread 128: [8..16)) of the LAST tensor kernel */
extern "C" int ude_debug_tc_prof(unsigned long long* out) {
    UDE_CUDA(cudaDeviceSynchronize());
    UDE_CUDA(cudaMemcpyFromSymbol(out, ude::g_tc_prof, sizeof(unsigned long long) * 16));
    return UDE_OK;
}
#endif
