/*
 * ude_b200.h — C ABI of the B200-native fused UDE solve.
 *
 * This is the drop-in boundary for ONE hot path of
 * M-Morris-95/Forecasting-Influenza-Using-Universal-Differential-Equations:
 *
 *     latent = odeint(self.ode, z, t, method='rk4', options=dict(step_size=h))
 *                                              /root/reference/lib/VAE.py:137
 *     loss.backward()                          /root/reference/lib/VAE.py:203
 *
 * i.e. the batched integration of the Fp / Fa / FaFp ODE funcs
 * (/root/reference/lib/models.py:109-265) by torchdiffeq's fixed-grid RK4
 * (3/8 rule) or adaptive Dopri5, and the discrete adjoint of that integration.
 * The reference reaches this path through Python (torchdiffeq.odeint +
 * torch.autograd); a binding is a ctypes/cffi stub over the entry points below
 * (see INTEGRATION.md).  The Python package in this repo is exactly such a stub.
 *
 * Conventions
 *   - plain C, no torch types; every pointer marked "dev" is a device pointer
 *     owned by the caller, everything else is host memory;
 *   - nothing persistent is allocated; scratch comes in through `workspace`
 *     (size from ude_plan()); all work is enqueued on `stream` (a cudaStream_t
 *     passed as void*) and is asynchronous;
 *   - every function returns 0 (UDE_OK) or a negative UdeStatus, never throws;
 *   - arithmetic is fp32 (side-channel sums are fp64), as the reference
 *     (/root/reference/run_ode.py:30).
 *
 * Tensor layouts (row-major, innermost last)
 *   y0, g_y0      [B, R, L]       the reference's flattened batch axis
 *                                 B = n_samples * batch (lib/models.py:23)
 *   y_out, g_out  [T, B, R, L]    what odeint returns (lib/VAE.py:137-138)
 *   theta, g_theta[G, n_theta]    per weight set: for every nn.Linear in
 *                                 ModuleList order (rate net, then aug net)
 *                                 W[out,in] row-major, then b[out]
 *                                 (lib/models.py:118-124, 167-172, 208-223)
 *   tape          opaque, ude_plan().tape_bytes; written by *_forward,
 *                 read by *_backward (stage inputs, SoA; rk4: one more record
 *                 for the final state, used by the decoded entry points)
 *   side          double[8]: sum beta, sum gamma, sum beta^2, sum gamma^2,
 *                 sum Fa^2, N rows, n RHS evals, sum of latent_init_loss
 *                 (decoded entry points, else 0) — the reductions
 *                 behind posterior() (lib/models.py:152-156) and the Fa_norm
 *                 loss (lib/VAE.py:179-183); per group: [G, 8]
 *   g_side        float[5] per group: dL/d(side[0..4])
 */
#ifndef UDE_B200_H
#define UDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDE_ABI_VERSION 1
#define UDE_MAX_LAYERS 8          /* nn.Linear layers per net, head included */
#define UDE_SIDE_LEN 8

typedef enum UdeStatus {
    UDE_OK = 0,
    UDE_ERR_INVALID = -1,         /* bad descriptor / null pointer / bad size   */
    UDE_ERR_UNSUPPORTED = -2,     /* valid request the kernels do not cover     */
    UDE_ERR_WORKSPACE = -3,       /* workspace / tape smaller than ude_plan()   */
    UDE_ERR_CUDA = -4,            /* CUDA runtime error (see ude_last_cuda_error) */
    UDE_ERR_NO_DEVICE = -5,       /* no sm_100 device visible                   */
    UDE_ERR_GRID = -6,            /* t not strictly increasing / bad step_size  */
    UDE_ERR_CAPACITY = -7         /* caller-provided table too small            */
} UdeStatus;

typedef enum UdeKind { UDE_KIND_FP = 0, UDE_KIND_FA = 1, UDE_KIND_FAFP = 2 } UdeKind;

/* Static description of the ODE func (constructor kwargs of
 * lib/models.py:110,159,200) and of the batch. */
typedef struct UdeDesc {
    int32_t kind;                         /* UdeKind                                         */
    int32_t n_regions;                    /* R                                               */
    int32_t latent_dim;                   /* L >= 3                                          */
    int32_t n_rate_layers;                /* Linear layers of the (beta,gamma) net, 0 for Fa */
    int32_t rate_sizes[UDE_MAX_LAYERS];   /* out width of each; last must be 2R              */
    int32_t n_aug_layers;                 /* Linear layers of the Fa net, 0 for Fp           */
    int32_t aug_sizes[UDE_MAX_LAYERS];    /* out width of each; last must be 3R              */
    float   fa_w;                         /* FaFp.Fa_w (lib/models.py:225,249); runtime      */
    int32_t n_groups;                     /* G independent weight sets (tuning sweeps); >= 1.
                                           * G > 1 (rk4 only): every array is GROUP-MAJOR --
                                           * theta [G][n_theta], y0 [G][B/G][R][L],
                                           * y_out / g_out [G][n_out][B/G][R][L], side [G][8],
                                           * g_side [G][5], g_y0 like y0, g_theta like theta,
                                           * tape / workspace [G][per-group bytes] -- and the call
                                           * equals G single-set calls (ude_plan returns the totals) */
    int32_t flags;                        /* UDE_FLAG_*                                      */
    int64_t batch;                        /* B, all groups together; B % G == 0              */
} UdeDesc;

/* lib/models.py puts NO activation between the last hidden Linear and the head
 * (Linear,[ELU,Linear]*,Linear, :118-124); the local Fp of tuning/tune_node.py:63-76 is an
 * nn.Sequential with an ELU before the head as well.  These flags select the latter. */
#define UDE_FLAG_RATE_HEAD_ELU 1
#define UDE_FLAG_AUG_HEAD_ELU 2
/* Kernel family of the RK4 solve.  By default the library picks the tcgen05 tensor-core kernels when
 * the configuration is one they cover AND the batch makes the layer products real dense contractions
 * (BASELINE north_star), else the FFMA kernels.  These flags pin the choice (parity tests, benchmarks);
 * FORCE_TENSOR on an uncovered configuration returns UDE_ERR_UNSUPPORTED. */
#define UDE_FLAG_FORCE_FFMA 4
#define UDE_FLAG_FORCE_TENSOR 8
/* ude_rk4_backward only: `workspace` is the buffer the matching ude_rk4_forward call (same desc, same theta) used and
 * nothing has written to it since, so it still holds the packed weight images -- the adjoint does not build them again.
 * (One launch less per training step; what loss.backward() right after odeint() looks like, lib/VAE.py:201-203.) */
#define UDE_FLAG_WEIGHTS_PACKED 16
/* ude_plan only: size the workspace for the ude_rk4_*_bayes entry points (set by them internally for the launch). */
#define UDE_FLAG_BAYES 32

typedef enum UdeMethod { UDE_METHOD_RK4 = 0, UDE_METHOD_DOPRI5 = 1 } UdeMethod;

/* Sizes the caller must provide (bytes). */
typedef struct UdePlan {
    int64_t n_theta;            /* floats per weight set                          */
    int64_t workspace_bytes;    /* scratch for either direction                   */
    int64_t tape_bytes;         /* for n_steps steps (rk4) / max_steps (dopri5)   */
    int32_t tile_traj;          /* trajectories per CTA tile                      */
    int32_t n_ctas;             /* persistent grid size                           */
    int32_t smem_bytes_fwd;     /* dynamic shared memory per CTA                  */
    int32_t smem_bytes_bwd;
    int32_t weights_in_smem;    /* 1: staged once per CTA with a TMA bulk copy    */
    int32_t path;               /* kernel family the call will run: 0 FFMA engine, 1 tcgen05 engine (ude_rk4_path),
                                 * 2 FFMA engine, all weight-set groups in one launch, 3 tcgen05 engine, all groups in one launch.  The two engines round
                                 * differently (both inside the 1e-5 bar): pin UDE_FLAG_FORCE_* for bit-reproducibility
                                 * across batch sizes / shard counts.                                            */
} UdePlan;

int         ude_abi_version(void);
const char* ude_strerror(int status);
const char* ude_last_cuda_error(void);

/* Number of floats in one weight set for `desc` (host only, no GPU needed). */
int64_t ude_theta_size(const UdeDesc* desc);

/* Algorithmic multiply-accumulates of one RHS evaluation of one trajectory. */
int64_t ude_macs_per_eval(const UdeDesc* desc);

/* Sizes for a solve of `n_steps` grid steps (rk4) or at most `n_steps`
 * accepted steps per trajectory (dopri5).  Host arithmetic only (assumes 148 SMs
 * when no GPU is visible). */
int ude_plan(const UdeDesc* desc, int method, int n_steps, UdePlan* plan);

/*
 * Fixed-grid construction, host only (no GPU needed).  Restates
 * torchdiffeq FixedGridODESolver._grid_constructor_from_step_size + the output
 * loop of FixedGridODESolver.integrate in fp32, bit for bit:
 *   niters = ceil((t[-1]-t[0])/step_size + 1); grid = arange(niters)*step_size + t[0];
 *   grid[-1] = t[-1];  outputs emitted while t1 >= t[j].
 * use_step_size == 0 -> the grid is t itself.
 *   dt[n]        = grid[n+1] - grid[n]                      n < *n_steps
 *   out_step[j]  = grid step whose end first reaches t[j]   (-1 for j == 0)
 *   out_kind[j]  = 0: copy y_n (t[j]==t0)  1: copy y_{n+1} (t[j]==t1)  2: lerp
 *   out_w[j]     = (t[j]-t0)/(t1-t0)  for kind 2
 * `cap_steps` is the capacity of dt[]; the per-output tables hold n_t entries.
 */
int ude_rk4_grid(const float* t, int32_t n_t, float step_size, int32_t use_step_size,
                 int32_t cap_steps, float* dt, int32_t* n_steps,
                 int32_t* out_step, int32_t* out_kind, float* out_w);

/*
 * Fixed-step RK4 (torchdiffeq 'rk4' = 3/8 rule) forward over the whole grid.
 * Replaces  odeint(func, y0, t, method='rk4', options={'step_size': h})
 * (/root/reference/lib/VAE.py:137).  `tape` may be NULL (no backward wanted).
 */
int ude_rk4_forward(const UdeDesc* desc,
                    const float* theta /*dev*/, const float* y0 /*dev*/,
                    int32_t n_steps, const float* dt /*dev [n_steps]*/,
                    int32_t n_out, const int32_t* out_step /*dev*/,
                    const int32_t* out_kind /*dev*/, const float* out_w /*dev*/,
                    float* y_out /*dev [n_out,B,R,L]*/, float* tape /*dev | NULL*/,
                    double* side /*dev [G,8]*/,
                    void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/*
 * Discrete adjoint of ude_rk4_forward = what torch.autograd computes for
 * loss.backward() through the solver (/root/reference/lib/VAE.py:203).
 */
int ude_rk4_backward(const UdeDesc* desc,
                     const float* theta /*dev*/, const float* y0 /*dev*/,
                     int32_t n_steps, const float* dt /*dev*/,
                     int32_t n_out, const int32_t* out_step /*dev*/,
                     const int32_t* out_kind /*dev*/, const float* out_w /*dev*/,
                     const float* tape /*dev*/,
                     const float* g_out /*dev [n_out,B,R,L]*/, const float* g_side /*dev [G,5] | NULL*/,
                     float* g_y0 /*dev [B,R,L]*/, float* g_theta /*dev [G,n_theta]*/,
                     void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/*
 * Fused producer of y0: the caller of the path builds the initial state element-wise from the
 * encoder's (mean, std) and a normal draw,
 *     z  = eps * std + mean;  y0 = [|z0|, |z1|, 1 - (|z0| + |z1|), z2, ...] + 1e-5
 * (/root/reference/lib/models.py:16-24 reparam, /root/reference/lib/VAE.py:120,130).  The *_reparam entry points do that
 * in the solve's prologue (the rows land in y_out[0], which IS y0) and, in the adjoint's epilogue, turn dL/dy0 into
 * dL/dz -- the sample-axis sums dL/dmean = sum_s dL/dz, dL/dstd = sum_s dL/dz * eps stay with the caller.
 * Trajectory index = s * batch + b, as reparam's reshape produces it (lib/models.py:23).  rk4, one weight set.
 */
typedef struct UdeReparam {
    const float* eps;           /* dev [n_samples, batch, R, L-1]; NULL together with std: z = mean (uncertainty=False) */
    const float* mean;          /* dev [batch, R, L-1]                                                                  */
    const float* std;           /* dev [batch, R, L-1]                                                                  */
    int32_t n_samples, batch;   /* n_samples * batch == desc->batch                                                     */
} UdeReparam;

int ude_rk4_forward_reparam(const UdeDesc* desc, const float* theta /*dev*/, const UdeReparam* rp,
                            int32_t n_steps, const float* dt /*dev*/, int32_t n_out, const int32_t* out_step /*dev*/,
                            const int32_t* out_kind /*dev*/, const float* out_w /*dev*/,
                            float* y_out /*dev [n_out,B,R,L]; y_out[0] = y0*/, float* tape /*dev | NULL*/, double* side /*dev*/,
                            void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/* y0 = the forward call's y_out[0]; g_y0 [B,R,L] is written as well (dL/dy0); g_z dev [B,R,L-1] */
int ude_rk4_backward_reparam(const UdeDesc* desc, const float* theta /*dev*/, const UdeReparam* rp, const float* y0 /*dev*/,
                             int32_t n_steps, const float* dt /*dev*/, int32_t n_out, const int32_t* out_step /*dev*/,
                             const int32_t* out_kind /*dev*/, const float* out_w /*dev*/, const float* tape /*dev*/,
                             const float* g_out /*dev*/, const float* g_side /*dev | NULL*/, float* g_y0 /*dev*/,
                             float* g_z /*dev*/, float* g_theta /*dev*/,
                             void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/*
 * Fused consumer of the outputs.  The caller of the path reduces `latent` right away,
 *     y_pred = self.dec(self.latent[..., :3])                        /root/reference/lib/VAE.py:138
 *              Decoder: nn.Linear(n_regions * 3, n_regions * input_dim) on the flattened [.., R, 3] rows
 *                                                                    /root/reference/lib/models.py:26-51
 *     reg_loss = 0.1 * latent_init_loss(self.latent[..., :3])        /root/reference/lib/VAE.py:186,
 *                                                                    /root/reference/lib/train_functions.py:116-126
 * so the *_decoded entry points emit those instead of the [n_out,B,R,L] trajectories (8x fewer bytes at L = 8, and no
 * eager loss passes over them):
 *     y_pred   dev [n_out, B, dec.n_out]    decoder output of every output time (row 0 decodes y0)
 *     side[7]  sum of latent_init_loss over all outputs / trajectories / regions / 3 components (fp64)
 *     out_mask dev [n_out, B, R] bytes      bit 2c: component c < 0, bit 2c+1: component c > 1 (d reg / d latent for the adjoint)
 * and the adjoint takes dL/dy_pred plus dL/dside[7] (= g_side[5]; g_side holds SIX floats here) and returns, next to
 * dL/dy0 (dL/dz) and dL/dtheta, the decoder's gradients g_dec = [dW (dec.n_out x 3R, row-major) | db (dec.n_out)].
 * The output states needed for dW are stage inputs already on the tape plus its last record (the final state).
 * `rp` may be NULL (y0 given) or the fused reparam producer (then y0_out [B,R,L] receives the initial state and is the
 * `y0` argument of the adjoint).  rk4, one weight set.  dec_partial: scratch, [ude_plan().n_ctas, dec.n_out*(3R+1)] floats.
 */
typedef struct UdeDecoder {
    const float* weight;    /* dev [n_out, 3R]  Decoder.decoder[1].weight  (lib/models.py:40) */
    const float* bias;      /* dev [n_out]                                                   */
    int32_t n_out;          /* R * input_dim                                                 */
} UdeDecoder;

int ude_rk4_forward_decoded(const UdeDesc* desc, const float* theta /*dev*/, const float* y0 /*dev | NULL with rp*/,
                            const UdeReparam* rp /*NULL: y0 given*/, int32_t n_steps, const float* dt /*dev*/, int32_t n_out,
                            const int32_t* out_step /*dev*/, const int32_t* out_kind /*dev*/, const float* out_w /*dev*/,
                            const UdeDecoder* dec, float* y_pred /*dev*/, uint8_t* out_mask /*dev*/, float* y0_out /*dev | NULL without rp*/,
                            float* tape /*dev | NULL*/, double* side /*dev [8]*/,
                            void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

int ude_rk4_backward_decoded(const UdeDesc* desc, const float* theta /*dev*/, const float* y0 /*dev*/, const UdeReparam* rp /*NULL*/,
                             int32_t n_steps, const float* dt /*dev*/, int32_t n_out, const int32_t* out_step /*dev*/,
                             const int32_t* out_kind /*dev*/, const float* out_w /*dev*/, const UdeDecoder* dec,
                             const float* tape /*dev*/, const uint8_t* out_mask /*dev*/, const float* g_pred /*dev [n_out,B,dec.n_out]*/,
                             const float* g_side /*dev [6] | NULL*/, float* g_y0 /*dev*/, float* g_z /*dev | NULL without rp*/,
                             float* g_theta /*dev*/, float* g_dec /*dev*/, float* dec_partial /*dev scratch*/,
                             void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/*
 * Bayesian ODE funcs: Bayes_Fp / Bayes_Fa / Bayes_FaFp of
 * /root/reference/lib/in_development/models_bayes.py:69-265 are the same three RHS functions over Dense_Variational
 * layers (:12-48), which draw NEW weights at every forward call, i.e. at every RHS evaluation of the solve:
 *     w = w_mean + z * |w_std|,  z ~ N(0, 1) per call, shared by the whole batch                    (:30-48)
 * The caller materialises the 4 * n_steps weight sets (evaluation e = 4 * step + stage, theta layout each):
 *     theta_evals dev [4 * n_steps, n_theta],   z_evals dev [4 * n_steps, n_theta] (adjoint only)
 * and the adjoint returns  g_mean_zsum dev [2, n_theta]:  row 0 = sum_e dL/dtheta_e (= dL/dw_mean),
 * row 1 = sum_e dL/dtheta_e * z_e (dL/dw_std = row 1 * sign(w_std)).  The KL term (get_kl, :118-128) stays with the caller.
 * FFMA engine; models whose weights and gradient accumulators fit in shared memory (else UDE_ERR_UNSUPPORTED);
 * workspace from ude_plan() with UDE_FLAG_BAYES set in desc->flags.  Everything else as ude_rk4_forward / _backward.
 */
int ude_rk4_forward_bayes(const UdeDesc* desc, const float* theta_evals /*dev*/, const float* y0 /*dev*/, int32_t n_steps,
                          const float* dt /*dev*/, int32_t n_out, const int32_t* out_step /*dev*/, const int32_t* out_kind /*dev*/,
                          const float* out_w /*dev*/, float* y_out /*dev*/, float* tape /*dev | NULL*/, double* side /*dev [8]*/,
                          void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

int ude_rk4_backward_bayes(const UdeDesc* desc, const float* theta_evals /*dev*/, const float* z_evals /*dev*/, const float* y0 /*dev*/,
                           int32_t n_steps, const float* dt /*dev*/, int32_t n_out, const int32_t* out_step /*dev*/,
                           const int32_t* out_kind /*dev*/, const float* out_w /*dev*/, const float* tape /*dev*/,
                           const float* g_out /*dev*/, const float* g_side /*dev | NULL*/, float* g_y0 /*dev*/,
                           float* g_mean_zsum /*dev [2, n_theta]*/, void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/* Bayes funcs with the fused consumers (and optionally the fused producer): ude_rk4_*_decoded over ude_rk4_*_bayes --
 * theta_evals / z_evals / g_mean_zsum as in the bayes pair, every other argument as in the decoded pair. */
int ude_rk4_forward_bayes_decoded(const UdeDesc* desc, const float* theta_evals /*dev*/, const float* y0 /*dev | NULL with rp*/,
                                  const UdeReparam* rp /*NULL*/, int32_t n_steps, const float* dt /*dev*/, int32_t n_out,
                                  const int32_t* out_step /*dev*/, const int32_t* out_kind /*dev*/, const float* out_w /*dev*/,
                                  const UdeDecoder* dec, float* y_pred /*dev*/, uint8_t* out_mask /*dev*/, float* y0_out /*dev | NULL*/,
                                  float* tape /*dev | NULL*/, double* side /*dev [8]*/,
                                  void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

int ude_rk4_backward_bayes_decoded(const UdeDesc* desc, const float* theta_evals /*dev*/, const float* z_evals /*dev*/,
                                   const float* y0 /*dev*/, const UdeReparam* rp /*NULL*/, int32_t n_steps, const float* dt /*dev*/,
                                   int32_t n_out, const int32_t* out_step /*dev*/, const int32_t* out_kind /*dev*/,
                                   const float* out_w /*dev*/, const UdeDecoder* dec, const float* tape /*dev*/,
                                   const uint8_t* out_mask /*dev*/, const float* g_pred /*dev*/, const float* g_side /*dev [6] | NULL*/,
                                   float* g_y0 /*dev*/, float* g_z /*dev | NULL*/, float* g_mean_zsum /*dev [2, n_theta]*/,
                                   float* g_dec /*dev*/, float* dec_partial /*dev scratch*/,
                                   void* workspace /*dev*/, int64_t workspace_bytes, void* stream);

/* Controller options of torchdiffeq's RKAdaptiveStepsizeODESolver. */
typedef struct UdeDopri5Opts {
    double  rtol, atol;
    double  first_step;          /* <= 0: torchdiffeq's _select_initial_step      */
    double  safety, ifactor, dfactor;   /* 0.9, 10, 0.2                            */
    int32_t max_steps;           /* cap on ACCEPTED steps per trajectory (tape)   */
    int32_t max_attempts;        /* torchdiffeq's max_num_steps: cap on attempted steps per trajectory and OUTPUT INTERVAL */
    int32_t shared_step;         /* 0: one controller per trajectory.  1: torchdiffeq's batch semantics -- ONE RMS error norm
                                  * over the whole [B,R,L] tensor, one dt, one accept / reject for all trajectories (cooperative
                                  * launch, norms summed across CTAs); side sums then cover EVERY evaluation (initial-step probe
                                  * and rejected attempts too, as the reference's lists do).  Needs batch <= tile_traj * n_ctas of
                                  * ude_plan (UDE_ERR_CAPACITY otherwise).                                                 */
    int32_t reserved;
} UdeDopri5Opts;

/*
 * Adaptive Dopri5 forward, one controller PER TRAJECTORY (= torchdiffeq called
 * with B == 1 for each row; DESIGN.md "Dopri5 batch semantics").
 *   t_out   dev [n_out] fp64 output times, increasing; y_out[0] = y0
 *   counts  dev [B,4] int32: accepted, rejected, status (0 ok, 1 max_steps,
 *           2 max_attempts, 3 dt underflow / non-finite), reserved
 *   steps   dev [max_steps,B] fp64 pairs are inside `tape`.
 */
int ude_dopri5_forward(const UdeDesc* desc, const UdeDopri5Opts* opts,
                       const float* theta /*dev*/, const float* y0 /*dev*/,
                       int32_t n_out, const double* t_out /*dev*/,
                       float* y_out /*dev*/, float* tape /*dev | NULL*/,
                       int32_t* counts /*dev [B,4]*/, double* side /*dev [G,8]*/,
                       void* workspace, int64_t workspace_bytes, void* stream);

int ude_dopri5_backward(const UdeDesc* desc, const UdeDopri5Opts* opts,
                        const float* theta /*dev*/, const float* y0 /*dev*/,
                        int32_t n_out, const double* t_out /*dev*/,
                        const float* tape /*dev*/, const int32_t* counts /*dev*/,
                        const float* g_out /*dev*/, const float* g_side /*dev | NULL*/,
                        float* g_y0 /*dev*/, float* g_theta /*dev*/,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Which kernel family ude_rk4_forward/backward will run for `desc`: 0 = FFMA engine, 1 = tcgen05
 * tensor-core path; negative UdeStatus on a bad descriptor.  Host only. */
int ude_rk4_path(const UdeDesc* desc);

/* Sustained fp32 FFMA throughput of the device in TFLOP/s (2 flops per FMA),
 * measured with a register-resident FMA chain kernel: the roofline denominator
 * bench.py reports against.  Synchronous. */
int ude_measure_ffma_tflops(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDE_B200_H */
