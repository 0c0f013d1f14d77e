// C-ABI entry points of the Dopri5 path (included at the end of ude_host.cu).
namespace ude {
static int dopri_opts(const UdeDopri5Opts* o, DopriOpts& d) {
    if (!o) return UDE_ERR_INVALID;
    if (!(o->rtol > 0.0) || !(o->atol >= 0.0) || o->max_steps < 1 || o->max_attempts < 1) return UDE_ERR_INVALID;
    d.rtol = (float)o->rtol;
    d.atol = (float)o->atol;
    d.first_step = o->first_step;
    d.safety = o->safety > 0 ? o->safety : 0.9;
    d.ifactor = o->ifactor > 0 ? o->ifactor : 10.0;
    d.dfactor = o->dfactor > 0 ? o->dfactor : 0.2;
    d.max_steps = o->max_steps;
    d.max_attempts = o->max_attempts;
    d.shared_step = o->shared_step != 0;
    return UDE_OK;
}
}  // namespace ude

extern "C" {

int ude_dopri5_forward(const UdeDesc* desc, const UdeDopri5Opts* opts, const float* theta, const float* y0,
                       int32_t n_out, const double* t_out, float* y_out, float* tape, int32_t* counts, double* side,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    DopriOpts o;
    if ((st = dopri_opts(opts, o)) != UDE_OK) return st;
    if (!theta || !y0 || !t_out || !y_out || !counts || !side || !workspace || n_out < 2) return UDE_ERR_INVALID;
    if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
    if (!device_info().ok) return UDE_ERR_NO_DEVICE;
    ModelP P;
    if ((st = choose_plan(desc, P, DF_NSTATE, DF_NRB, false)) != UDE_OK) return st;
    if (P.TB > kMaxTB) return UDE_ERR_UNSUPPORTED;
    const int n_ctas = grid_size(P, desc->batch);
    Workspace w = carve_workspace(workspace, P, n_ctas, P.cold_global ? P.cold_floats : 0);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if ((st = launch_pack(P, theta, w.wpacked, s)) != UDE_OK) return st;
    DopriFwdArgs a;
    a.wpacked = w.wpacked;
    a.y0 = y0;
    a.n_out = n_out;
    a.t_out = t_out;
    a.y_out = y_out;
    a.tape_x = tape;
    a.tape_t = tape ? reinterpret_cast<double*>(reinterpret_cast<char*>(tape) + dopri_tape_x_bytes(P, desc->batch, o.max_steps))
                    : nullptr;
    a.counts = counts;
    a.side_partial = w.side_partial;
    a.cold = w.cold;
    a.B = desc->batch;
    a.o = o;
    a.gpart = reinterpret_cast<double*>(w.slabs);      // the adjoint's slab area is idle in the forward (>= 4 * n_ctas doubles)
    if (o.shared_step && (desc->batch + P.TB - 1) / P.TB > n_ctas) return UDE_ERR_CAPACITY;     // one resident tile per CTA
    const int smem = P.smem_floats * 4;
    void* kargs[2] = {&P, &a};
#define UDE_LAUNCH_DFWD(W, C)                                                             \
    do {                                                                                  \
        if ((st = set_smem(dopri5_fwd_kernel<W, C>, smem)) != UDE_OK) return st;          \
        if (o.shared_step)                                                                \
            UDE_CUDA(cudaLaunchCooperativeKernel((const void*)dopri5_fwd_kernel<W, C>, dim3(n_ctas), dim3(kThreads), kargs, smem, s)); \
        else                                                                              \
            dopri5_fwd_kernel<W, C><<<n_ctas, kThreads, smem, s>>>(P, a);                 \
    } while (0)
    if (P.cold_global) UDE_LAUNCH_DFWD(false, true);
    else if (P.w_in_smem) UDE_LAUNCH_DFWD(true, false);
    else UDE_LAUNCH_DFWD(false, false);
#undef UDE_LAUNCH_DFWD
    UDE_CUDA(cudaGetLastError());
    reduce_side_kernel<<<1, 32, 0, s>>>(w.side_partial, n_ctas, side);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

int ude_dopri5_backward(const UdeDesc* desc, const UdeDopri5Opts* opts, const float* theta, const float* y0,
                        int32_t n_out, const double* t_out, const float* tape, const int32_t* counts,
                        const float* g_out, const float* g_side, float* g_y0, float* g_theta, void* workspace,
                        int64_t workspace_bytes, void* stream) {
    int st = check_desc(desc);
    if (st != UDE_OK) return st;
    DopriOpts o;
    if ((st = dopri_opts(opts, o)) != UDE_OK) return st;
    if (!theta || !y0 || !t_out || !tape || !counts || !g_out || !g_y0 || !g_theta || !workspace || n_out < 2)
        return UDE_ERR_INVALID;
    if (desc->n_groups != 1) return UDE_ERR_UNSUPPORTED;
    if (!device_info().ok) return UDE_ERR_NO_DEVICE;
    ModelP P;
    if ((st = choose_plan(desc, P, DB_NSTATE, 0, true)) != UDE_OK) return st;
    if (P.TB > kMaxTB) return UDE_ERR_UNSUPPORTED;
    const int n_ctas = grid_size(P, desc->batch);
    Workspace w = carve_workspace(workspace, P, n_ctas, P.cold_global ? P.cold_floats : 0);
    if (w.bytes > workspace_bytes) return UDE_ERR_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if ((st = launch_pack(P, theta, w.wpacked, s)) != UDE_OK) return st;
    DopriBwdArgs a;
    a.wpacked = w.wpacked;
    a.y0 = y0;
    a.n_out = n_out;
    a.t_out = t_out;
    a.tape_x = tape;
    a.tape_t = reinterpret_cast<const double*>(reinterpret_cast<const char*>(tape) + dopri_tape_x_bytes(P, desc->batch, o.max_steps));
    a.counts = counts;
    a.g_out = g_out;
    a.g_side = g_side;
    a.g_y0 = g_y0;
    a.slabs = w.slabs;
    a.cold = w.cold;
    a.B = desc->batch;
    a.o = o;
    const int smem = P.smem_floats * 4;
#define UDE_LAUNCH_DBWD(W, G, C)                                                          \
    do {                                                                                  \
        if ((st = set_smem(dopri5_bwd_kernel<W, G, C>, smem)) != UDE_OK) return st;       \
        dopri5_bwd_kernel<W, G, C><<<n_ctas, kThreads, smem, s>>>(P, a);                  \
    } while (0)
    if (P.cold_global) UDE_LAUNCH_DBWD(false, false, true);
    else if (P.w_in_smem && P.g_in_smem) UDE_LAUNCH_DBWD(true, true, false);
    else if (P.w_in_smem) UDE_LAUNCH_DBWD(true, false, false);
    else UDE_LAUNCH_DBWD(false, false, false);
#undef UDE_LAUNCH_DBWD
    UDE_CUDA(cudaGetLastError());
    reduce_dtheta_kernel<<<(P.n_theta + 255) / 256, 256, 0, s>>>(P, w.slabs, n_ctas, g_theta);
    UDE_CUDA(cudaGetLastError());
    return UDE_OK;
}

}  // extern "C"
