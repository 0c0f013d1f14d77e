// C-ABI entry points of the Dopri5 path.
extern "C" {
int ude_dopri5_forward(const UdeDesc*, const UdeDopri5Opts*, const float*, const float*, int32_t, const double*, float*,
                       float*, int32_t*, double*, void*, int64_t, void*) {
    return UDE_ERR_UNSUPPORTED;
}
int ude_dopri5_backward(const UdeDesc*, const UdeDopri5Opts*, const float*, const float*, int32_t, const double*,
                        const float*, const int32_t*, const float*, const float*, float*, float*, void*, int64_t, void*) {
    return UDE_ERR_UNSUPPORTED;
}
}
