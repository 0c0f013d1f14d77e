// Shared-memory wavefront micro-benchmark: cycles per LDS for several lane->address patterns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/smem_bench tools/smem_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC>
__global__ void k(const int* __restrict__ lane_off /*[32] float index*/, int iters, float* out, long long* cyc) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int base = lane_off[lane];
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = (base + u * 256) & 8191 & ~(VEC - 1);
            const unsigned addr = (unsigned)__cvta_generic_to_shared(&sm[idx]);
            if (VEC == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); acc += v.x; }
            else if (VEC == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr)); acc += v.x; }
            else { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); acc += v; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 1.2345f) out[0] = acc;
}
template <int VEC>
double run(const int* h_off, int warps) {
    int* d_off; float* d_out; long long* d_cyc;
    cudaMalloc(&d_off, 128); cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, 8 * 148);
    cudaMemcpy(d_off, h_off, 128, cudaMemcpyHostToDevice);
    const int iters = 2000;
    k<VEC><<<148, warps * 32, 32768>>>(d_off, 10, d_out, d_cyc);
    k<VEC><<<148, warps * 32, 32768>>>(d_off, iters, d_out, d_cyc);
    long long c[148];
    cudaMemcpy(c, d_cyc, 8 * 148, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += c[i]; avg /= 148;
    cudaFree(d_off); cudaFree(d_out); cudaFree(d_cyc);
    return avg / (double)(iters * 16 * warps);       // SM cycles per warp-level LDS
}
int main() {
    int off[32];
    struct P { const char* name; int vec; int (*f)(int); };
    P ps[] = {
        {"LDS.128 all 32 lanes distinct contiguous (512B)", 4, [](int l) { return l * 4; }},
        {"LDS.128 uniform (all lanes same 16B)", 4, [](int l) { return 0; }},
        {"LDS.128 A-type now: 8 distinct float4 (tx=l&7) x4 replicas", 4, [](int l) { return (l & 7) * 4; }},
        {"LDS.128 W-type now: 4 distinct rows (ty=l>>3), stride 68", 4, [](int l) { return (l >> 3) * 68; }},
        {"LDS.128 16 distinct float4 (l&15) x2 replicas", 4, [](int l) { return (l & 15) * 4; }},
        {"LDS.128 8 distinct float4 replicated as l>>2 (4 consecutive lanes same)", 4, [](int l) { return (l >> 2) * 4; }},
        {"LDS.128 4 distinct float4 as l>>3 contiguous", 4, [](int l) { return (l >> 3) * 4; }},
        {"LDS.128 2 distinct float4 as l>>4", 4, [](int l) { return (l >> 4) * 4; }},
        {"LDS.128 4 distinct as l&3 (interleaved)", 4, [](int l) { return (l & 3) * 4; }},
        {"LDS.128 8 rows stride 68 (tx=l&7) x4 replicas (dW A-type)", 4, [](int l) { return (l & 7) * 68; }},
        {"LDS.64 all 32 distinct contiguous (256B)", 2, [](int l) { return l * 2; }},
        {"LDS.64 uniform", 2, [](int l) { return 0; }},
        {"LDS.64 16 distinct (l&15) x2", 2, [](int l) { return (l & 15) * 2; }},
        {"LDS.32 all distinct contiguous (128B)", 1, [](int l) { return l; }},
        {"LDS.32 uniform", 1, [](int l) { return 0; }},
        {"LDS.32 8 distinct (l&7) x4", 1, [](int l) { return l & 7; }},
    };
    for (auto& p : ps) {
        for (int l = 0; l < 32; ++l) off[l] = p.f(l);
        double c8 = p.vec == 4 ? run<4>(off, 8) : p.vec == 2 ? run<2>(off, 8) : run<1>(off, 8);
        double c16 = p.vec == 4 ? run<4>(off, 16) : p.vec == 2 ? run<2>(off, 16) : run<1>(off, 16);
        printf("%-75s  cyc/LDS: %.2f (8 warps)  %.2f (16 warps)\n", p.name, c8, c16);
    }
    return 0;
}
