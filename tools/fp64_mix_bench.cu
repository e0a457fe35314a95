// Does scalar FP64 work (DFMA) interfere with DMMA on the shared FP64 unit beyond its own issue slots?
// Variants: (a) all warps DMMA; (b) all warps DFMA; (c) half the warps of every SM sub-partition DMMA, half DFMA;
// (d) every warp alternates blocks of NB DMMAs and NF DFMAs (the K1 pattern: 24 DMMAs then ~40 scalar FP64 ops).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// mode 0: dmma only, 1: dfma only, 2: warp-split (warps with (warp>>2)&1 do dfma), 3: alternate in every warp
template <int MODE, int NB, int NF, bool DEP>
__global__ void k(double* out, int iters, double s) {
    const int warp = threadIdx.x >> 5;
    double c[4][2] = {{0,0},{0,0},{0,0},{0,0}};
    double f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
    const double a = s * (threadIdx.x & 7), b = s * (threadIdx.x & 3), m = 1.0 + s;
    const bool do_dmma = (MODE == 0) || (MODE == 3) || (MODE == 2 && !((warp >> 2) & 1));
    const bool do_dfma = (MODE == 1) || (MODE == 3) || (MODE == 2 && ((warp >> 2) & 1));
    for (int it = 0; it < iters; ++it) {
        if (do_dmma) {
#pragma unroll
            for (int i = 0; i < NB; ++i) dmma(c[i & 1][0], c[i & 1][1], a, b);
        }
        if (do_dfma) {
            if (DEP) {   // one dependent chain, as in a pointwise evaluation
#pragma unroll
                for (int i = 0; i < NF; ++i) f[0] = fma(f[0], m, s);
            } else {
#pragma unroll
                for (int i = 0; i < NF; ++i) f[i & 7] = fma(f[i & 7], m, s);
            }
        }
    }
    double r = 0;
    for (int i = 0; i < 8; ++i) r += f[i];
    for (int i = 0; i < 4; ++i) r += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE, int NB, int NF, bool DEP>
void run(const char* name, double* out, int sms, int wps) {
    int iters = 4000;
    k<MODE, NB, NF, DEP><<<sms, wps * 32>>>(out, iters, 1e-9); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE, NB, NF, DEP><<<sms, wps * 32>>>(out, iters, 1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double nd = (MODE == 1) ? 0 : (MODE == 2 ? 0.5 : 1.0), nf = (MODE == 0) ? 0 : (MODE == 2 ? 0.5 : 1.0);
    double cyc = ms * 1e-3 * 1.965e9;
    double dmma_cyc_per_smsp = nd * NB * (double)iters * wps / 4 * 16.0;     // pipe cycles the DMMAs need
    double dfma_cyc_per_smsp = nf * NF * (double)iters * wps / 4 * 2.2;
    printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.3f, \"cycles\": %.0f, \"dmma_pipe_cycles\": %.0f, \"dfma_pipe_cycles\": %.0f, \"pipe_demand_over_elapsed\": %.3f}\n",
           name, wps, ms, cyc, dmma_cyc_per_smsp, dfma_cyc_per_smsp, (dmma_cyc_per_smsp + dfma_cyc_per_smsp) / cyc);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; CK(cudaMalloc(&out, 8 * 148 * 1024));
    int sms = p.multiProcessorCount;
    for (int wps = 8; wps <= 16; wps *= 2) {
        run<0, 24, 0, false>("dmma_only", out, sms, wps);
        run<1, 0, 40, false>("dfma_only_indep", out, sms, wps);
        run<1, 0, 40, true>("dfma_only_dep", out, sms, wps);
        run<2, 24, 40, false>("warp_split_indep", out, sms, wps);
        run<2, 24, 40, true>("warp_split_dep", out, sms, wps);
        run<3, 24, 40, false>("alternate_24dmma_40dfma_indep", out, sms, wps);
        run<3, 24, 40, true>("alternate_24dmma_40dfma_dep", out, sms, wps);
        run<3, 24, 10, true>("alternate_24dmma_10dfma_dep", out, sms, wps);
        run<3, 2, 2, true>("alternate_2dmma_2dfma_dep", out, sms, wps);
    }
    return 0;
}
