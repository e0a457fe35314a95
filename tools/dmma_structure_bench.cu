// Microbenchmark of the K1 inner structure: per step, NCHAIN interleaved accumulate chains of 12 DMMAs whose A and B
// operands come from distinct registers (A: sliding register window, B: "fragment" registers), optional window shift,
// optional B from shared memory.  Reports achieved fraction of the DMMA peak for W warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NCHAIN, bool SHIFT, bool BSMEM, int SPLIT>
__global__ void k(double* out, const double* in, int steps) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 12 * NCHAIN; i += blockDim.x) sm[i] = 1e-9 * i;
    __syncthreads();
    double w[NCHAIN][13], f[NCHAIN][12];
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c) {
#pragma unroll
        for (int i = 0; i < 13; ++i) w[c][i] = in[(c * 13 + i) * 32 + lane];
#pragma unroll
        for (int i = 0; i < 12; ++i) f[c][i] = in[(c * 12 + i) * 32 + lane + 1024];
    }
    double tot = 0.0;
    for (int s = 0; s < steps; ++s) {
        double acc[NCHAIN][SPLIT][2];
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c)
#pragma unroll
            for (int u = 0; u < SPLIT; ++u) acc[c][u][0] = acc[c][u][1] = 0.0;
#pragma unroll
        for (int hh = 0; hh < 12; ++hh)
#pragma unroll
            for (int c = 0; c < NCHAIN; ++c) {
                double b = BSMEM ? sm[(c * 12 + hh) * 32 + lane] : f[c][hh];
                dmma(acc[c][hh % SPLIT][0], acc[c][hh % SPLIT][1], w[c][hh], b);
            }
#pragma unroll
        for (int c = 0; c < NCHAIN; ++c) {
            double r0 = 0, r1 = 0;
#pragma unroll
            for (int u = 0; u < SPLIT; ++u) { r0 += acc[c][u][0]; r1 += acc[c][u][1]; }
            if (SHIFT) {
#pragma unroll
                for (int i = 0; i < 11; ++i) w[c][i] = w[c][i + 2];
                w[c][11] = r0 * 1e-30 + 1.0; w[c][12] = r1 * 1e-30 + 1.0;
            } else tot += r0 + r1;
        }
    }
#pragma unroll
    for (int c = 0; c < NCHAIN; ++c)
#pragma unroll
        for (int i = 0; i < 13; ++i) tot += w[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
}
template <int NCHAIN, bool SHIFT, bool BSMEM, int SPLIT>
void run(const char* name, double* out, double* in, int sms) {
    for (int wps = 4; wps <= 16; wps *= 2) {
        int steps = 2000;
        size_t smb = 32 * 12 * NCHAIN * 8;
        k<NCHAIN, SHIFT, BSMEM, SPLIT><<<sms, wps * 32, smb>>>(out, in, steps);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<NCHAIN, SHIFT, BSMEM, SPLIT><<<sms, wps * 32, smb>>>(out, in, steps);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 256 * 12 * NCHAIN * (double)steps * sms * wps / ms * 1e-9;
        printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"tflops\": %.2f, \"frac_of_37.1\": %.3f}\n", name, wps, tf, tf / 37.1);
    }
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double *out, *in; CK(cudaMalloc(&out, 8 * 148 * 1024)); CK(cudaMalloc(&in, 8 * 4096)); CK(cudaMemset(in, 0, 8 * 4096));
    run<2, false, false, 1>("2chains_noshift_breg", out, in, p.multiProcessorCount);
    run<2, true, false, 1>("2chains_shift_breg", out, in, p.multiProcessorCount);
    run<2, true, true, 1>("2chains_shift_bsmem", out, in, p.multiProcessorCount);
    run<2, true, false, 2>("2chains_split2_shift_breg", out, in, p.multiProcessorCount);
    run<2, true, true, 2>("2chains_split2_shift_bsmem", out, in, p.multiProcessorCount);
    run<1, true, true, 2>("1chain_split2_shift_bsmem", out, in, p.multiProcessorCount);
    run<1, true, true, 4>("1chain_split4_shift_bsmem", out, in, p.multiProcessorCount);
    run<4, true, true, 1>("4chains_shift_bsmem", out, in, p.multiProcessorCount);
    return 0;
}
