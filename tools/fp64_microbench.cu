// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) issue rates,
// plus a numerical check of the fragment layouts used by the MAGI kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void k_dfma(double *out, int iters, double s) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    double m = 1.0 + s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], m, s);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CH>
__global__ void k_dmma884(double *out, int iters, double s) {
    double d0[CH], d1[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { d0[i] = 0; d1[i] = 0; }
    double a = s * (threadIdx.x & 7), b = s * (threadIdx.x & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) dmma884(d0[i], d1[i], a, b);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += d0[i] + d1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int CH, int KK>
__global__ void k_dmma16(double *out, int iters, double s) {
    double d[CH][4];
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0;
    double a[8], b[4];
    for (int j = 0; j < 8; ++j) a[j] = s * ((threadIdx.x + j) & 7);
    for (int j = 0; j < 4; ++j) b[j] = s * ((threadIdx.x + j) & 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (KK == 4) { double a2[2] = {a[0], a[1]}; dmma1684(d[i], a2, b[0]); }
            else if (KK == 8) { double a4[4] = {a[0], a[1], a[2], a[3]}; double b2[2] = {b[0], b[1]}; dmma1688(d[i], a4, b2); }
            else dmma16816(d[i], a, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) for (int j = 0; j < 4; ++j) r += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// DMMA fed by one LDS.64 per MMA (B operand from shared memory, conflict-free), A in registers:
// models the banded kernel's inner loop with c chain-groups sharing one table fragment.
template <int C>
__global__ void k_dmma_lds(double *out, int iters, double s) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) sm[i] = s * (i & 15);
    __syncthreads();
    double d0[C][2], d1[C][2];
#pragma unroll
    for (int i = 0; i < C; ++i) { d0[i][0] = d0[i][1] = d1[i][0] = d1[i][1] = 0; }
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = s * ((threadIdx.x + i) & 7);
    int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 64; k += 2) {
            double b0 = sm[k * 32 + lane];
            double b1 = sm[(k + 1) * 32 + lane];
#pragma unroll
            for (int i = 0; i < C; ++i) { dmma884(d0[i][0], d1[i][0], a[i], b0); dmma884(d0[i][1], d1[i][1], a[i], b1); }
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) r += d0[i][0] + d1[i][0] + d0[i][1] + d1[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---- layout check: D(16x8) = A(16x16) * B(16x8) through each shape ----
__global__ void k_layout(const double *A, const double *B, double *D884, double *D1684, double *D1688, double *D16816) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    // m8n8k4: two row blocks, four k chunks
    for (int rb = 0; rb < 2; ++rb) {
        double c0 = 0, c1 = 0;
        for (int kc = 0; kc < 4; ++kc) dmma884(c0, c1, A[(rb * 8 + g) * 16 + kc * 4 + t], B[(kc * 4 + t) * 8 + g]);
        D884[(rb * 8 + g) * 8 + 2 * t] = c0; D884[(rb * 8 + g) * 8 + 2 * t + 1] = c1;
    }
    {
        double c[4] = {0, 0, 0, 0};
        for (int kc = 0; kc < 4; ++kc) { double a[2] = {A[g * 16 + kc * 4 + t], A[(g + 8) * 16 + kc * 4 + t]}; dmma1684(c, a, B[(kc * 4 + t) * 8 + g]); }
        D1684[g * 8 + 2 * t] = c[0]; D1684[g * 8 + 2 * t + 1] = c[1]; D1684[(g + 8) * 8 + 2 * t] = c[2]; D1684[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
    {
        double c[4] = {0, 0, 0, 0};
        for (int kc = 0; kc < 2; ++kc) {
            double a[4] = {A[g * 16 + kc * 8 + t], A[(g + 8) * 16 + kc * 8 + t], A[g * 16 + kc * 8 + t + 4], A[(g + 8) * 16 + kc * 8 + t + 4]};
            double b[2] = {B[(kc * 8 + t) * 8 + g], B[(kc * 8 + t + 4) * 8 + g]};
            dmma1688(c, a, b);
        }
        D1688[g * 8 + 2 * t] = c[0]; D1688[g * 8 + 2 * t + 1] = c[1]; D1688[(g + 8) * 8 + 2 * t] = c[2]; D1688[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
    {
        double c[4] = {0, 0, 0, 0};
        double a[8], b[4];
        for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
        for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
        dmma16816(c, a, b);
        D16816[g * 8 + 2 * t] = c[0]; D16816[g * 8 + 2 * t + 1] = c[1]; D16816[(g + 8) * 8 + 2 * t] = c[2]; D16816[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
}

template <typename F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    // layout check
    {
        std::vector<double> A(256), B(128), ref(128, 0.0);
        for (int i = 0; i < 256; ++i) A[i] = sin(0.37 * i) + 0.01 * i;
        for (int i = 0; i < 128; ++i) B[i] = cos(0.91 * i) - 0.02 * i;
        for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 16; ++k) s += A[i * 16 + k] * B[k * 8 + j]; ref[i * 8 + j] = s; }
        double *dA, *dB, *dD; CK(cudaMalloc(&dA, 256 * 8)); CK(cudaMalloc(&dB, 128 * 8)); CK(cudaMalloc(&dD, 4 * 128 * 8));
        CK(cudaMemcpy(dA, A.data(), 256 * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), 128 * 8, cudaMemcpyHostToDevice));
        k_layout<<<1, 32>>>(dA, dB, dD, dD + 128, dD + 256, dD + 384); CK(cudaDeviceSynchronize());
        std::vector<double> D(512); CK(cudaMemcpy(D.data(), dD, 512 * 8, cudaMemcpyDeviceToHost));
        const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
        for (int v = 0; v < 4; ++v) { double e = 0; for (int i = 0; i < 128; ++i) e = fmax(e, fabs(D[v * 128 + i] - ref[i])); printf("{\"layout_check\": \"%s\", \"max_abs_err\": %.3e}\n", names[v], e); }
    }
    const int threads = 256;
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int blocks = sms * bps; int warps = blocks * threads / 32;
        { double ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1e-9); }, 20);
          double fl = 2.0 * 8 * iters * (double)blocks * threads; printf("{\"test\": \"dfma\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1e-9); }, 20);
          double fl = 2.0 * 256 * 8 * iters * (double)warps; printf("{\"test\": \"dmma_m8n8k4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma16<4, 4><<<blocks, threads>>>(out, iters, 1e-9); }, 20);
          double fl = 2.0 * 512 * 4 * iters * (double)warps; printf("{\"test\": \"dmma_m16n8k4\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma16<4, 8><<<blocks, threads>>>(out, iters, 1e-9); }, 20);
          double fl = 2.0 * 1024 * 4 * iters * (double)warps; printf("{\"test\": \"dmma_m16n8k8\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma16<4, 16><<<blocks, threads>>>(out, iters / 4, 1e-9); }, 20);
          double fl = 2.0 * 2048 * 4 * (iters / 4) * (double)warps; printf("{\"test\": \"dmma_m16n8k16\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", bps, ms, fl / ms * 1e-9); }
    }
    // occupancy sweep for m8n8k4: warps per SM 4, 8, 16 with 1..4 independent chains
    for (int w = 4; w <= 32; w *= 2) {
        int blocks = sms, th = w * 32;
        { double ms = time_ms([&] { k_dmma884<1><<<blocks, th>>>(out, iters, 1e-9); }, 20);
          printf("{\"test\": \"dmma_m8n8k4_chains1\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", w, 2.0 * 256 * 1 * iters * (double)blocks * w / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma884<2><<<blocks, th>>>(out, iters, 1e-9); }, 20);
          printf("{\"test\": \"dmma_m8n8k4_chains2\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", w, 2.0 * 256 * 2 * iters * (double)blocks * w / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma884<4><<<blocks, th>>>(out, iters, 1e-9); }, 20);
          printf("{\"test\": \"dmma_m8n8k4_chains4\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", w, 2.0 * 256 * 4 * iters * (double)blocks * w / ms * 1e-9); }
    }
    // LDS-fed DMMA: C chain-groups per table fragment
    for (int w = 4; w <= 16; w *= 2) {
        int blocks = sms, th = w * 32; size_t smb = 32 * 64 * 8;
        { double ms = time_ms([&] { k_dmma_lds<1><<<blocks, th, smb>>>(out, 64, 1e-9); }, 20);
          printf("{\"test\": \"dmma_lds_c1\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", w, 2.0 * 256 * 64 * 1 * 64 * (double)blocks * w / ms * 1e-9); }
        { double ms = time_ms([&] { k_dmma_lds<2><<<blocks, th, smb>>>(out, 64, 1e-9); }, 20);
          printf("{\"test\": \"dmma_lds_c2\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", w, 2.0 * 256 * 64 * 2 * 64 * (double)blocks * w / ms * 1e-9); }
    }
    // sustained: ~2 s of dmma and dfma back to back
    {
        int blocks = sms * 2;
        double ms = time_ms([&] { k_dmma884<8><<<blocks, threads>>>(out, iters * 8, 1e-9); }, 200);
        printf("{\"test\": \"dmma_m8n8k4_sustained\", \"ms_per_launch\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 256 * 8 * iters * 8 * (double)blocks * threads / 32 / ms * 1e-9);
        ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters * 8, 1e-9); }, 200);
        printf("{\"test\": \"dfma_sustained\", \"ms_per_launch\": %.3f, \"tflops\": %.2f}\n", ms, 2.0 * 8 * iters * 8 * (double)blocks * threads / ms * 1e-9);
    }
    return 0;
}
