// FP32 pipe throughput on sm_100a: scalar FFMA / FMUL / FADD vs packed fma.rn.f32x2 (FFMA2).
// Prints G lane-ops/s per variant; peak nominal = 148 SMs x 128 lanes x clock.
#include <cuda_runtime.h>
#include <cstdio>

#ifndef ITERS
#define ITERS 4096
#endif
#define CHAINS 8

__global__ void k_ffma(float* out, float a, float b) {
    float x[CHAINS];
    for (int i = 0; i < CHAINS; i++) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct register operands per FFMA (no constant-bank / immediate operand)
__global__ void k_ffma3(float* out, float a, float b) {
    float x[CHAINS], y[CHAINS], z[CHAINS];
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = a + i * 1e-3f; z[i] = b + i * 1e-3f + threadIdx.x * 1e-6f; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = fmaf(x[i], y[i], z[i]);
    }
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmul(float* out, float a, float b) {
    float x[CHAINS], y[CHAINS];
    for (int i = 0; i < CHAINS; i++) { x[i] = 1.0f + threadIdx.x * 1e-6f + i * 1e-3f; y[i] = a + i * 1e-7f; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) x[i] = x[i] * y[i];
    }
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[CHAINS], y[CHAINS], z[CHAINS];
    for (int i = 0; i < CHAINS; i++) {
        float2 fx = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i), fy = make_float2(a + i * 1e-3f, a - i * 1e-3f), fz = make_float2(b + i * 1e-3f, b + threadIdx.x * 1e-6f);
        x[i] = *reinterpret_cast<unsigned long long*>(&fx);
        y[i] = *reinterpret_cast<unsigned long long*>(&fy);
        z[i] = *reinterpret_cast<unsigned long long*>(&fz);
    }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y[i]), "l"(z[i]));
    }
    float s = 0;
    for (int i = 0; i < CHAINS; i++) { float2 f = *reinterpret_cast<float2*>(&x[i]); s += f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(float* out, float a, float b) {   // FFMA3 interleaved with FMUL 1:1 (our loop's mix)
    float x[CHAINS], y[CHAINS], z[CHAINS], w[CHAINS];
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = a + i * 1e-3f; z[i] = b + i * 1e-3f + threadIdx.x * 1e-6f; w[i] = 1.0f + i * 1e-6f; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) { x[i] = fmaf(x[i], y[i], z[i]); w[i] = w[i] * y[i]; }
    }
    float s = 0;
    for (int i = 0; i < CHAINS; i++) s += x[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kernel, double ops_per_thread_iter, int blocks, int threads, float* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kernel<<<blocks, threads>>>(out, 1.0001f, 0.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) kernel<<<blocks, threads>>>(out, 1.0001f, 0.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 5.0 * blocks * threads * (double)ITERS * ops_per_thread_iter;
    printf("%-10s blocks/SM %2d  threads %4d : %8.1f G lane-ops/s  (%.3f ms)\n", name, blocks / 148, threads, ops / (ms * 1e-3) / 1e9, ms / 5);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 16 * 1024 * sizeof(float));
    for (int bps : {1, 2, 4, 8}) {
        int blocks = 148 * bps, threads = 256;
        run("ffma_c", k_ffma, CHAINS, blocks, threads, out);
        run("ffma_3reg", k_ffma3, CHAINS, blocks, threads, out);
        run("fmul", k_fmul, CHAINS, blocks, threads, out);
        run("ffma2", k_ffma2, 2.0 * CHAINS, blocks, threads, out);
        run("mix_fma_mul", k_mix, 2.0 * CHAINS, blocks, threads, out);
    }
    printf("nominal: 148 x 128 x 1.965 GHz = %.1f G lane-ops/s\n", 148 * 128 * 1.965);
    return 0;
}
