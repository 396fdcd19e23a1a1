// Micro-benchmark: issue/pipe rate of packed FP32x2 (FFMA2/FMUL2/FADD2) vs scalar FFMA on sm_100a,
// alone and interleaved with integer (ALU-pipe) work.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_ffma(float *out, float a, float b)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2(float *out, float a, float b)
{
    float2 x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same FLOPs as k_ffma, plus 8 integer ops per iteration
__global__ void k_ffma_int(float *out, float a, float b, int m)
{
    float x[8];
    int y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; y[i] = threadIdx.x * i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { x[i] = fmaf(x[i], a, b); y[i] = (y[i] ^ m) + it; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma2_int(float *out, float a, float b, int m)
{
    float2 x[4];
    int y[8];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = threadIdx.x * i;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = __ffma2_rn(x[i], a2, b2);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = (y[i] ^ m) + it;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
#pragma unroll
    for (int i = 0; i < 8; i++) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mul2 / add2 mix
__global__ void k_mix2(float *out, float a, float b)
{
    float2 x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] = __fmul2_rn(x[i], a2); x[i] = __fadd2_rn(x[i], b2); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main()
{
    const int blocks = 148 * 8, threads = 256;
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    const double fma_lane = (double)blocks * threads * ITERS * 8;   // FMAs (lane level) per launch, all kernels
    auto report = [&](const char *n, float ms, double extra) {
        printf("%-14s %8.3f ms  %7.2f TFLOP/s fp32   (%.2f lane-FMA/clk/SM @1.965GHz)%s\n", n, ms, 2 * fma_lane / ms / 1e9,
               fma_lane / (ms * 1e-3) / 148 / 1.965e9, extra > 0 ? "  [+8 int ops/iter]" : "");
    };
    report("FFMA", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 0);
    report("FFMA2", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 0);
    report("FFMA+int", time_ms([&] { k_ffma_int<<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345); }), 1);
    report("FFMA2+int", time_ms([&] { k_ffma2_int<<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345); }), 1);
    report("FMUL2+FADD2", time_ms([&] { k_mix2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 0);
    cudaDeviceSynchronize();
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
