// Micro-benchmark: how many issue cycles of an SM sub-partition does a packed FP32x2 instruction take, and can an
// ALU-pipe instruction issue in the shadow of it?  (sm_100a; nvcc -gencode arch=compute_100a,code=sm_100a -O3)
// Every kernel runs 16 warps per sub-partition, all-register operands, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int MODE, int NINT>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, int m)
{
    float2 x[8], y[8], z[8];
    int w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
        y[i] = make_float2(a + i * 1e-7f, a - i * 1e-7f);
        z[i] = make_float2(b + i, b - i);
        w[i] = threadIdx.x * (i + 1);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) x[i] = __ffma2_rn(x[i], y[i], z[i]);                       // packed, three register pairs
            if (MODE == 1) { x[i].x = fmaf(x[i].x, y[i].x, z[i].x); x[i].y = fmaf(x[i].y, y[i].y, z[i].y); }   // two scalar
            if (MODE == 2) x[i] = __fmul2_rn(x[i], y[i]);                             // packed, two register pairs
            if (MODE == 3) x[i] = __ffma2_rn(x[i], make_float2(a, a), make_float2(b, b));   // packed, uniform scalar operands
            if (MODE == 4) x[i] = __ffma2_rn(make_float2(y[i].x, y[i].x), x[i], z[i]);      // packed: per-thread scalar register, pair, pair
            if (MODE == 5) x[i] = __ffma2_rn(x[i], y[i], make_float2(2.f, 2.f));             // packed: pair, pair, immediate
            if (MODE == 6) x[i] = __fmul2_rn(make_float2(y[i].x, y[i].x), x[i]);             // packed: per-thread scalar register, pair
            if (MODE == 7) { x[i].x = fmaf(y[i].x, x[i].x, z[i].x); x[i].y = fmaf(y[i].x, x[i].y, z[i].y); }   // two scalar, shared multiplier
#pragma unroll
            for (int q = 0; q < NINT; q++) w[(i + q) & 7] = (w[(i + q) & 7] ^ m) + it;   // LOP3 + IADD3 (ALU pipe)
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i].x + x[i].y + w[i];
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
    const int blocks = 148 * 8, threads = 256;    // 64 warps per SM = 16 per sub-partition
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    int dev_clock_khz = 0; cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
    const double hz = dev_clock_khz * 1e3;
    auto report = [&](const char *name, float ms, int fp_per_iter, int int_per_iter) {
        // cycles one sub-partition spends per loop iteration of ONE warp
        const double cyc = ms * 1e-3 * hz / (16.0 * ITERS);
        printf("%-44s %8.3f ms  %6.2f cycles per warp-iteration: %d FP + %d ALU instructions -> %.2f cycles per FP instruction if ALU were free, %.2f per instruction\n",
               name, ms, cyc, fp_per_iter, int_per_iter, cyc / fp_per_iter, cyc / (fp_per_iter + int_per_iter));
    };
#define RUN(MODE, NINT, name, nfp) report(name, time_ms([&] { k<MODE, NINT><<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345); }), nfp, 16 * NINT)
    RUN(0, 0, "FFMA2 r,r,r", 8);
    RUN(1, 0, "2 x FFMA r,r,r", 16);
    RUN(2, 0, "FMUL2 r,r", 8);
    RUN(3, 0, "FFMA2 r,scalar,scalar", 8);
    RUN(4, 0, "FFMA2 s(reg),pair,pair", 8);
    RUN(5, 0, "FFMA2 pair,pair,imm", 8);
    RUN(6, 0, "FMUL2 s(reg),pair", 8);
    RUN(7, 0, "2 x FFMA s(reg),r,r", 16);
    RUN(0, 1, "FFMA2 r,r,r + 2 ALU each", 8);
    RUN(1, 1, "2 x FFMA r,r,r + 2 ALU per pair", 16);
    RUN(2, 1, "FMUL2 r,r + 2 ALU each", 8);
    RUN(0, 2, "FFMA2 r,r,r + 4 ALU each", 8);
    RUN(3, 1, "FFMA2 r,s,s + 2 ALU each", 8);
    cudaDeviceSynchronize();
    printf("clock attribute %.0f MHz; last error: %s\n", hz / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
