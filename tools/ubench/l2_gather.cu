// Micro-benchmark: what the B200 L2 delivers for the access pattern of attenuate_kernel (K1):
// groups of 8 lanes read 128 contiguous bytes (one float4 per lane) from pseudo-random 512-byte
// rows of an L2-resident slab (3 consecutive source rows + 1 sigT row per "segment"), and reduce
// one float4 per lane into a second L2-resident slab (red.global.add.v4.f32).
// Reports bytes/s at the SM<->L2 interface for: loads only, loads + reductions (the K1 mix).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather l2_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <bool RED, int ROWS_PER_SEG>
__global__ void __launch_bounds__(128) gather_kernel(const float4 *__restrict__ src, float4 *flux, uint32_t n_regions,
                                                     uint32_t fai, int iters, float4 *sink)
{
    const int lane8 = threadIdx.x & 7;
    const uint32_t track = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    float4 acc = make_float4(0, 0, 0, 0);
    // row pitch: 128 floats = 32 float4; slab = [n_regions][fai] rows, then [n_regions] sigT rows
    const float4 *sig = src + (size_t)n_regions * fai * 32;
#pragma unroll 2
    for (int it = 0; it < iters; it++) {
        const uint32_t h = hash32(track * 0x9E3779B9u + it);
        const uint32_t region = h % n_regions;
        const uint32_t r0 = (h >> 24) % (fai - 2);
        const float4 *row = src + ((size_t)region * fai + r0) * 32;
#pragma unroll
        for (int v = 0; v < 3; v++) {          // 3 x 128 B per row per 8 lanes (G = 96 of 104 groups)
            float4 t = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int r = 0; r < ROWS_PER_SEG; r++) {
                const float4 y = __ldg(row + r * 32 + lane8 + 8 * v);
                t.x += y.x; t.y += y.y; t.z += y.z; t.w += y.w;
            }
            const float4 s = __ldg(sig + (size_t)region * 32 + lane8 + 8 * v);
            t.x += s.x; t.y += s.y; t.z += s.z; t.w += s.w;
            if (RED) {
                float *addr = reinterpret_cast<float *>(flux + ((size_t)region * fai + r0 + 1) * 32 + lane8 + 8 * v);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(t.x), "f"(t.y),
                             "f"(t.z), "f"(t.w));
            }
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
    }
    if (acc.x == 123.456f) sink[0] = acc;
}

template <bool RED>
static double run(const float4 *src, float4 *flux, uint32_t n_regions, uint32_t fai, int iters, float4 *sink,
                  int blocks, const char *name)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<RED, 3><<<blocks, 128>>>(src, flux, n_regions, fai, iters, sink);   // warm-up: slab into L2
    cudaEventRecord(e0);
    gather_kernel<RED, 3><<<blocks, 128>>>(src, flux, n_regions, fai, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double segs = (double)blocks * 16 * iters;                 // 16 tracks per CTA
    const double rd = segs * 3 * 128.0 * 4;                           // 3 quads x (3 rows + sigT) x 128 B
    const double red = RED ? segs * 3 * 128.0 : 0.0;
    printf("%-28s %8.3f ms  read %7.2f TB/s  red %6.2f TB/s  total %7.2f TB/s\n", name, ms, rd / ms / 1e9,
           red / ms / 1e9, (rd + red) / ms / 1e9);
    return (rd + red) / ms / 1e9;
}

int main()
{
    const uint32_t n_regions = 6750, fai = 5;                        // the default problem: 17.3 + 3.5 MB slab
    const size_t rows = (size_t)n_regions * fai + n_regions;
    float4 *src, *flux, *sink;
    cudaMalloc(&src, rows * 512);
    cudaMalloc(&flux, (size_t)n_regions * fai * 512);
    cudaMalloc(&sink, 64);
    cudaMemset(src, 0, rows * 512);
    cudaMemset(flux, 0, (size_t)n_regions * fai * 512);
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    for (int occ : {5, 8, 16}) {
        printf("-- %d CTAs of 128 threads per SM (%d SMs)\n", occ, sm);
        run<false>(src, flux, n_regions, fai, 2000, sink, sm * occ * 4, "gather only");
        run<true>(src, flux, n_regions, fai, 2000, sink, sm * occ * 4, "gather + red.v4 (K1 mix)");
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
