// write_bw.cu — what write-only HBM bandwidth does the render kernel's store pattern reach?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu && ./write_bw
// A: linear (every warp writes consecutive 512-byte pieces of one contiguous span)
// B: the render kernel's pattern: warp w owns rows 32w .. 32w+31 of a [65536][4096] f32 matrix and visits them tile by
//    tile: 8 STG.128 per lane per 32-frame tile, each covering 4 rows x 128 contiguous bytes
// C / D: the same with 64 / 128 contiguous frames (256 / 512 bytes) per row per visit
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int V = 65536, T = 4096;

__global__ void __launch_bounds__(32) k_linear(float4* __restrict__ out, size_t n4_per_warp, float4 v) {
    float4* p = out + (size_t)blockIdx.x * n4_per_warp;
    for (size_t i = threadIdx.x; i < n4_per_warp; i += 32) __stcs(p + i, v);
}

template <int SEG>     // frames per row per visit: 32, 64, 128
__global__ void __launch_bounds__(32) k_tile(float* __restrict__ out, float4 v, int spin) {
    const int lane = threadIdx.x;
    constexpr int LPR = SEG / 4;                 // lanes per row
    constexpr int RPS = 32 / LPR;                // rows per store
    const int q = lane / LPR, c4 = (lane % LPR) * 4;
    float* base = out + (size_t)blockIdx.x * 32 * T;
    float acc = v.x;
    for (int t = 0; t < T; t += SEG) {
        for (int s = 0; s < spin; s++) acc = acc * 1.0000001f + 1e-9f;       // stand-in for the render work between stores
#pragma unroll
        for (int i = 0; i < 32 / RPS; i++) {
            float4 w = v; w.x = acc;
            __stcs(reinterpret_cast<float4*>(base + (size_t)(RPS * i + q) * T + t + c4), w);
        }
    }
}

int main() {
    float* buf[2];
    const size_t bytes = (size_t)V * T * 4;
    cudaMalloc(&buf[0], bytes); cudaMalloc(&buf[1], bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; i++) launch(buf[i & 1]);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; i++) launch(buf[i & 1]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.1f us/GiB  %7.1f GB/s   %s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run("A linear, 2048 warps", [&](float* b) { k_linear<<<2048, 32>>>((float4*)b, bytes / 16 / 2048, v); });
    run("A linear, 16384 warps", [&](float* b) { k_linear<<<16384, 32>>>((float4*)b, bytes / 16 / 16384, v); });
    for (int spin : {0, 64, 256, 512}) {
        char name[96];
        snprintf(name, sizeof name, "B tiles of 32 frames (128 B/row), spin %d", spin);
        run(name, [&](float* b) { k_tile<32><<<2048, 32>>>(b, v, spin); });
        snprintf(name, sizeof name, "C tiles of 64 frames (256 B/row), spin %d", 2 * spin);
        run(name, [&](float* b) { k_tile<64><<<2048, 32>>>(b, v, 2 * spin); });
        snprintf(name, sizeof name, "D tiles of 128 frames (512 B/row), spin %d", 4 * spin);
        run(name, [&](float* b) { k_tile<128><<<2048, 32>>>(b, v, 4 * spin); });
    }
    return 0;
}
