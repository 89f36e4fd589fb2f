// dfma_rate.cu — binary64 throughput and latency on this part (the moving-cutoff chunk evaluates exp2 / sin / cos
// in binary64 per frame).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rate dfma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int UNR, int OP>
__global__ void k(double* out, double a, double b, long long* cyc) {
    double x[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) x[i] = a + i + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < 1024; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) x[i] = fma(x[i], a, b);
            if (OP == 1) x[i] = (double)(float)x[i] + a;          // F2F down + up + DADD
            if (OP == 2) x[i] = rint(x[i] * a);                   // DMUL + FRND.F64
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int UNR, int OP>
void run(const char* name, int warps_per_smsp) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    for (int rep = 0; rep < 2; rep++) { k<UNR, OP><<<148, 128 * warps_per_smsp>>>(out, 1.0000001, 1e-9, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s UNR=%d %dw/SMSP: %.3f ops/clk/SMSP, %.1f cycles per dependent step\n", name, UNR, warps_per_smsp,
           1024.0 * UNR * warps_per_smsp / c, (double)c / 1024 / (UNR == 1 ? 1 : UNR) * (UNR == 1 ? 1 : 0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, 0>("DFMA chain (latency)", 1);
    run<8, 0>("DFMA x8 independent", 1);
    run<8, 0>("DFMA x8 independent", 4);
    run<8, 1>("F2F.F32.F64+F2F.F64.F32+DADD", 4);
    run<8, 2>("DMUL+FRND.F64", 4);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
