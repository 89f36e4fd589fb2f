// pipe_rates.cu — issue/pipe throughput of the instructions the render loop is made of, on one B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe_rates pipe_rates.cu
// Each kernel runs ITER trips of UNR independent ops per thread; reports warp-instructions per
// cycle per SM sub-partition (SMSP) at 1..4 warps per SMSP (blocks of 128 threads = 1 warp/SMSP).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 1024;
constexpr int UNR = 8;

template <int OP>
__global__ void k(float* out, float a, float b, long long* cyc) {
    float x[UNR]; float2 y[UNR]; unsigned u[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) { x[i] = a + i + threadIdx.x; y[i] = make_float2(a + i, b + threadIdx.x); u[i] = threadIdx.x * 7 + i; }
    const float2 a2 = make_float2(a, b), b2 = make_float2(b, a);
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) x[i] = __fmaf_rn(x[i], a, b);                 // FFMA
            if (OP == 1) y[i] = __ffma2_rn(y[i], a2, b2);              // FFMA2
            if (OP == 2) x[i] = __fadd_rn(x[i], a);                    // FADD
            if (OP == 3) y[i] = __fadd2_rn(y[i], a2);                  // FADD2
            if (OP == 4) x[i] = __fmul_rn(x[i], a);                    // FMUL
            if (OP == 5) y[i] = __fmul2_rn(y[i], a2);                  // FMUL2
            if (OP == 6) { u[i] = (u[i] ^ 0x1234567u) * 0x9e3779b9u; } // LOP3 + IMAD
            if (OP == 7) { x[i] = __uint2float_rn(__float_as_uint(x[i]) & 0xffffu); }   // LOP + I2F
            if (OP == 8) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ 0x1234567u) + 0x9e3779b9u; }  // FFMA2 + LOP3 + IADD
            if (OP == 9) { x[i] = __fmaf_rn(x[i], a, b); u[i] = (u[i] ^ 0x1234567u) + 0x9e3779b9u; }     // FFMA + LOP3 + IADD
            if (OP == 10) { x[i] = x[i] >= 1.0f ? __fadd_rn(x[i], -1.0f) : x[i]; }                        // FSETP + @P FADD
            if (OP == 11) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = __fmaf_rn(x[i], a, b); }             // FFMA2 + FFMA
            if (OP == 12) {   // 17 scalar FP (the render frame, all-scalar form)
#pragma unroll
                for (int r = 0; r < 17; r++) x[i] = (r % 3 == 0) ? __fmaf_rn(x[i], a, b) : (r % 3 == 1) ? __fadd_rn(x[i], a) : __fmul_rn(x[i], b);
            }
            if (OP == 13) {   // 9 scalar FP + 4 packed (time-packed render frame)
#pragma unroll
                for (int r = 0; r < 9; r++) x[i] = (r % 3 == 0) ? __fmaf_rn(x[i], a, b) : (r % 3 == 1) ? __fadd_rn(x[i], a) : __fmul_rn(x[i], b);
                y[i] = __ffma2_rn(y[i], a2, b2); y[i] = __fadd2_rn(y[i], a2); y[i] = __fmul2_rn(y[i], b2); y[i] = __ffma2_rn(y[i], a2, b2);
            }
            if (OP == 14) {   // OP 13 + IMAD + 2 LOP3 + I2F (full time-packed frame mix)
#pragma unroll
                for (int r = 0; r < 9; r++) x[i] = (r % 3 == 0) ? __fmaf_rn(x[i], a, b) : (r % 3 == 1) ? __fadd_rn(x[i], a) : __fmul_rn(x[i], b);
                y[i] = __ffma2_rn(y[i], a2, b2); y[i] = __fadd2_rn(y[i], a2); y[i] = __fmul2_rn(y[i], b2);
                u[i] = (u[i] ^ 0x1234567u) * 0x79b9u; u[i] ^= 0x55u;
                y[i].x = __fmaf_rn(__uint2float_rn(u[i] & 0xffffu), a, y[i].x);
                y[i] = __ffma2_rn(y[i], a2, b2);
            }
            if (OP == 15) {   // all-scalar frame mix: 17 FP + IMAD + 2 LOP3 + I2F
#pragma unroll
                for (int r = 0; r < 16; r++) x[i] = (r % 3 == 0) ? __fmaf_rn(x[i], a, b) : (r % 3 == 1) ? __fadd_rn(x[i], a) : __fmul_rn(x[i], b);
                u[i] = (u[i] ^ 0x1234567u) * 0x79b9u; u[i] ^= 0x55u;
                x[i] = __fmaf_rn(__uint2float_rn(u[i] & 0xffffu), a, x[i]);
            }
        }
    }
    float s = 0; unsigned us = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) { s += x[i] + y[i].x + y[i].y; us += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + us;
    __syncthreads();
    long long t2 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t2 - t0;   // until the last warp of the block is done
}

template <int OP>
void run(const char* name, int inst_per_op, float* d, long long* dc) {
    printf("%-28s", name);
    // one block per SM (148 blocks), 4*wps warps per block -> wps warps on every SMSP (wid % 4)
    for (int wps = 1; wps <= 8; wps *= 2) {
        const int threads = 128 * wps;
        k<OP><<<148, threads>>>(d, 1.0001f, 0.5f, dc);   // warm
        cudaDeviceSynchronize();
        k<OP><<<148, threads>>>(d, 1.0001f, 0.5f, dc);
        cudaDeviceSynchronize();
        long long c[148], mx = 0; cudaMemcpy(c, dc, sizeof c, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 148; i++) mx = c[i] > mx ? c[i] : mx;
        printf("  %dw: %.3f", wps, (double)wps * ITER * UNR * inst_per_op / (double)mx);
    }
    printf("\n");
}

int main() {
    float* d; long long* dc;
    cudaMalloc(&d, 148 * 1024 * sizeof(float));
    cudaMalloc(&dc, 148 * sizeof(long long));
    printf("warp-instructions issued per cycle per SMSP at 1/2/4/8 warps per SMSP (%d independent ops per warp)\n", UNR);
    run<0>("FFMA", 1, d, dc);
    run<1>("FFMA2 (per packed instr)", 1, d, dc);
    run<2>("FADD", 1, d, dc);
    run<3>("FADD2", 1, d, dc);
    run<4>("FMUL", 1, d, dc);
    run<5>("FMUL2", 1, d, dc);
    run<6>("LOP3+IMAD", 2, d, dc);
    run<7>("LOP3+I2F", 2, d, dc);
    run<8>("FFMA2+LOP3+IADD3", 3, d, dc);
    run<9>("FFMA+LOP3+IADD3", 3, d, dc);
    run<10>("FSETP+@P FADD", 2, d, dc);
    run<11>("FFMA2+FFMA", 2, d, dc);
    printf("frame mixes: 'ipc' here = frame-equivalents per cycle per SMSP x 100 (higher is better)\n");
    run<12>("17 scalar FP", 100, d, dc);
    run<13>("9 scalar + 4 packed", 100, d, dc);
    run<14>("9 sc + 4 pk + IMAD,2LOP,I2F", 100, d, dc);
    run<15>("17 sc + IMAD,2LOP,I2F", 100, d, dc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
