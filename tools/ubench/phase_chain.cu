// Latency of one step of the exact phase recurrence, phase' = fmod(phase + d, 1) for phase, d in [0, 1),
// in five bit-identical formulations.  One warp, one block: pure dependent-chain latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o phase_chain phase_chain.cu && ./phase_chain
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void chain(float ph, float d, int steps, float* out, long long* cycles) {
    ph += threadIdx.x * 1e-3f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < steps; j += 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float t = __fadd_rn(ph, d);
            if (V == 0) {                                   // predicated subtract
                ph = t >= 1.0f ? __fadd_rn(t, -1.0f) : t;
            } else if (V == 1) {                            // both candidates, then select
                const float w = __fadd_rn(t, -1.0f);
                asm("{ .reg .pred p; setp.ge.f32 p, %1, 0f3F800000; selp.f32 %0, %2, %1, p; }" : "=f"(ph) : "f"(t), "f"(w));
            } else if (V == 2) {                            // select on the sign of the wrapped candidate
                const float w = __fadd_rn(t, -1.0f);
                asm("{ .reg .pred p; setp.lt.s32 p, %2, 0; selp.f32 %0, %1, %3, p; }"
                    : "=f"(ph) : "f"(t), "r"(__float_as_int(w)), "f"(w));
            } else if (V == 3) {                            // integer compare of t, in parallel with the subtract
                const float w = __fadd_rn(t, -1.0f);
                asm("{ .reg .pred p; setp.ge.s32 p, %2, 0x3F800000; selp.f32 %0, %3, %1, p; }"
                    : "=f"(ph) : "f"(t), "r"(__float_as_int(t)), "f"(w));
            } else {                                        // unsigned min of the bit patterns: w < 0 has the sign bit
                const float w = __fadd_rn(t, -1.0f);
                ph = __uint_as_float(min(__float_as_uint(w), __float_as_uint(t)));
            }
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = ph;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    const int steps = 1 << 20;
    float res[5][32];
    for (int v = 0; v < 5; v++) {
        long long c = 0;
        for (int rep = 0; rep < 2; rep++) {
            if (v == 0) chain<0><<<1, 32>>>(0.25f, 0.0123457f, steps, out, cyc);
            if (v == 1) chain<1><<<1, 32>>>(0.25f, 0.0123457f, steps, out, cyc);
            if (v == 2) chain<2><<<1, 32>>>(0.25f, 0.0123457f, steps, out, cyc);
            if (v == 3) chain<3><<<1, 32>>>(0.25f, 0.0123457f, steps, out, cyc);
            if (v == 4) chain<4><<<1, 32>>>(0.25f, 0.0123457f, steps, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(res[v], out, 128, cudaMemcpyDeviceToHost);
        printf("variant %d: %.2f cycles per step\n", v, (double)c / steps);
    }
    bool same = true;
    for (int i = 0; i < 32; i++) for (int v = 1; v < 5; v++) same &= res[0][i] == res[v][i];
    printf("results identical: %s\n", same ? "yes" : "NO");
    return 0;
}
