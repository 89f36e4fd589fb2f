// fuse_check.cu — shows that ptxas 12.9 (sm_100a) contracts packed mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// (kernels k1: inline PTX, k2: builtins) while scalar __fmul_rn + __fadd_rn (k3) stays FMUL + FADD:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o fuse.cubin fuse_check.cu && cuobjdump -sass fuse.cubin | grep -E "Function|FFMA|FMUL|FADD"
#include <cuda_runtime.h>
__device__ __forceinline__ float2 padd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 pmul2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__global__ void k1(const float2* a, const float2* b, const float2* c, float2* d) { d[threadIdx.x] = padd2(pmul2(a[threadIdx.x], b[threadIdx.x]), c[threadIdx.x]); }
__global__ void k2(const float2* a, const float2* b, const float2* c, float2* d) { d[threadIdx.x] = __fadd2_rn(__fmul2_rn(a[threadIdx.x], b[threadIdx.x]), c[threadIdx.x]); }
__global__ void k3(const float* a, const float* b, const float* c, float* d) { d[threadIdx.x] = __fadd_rn(__fmul_rn(a[threadIdx.x], b[threadIdx.x]), c[threadIdx.x]); }
