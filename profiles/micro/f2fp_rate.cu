// Issue-rate microbenchmark (sm_100a): cycles per warp instruction per SM sub-partition (SMSP) for the conversion /
// packing instructions the tensor-core epilogues are made of.  Four independent loop-carried chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f2fp_rate f2fp_rate.cu && ./f2fp_rate
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 8192
#define F(x) __uint_as_float(x)
template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b) {
    unsigned r;
    if (OP == 0) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(F(a)), "f"(F(b)));
    else if (OP == 1) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(F(a)), "f"(F(b)));
    else if (OP == 2) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 3) asm volatile("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else if (OP == 4) { float f; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"((unsigned short)a)); r = __float_as_uint(f) ^ b; }
    else if (OP == 5) { float f; asm volatile("max.f32 %0, %1, %2;" : "=f"(f) : "f"(F(a)), "f"(F(b))); r = __float_as_uint(f); }
    else if (OP == 6) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    else { float f; asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=f"(f) : "f"(F(a)), "f"(F(b))); r = __float_as_uint(f); }
    return r;
}
template <int OP>
__global__ void k(unsigned* out, long long* cyc, unsigned seed) {
    unsigned r0 = seed + threadIdx.x, r1 = r0 * 3u + 0x3f800000u, r2 = r0 * 5u + 0x3f000000u, r3 = r0 * 7u + 0x40000000u;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            unsigned n0 = op<OP>(r0, r1), n1 = op<OP>(r1, r2), n2 = op<OP>(r2, r3), n3 = op<OP>(r3, r0);
            r0 = n0; r1 = n1; r2 = n2; r3 = n3;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = r0 ^ r1 ^ r2 ^ r3;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[OP] = t1 - t0;
}
int main() {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 128);
    const char* names[] = {"cvt.rn.relu.bf16x2.f32 (F2FP.RELU)", "cvt.rn.f16x2.f32 (F2FP)", "prmt.b32", "max.bf16x2 (HMNMX2)", "cvt.f32.f16 (+ xor)",
                           "max.f32 (FMNMX)", "add.u32 (IADD3)", "fma.rn.f32 (FFMA)"};
    for (int warps = 4; warps <= 16; warps *= 2) {
        k<0><<<148, warps * 32>>>(out, cyc, 1u); k<1><<<148, warps * 32>>>(out, cyc, 1u); k<2><<<148, warps * 32>>>(out, cyc, 1u);
        k<3><<<148, warps * 32>>>(out, cyc, 1u); k<4><<<148, warps * 32>>>(out, cyc, 1u); k<5><<<148, warps * 32>>>(out, cyc, 1u);
        k<6><<<148, warps * 32>>>(out, cyc, 1u); k<7><<<148, warps * 32>>>(out, cyc, 1u);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        for (int o = 0; o < 8; ++o)
            printf("%2d warps/SM (%d per SMSP): %-38s %.2f cycles per warp instruction per SMSP\n", warps, warps / 4, names[o],
                   (double)h[o] / (ITERS * 16.0 * (warps / 4)));
    }
    return 0;
}
