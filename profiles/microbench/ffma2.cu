// Microbenchmark: issue rate of packed fp32 (FFMA2/FADD2/FMUL2, PTX *.f32x2) vs scalar FFMA on B200 (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int PACKED>
__global__ void k(float* out, float s, int iters)
{
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    u64 p0 = pack2(a0, a1), p1 = pack2(a2, a3), p2 = pack2(a4, a5), p3 = pack2(a6, a7), p4 = pack2(a1, a3), p5 = pack2(a5, a7),
        p6 = pack2(a0, a2), p7 = pack2(a4, a6);
    u64 ss = pack2(s, s), cc = pack2(0.5f, 0.25f);
    for (int i = 0; i < iters; i++) {
        if (PACKED) {
            p0 = fma2(p0, ss, cc); p1 = fma2(p1, ss, cc); p2 = fma2(p2, ss, cc); p3 = fma2(p3, ss, cc);
            p4 = fma2(p4, ss, cc); p5 = fma2(p5, ss, cc); p6 = fma2(p6, ss, cc); p7 = fma2(p7, ss, cc);
        } else {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a4) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a5) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a6) : "f"(s), "f"(0.5f));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a7) : "f"(s), "f"(0.5f));
        }
    }
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7, x, y;
    unpack2(p0, x, y); r += x + y; unpack2(p1, x, y); r += x + y; unpack2(p2, x, y); r += x + y; unpack2(p3, x, y); r += x + y;
    unpack2(p4, x, y); r += x + y; unpack2(p5, x, y); r += x + y; unpack2(p6, x, y); r += x + y; unpack2(p7, x, y); r += x + y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16;
    for (int packed = 0; packed < 2; packed++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (packed) k<1><<<148 * 8, 256>>>(out, 0.999f, iters); else k<0><<<148 * 8, 256>>>(out, 0.999f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double inst = (double)148 * 8 * 8 * iters * 8;          // warp instructions
            if (rep) printf("%s: %.3f ms, %.2f warp-inst/clk/SM @1.965GHz, %.1f TFLOP/s\n", packed ? "FFMA2 (f32x2)" : "FFMA  (f32)  ", ms,
                            inst / 148 / (ms * 1e-3 * 1.965e9), inst * 32 * (packed ? 4 : 2) / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
