// Micro-benchmark: issue/pipe throughput of unfused multiply+add on sm_100a —
// scalar FMUL+FADD against packed FFMA2 pairs (mul as fma(x,w,-0), add as fma(r,1,y), with the
// constants opaque to ptxas so it cannot contract them into one FFMA2), and plain FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int T = 16;   // taps per iteration
constexpr int NACC = 4; // float4 accumulators per thread (independent chains)

template <int MODE>
__global__ void __launch_bounds__(256) k(const float4* in, float4* out, const float* w, int iters, float one, float nzero) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    float4 v[NACC];
    for (int a = 0; a < NACC; ++a) v[a] = in[gid * NACC + a];
    float wt[T];
    for (int t = 0; t < T; ++t) wt[t] = w[t];
    float4 acc[NACC];
    for (int a = 0; a < NACC; ++a) acc[a] = make_float4(0, 0, 0, 0);
    const u64 one2 = pk(one, one), nz2 = pk(nzero, nzero);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
                if (MODE == 0) {  // scalar unfused (what -fmad=false emits today)
                    acc[a].x += wt[t] * v[a].x; acc[a].y += wt[t] * v[a].y; acc[a].z += wt[t] * v[a].z; acc[a].w += wt[t] * v[a].w;
                } else if (MODE == 1) {  // scalar fused
                    acc[a].x = __fmaf_rn(wt[t], v[a].x, acc[a].x); acc[a].y = __fmaf_rn(wt[t], v[a].y, acc[a].y);
                    acc[a].z = __fmaf_rn(wt[t], v[a].z, acc[a].z); acc[a].w = __fmaf_rn(wt[t], v[a].w, acc[a].w);
                } else {  // packed, unfused semantics
                    const u64 w2 = pk(wt[t], wt[t]);
                    u64 lo = pk(acc[a].x, acc[a].y), hi = pk(acc[a].z, acc[a].w);
                    lo = fma2(fma2(pk(v[a].x, v[a].y), w2, nz2), one2, lo);
                    hi = fma2(fma2(pk(v[a].z, v[a].w), w2, nz2), one2, hi);
                    upk(lo, acc[a].x, acc[a].y); upk(hi, acc[a].z, acc[a].w);
                }
            }
        }
        for (int a = 0; a < NACC; ++a) { v[a].x += 1e-9f; v[a].y += 1e-9f; v[a].z += 1e-9f; v[a].w += 1e-9f; }
    }
    for (int a = 0; a < NACC; ++a) out[gid * NACC + a] = acc[a];
}

template <int MODE>
static void run(const char* name, float4* in, float4* out, float* w, int blocks) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(in, out, w, 10, 1.0f, -0.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(in, out, w, iters, 1.0f, -0.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)blocks * 256 * NACC * 4 * T * iters;
    float4 h; cudaMemcpy(&h, out + 5, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-22s %8.3f ms  %8.2f G mul-add/s   (out %.9g)\n", name, ms, macs / ms * 1e-6, h.x);
}

int main() {
    const int blocks = 148 * 8;
    float4 *in, *out; float* w;
    cudaMalloc(&in, (size_t)blocks * 256 * NACC * 16); cudaMalloc(&out, (size_t)blocks * 256 * NACC * 16); cudaMalloc(&w, T * 4);
    float hw[T]; for (int t = 0; t < T; ++t) hw[t] = 0.01f + 0.003f * t;
    cudaMemcpy(w, hw, sizeof hw, cudaMemcpyHostToDevice);
    float4* hin = new float4[(size_t)blocks * 256 * NACC];
    for (size_t i = 0; i < (size_t)blocks * 256 * NACC; ++i) hin[i] = make_float4(0.1f + 1e-6f * (i % 977), 0.2f, 0.3f, 0.4f);
    cudaMemcpy(in, hin, (size_t)blocks * 256 * NACC * 16, cudaMemcpyHostToDevice);
    run<0>("scalar FMUL+FADD", in, out, w, blocks);
    run<1>("scalar FFMA (fused)", in, out, w, blocks);
    run<2>("packed 2xFFMA2 unfused", in, out, w, blocks);
    return 0;
}
