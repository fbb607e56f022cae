// Micro-benchmark: what HBM rate does a 2-read / 2-write float4 stream reach on this GPU
// (a) flat elementwise, (b) marched along z like blur_march_kernel (thread = float4 column, plane stride)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream4_bench stream4_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef long long ll;
__global__ void __launch_bounds__(256) flat(const float4* a, const float4* b, float4* c, float4* d, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 x = a[i], y = b[i];
        c[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        d[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    }
}
template <int U, int TPB>
__global__ void __launch_bounds__(TPB) march(const float4* a, const float4* b, float4* c, float4* d, int plane4, int nz, int seg) {
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / plane4, col = gid - s * plane4;
    const int p0 = s * seg, p1 = min(nz, p0 + seg);
    for (int p = p0; p < p1; p += U) {
        float4 x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { x[u] = a[(size_t)(p + u) * plane4 + col]; y[u] = b[(size_t)(p + u) * plane4 + col]; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            c[(size_t)(p + u) * plane4 + col] = make_float4(x[u].x + y[u].x, x[u].y + y[u].y, x[u].z + y[u].z, x[u].w + y[u].w);
            d[(size_t)(p + u) * plane4 + col] = make_float4(x[u].x - y[u].x, x[u].y - y[u].y, x[u].z - y[u].z, x[u].w - y[u].w);
        }
    }
}
// prefetch-ring variant with limited occupancy (dynamic shared memory pads the CTA): closer to blur_march_kernel
template <int PF>
__global__ void __launch_bounds__(128) march_pf(const float4* a, const float4* b, float4* c, float4* d, int plane4, int nz, int seg, int halo) {
    extern __shared__ float4 pad[];
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / plane4, col = gid - s * plane4;
    const int p0 = s * seg, p1 = min(nz, p0 + seg);
    float4 ra[PF], rb[PF];
    float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = max(0, p0 - halo); q < p0; ++q) { float4 t = a[(size_t)q * plane4 + col]; h.x += t.x; h.y += t.y; h.z += t.z; h.w += t.w; }
    for (int q = p1; q < min(nz, p1 + halo); ++q) { float4 t = a[(size_t)q * plane4 + col]; h.x += t.x; h.y += t.y; h.z += t.z; h.w += t.w; }
#pragma unroll
    for (int u = 0; u < PF; ++u) { ra[u] = a[(size_t)min(p0 + u, nz - 1) * plane4 + col]; rb[u] = b[(size_t)min(p0 + u, nz - 1) * plane4 + col]; }
    for (int p = p0; p < p1; p += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const float4 x = ra[u], y = rb[u];
            const int q = min(p + u + PF, nz - 1);
            ra[u] = a[(size_t)q * plane4 + col]; rb[u] = b[(size_t)q * plane4 + col];
            c[(size_t)(p + u) * plane4 + col] = make_float4(x.x + y.x + h.x, x.y + y.y + h.y, x.z + y.z + h.z, x.w + y.w + h.w);
            d[(size_t)(p + u) * plane4 + col] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
        }
    }
    if (gid == 0xffffffffu) pad[0] = h;
}
// cp.async variant: per-thread private ring of D slots per input in shared memory (no CTA barrier; the thread that
// issued the copy is the one that reads it), depth tracked by commit groups instead of the 6 register scoreboards
template <int D>
__global__ void __launch_bounds__(128) march_cp(const float4* a, const float4* b, float4* c, float4* d, int plane4, int nz, int seg) {
    extern __shared__ float4 ring[];  // [D][2][128]
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = gid / plane4, col = gid - s * plane4;
    const int p0 = s * seg, p1 = min(nz, p0 + seg);
    auto issue = [&](int p) {
        if (p < p1) {
            const int slot = (p - p0) % D;
            unsigned sa = (unsigned)__cvta_generic_to_shared(&ring[(slot * 2 + 0) * 128 + threadIdx.x]);
            unsigned sb = (unsigned)__cvta_generic_to_shared(&ring[(slot * 2 + 1) * 128 + threadIdx.x]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(a + (size_t)p * plane4 + col) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sb), "l"(b + (size_t)p * plane4 + col) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int u = 0; u < D - 1; ++u) issue(p0 + u);
    for (int p = p0; p < p1; ++p) {
        issue(p + D - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        const int slot = (p - p0) % D;
        const float4 x = ring[(slot * 2 + 0) * 128 + threadIdx.x], y = ring[(slot * 2 + 1) * 128 + threadIdx.x];
        c[(size_t)p * plane4 + col] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        d[(size_t)p * plane4 + col] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    }
}
template <typename F>
static void timeit(const char* name, F f, double bytes) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    printf("%-34s %8.1f us  %7.1f GB/s  (%s)\n", name, best * 1e3, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int n = 512; const size_t N = (size_t)n * n * n, n4 = N / 4; const int plane4 = n * n / 4;
    float4 *a, *b, *c, *d;
    cudaMalloc(&a, N * 4); cudaMalloc(&b, N * 4); cudaMalloc(&c, N * 4); cudaMalloc(&d, N * 4);
    cudaMemset(a, 0, N * 4); cudaMemset(b, 0, N * 4);
    const double bytes = 16.0 * N;
    timeit("flat 2r2w, 256 thr", [&] { flat<<<(unsigned)((n4 + 255) / 256), 256>>>(a, b, c, d, n4); }, bytes);
    timeit("memcpy d2d 1r1w (512 MiB)", [&] { cudaMemcpyAsync(c, a, N * 4, cudaMemcpyDeviceToDevice); }, 8.0 * N);
    for (int seg : {512, 128, 64, 32}) {
        const int nseg = n / seg; const unsigned thr = (unsigned)plane4 * nseg;
        char nm[64];
        snprintf(nm, 64, "march seg=%d U=1 128thr", seg); timeit(nm, [&] { march<1, 128><<<thr / 128, 128>>>(a, b, c, d, plane4, n, seg); }, bytes);
        snprintf(nm, 64, "march seg=%d U=2 128thr", seg); timeit(nm, [&] { march<2, 128><<<thr / 128, 128>>>(a, b, c, d, plane4, n, seg); }, bytes);
        snprintf(nm, 64, "march seg=%d U=4 128thr", seg); timeit(nm, [&] { march<4, 128><<<thr / 128, 128>>>(a, b, c, d, plane4, n, seg); }, bytes);
        snprintf(nm, 64, "march seg=%d U=4 256thr", seg); timeit(nm, [&] { march<4, 256><<<thr / 256, 256>>>(a, b, c, d, plane4, n, seg); }, bytes);
    }
    cudaFuncSetAttribute(march_pf<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int occ : {16, 7, 5, 4}) {
        const size_t smem = occ >= 16 ? 0 : (size_t)(220 * 1024 / occ - 1024);
        for (int halo : {0, 4}) {
            const int seg = 64, nseg = n / seg; const unsigned thr = (unsigned)plane4 * nseg;
            char nm[64];
            snprintf(nm, 64, "march_pf4 seg=64 occ=%d halo=%d", occ, halo);
            timeit(nm, [&] { march_pf<4><<<thr / 128, 128, smem>>>(a, b, c, d, plane4, n, seg, halo); }, bytes);
        }
    }
    cudaFuncSetAttribute(march_cp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(march_cp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(march_cp<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int occ : {6, 5, 4, 3}) {
        const int seg = 64, nseg = n / seg; const unsigned thr = (unsigned)plane4 * nseg;
        const size_t smem = (size_t)(220 * 1024 / occ - 1024);
        if (smem < 8 * 2 * 128 * 16) continue;
        char nm[64];
        snprintf(nm, 64, "march_cp D=4 seg=64 occ=%d", occ); timeit(nm, [&] { march_cp<4><<<thr / 128, 128, smem>>>(a, b, c, d, plane4, n, seg); }, bytes);
        snprintf(nm, 64, "march_cp D=8 seg=64 occ=%d", occ); timeit(nm, [&] { march_cp<8><<<thr / 128, 128, smem>>>(a, b, c, d, plane4, n, seg); }, bytes);
        if (smem >= 12 * 2 * 128 * 16) { snprintf(nm, 64, "march_cp D=12 seg=64 occ=%d", occ); timeit(nm, [&] { march_cp<12><<<thr / 128, 128, smem>>>(a, b, c, d, plane4, n, seg); }, bytes); }
    }
    return 0;
}
