// s3d_kernels.cuh — device code of the extraction path (dense pyramid + sparse keypoint stages).
//
// COMPILED WITH -fmad=false.  The dense stages are bit-exact against the reference, which is
// built for baseline x86-64 where `acc += w * v` is an IEEE multiply followed by an IEEE add.
// With FMA contraction off, IEEE division/sqrt (nvcc defaults) and the same operation order,
// FP32 expressions below round exactly like the reference's.  s3d_selftest() verifies the flags.
//
// Layout: every volume is float32, x fastest (xs=1, ys=nx, zs=nx*ny), as in
// /root/reference/3DSIFT/Src/Util/cTexImage.cc:28-30.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <type_traits>

#include "expf_ref.h"
#include "sift3d_b200.h"

namespace s3d {

constexpr int kMaxHW = 16;   // taps up to 33 (sigma <= 5.33); defaults use hw in {2,3,4,5,6,8}
constexpr int kMaxOct = 16;
constexpr int kMaxG = 12;    // num_kp_levels + 3 <= 12

struct Taps {
    int hw;
    float w[2 * kMaxHW + 1];
    // extended-line table for the axis this pass runs along (length n): sample q = n-1+e,
    // e = 0..hw, is (1-frac)*in[il] + frac*in[il+1]  (filled on the host by fill_ext)
    int ext_il[kMaxHW + 1];
    float ext_frac[kMaxHW + 1];
};

typedef long long ll;

__device__ __forceinline__ float ld_clamped(const float* __restrict__ p, ll i, ll total) {
    i = i < 0 ? 0 : (i >= total ? total - 1 : i);
    return p[i];
}

// One output of the reference's boundary sweep (GaussianSmooth_3D_Imp second pass,
// Src/cSIFT3D.cc:722-788): mirror about 0 without repeating the edge (:747-750), right edge
// 2*(n-1)-c-0.1 (:751-755), (int) truncation and linear blend of in[lo], in[lo+1] (:757-764).
// line0 = flat index of the element with coordinate 0 on this line, st = stride along the axis.
// Reads are flat and unchecked in the reference (an out-of-row index lands in the neighbouring
// row); indices outside the whole buffer are clamped to its ends.
__device__ __forceinline__ float blur_boundary_one(const float* __restrict__ buf, ll line0, ll st, int n, int p,
                                                   const Taps& t, ll total) {
    float acc = 0.0f;
    const int dim_end = n - 1;
    for (int d = -t.hw; d <= t.hw; ++d) {
        float c = (float)p - (float)d;
        if (c < 0)
            c = -1 * c;
        else if (c >= dim_end)
            c = (float)(2 * dim_end) - c - 0.1f;
        int il = (int)c;
        float frac = c - (float)il;
        float lo = ld_clamped(buf, line0 + (ll)il * st, total);
        float hi = ld_clamped(buf, line0 + (ll)(il + 1) * st, total);
        acc += t.w[d + t.hw] * ((1.0f - frac) * lo + frac * hi);
    }
    return acc;
}

__device__ __forceinline__ bool blur_is_interior(int p, int n, int hw) { return p >= hw && p <= n - hw - 2; }

// The sample the reference uses for tap coordinate c = q (an integer) of a BOUNDARY output
// (Src/cSIFT3D.cc:747-764), as a function of q alone: q < 0 mirrors to in[-q] (frac = 0, so
// 1.0f*in[-q] + 0.0f*in[-q+1] == in[-q]); q >= n-1 is the 0.1/0.9-style blend at
// c' = 2(n-1) - q - 0.1f; otherwise in[q].  Interior outputs only ever touch 0 <= q <= n-2, where
// this is in[q] as well, so with n >= 2*hw+2 EVERY output of a line is the plain ordered
// correlation over this extended line — no divergent boundary path in the fast kernels.
// (il, frac) for q >= n-1 come from Taps::ext_il / ext_frac, computed on the host with the same
// FP32 operations.
__device__ __forceinline__ float ext_sample1(const float* __restrict__ line, int n, int q, const Taps& t) {
    if (q < 0) return line[-q];
    if (q < n - 1) return line[q];
    const int e = q - (n - 1);
    const int il = t.ext_il[e];
    const float frac = t.ext_frac[e];
    return (1.0f - frac) * line[il] + frac * line[il + 1];
}

__device__ __forceinline__ float4 ext_sample4(const float* __restrict__ col, ll st, int n, int q, const Taps& t) {
    if (q < n - 1) return *reinterpret_cast<const float4*>(col + (ll)(q < 0 ? -q : q) * st);
    const int e = q - (n - 1);
    const int il = t.ext_il[e];
    const float frac = t.ext_frac[e];
    const float4 lo = *reinterpret_cast<const float4*>(col + (ll)il * st);
    const float4 hi = *reinterpret_cast<const float4*>(col + (ll)(il + 1) * st);
    float4 r;
    r.x = (1.0f - frac) * lo.x + frac * hi.x; r.y = (1.0f - frac) * lo.y + frac * hi.y;
    r.z = (1.0f - frac) * lo.z + frac * hi.z; r.w = (1.0f - frac) * lo.w + frac * hi.w;
    return r;
}

__device__ __forceinline__ void atomic_max_abs(unsigned* slot, float m) {
    // non-negative floats order like their bit patterns
    atomicMax(slot, __float_as_uint(m));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// K1  max|v| (im_max_abs / data_scale first sweep, Src/cUtil.cc:538-550,587-605) and the IEEE
//     division sweep (Src/cUtil.cc:552-561).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxabs_kernel(const float* __restrict__ src, size_t n, unsigned* slot) {
    float m = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // a shard's owned planes may start at any element: peel to 16-byte alignment
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15) >> 2;
    head = head < n ? head : n;
    if (gid < head) m = fabsf(src[gid]);
    const float* body = src + head;
    const size_t nb = n - head, n4 = nb >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(body);
    for (size_t i = gid; i < n4; i += stride) {
        float4 v = s4[i];
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = (n4 << 2) + gid; i < nb; i += stride) m = fmaxf(m, fabsf(body[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomic_max_abs(slot, m);
}

__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n,
                                                        const unsigned* __restrict__ slot) {
    const float mx = __uint_as_float(*slot);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = s4[i];
        v.x = v.x / mx; v.y = v.y / mx; v.z = v.z / mx; v.w = v.w / mx;
        d4[i] = v;
    }
    for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i] / mx;
}

// ---------------------------------------------------------------------------------------------
// K2  separable Gaussian pass.  GaussianSmooth_3D_Imp, Src/cSIFT3D.cc:624-790.
//     Per output: taps d = -hw..+hw ascending, acc = acc + w[d+hw] * in[p-d] (interior samples
//     are 1.0f*in[c] + 0.0f*in[c+1] == in[c]); boundary outputs use blur_boundary_one.
//     DOG variants additionally write dog = (cur - prev) * (-1) (Sub, Src/cSIFT3D.cc:875) and
//     fold max|dog| into *maxslot (im_max_abs for Detect_KeyPoints, Src/cSIFT3D.cc:384).
// ---------------------------------------------------------------------------------------------

// Generic: one thread per output, any dims, any hw <= kMaxHW, any axis.
__global__ void __launch_bounds__(256) blur_generic_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx,
                                                           int ny, int nz, int axis, Taps t,
                                                           const float* __restrict__ prev, float* __restrict__ dog,
                                                           unsigned* maxslot, int gn, int goff, int olo, int ohi) {
    // gn > 0: the buffer holds planes [goff, goff+nz) of a z line of gn planes (a z-slab shard whose
    // local extent is not the whole line).  The boundary rule is then evaluated in GLOBAL coordinates
    // (its FP32 blend fraction depends on the magnitude of the coordinate) and samples are fetched
    // from the local planes (clamped: outputs whose taps would leave the buffer are not requested).
    // Only the outputs on global planes [olo, ohi) are produced.
    const ll total = (ll)nx * ny * nz;
    const ll idx = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    const int zq = (int)(idx / ((ll)nx * ny));
    if (idx < total && !(gn > 0 && axis == 2 && (zq + goff < olo || zq + goff >= ohi))) {
        const int x = (int)(idx % nx), y = (int)((idx / nx) % ny), z = zq;
        const int p = axis == 0 ? x : (axis == 1 ? y : z);
        const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
        const ll st = axis == 0 ? 1 : (axis == 1 ? (ll)nx : (ll)nx * ny);
        const ll line0 = idx - (ll)p * st;
        float acc;
        if (gn > 0 && axis == 2) {
            acc = 0.0f;
            const int dim_end = gn - 1;
            for (int d = -t.hw; d <= t.hw; ++d) {
                float c = (float)(p + goff) - (float)d;
                if (c < 0)
                    c = -1 * c;
                else if (c >= dim_end)
                    c = (float)(2 * dim_end) - c - 0.1f;
                const int il = (int)c;
                const float frac = c - (float)il;
                const int l0 = min(max(il - goff, 0), nz - 1), l1 = min(max(il + 1 - goff, 0), nz - 1);
                acc += t.w[d + t.hw] * ((1.0f - frac) * src[line0 + (ll)l0 * st] + frac * src[line0 + (ll)l1 * st]);
            }
        } else if (blur_is_interior(p, n, t.hw)) {
            acc = 0.0f;
            for (int d = -t.hw; d <= t.hw; ++d) acc += t.w[d + t.hw] * src[line0 + (ll)(p - d) * st];
        } else {
            acc = blur_boundary_one(src, line0, st, n, p, t, total);
        }
        dst[idx] = acc;
        if (dog) {
            float dv = (acc - prev[idx]) * (-1);
            dog[idx] = dv;
            m = fabsf(dv);
        }
    }
    if (dog) {
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) atomic_max_abs(maxslot, m);
    }
}

// X pass, 4 outputs per thread (float4 store).  Requires nx % 4 == 0 and nx >= 2*HW+2.
template <int HW>
__global__ void __launch_bounds__(256) blur_x_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx,
                                                     unsigned nthreads, Taps t) {
    constexpr int PAD = (HW + 3) / 4 * 4;
    constexpr int NV = (2 * PAD + 4) / 4;
    const unsigned nx4 = (unsigned)nx >> 2;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nthreads) return;
    const unsigned row = gid / nx4;
    const int x0 = (int)(gid - row * nx4) * 4;
    const float* r = src + (size_t)row * nx;
    float v[2 * PAD + 4];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const int xx = x0 - PAD + 4 * q;
        if (xx >= 0 && xx + 3 <= nx - 2) {
            const float4 f = *reinterpret_cast<const float4*>(r + xx);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        } else {
            // row ends: mirrored / blended samples of the extended line (only the edge threads);
            // positions farther than HW from the row are never used by any tap
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int qq = xx + e;
                v[4 * q + e] = (qq >= -HW && qq <= nx - 1 + HW) ? ext_sample1(r, nx, qq, t) : 0.0f;
            }
        }
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k <= 2 * HW; ++k) acc += t.w[k] * v[PAD + j + HW - k];
        o[j] = acc;
    }
    *reinterpret_cast<float4*>(dst + (size_t)row * nx + x0) = make_float4(o[0], o[1], o[2], o[3]);
}

// Y / Z pass: each thread owns a float4 column (4 consecutive x) and marches along the axis over
// a segment, keeping the 2*HW+1 most recent inputs in a register ring (static indices through a
// (2*HW+1)-way unrolled loop), so every input is loaded once per segment.  Line ends use the
// extended-line samples (ext_sample4), so there is no separate boundary path.
// Requires nx % 4 == 0 and n >= 2*HW+2.  n = length of the marched axis, st = its stride, n_other / st_other =
// the remaining non-x axis.
// PFX = extra prefetch distance of the marched samples, PP = prefetch distance of the DoG pass's `prev`
// operand (its loads are issued PP steps before their use; PP = 1 left the Z pass waiting on every one:
// ncu long_scoreboard = 20 warp-cycles per issue at 63 % of DRAM throughput).
template <int HW, bool DOG, int PFX = 0, int PP = 1>
__global__ void __launch_bounds__(128) blur_march_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx,
                                                         int n, ll st, int n_other, ll st_other, int seg, Taps t,
                                                         const float* __restrict__ prev, float* __restrict__ dog,
                                                         unsigned* maxslot, int zlo, int zhi) {
    // [zlo, zhi): the output positions along the marched axis (the whole line: 0, n).  A z-slab shard passes pointers
    // offset to the line's virtual origin, the GLOBAL line length n and its own sub-range, so the boundary rule is
    // evaluated in global coordinates and interior taps read the shard's halo planes.
    constexpr int PF = (HW >= 5 ? 2 : 4) + PFX;  // loads are issued PF steps ahead of their first use
    constexpr int WR = 2 * HW + 1 + PF;       // ring size == unroll factor (static slot indices)
    const unsigned nx4 = (unsigned)nx >> 2;
    const int nseg = (zhi - zlo + seg - 1) / seg;
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (gid < nx4 * (unsigned)n_other * (unsigned)nseg) {
        const unsigned tt = gid / nx4;
        const int x4 = (int)(gid - tt * nx4);
        const int s = (int)(tt / (unsigned)n_other);
        const int other = (int)(tt - (unsigned)s * (unsigned)n_other);
        const int p0 = zlo + s * seg;
        const int p1 = min(zhi, p0 + seg);
        const int qmax = p1 - 1 + HW;  // last extended-line sample this segment needs (<= n-1+HW)
        const ll line0 = (ll)x4 * 4 + (ll)other * st_other;  // flat index of coordinate 0 on this line
        const float* col = src + line0;
        float4 win[WR];
        // prologue: samples p0-HW .. p0+HW+PF-1  (relative r = 0 .. 2HW+PF-1)
#pragma unroll
        for (int r = 0; r < 2 * HW + PF; ++r) {
            const int q = p0 - HW + r;
            win[r] = (q <= qmax) ? ext_sample4(col, st, n, q, t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        win[2 * HW + PF] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 pvq[PP];  // prev samples p .. p+PP-1
#pragma unroll
        for (int k = 0; k < PP; ++k) {
            pvq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (DOG && p0 + k < p1) pvq[k] = *reinterpret_cast<const float4*>(prev + line0 + (ll)(p0 + k) * st);
        }
        for (int ib = 0; p0 + ib < p1; ib += WR) {
#pragma unroll
            for (int j = 0; j < WR; ++j) {
                const int p = p0 + ib + j;
                if (p < p1) {
                    // prefetch sample p+HW+PF into the slot of the oldest one (relative r = ib+j+2HW+PF)
                    const int q = p + HW + PF;
                    if (q <= qmax) win[(j + 2 * HW + PF) % WR] = ext_sample4(col, st, n, q, t);
                    float4 pvn = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (DOG && p + PP < p1) pvn = *reinterpret_cast<const float4*>(prev + line0 + (ll)(p + PP) * st);
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k <= 2 * HW; ++k) {
                        const float4 v = win[(j + 2 * HW - k) % WR];  // extended-line sample p+HW-k
                        const float w = t.w[k];
                        acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
                    }
                    const ll oidx = line0 + (ll)p * st;
                    *reinterpret_cast<float4*>(dst + oidx) = acc;
                    if (DOG) {
                        float4 dv;
                        const float4 pv = pvq[0];
                        dv.x = (acc.x - pv.x) * (-1); dv.y = (acc.y - pv.y) * (-1);
                        dv.z = (acc.z - pv.z) * (-1); dv.w = (acc.w - pv.w) * (-1);
                        *reinterpret_cast<float4*>(dog + oidx) = dv;
                        m = fmaxf(m, fmaxf(fmaxf(fabsf(dv.x), fabsf(dv.y)), fmaxf(fabsf(dv.z), fabsf(dv.w))));
#pragma unroll
                        for (int k = 0; k + 1 < PP; ++k) pvq[k] = pvq[k + 1];
                        pvq[PP - 1] = pvn;
                    }
                }
            }
        }
    }
    if (DOG) {
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) atomic_max_abs(maxslot, m);
    }
}

constexpr int kXYBufs = 4;  // shared-memory row ring of the fused X+Y pass (cp.async runs kXYBufs-1 rows ahead)

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The same march with the loads issued as cp.async copies into a per-thread ring in shared memory.
// blur_march_kernel prefetches into registers, and a warp has only six scoreboards to track its
// outstanding loads: waiting for the oldest one also waits for younger ones that share its scoreboard, so
// the prefetch distance collapses once occupancy is register-limited (scripts/micro/stream4_bench.cu: a
// 2-read/2-write march at 4-7 CTAs/SM moves 5.4-5.6 TB/s with a register ring and 6.4-6.5 TB/s with this
// ring).  cp.async groups are counted, not score-boarded: the copies of step r + D - 1 are issued before
// step r's are consumed, whatever the occupancy.  Every slot is written and read by the same thread, so no
// CTA barrier is needed.  One loop covers prologue and body: step r takes extended-line sample
// q = p0 - HW + r into the register window and, from r = 2*HW on, emits output p = p0 + r - 2*HW.
// Samples beyond the line's end (q >= n-1, the reference's right-edge blend) are loaded directly.
template <int HW, bool DOG>
__global__ void __launch_bounds__(128) blur_marchc_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx,
                                                          int n, ll st, int n_other, ll st_other, int seg, Taps t,
                                                          const float* __restrict__ prev, float* __restrict__ dog,
                                                          unsigned* maxslot, int zlo, int zhi) {
    constexpr int D = 4;             // copies in flight per thread and stream
    constexpr int WR = 2 * HW + 1;   // register window == unroll factor (static slot indices)
    constexpr int NS = DOG ? 2 : 1;
    __shared__ float4 ring[D][NS][128];
    const unsigned nx4 = (unsigned)nx >> 2;
    const int nseg = (zhi - zlo + seg - 1) / seg;  // [zlo, zhi): output positions, see blur_march_kernel
    const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.0f;
    if (gid < nx4 * (unsigned)n_other * (unsigned)nseg) {
        const unsigned tt = gid / nx4;
        const int x4 = (int)(gid - tt * nx4);
        const int s = (int)(tt / (unsigned)n_other);
        const int other = (int)(tt - (unsigned)s * (unsigned)n_other);
        const int p0 = zlo + s * seg;
        const int p1 = min(zhi, p0 + seg);
        const int nsamp = p1 - p0 + 2 * HW;  // samples q = p0-HW .. p1-1+HW
        const ll line0 = (ll)x4 * 4 + (ll)other * st_other;
        const float* col = src + line0;
        const float* pcol = DOG ? prev + line0 : nullptr;
        const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(&ring[0][0][threadIdx.x]);
        constexpr uint32_t kSlotBytes = NS * 128 * 16, kPrevOff = 128 * 16;
        auto issue = [&](int r, uint32_t slot_addr) {
            if (r < nsamp) {
                const int q = p0 - HW + r;
                if (q < n - 1)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot_addr), "l"(col + (ll)(q < 0 ? -q : q) * st) : "memory");
                if (DOG && r >= 2 * HW)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slot_addr + kPrevOff), "l"(pcol + (ll)(q - HW) * st) : "memory");
            }
            cp_async_commit();
        };
#pragma unroll
        for (int r = 0; r < D - 1; ++r) issue(r, ring0 + r * kSlotBytes);
        int slot = 0;  // slot of step r; step r + D - 1 goes into the slot before it
        float4 win[WR];
#pragma unroll
        for (int r = 0; r < WR; ++r) win[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ib = 0; ib < nsamp; ib += WR) {
#pragma unroll
            for (int j = 0; j < WR; ++j) {
                const int r = ib + j;
                if (r < nsamp) {
                    const int pslot = slot == 0 ? D - 1 : slot - 1;
                    issue(r + D - 1, ring0 + pslot * kSlotBytes);
                    cp_async_wait<D - 1>();
                    const int q = p0 - HW + r;
                    float4 sv;
                    if (q < n - 1) {
                        sv = ring[slot][0][threadIdx.x];
                    } else {
                        sv = ext_sample4(col, st, n, q, t);
                    }
                    win[j] = sv;
                    if (r >= 2 * HW) {
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k <= 2 * HW; ++k) {
                            const float4 v = win[(j - k + WR) % WR];  // sample q - k = p + HW - k
                            const float w = t.w[k];
                            acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
                        }
                        const ll oidx = line0 + (ll)(q - HW) * st;
                        *reinterpret_cast<float4*>(dst + oidx) = acc;
                        if (DOG) {
                            const float4 pv = ring[slot][NS - 1][threadIdx.x];
                            float4 dv;
                            dv.x = (acc.x - pv.x) * (-1); dv.y = (acc.y - pv.y) * (-1);
                            dv.z = (acc.z - pv.z) * (-1); dv.w = (acc.w - pv.w) * (-1);
                            *reinterpret_cast<float4*>(dog + oidx) = dv;
                            m = fmaxf(m, fmaxf(fmaxf(fabsf(dv.x), fabsf(dv.y)), fmaxf(fabsf(dv.z), fabsf(dv.w))));
                        }
                    }
                    slot = slot == D - 1 ? 0 : slot + 1;
                }
            }
        }
        cp_async_wait<0>();
    }
    if (DOG) {
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) atomic_max_abs(maxslot, m);
    }
}

// Fused X + Y pass.  The reference blurs along X over the whole volume, then along Y over the result
// (GaussianSmooth_3D Src/cSIFT3D.cc:609-617, through two Im_permute transposes); every Y output only
// needs the 2*HW+1 X-blurred rows around it, so one kernel can produce them on the fly and the
// X-blurred volume is never written to HBM (8 B/voxel instead of 16 for the two passes).
//   * A LINE is one x-row position marching along y over a segment of one z plane: nx/4 threads, each
//     owning a float4 column (4 consecutive x).  A CTA of 128 threads runs 128/(nx/4) lines of
//     consecutive z planes and the same y segment, so control flow is CTA-uniform.
//   * X(r): the line's threads park raw row r in shared memory (double-buffered), the edge threads add
//     the reference's extended-line samples (mirror on the left :747-750, 0.1/0.9-style blend on the
//     right :751-764, same FP32 operations as ext_sample1), and every thread takes its 4 outputs as the
//     ordered correlation over its window — the arithmetic of blur_x_kernel, bit for bit.
//   * The X-blurred float4 goes into a register ring of the last 2*HW+1 rows (static slot indices
//     through a (2*HW+1)-way unrolled loop, as in blur_march_kernel); the Y output is the ordered
//     correlation over the ring.  Rows beyond the ends of the y line are the extended-line samples of
//     the X-BLURRED rows (ext_il/ext_frac of the y axis), as in the reference's second sweep.
// Requires nx % 4 == 0, (nx/4) a divisor of 128 and > HW + 1, nx and ny >= 2*HW+2.
template <int HW>
__global__ void __launch_bounds__(128, 4) blur_xy_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx, int ny,
                                                         int nz, int seg, Taps tx, Taps ty) {
    constexpr int PAD = (HW + 3) / 4 * 4;
    constexpr int NV = (2 * PAD + 4) / 4;
    constexpr int WR = 2 * HW + 1;
    constexpr int NB = kXYBufs;
    extern __shared__ __align__(16) float xy_sm[];
    const int T = nx >> 2, lpc = 128 / T;
    const int line = threadIdx.x / T, x4 = threadIdx.x - line * T;
    const int zgroups = (nz + lpc - 1) / lpc;
    const int s = blockIdx.x / zgroups, zg = blockIdx.x - s * zgroups;
    const int zraw = zg * lpc + line;
    const bool zok = zraw < nz;
    const int z = zok ? zraw : nz - 1;  // surplus lines of the last group shadow the last plane (no stores)
    const int p0 = s * seg, p1 = min(ny, p0 + seg);
    const int rowlen = nx + 2 * PAD;
    float* const bufs = xy_sm + (size_t)line * NB * rowlen;
    const float* const plane = src + (size_t)z * nx * ny;
    float* const oplane = dst + (size_t)z * nx * ny;

    // The segment consumes the extended-line samples q = p0-HW .. p1-1+HW of the X-blurred y line.
    // Samples q < ny-1 need one X-blurred row (|q|); the others are blends of two (ext_il, ext_il+1).
    // rowof(i) = raw row of the i-th X() call, so the cp.async producer can run ahead of the consumer.
    const int nsamp = (p1 - 1 + HW) - (p0 - HW) + 1;
    const int nreg = max(0, min(nsamp, (ny - 1) - (p0 - HW)));
    const int ncalls = nreg + 2 * (nsamp - nreg);
    auto rowof = [&](int i) -> int {
        if (i < nreg) {
            const int q = p0 - HW + i;
            return q < 0 ? -q : q;
        }
        const int j = i - nreg;
        return ty.ext_il[j >> 1] + (j & 1);
    };
    auto issue = [&](int i) {
        if (i < ncalls) cp_async16(bufs + (i % NB) * rowlen + PAD + 4 * x4, plane + (size_t)rowof(i) * nx + 4 * x4);
        cp_async_commit();  // an empty group keeps the group count uniform
    };
    int ci = 0;
#pragma unroll
    for (int i = 0; i < NB - 1; ++i) issue(i);

    // X-blurred float4 (this thread's 4 columns) of the next row in sequence; cooperative (CTA-uniform)
    auto X = [&]() -> float4 {
        float* b = bufs + (ci % NB) * rowlen;
        cp_async_wait<NB - 2>();  // this thread's piece of row ci has landed
        __syncthreads();          // ... everyone's has; and everyone is done reading row ci-1's buffer
        issue(ci + NB - 1);       // refill that buffer
        ++ci;
        // edge samples: every write below lands outside [PAD+1, PAD+nx-2], the only cells other threads
        // read here; cell PAD+nx-1 (raw, then the e = 0 blend) is read and written by the same thread
        if (x4 >= 1 && x4 <= HW) b[PAD - x4] = b[PAD + x4];  // left: q = -e mirrors to line[e]
        const int e = T - 1 - x4;                             // right: q = n-1+e is the blend at (il, frac)
        if (e <= HW) {
            const int il = tx.ext_il[e];
            const float frac = tx.ext_frac[e];
            const float lo = b[PAD + il], hi = b[PAD + il + 1];
            b[PAD + nx - 1 + e] = (1.0f - frac) * lo + frac * hi;
        }
        __syncthreads();
        float v[2 * PAD + 4];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 f = *reinterpret_cast<const float4*>(b + 4 * x4 + 4 * q);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k <= 2 * HW; ++k) acc += tx.w[k] * v[PAD + j + HW - k];
            o[j] = acc;
        }
        return make_float4(o[0], o[1], o[2], o[3]);
    };
    // extended-line sample q of the X-blurred y line (consumes one or two rows of the sequence)
    auto sample = [&](int q) -> float4 {
        if (q < ny - 1) return X();
        const int e = q - (ny - 1);
        const float frac = ty.ext_frac[e];
        const float4 lo = X(), hi = X();
        float4 r;
        r.x = (1.0f - frac) * lo.x + frac * hi.x; r.y = (1.0f - frac) * lo.y + frac * hi.y;
        r.z = (1.0f - frac) * lo.z + frac * hi.z; r.w = (1.0f - frac) * lo.w + frac * hi.w;
        return r;
    };

    float4 win[WR];
#pragma unroll
    for (int r = 0; r < 2 * HW; ++r) win[r] = sample(p0 - HW + r);
    win[2 * HW] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ib = 0; p0 + ib < p1; ib += WR) {
#pragma unroll
        for (int j = 0; j < WR; ++j) {
            const int p = p0 + ib + j;
            if (p < p1) {  // CTA-uniform
                win[(j + 2 * HW) % WR] = sample(p + HW);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k <= 2 * HW; ++k) {
                    const float4 vv = win[(j + 2 * HW - k) % WR];  // extended-line sample p+HW-k
                    const float w = ty.w[k];
                    acc.x += w * vv.x; acc.y += w * vv.y; acc.z += w * vv.z; acc.w += w * vv.w;
                }
                if (zok) *reinterpret_cast<float4*>(oplane + (size_t)p * nx + 4 * x4) = acc;
            }
        }
    }
    cp_async_wait<0>();
}

// The same pass with COMPACT CODE for the wide taps.  blur_xy_kernel inlines X() at every slot of its
// (2*HW+1)-way unrolled ring (three bodies per slot through sample()), which at HW >= 4 is tens of
// thousands of instructions: ncu shows the warps waiting on instruction fetch (no_instruction is the top
// stall at HW 4/6/8).  Here the row loop is NOT unrolled: one X body, one Y body, the ring advances by
// register moves (win[r] = win[r+1]); the wide passes are bound by the FP32 pipe (-fmad=false: an FMUL and
// an FADD per tap, each two issue cycles per warp), so the moves ride in otherwise idle issue slots.
// Arithmetic, operand order and the row sequence are those of blur_xy_kernel — results are bit-identical.
template <int HW>
__global__ void __launch_bounds__(128, 4) blur_xyc_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx, int ny,
                                                          int nz, int seg, Taps tx, Taps ty) {
    constexpr int PAD = (HW + 3) / 4 * 4;
    constexpr int NV = (2 * PAD + 4) / 4;
    constexpr int WR = 2 * HW + 1;
    constexpr int NB = kXYBufs;
    extern __shared__ __align__(16) float xy_sm[];
    const int T = nx >> 2, lpc = 128 / T;
    const int line = threadIdx.x / T, x4 = threadIdx.x - line * T;
    const int zgroups = (nz + lpc - 1) / lpc;
    const int s = blockIdx.x / zgroups, zg = blockIdx.x - s * zgroups;
    const int zraw = zg * lpc + line;
    const bool zok = zraw < nz;
    const int z = zok ? zraw : nz - 1;
    const int p0 = s * seg, p1 = min(ny, p0 + seg);
    const int rowlen = nx + 2 * PAD;
    float* const bufs = xy_sm + (size_t)line * NB * rowlen;
    const float* const plane = src + (size_t)z * nx * ny;
    float* const oplane = dst + (size_t)z * nx * ny;
    const int nsamp = (p1 - 1 + HW) - (p0 - HW) + 1;
    const int nreg = max(0, min(nsamp, (ny - 1) - (p0 - HW)));
    const int ncalls = nreg + 2 * (nsamp - nreg);
    auto rowof = [&](int i) -> int {
        if (i < nreg) {
            const int q = p0 - HW + i;
            return q < 0 ? -q : q;
        }
        const int j = i - nreg;
        return ty.ext_il[j >> 1] + (j & 1);
    };
    auto issue = [&](int i) {
        if (i < ncalls) cp_async16(bufs + (i % NB) * rowlen + PAD + 4 * x4, plane + (size_t)rowof(i) * nx + 4 * x4);
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < NB - 1; ++i) issue(i);
    // right-edge blend of the x axis: per-thread constants
    const int xe = T - 1 - x4;
    const bool xr = xe <= HW;
    const int xil = xr ? tx.ext_il[xe] : 0;
    const float xfrac = xr ? tx.ext_frac[xe] : 0.0f;

    float4 win[WR];
#pragma unroll
    for (int r = 0; r < WR; ++r) win[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    int ci = 0;
#pragma unroll 1
    for (int i = 0; i < nsamp; ++i) {
        const bool blend = i >= nreg;  // CTA-uniform: this sample is a blend of two X-blurred rows
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), xr4 = lo;
#pragma unroll 1
        for (int c = 0; c < (blend ? 2 : 1); ++c) {
            float* b = bufs + (ci % NB) * rowlen;
            cp_async_wait<NB - 2>();
            __syncthreads();
            issue(ci + NB - 1);
            ++ci;
            if (x4 >= 1 && x4 <= HW) b[PAD - x4] = b[PAD + x4];
            if (xr) {
                const float l = b[PAD + xil], h = b[PAD + xil + 1];
                b[PAD + nx - 1 + xe] = (1.0f - xfrac) * l + xfrac * h;
            }
            __syncthreads();
            float v[2 * PAD + 4];
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const float4 f = *reinterpret_cast<const float4*>(b + 4 * x4 + 4 * q);
                v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
            }
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float acc = 0.0f;
#pragma unroll
                for (int k = 0; k <= 2 * HW; ++k) acc += tx.w[k] * v[PAD + j + HW - k];
                o[j] = acc;
            }
            lo = xr4;
            xr4 = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (blend) {  // (lo, xr4) = rows (ext_il, ext_il + 1) of the y axis
            const float frac = ty.ext_frac[i - nreg];
            float4 r;
            r.x = (1.0f - frac) * lo.x + frac * xr4.x; r.y = (1.0f - frac) * lo.y + frac * xr4.y;
            r.z = (1.0f - frac) * lo.z + frac * xr4.z; r.w = (1.0f - frac) * lo.w + frac * xr4.w;
            xr4 = r;
        }
#pragma unroll
        for (int r = 0; r < 2 * HW; ++r) win[r] = win[r + 1];
        win[2 * HW] = xr4;  // win[r] = extended-line sample p - HW + r of output row p = p0 + i - 2*HW
        if (i >= 2 * HW) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k <= 2 * HW; ++k) {
                const float4 vv = win[2 * HW - k];  // sample p + HW - k
                const float w = ty.w[k];
                acc.x += w * vv.x; acc.y += w * vv.y; acc.z += w * vv.z; acc.w += w * vv.w;
            }
            if (zok) *reinterpret_cast<float4*>(oplane + (size_t)(p0 + i - 2 * HW) * nx + 4 * x4) = acc;
        }
    }
    cp_async_wait<0>();
}

// K3  DownSample_3D, Src/cSIFT3D.cc:506-533: dst(n,m,k) = src(2n,2m,2k).
__global__ void __launch_bounds__(256) downsample_kernel(const float* __restrict__ src, int sx, int sy, float* __restrict__ dst,
                                                         int dx, int dy, int dz) {
    const ll total = (ll)dx * dy * dz;
    const ll idx = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx % dx), mm = (int)((idx / dx) % dy), k = (int)(idx / ((ll)dx * dy));
    dst[idx] = src[(ll)(2 * n) + (ll)(2 * mm) * sx + (ll)(2 * k) * sx * sy];
}

// ---------------------------------------------------------------------------------------------
// K4  Detection.  Detect_KeyPoints Src/cSIFT3D.cc:362-425 + IsExtrema_neighbor :884-911.
//     thres = peak_thresh * max|D[j]| (:384-385); candidate iff |val| > thres (strict) and val is
//     strictly below or strictly above its 6 face neighbours in D[j] and the same voxel in
//     D[j-1], D[j+1] (8 neighbours, App. B Q1); x,y,z in [1, n-2].
//     Ordered compaction: each CTA owns 4096 consecutive voxels (16 per thread); flagged voxels are
//     ranked inside the CTA (popc prefix over 16 flags/thread + warp scan), staged under an atomically
//     claimed segment, and a later scan of per-CTA counts turns (cta, rank) into the raster
//     position — the reference's (octave, level, z, y, x) emission order (App. B Q16).
// ---------------------------------------------------------------------------------------------
struct StageEntry {
    uint32_t key;   // linear voxel index inside the level
    uint32_t gb;    // global CTA index over all (octave, level) detect units
    uint32_t rank;  // rank inside the CTA
    uint32_t unit;  // octave * L + (level - 1)
};

constexpr int kDetectPer = 16;                    // voxels per thread: 4 float4 groups
constexpr int kDetectChunk = 256 * kDetectPer;    // voxels per CTA (8 warps x 512 consecutive voxels)

// Voxel layout inside a CTA: warp w owns voxels [w*512, w*512+512) of the CTA's chunk; its load q
// (0..3) is the float4 at index q*32 + lane of that range, so every request is one fully coalesced
// 512-byte line group.  A thread therefore holds 4 groups of 4 consecutive voxels; raster order
// inside the warp is (q, lane, j).
__global__ void __launch_bounds__(256) detect_kernel(const float* __restrict__ Dm, const float* __restrict__ D0,
                                                     const float* __restrict__ Dp, int nx, int ny, int nz,
                                                     const unsigned* __restrict__ maxslot, float peak, uint32_t unit,
                                                     uint32_t gb_base, int* __restrict__ blk_cnt,
                                                     StageEntry* __restrict__ stage, unsigned* stage_count,
                                                     unsigned stage_cap, float* thres_out, ll vox_begin, ll vox_end) {
    // [vox_begin, vox_end): the voxels this launch owns, as linear indices of the WHOLE level
    // (nx, ny, nz are the level's global dims).  A z-slab shard passes pointers offset to a virtual
    // origin so that global indices address its local planes; the unsharded run passes [0, total).
    __shared__ int warp_tot[8];
    __shared__ unsigned seg_base;
    const ll total = vox_end;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const ll wbase_vox = vox_begin + (ll)blockIdx.x * kDetectChunk + (ll)wid * 512;
    const float thres = peak * __uint_as_float(*maxslot);
    if (blockIdx.x == 0 && threadIdx.x == 0 && thres_out) *thres_out = thres;
    // Phase 1 (every voxel, HBM-bound): |val| > thres  (val > thres || val < -thres, Src/cSIFT3D.cc:394)
    unsigned above = 0;
    if (wbase_vox + 512 <= total && (reinterpret_cast<uintptr_t>(D0 + wbase_vox) & 15) == 0) {
        float4 f[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) f[q] = *(reinterpret_cast<const float4*>(D0 + wbase_vox) + q * 32 + lane);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            above |= ((fabsf(f[q].x) > thres ? 1u : 0u) | (fabsf(f[q].y) > thres ? 2u : 0u) | (fabsf(f[q].z) > thres ? 4u : 0u) |
                      (fabsf(f[q].w) > thres ? 8u : 0u)) << (4 * q);
        // Pre-filter in registers: an extremum is in particular a strict extremum against its two x
        // neighbours, which sit in the same float4 or in the adjacent lane's (flat index +-1; voxels on a
        // row end are rejected by the interior test below whatever this says).  Elements whose neighbour
        // lies outside the warp's 512 voxels stay candidates.  This drops most above-threshold voxels
        // before the divergent per-voxel loop of phase 2.
        if (__any_sync(0xffffffffu, above != 0)) {
            unsigned keep = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float lw = __shfl_up_sync(0xffffffffu, f[q].w, 1), rx = __shfl_down_sync(0xffffffffu, f[q].x, 1);
                const float pl = __shfl_sync(0xffffffffu, f[q > 0 ? q - 1 : 0].w, 31), nr = __shfl_sync(0xffffffffu, f[q < 3 ? q + 1 : 3].x, 0);
                const bool hasl = lane > 0 || q > 0, hasr = lane < 31 || q < 3;
                const float left = lane == 0 ? pl : lw, right = lane == 31 ? nr : rx;
                auto ext = [](float v, float a, float b) { return (v < a && v < b) || (v > a && v > b); };
                const unsigned e0 = (!hasl || ext(f[q].x, left, f[q].y)) ? 1u : 0u;
                const unsigned e1 = ext(f[q].y, f[q].x, f[q].z) ? 2u : 0u;
                const unsigned e2 = ext(f[q].z, f[q].y, f[q].w) ? 4u : 0u;
                const unsigned e3 = (!hasr || ext(f[q].w, f[q].z, right)) ? 8u : 0u;
                keep |= (e0 | e1 | e2 | e3) << (4 * q);
            }
            above &= keep;
        }
    } else {
#pragma unroll
        for (int b = 0; b < kDetectPer; ++b) {
            const ll i = wbase_vox + (ll)((b >> 2) * 32 + lane) * 4 + (b & 3);
            if (i < total && fabsf(D0[i]) > thres) above |= 1u << b;
        }
    }
    // Phase 2 (rare): the 8-neighbour test for the voxels above the threshold
    unsigned flags = 0;
    if (above) {
        const ll ys = nx, zs = (ll)nx * ny;
        const uint32_t unx = (uint32_t)nx, uny = (uint32_t)ny;
        for (unsigned m = above; m; m &= m - 1) {
            const int b = __ffs(m) - 1;
            const ll i = wbase_vox + (ll)((b >> 2) * 32 + lane) * 4 + (b & 3);
            // total < 2^32 (checked at create): 32-bit divisions
            const uint32_t bi = (uint32_t)i, tq = bi / unx;
            const int x = (int)(bi - tq * unx), z = (int)(tq / uny), y = (int)(tq - (uint32_t)z * uny);
            if (x >= 1 && x <= nx - 2 && y >= 1 && y <= ny - 2 && z >= 1 && z <= nz - 2) {
                // strict extremum against all 8 neighbours (:889-902); cheapest (cached) neighbours
                // first, the conjunction is order-independent
                const float val = D0[i];
                const float t1 = D0[i - 1], t2 = D0[i + 1];
                bool mn = val < t1 && val < t2, mx = val > t1 && val > t2;
                if (!(mn || mx)) continue;
                const float t3 = D0[i + ys], t4 = D0[i - ys];
                mn = mn && val < t3 && val < t4; mx = mx && val > t3 && val > t4;
                if (!(mn || mx)) continue;
                const float t5 = D0[i + zs], t6 = D0[i - zs];
                mn = mn && val < t5 && val < t6; mx = mx && val > t5 && val > t6;
                if (!(mn || mx)) continue;
                const float t0 = Dm[i], t7 = Dp[i];
                mn = mn && val < t0 && val < t7; mx = mx && val > t0 && val > t7;
                if (mn || mx) flags |= 1u << b;
            }
        }
    }
    // CTAs without any detection (the majority) leave here
    if (!__syncthreads_or(flags != 0)) {
        if (threadIdx.x == 0) blk_cnt[gb_base + blockIdx.x] = 0;
        return;
    }
    // Ranks in raster order: warp base + groups before mine + lanes before me in my group + bits below
    int excl[4], run = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = __popc((flags >> (4 * q)) & 15u);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        excl[q] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) warp_tot[wid] = run;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < wid) wbase += warp_tot[w];
        tot += warp_tot[w];
    }
    if (threadIdx.x == 0) {
        blk_cnt[gb_base + blockIdx.x] = tot;
        seg_base = atomicAdd(stage_count, (unsigned)tot);
    }
    __syncthreads();
    if (flags) {
        const unsigned sb = seg_base;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned rank = (unsigned)(wbase + excl[q]);
            for (unsigned m = (flags >> (4 * q)) & 15u; m; m &= m - 1) {
                const int j = __ffs(m) - 1;
                const unsigned pos = sb + rank;
                if (pos < stage_cap) {
                    StageEntry e;
                    e.key = (uint32_t)(wbase_vox + (ll)(q * 32 + lane) * 4 + j);
                    e.gb = gb_base + blockIdx.x; e.rank = rank; e.unit = unit;
                    stage[pos] = e;
                }
                ++rank;
            }
        }
    }
}

// Exclusive scan of n ints by ONE CTA (1024 threads); total -> *total_out.  Each warp owns a contiguous
// range and walks it with coalesced loads (lane-contiguous, 4 x 32 entries in flight per step): the first
// version gave every THREAD a contiguous range, i.e. 32 different cache lines per warp load (111 us for the
// 112k per-CTA counts of a 512^3 detection).
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n, int* total_out) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per = (((n + 31) / 32) + 127) / 128 * 128;  // entries per warp, a multiple of 128
    const int b = min(n, wid * per), e = min(n, b + per);
    int s = 0;
    for (int i = b + lane; i < e; i += 128) {
        const int v0 = in[i], v1 = i + 32 < e ? in[i + 32] : 0, v2 = i + 64 < e ? in[i + 64] : 0, v3 = i + 96 < e ? in[i + 96] : 0;
        s += (v0 + v1) + (v2 + v3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[wid] = s;
    __syncthreads();
    int carry = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
        const int t = wsum[w];
        if (w < wid) carry += t;
        tot += t;
    }
    for (int i0 = b; i0 < e; i0 += 128) {
        int v[4], inc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * 32 + lane;
            v[q] = i < e ? in[i] : 0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int x = v[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += t;
            }
            inc[q] = x;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = i0 + q * 32 + lane;
            if (i < e) out[i] = carry + inc[q] - v[q];
            carry += __shfl_sync(0xffffffffu, inc[q], 31);
        }
    }
    if (threadIdx.x == 0 && total_out) *total_out = tot;
}

struct Cand {
    uint32_t key;
    uint32_t unit;
};

// (grid-stride over the staged entries: the count lives on the device; cand_cap = capacity of cand[])
__global__ void __launch_bounds__(256) scatter_kernel(const StageEntry* __restrict__ stage, const unsigned* stage_count,
                                                      unsigned stage_cap, const int* __restrict__ blk_off,
                                                      Cand* __restrict__ cand, int cand_cap) {
    const unsigned n = min(*stage_count, stage_cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const StageEntry e = stage[i];
        Cand c;
        c.key = e.key; c.unit = e.unit;
        const int pos = blk_off[e.gb] + e.rank;
        if (pos < cand_cap) cand[pos] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// Level table shared by the sparse kernels.
// ---------------------------------------------------------------------------------------------
struct LevelTable {
    const float* gss[kMaxOct * kMaxG];
    float scale[kMaxOct * kMaxG];  // (float)(2^(o+s/L) * sigma0), Src/cUtil.cc:209-210
    int dims[kMaxOct][3];
    int L;  // num_kp_levels
    int G;  // L + 3
    // Gaussian window weights as tables (see wtab_kernel): offsets into `wtab`, entries per table
    const float* wtab;
    int ori_off[kMaxOct * kMaxG], ori_n[kMaxOct * kMaxG];
    int desc_off[kMaxOct * kMaxG], desc_n[kMaxOct * kMaxG];
};

// Window weights depend on the voxel only through m = i^2+j^2+k^2 (integer voxel offsets from the
// integer keypoint position): disp = offset * u with u = 2^octave, so
// sq = dx*dx+dy*dy+dz*dz = u*u*m EXACTLY in FP32 (all terms are small integers times a power of two).
// The tables hold, per (octave, level), the reference's expression evaluated at every m:
//   orientation  expf((float)(-0.5 * (double)sq / (double)(sigma*sigma)))   Src/cSIFT3D.cc:971
//   descriptor   expf(-0.5f * sq / (sigma*sigma))                            Src/cSIFT3D.cc:1312
// with expf = s3d_expf_ref (glibc's algorithm), so a lookup returns bit for bit what the
// per-voxel evaluation did.  One thread per entry.
__global__ void __launch_bounds__(256) wtab_kernel(LevelTable tab, int noct, float* __restrict__ out) {
    const int lv = blockIdx.y;                 // o * G + level
    const int o = lv / tab.G;
    if (o >= noct) return;
    const float u = (float)(1 << o);
    const float scale = tab.scale[lv];
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < tab.ori_n[lv] + tab.desc_n[lv]; m += gridDim.x * blockDim.x) {
        if (m < tab.ori_n[lv]) {
            const float sigma = 1.5f * scale;
            const float sq = u * u * (float)m;
            out[tab.ori_off[lv] + m] = s3d_expf_ref((float)(-0.5 * (double)sq / (double)(sigma * sigma)));
        } else {
            const int md = m - tab.ori_n[lv];
            const float sigma = scale * 7.071067812f;
            const float sq = u * u * (float)md;
            out[tab.desc_off[lv] + md] = s3d_expf_ref(-0.5f * sq / (sigma * sigma));
        }
    }
}

// Symmetric 3x3 eigen-decomposition in double (cyclic Jacobi), columns of V = unit eigenvectors.
// Stands in for Eigen::EigenSolver<Matrix3d> (Src/cSIFT3D.cc:1016-1029): eigenvalues agree to
// ~1e-15 relative, eigenvectors up to sign, and the reference fixes signs itself (:1089-1108).
__device__ inline void eig3_jacobi(const double A[9], double val[3], double V[9]) {
    double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        val[i] = a[i][i];
        double nn = sqrt(v[0][i] * v[0][i] + v[1][i] * v[1][i] + v[2][i] * v[2][i]);
#pragma unroll
        for (int k = 0; k < 3; k++) V[k * 3 + i] = v[k][i] / nn;
    }
}

// Window bounds, Src/cSIFT3D.cc:939-955 and :1182-1198 (App. B Q23).
__device__ __forceinline__ void window_bounds(float c, float r_over_u, int n, int& s, int& e) {
    int a = (int)floorf(c - r_over_u);
    s = a > 1 ? a : 1;
    int b = (int)ceilf(c + r_over_u);
    e = b < (n - 2) ? b : n - 1 - 1;
}

struct OrientSums {
    float t00, t01, t02, t11, t12, t22, wx, wy, wz;
};

// One window voxel of Assign_Orientation_Imp, Src/cSIFT3D.cc:964-994 (same operation order).
// (0.5 * (double)(a - b) rounded to float == 0.5f * (a - b): halving is exact.)
__device__ __forceinline__ void orient_voxel(const float* __restrict__ g, ll i, ll ys, ll zs, float dx, float dy, float dz,
                                             float iu, float inv_u2, const float* __restrict__ wt, float r2,
                                             OrientSums& S) {
    const float sq = dx * dx + dy * dy + dz * dz;
    if (sq > r2) return;
    const float weight = wt[(int)(sq * inv_u2)];
    float vx = 0.5f * (g[i + 1] - g[i - 1]);
    float vy = 0.5f * (g[i + ys] - g[i - ys]);
    float vz = 0.5f * (g[i + zs] - g[i - zs]);
    vx *= iu; vy *= iu; vz *= iu;
    S.t00 += vx * vx * weight;
    S.t01 += vx * vy * weight;
    S.t02 += vx * vz * weight;
    S.t11 += vy * vy * weight;
    S.t12 += vy * vz * weight;
    S.t22 += vz * vz * weight;
    S.wx += vx * weight; S.wy += vy * weight; S.wz += vz * weight;
}

// Everything after the window loop of Assign_Orientation_Imp, Src/cSIFT3D.cc:1000-1137.
// Fills kp (str_tensor, win, eigvalue, eigvector, Rotation) and returns 1 / -1 / -2 / -3.
// margin_out (may be null) receives the smallest relative distance of any accept/reject test to
// its threshold (used to pick candidates for the exact serial re-evaluation).
__device__ inline int orient_finish(const OrientSums& S, s3d_keypoint& kp, float max_eig_ratio, float corner_thresh,
                                    float* margin_out) {
    float* T = kp.str_tensor;
    T[0] = S.t00; T[1] = S.t01; T[2] = S.t02; T[4] = S.t11; T[5] = S.t12; T[8] = S.t22;
    T[3] = T[1]; T[6] = T[2]; T[7] = T[5];
    kp.win[0] = S.wx; kp.win[1] = S.wy; kp.win[2] = S.wz;
    float margin = FLT_MAX;
    const float wn2 = S.wx * S.wx + S.wy * S.wy + S.wz * S.wz;
    margin = fminf(margin, fabsf(wn2 - 1E-10f) / 1E-10f);
    if (wn2 < 1E-10f) {
        if (margin_out) *margin_out = margin;
        return -1;
    }
    double A[9], val[3], V[9];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = (double)T[i];
    eig3_jacobi(A, val, V);
    float ev[3], evec[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        ev[i] = (float)val[i];
#pragma unroll
        for (int k = 0; k < 3; k++) evec[i][k] = (float)V[k * 3 + i];
    }
    // ascending by value (std::sort with cmp, :1050)
#pragma unroll
    for (int i = 1; i < 3; i++)
#pragma unroll
        for (int j = i; j > 0; j--)
            if (ev[j] < ev[j - 1]) {
                float tv = ev[j]; ev[j] = ev[j - 1]; ev[j - 1] = tv;
#pragma unroll
                for (int k = 0; k < 3; k++) { float tk = evec[j][k]; evec[j][k] = evec[j - 1][k]; evec[j - 1][k] = tk; }
            }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        kp.eigvalue[i] = ev[i];
#pragma unroll
        for (int k = 0; k < 3; k++) kp.eigvector[i * 3 + k] = evec[i][k];
    }
    const float r01 = fabsf(ev[0] / ev[1]), r12 = fabsf(ev[1] / ev[2]);
    margin = fminf(margin, fminf(fabsf(r01 - max_eig_ratio), fabsf(r12 - max_eig_ratio)));
    int code = 1;
    if (r01 > max_eig_ratio || r12 > max_eig_ratio) code = -2;
    // DistinctEig, :1140-1150
    if (code == 1 && (fabs((double)(ev[0] - ev[1])) < DBL_EPSILON || fabs((double)(ev[0] - ev[2])) < DBL_EPSILON ||
                      fabs((double)(ev[2] - ev[1])) < DBL_EPSILON))
        code = -2;
    if (code != 1) {
        if (margin_out) *margin_out = margin;
        return code;
    }
    const float d_NORM = sqrtf(wn2);
    float corner_score = FLT_MAX;
#pragma unroll
    for (int i = 2; i > 0; i--) {
        const float ex = evec[i][0], ey = evec[i][1], ez = evec[i][2];
        const float d = ex * S.wx + ey * S.wy + ez * S.wz;
        const float q_NORM = sqrtf(ex * ex + ey * ey + ez * ez);
        const float cos_ang = d / (d_NORM * q_NORM);
        const float abs_cos_ang = fabsf(cos_ang);
        corner_score = corner_score < abs_cos_ang ? corner_score : abs_cos_ang;
        const float sgn = d > 0.0f ? 1.0f : -1.0f;
        evec[i][0] *= sgn; evec[i][1] *= sgn; evec[i][2] *= sgn;
    }
    margin = fminf(margin, fabsf(corner_score - corner_thresh));
    if (margin_out) *margin_out = margin;
    if (corner_score < corner_thresh) return -3;
    const float* v1 = evec[2];
    const float* v2 = evec[1];
    const float vr0 = v1[1] * v2[2] - v1[2] * v2[1];
    const float vr1 = v1[2] * v2[0] - v1[0] * v2[2];
    const float vr2 = v1[0] * v2[1] - v1[1] * v2[0];
    float* R = kp.Rotation;
    R[0] = v1[0]; R[1] = v2[0]; R[2] = vr0;
    R[3] = v1[1]; R[4] = v2[1]; R[5] = vr1;
    R[6] = v1[2]; R[7] = v2[2]; R[8] = vr2;
    return 1;
}

__device__ __forceinline__ void cand_decode(const Cand c, const LevelTable& tab, int& o, int& lvl, int& x, int& y, int& z) {
    o = (int)(c.unit / (uint32_t)tab.L);
    lvl = (int)(c.unit % (uint32_t)tab.L) + 1;
    const int nx = tab.dims[o][0], ny = tab.dims[o][1];
    x = (int)(c.key % (uint32_t)nx);
    y = (int)((c.key / (uint32_t)nx) % (uint32_t)ny);
    z = (int)(c.key / ((uint32_t)nx * (uint32_t)ny));
}

__device__ __forceinline__ void kp_init(s3d_keypoint& kp, int o, int lvl, int x, int y, int z, float scale) {
    // Detect_KeyPoints :401-408 + Initialize_Keypoint Src/cUtil.cc:449-463
    kp.x = (float)x; kp.y = (float)y; kp.z = (float)z;
    kp.scale = scale; kp.octave = o; kp.level = lvl;
    kp.rx = kp.ry = kp.rz = -1.0f;
#pragma unroll
    for (int i = 0; i < 3; i++) { kp.win[i] = 0.f; kp.eigvalue[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 9; i++) { kp.eigvector[i] = 0.f; kp.Rotation[i] = 0.f; kp.str_tensor[i] = 0.f; }
    kp.desc = nullptr;
}

// K5  Orientation: one warp per detection.  Assign_Orientation Src/cSIFT3D.cc:427-482.
//     Lanes stride over the (y,x) plane of each z slice, keep 9 partial sums, butterfly-reduce
//     (fixed tree, so results are run-to-run deterministic), then finish redundantly per lane.
//     The FP32 summation order differs from the reference's serial z,y,x order; detections whose
//     tests land within `recheck_margin` of a threshold are appended to recheck_list for the exact
//     serial re-evaluation kernel below.
// (measured: 3 CTAs/SM without spills, with one or two window voxels per lane in flight, is 16-20 % slower
// than these 4 CTAs/SM with a few spilled words — occupancy hides the L2 latency better)
__global__ void __launch_bounds__(256, 4) orient_kernel(const Cand* __restrict__ cand, const int* __restrict__ ncand_dev, int cand_cap, LevelTable tab,
                                                     s3d_keypoint* __restrict__ out, int* __restrict__ codes,
                                                     int* __restrict__ xyz5, float max_eig, float corner,
                                                     float recheck_margin, int* __restrict__ recheck_list,
                                                     int* __restrict__ n_recheck) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    const int ncand = min(*ncand_dev, cand_cap);  // the detection count never visits the host
    for (int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < ncand; ci += warps_per_grid) {
        int o, lvl, x, y, z;
        cand_decode(cand[ci], tab, o, lvl, x, y, z);
        const int nx = tab.dims[o][0], ny = tab.dims[o][1], nz = tab.dims[o][2];
        const int lv = o * tab.G + lvl;
        const float* g = tab.gss[lv];
        const float* wt = tab.wtab + tab.ori_off[lv];
        const float scale = tab.scale[lv];
        const float u = (float)(1 << o);
        const float iu = 1.0f / u, inv_u2 = iu * iu;
        s3d_keypoint kp;
        kp_init(kp, o, lvl, x, y, z, scale);
        const float sigma = 1.5f * scale;          // ori_sig_fctr, :27,:442
        const float win_radius = sigma * 3.0f;     // ori_rad_fctr, :28,:915
        const float r2 = win_radius * win_radius;
        int xs, xe, y0, y1, z0, z1;
        window_bounds(kp.x, win_radius / u, nx, xs, xe);
        window_bounds(kp.y, win_radius / u, ny, y0, y1);
        window_bounds(kp.z, win_radius / u, nz, z0, z1);
        const int wxn = xe - xs + 1, wyn = y1 - y0 + 1;
        const int plane = wxn > 0 && wyn > 0 ? wxn * wyn : 0;
        const float inv_wxn = 1.0f / (float)(wxn > 0 ? wxn : 1);
        const ll ys = nx, zs = (ll)nx * ny;
        OrientSums S = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int zz = z0; zz <= z1; ++zz) {
            const float dz = ((float)zz - kp.z) * u;
            const float* gz = g + (ll)zz * zs;
            for (int tI = lane; tI < plane; tI += 32) {
                // tI / wxn without an integer division: (tI + 0.5) / wxn is at least 0.5/wxn away from
                // an integer, far more than the FP32 error for tI < 2^20
                const int ry = (int)(((float)tI + 0.5f) * inv_wxn);
                const int yy = y0 + ry, xx = xs + (tI - ry * wxn);
                const float dx = ((float)xx - kp.x) * u, dy = ((float)yy - kp.y) * u;
                orient_voxel(gz, (ll)xx + (ll)yy * ys, ys, zs, dx, dy, dz, iu, inv_u2, wt, r2, S);
            }
        }
        float* sp = reinterpret_cast<float*>(&S);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            float v = sp[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            sp[k] = v;
        }
        float margin;
        const int code = orient_finish(S, kp, max_eig, corner, &margin);
        if (lane == 0) {
            if (code < 1) kp.x = kp.y = kp.z = -1.0f;  // :446-450
            out[ci] = kp;
            codes[ci] = code;
            if (recheck_list && margin < recheck_margin) recheck_list[atomicAdd(n_recheck, 1)] = ci;
            xyz5[5 * ci + 0] = x; xyz5[5 * ci + 1] = y; xyz5[5 * ci + 2] = z; xyz5[5 * ci + 3] = o; xyz5[5 * ci + 4] = lvl;
        }
    }
}

// One window voxel's nine addends (no accumulation), same arithmetic as orient_voxel.
__device__ __forceinline__ bool orient_terms(const float* __restrict__ g, ll i, ll ys, ll zs, float dx, float dy, float dz,
                                             float iu, float inv_u2, const float* __restrict__ wt, float r2,
                                             float tm[9]) {
    const float sq = dx * dx + dy * dy + dz * dz;
    if (sq > r2) return false;
    const float weight = wt[(int)(sq * inv_u2)];
    float vx = 0.5f * (g[i + 1] - g[i - 1]);
    float vy = 0.5f * (g[i + ys] - g[i - ys]);
    float vz = 0.5f * (g[i + zs] - g[i - zs]);
    vx *= iu; vy *= iu; vz *= iu;
    tm[0] = vx * vx * weight; tm[1] = vx * vy * weight; tm[2] = vx * vz * weight;
    tm[3] = vy * vy * weight; tm[4] = vy * vz * weight; tm[5] = vz * vz * weight;
    tm[6] = vx * weight; tm[7] = vy * weight; tm[8] = vz * weight;
    return true;
}

// Exact re-evaluation of the flagged detections in the reference's own summation order
// (z, y, x serial FP32 accumulation, Src/cSIFT3D.cc:958-998), so near-threshold accept/reject
// decisions follow the reference's rounding instead of the warp-parallel tree's.
// One CTA per re-checked detection: warps 1..7 PRODUCE the nine addends of consecutive window positions
// (the reference's z, y, x order, linearised) into a double-buffered shared-memory chunk while warp 0
// CONSUMES the previous chunk — lane k < 9 adds component k's values one by one, which is the reference's
// serial FP32 summation (Src/cSIFT3D.cc:957-1010).  Positions outside the sphere contribute +0.0f, which
// leaves a running sum that started at +0.0f unchanged bit for bit (x + 0 == x; the sum can never be -0).
// The first version gave a whole window to one warp, which waited a full L2 round trip per row before
// its 24 dependent adds: 0.40 ms for the slowest detection, whatever the GPU's width.
constexpr int kExactWarps = 8;
constexpr int kExactChunk = 512;  // window positions per chunk
__global__ void __launch_bounds__(kExactWarps * 32) orient_exact_kernel(const Cand* __restrict__ cand, LevelTable tab,
                                                                        s3d_keypoint* __restrict__ out,
                                                                        int* __restrict__ codes,
                                                                        const int* __restrict__ recheck_list,
                                                                        const int* __restrict__ n_recheck, float max_eig,
                                                                        float corner, int* n_flipped) {
    __shared__ float buf[2][9][kExactChunk];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nre = *n_recheck;
    const int row9 = lane < 9 ? lane : 0;
    constexpr int NP = (kExactWarps - 1) * 32;  // producer threads
    for (int ri = blockIdx.x; ri < nre; ri += gridDim.x) {
        const int ci = recheck_list[ri];
        int o, lvl, x, y, z;
        cand_decode(cand[ci], tab, o, lvl, x, y, z);
        const int nx = tab.dims[o][0], ny = tab.dims[o][1], nz = tab.dims[o][2];
        const int lv = o * tab.G + lvl;
        const float* g = tab.gss[lv];
        const float* wt = tab.wtab + tab.ori_off[lv];
        const float scale = tab.scale[lv];
        const float u = (float)(1 << o);
        const float iu = 1.0f / u, inv_u2 = iu * iu;
        s3d_keypoint kp;
        kp_init(kp, o, lvl, x, y, z, scale);
        const float sigma = 1.5f * scale;
        const float win_radius = sigma * 3.0f;
        const float r2 = win_radius * win_radius;
        int xs, xe, y0, y1, z0, z1;
        window_bounds(kp.x, win_radius / u, nx, xs, xe);
        window_bounds(kp.y, win_radius / u, ny, y0, y1);
        window_bounds(kp.z, win_radius / u, nz, z0, z1);
        const ll ys = nx, zs = (ll)nx * ny;
        const int ncol = max(0, xe - xs + 1), nrow = max(0, y1 - y0 + 1), nsl = max(0, z1 - z0 + 1);
        const int npos = ncol * nrow * nsl;
        const int nchunk = (npos + kExactChunk - 1) / kExactChunk;
        float acc = 0.0f;  // warp 0, lane k < 9: component k
        for (int c = 0; c <= nchunk; ++c) {
            if (wid > 0) {
                if (c < nchunk) {
                    float(*bw)[kExactChunk] = buf[c & 1];
                    for (int t = (int)threadIdx.x - 32; t < kExactChunk; t += NP) {
                        const int idx = c * kExactChunk + t;
                        float tm[9];
                        bool in = false;
                        if (idx < npos) {
                            const int rowi = idx / ncol;
                            const int xx = xs + (idx - rowi * ncol);
                            const int sl = rowi / nrow;
                            const int yy = y0 + (rowi - sl * nrow), zz = z0 + sl;
                            in = orient_terms(g, (ll)xx + (ll)yy * ys + (ll)zz * zs, ys, zs, ((float)xx - kp.x) * u,
                                              ((float)yy - kp.y) * u, ((float)zz - kp.z) * u, iu, inv_u2, wt, r2, tm);
                        }
#pragma unroll
                        for (int k = 0; k < 9; ++k) bw[k][t] = in ? tm[k] : 0.0f;
                    }
                }
            } else if (c > 0) {
                const float* br = buf[(c - 1) & 1][row9];
                const int cnt = min(kExactChunk, npos - (c - 1) * kExactChunk);
                int j = 0;
                for (; j + 8 <= cnt; j += 8) {
                    const float4 a = *reinterpret_cast<const float4*>(br + j), b = *reinterpret_cast<const float4*>(br + j + 4);
                    acc += a.x; acc += a.y; acc += a.z; acc += a.w;
                    acc += b.x; acc += b.y; acc += b.z; acc += b.w;
                }
                for (; j < cnt; ++j) acc += br[j];
            }
            __syncthreads();
        }
        if (wid == 0) {
            OrientSums S;
            S.t00 = __shfl_sync(0xffffffffu, acc, 0); S.t01 = __shfl_sync(0xffffffffu, acc, 1);
            S.t02 = __shfl_sync(0xffffffffu, acc, 2); S.t11 = __shfl_sync(0xffffffffu, acc, 3);
            S.t12 = __shfl_sync(0xffffffffu, acc, 4); S.t22 = __shfl_sync(0xffffffffu, acc, 5);
            S.wx = __shfl_sync(0xffffffffu, acc, 6); S.wy = __shfl_sync(0xffffffffu, acc, 7);
            S.wz = __shfl_sync(0xffffffffu, acc, 8);
            const int code = orient_finish(S, kp, max_eig, corner, nullptr);
            if (lane == 0) {
                if (code < 1) kp.x = kp.y = kp.z = -1.0f;
                if (code != codes[ci]) atomicAdd(n_flipped, 1);
                out[ci] = kp;
                codes[ci] = code;
            }
        }
    }
}

// Ordered compaction of the survivors (serial loop Src/cSIFT3D.cc:459-466): one CTA.
__global__ void __launch_bounds__(1024) survivors_kernel(const int* __restrict__ codes, const int* __restrict__ n_dev, int cap,
                                                         int* __restrict__ surv, int* total_out) {
    __shared__ int part[1024];
    const int n = min(*n_dev, cap);
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += codes[i] == 1;
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = b; i < e; ++i)
        if (codes[i] == 1) surv[run++] = i;
    if (threadIdx.x == 1023) *total_out = part[1023];
}

// Launch order of the descriptor CTAs: heaviest keypoints first.  A keypoint's window volume grows with
// the cube of its level's scale (levels 1..L: about 35k / 68k / 137k contributing voxels), and the
// survivor list is in (octave, level, z, y, x) order, so a grid launched in list order ends on the
// slowest CTAs with most SMs idle.  order[] = survivor indices by DESCENDING level, stable within a level
// (one CTA, one ordered block scan per level).  Results do not depend on the order: every CTA writes its
// own keypoint's slot.
__global__ void __launch_bounds__(1024) desc_order_kernel(const s3d_keypoint* __restrict__ extre, const int* __restrict__ surv,
                                                          const int* __restrict__ n_dev, int max_level, int* __restrict__ order) {
    constexpr int NL = kMaxG;  // classes: level 1..NL-1, class 0 = anything else (goes last)
    constexpr int CH = 16;     // levels cached per thread
    __shared__ int wtot[32][NL];
    __shared__ int cbase[NL];
    const int n = *n_dev;
    const int per = (n + 1023) / 1024;
    const int b = min(n, (int)threadIdx.x * per), e = min(n, b + per);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    auto cls = [&](int i) -> int {
        const int l = extre[surv[i]].level;
        return (l >= 1 && l <= max_level && l < NL) ? l : 0;
    };
    unsigned char lv[CH];
    int cnt[NL];
#pragma unroll
    for (int c = 0; c < NL; ++c) cnt[c] = 0;
    for (int i = b; i < e; ++i) {
        const int c = cls(i);
        if (i - b < CH) lv[i - b] = (unsigned char)c;
#pragma unroll
        for (int k = 0; k < NL; ++k) cnt[k] += (c == k);
    }
    // exclusive block scan of every class count: warp shuffles, then the warp totals
    int excl[NL];
#pragma unroll
    for (int c = 0; c < NL; ++c) {
        int v = cnt[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        excl[c] = v - cnt[c];
        if (lane == 31) wtot[wid][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < NL) {  // per class: exclusive scan over the 32 warps, and the class base (descending level)
        const int c = threadIdx.x;
        int run = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = wtot[w][c];
            wtot[w][c] = run;
            run += t;
        }
        cbase[c] = run;  // total, turned into a base below
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int c = NL - 1; c >= 1; --c) { const int t = cbase[c]; cbase[c] = run; run += t; }
        cbase[0] = run;
    }
    __syncthreads();
    int pos[NL];
#pragma unroll
    for (int c = 0; c < NL; ++c) pos[c] = cbase[c] + wtot[wid][c] + excl[c];
    for (int i = b; i < e; ++i) {
        const int c = (i - b < CH) ? (int)lv[i - b] : cls(i);
#pragma unroll
        for (int k = 0; k < NL; ++k)
            if (c == k) order[pos[k]++] = i;
    }
}

// ---------------------------------------------------------------------------------------------
// K6  Descriptor.  Extract_Descriptor_Imp Src/cSIFT3D.cc:1152-1381 +
//     Trilinear_interpolation_over_desc_debug :1450-1540 + Check_intersect_faces :1542-1573 +
//     cart2bary :1592-1637 + normailize_desc :1639-1656.
// ---------------------------------------------------------------------------------------------
struct MeshConst {
    float e1[20][3], e2[20][3], t[20][3], q[20][3];  // per face: V1-V0, V2-V0, -V0, t x e1
    float qe2[20];                                   // q . e2
    float cen10[10][3];                              // centroid of one face of each antipodal pair
    int pos[10], neg[10];                            // face with centroid +cen10[i] / -cen10[i]
    int idx[20][3];                                  // vertex ids (NOT swapped, App. B Q13)
    // Octant pre-selection (host_mesh): an octant of direction space is covered by ONE whole face (a
    // vertex on each coordinate plane) and the halves of the three faces across its edges.  octn[o][k]
    // is the inward normal of edge k's plane through the origin, octface[o] = {octant face, neighbour
    // across edge 0, 1, 2}.  o = (gx<0) | (gy<0)<<1 | (gz<0)<<2.
    float octn[8][3][3];
    int octface[8][4];
};

// cart2bary for one face, Src/cSIFT3D.cc:1592-1637, with the face-only terms precomputed on the
// host in the same FP32 arithmetic.  Returns false for |det| < bary_eps.
__device__ __forceinline__ bool face_bary(const MeshConst& M, int f, float gx, float gy, float gz, float bary_eps,
                                          float& b0, float& b1, float& b2, float& k) {
    const float e2x = M.e2[f][0], e2y = M.e2[f][1], e2z = M.e2[f][2];
    const float px = gy * e2z - gz * e2y;
    const float py = gz * e2x - gx * e2z;
    const float pz = gx * e2y - gy * e2x;
    const float det = M.e1[f][0] * px + M.e1[f][1] * py + M.e1[f][2] * pz;
    if (fabsf(det) < bary_eps) return false;
    const float det_inv = (float)(1.0 / (double)det);
    b1 = det_inv * (px * M.t[f][0] + py * M.t[f][1] + pz * M.t[f][2]);
    b2 = det_inv * (gx * M.q[f][0] + gy * M.q[f][1] + gz * M.q[f][2]);
    b0 = 1 - b1 - b2;
    k = det_inv * M.qe2[f];
    return true;
}

// The same with fused multiply-adds and a correctly rounded FP32 reciprocal (the reference's
// (float)(1.0/(double)det) differs from it at most in the last bit, in ~1 of 2^29 cases): used only for
// the pre-selected face of find_face, whose result is accepted when all coordinates are comfortably
// inside the face — 1e-4 away from any decision the last bits could change.
__device__ __forceinline__ bool face_bary_fast(const MeshConst& M, int f, float gx, float gy, float gz, float bary_eps,
                                               float& b0, float& b1, float& b2, float& k) {
    const float e2x = M.e2[f][0], e2y = M.e2[f][1], e2z = M.e2[f][2];
    const float px = __fmaf_rn(gy, e2z, -__fmul_rn(gz, e2y));
    const float py = __fmaf_rn(gz, e2x, -__fmul_rn(gx, e2z));
    const float pz = __fmaf_rn(gx, e2y, -__fmul_rn(gy, e2x));
    const float det = __fmaf_rn(M.e1[f][0], px, __fmaf_rn(M.e1[f][1], py, __fmul_rn(M.e1[f][2], pz)));
    if (fabsf(det) < bary_eps) return false;
    const float det_inv = __frcp_rn(det);
    b1 = det_inv * __fmaf_rn(px, M.t[f][0], __fmaf_rn(py, M.t[f][1], __fmul_rn(pz, M.t[f][2])));
    b2 = det_inv * __fmaf_rn(gx, M.q[f][0], __fmaf_rn(gy, M.q[f][1], __fmul_rn(gz, M.q[f][2])));
    b0 = 1 - b1 - b2;
    k = det_inv * M.qe2[f];
    return true;
}

// Check_intersect_faces, Src/cSIFT3D.cc:1542-1573: the FIRST face (index order) whose barycentric
// coordinates are all >= -bary_eps with k >= 0 wins (App. B Q14).
// Fast path: the icosahedron's vertices lie on the coordinate planes, so each octant of direction
// space holds one whole face and halves of the three faces across its edges: the sign bits and
// three plane tests (fused: this is only a pre-selection) name the face.  If the selected face's
// barycentric coordinates are all comfortably positive, no other face can pass the reference's
// tolerance test and (face, bary) is what the sequential scan returns.  Directions within `margin`
// of an edge/vertex — and any direction the pre-selection got wrong — take the reference's scan.
__device__ __forceinline__ int find_face(const MeshConst& M, float gx, float gy, float gz, float bary_eps, float& b0,
                                         float& b1, float& b2) {
    // three plane tests against the edges of the direction's octant face pick one of four faces (a
    // pre-selection only: the margin test below decides whether it stands)
    const int oct = (gx < 0.0f ? 1 : 0) | (gy < 0.0f ? 2 : 0) | (gz < 0.0f ? 4 : 0);
    const float(*N)[3] = M.octn[oct];
    const float d0 = __fmaf_rn(N[0][0], gx, __fmaf_rn(N[0][1], gy, __fmul_rn(N[0][2], gz)));
    const float d1 = __fmaf_rn(N[1][0], gx, __fmaf_rn(N[1][1], gy, __fmul_rn(N[1][2], gz)));
    const float d2 = __fmaf_rn(N[2][0], gx, __fmaf_rn(N[2][1], gy, __fmul_rn(N[2][2], gz)));
    const int region = d0 < 0.0f ? 1 : (d1 < 0.0f ? 2 : (d2 < 0.0f ? 3 : 0));
    const int fs = M.octface[oct][region];
    float k;
    const float margin = 1e-4f;
    if (face_bary_fast(M, fs, gx, gy, gz, bary_eps, b0, b1, b2, k) && b0 > margin && b1 > margin && b2 > margin && k > 0.0f)
        return fs;
    for (int f = 0; f < 20; ++f) {
        float c0, c1, c2, kk;
        if (!face_bary(M, f, gx, gy, gz, bary_eps, c0, c1, c2, kk)) continue;
        if (c0 < -bary_eps || c1 < -bary_eps || c2 < -bary_eps || kk < 0) continue;
        b0 = c0; b1 = c1; b2 = c2;
        return f;
    }
    return -1;
}

constexpr int kDescWarps = 7;
constexpr int kDescThreads = kDescWarps * 32;
// Histogram bin (x,y,z,v) lives at (x + 4y)*12 + z*kHistZ + v: the z stride is padded from 192 to
// 200 words so the 8 trilinear cells of a voxel fall into 8 different bank groups.
constexpr int kHistZ = 200;
constexpr int kHistDump = 3 * kHistZ + 192;      // 12 slots (one per vertex) for out-of-grid cells
constexpr int kHistStride = kHistDump + 16;
constexpr int kStageCols = 33;                   // 32 voxels + 1 pad: conflict-free 64-bit rows and columns
constexpr int kDescWtCap = 1312;                 // default parameters need <= 1292 entries

struct DescSmem {
    float hist[kDescWarps][kHistStride];
    uint2 stage[kDescWarps][24][kStageCols];         // (bin address, value bits); column = rank in the batch
    uint32_t queue[kDescWarps][64];                  // per-warp ring of voxels that passed the cheap tests
    MeshConst M;
    s3d_keypoint kp;
    float red[kDescWarps + 1];
    float wt[kDescWtCap];                            // Gaussian window weight by m = i^2+j^2+k^2 (wtab_kernel)
};
// Fixed-point variant: ONE histogram of 32-bit counters per CTA, accumulated with native shared-memory
// integer atomics (ATOMS.ADD; FP32 shared atomics are CAS loops on sm_100) — no staging, no replay.
// kQCopies interleaved copies (lane & (kQCopies-1) picks one): neighbouring voxels of a row mostly hit
// the SAME bins, and same-address atomics of one warp instruction are serialised; with the copies
// kQCopyStride words apart (== 4 mod 32 banks) the lanes of a group also land in different banks.
// (measured: 16 copies, which fit 3 CTAs/SM instead of 4, halve the same-address serialisation but lose more to the
// lower occupancy: describe 10.14 ms against 9.52 ms with 8 copies; build with -DS3D_QCOPIES=16 to repeat it)
#ifndef S3D_QCOPIES
#define S3D_QCOPIES 8
#endif
constexpr int kQCopies = S3D_QCOPIES;
constexpr int kQCopyStride = kQCopies == 8 ? 836 : 834;  // == 32 / kQCopies mod 32 banks
constexpr int kQMinCtas = kQCopies == 8 ? 4 : 3;          // CTAs per SM the shared-memory footprint allows
static_assert(kQCopyStride >= kHistStride && kQCopyStride % 32 == 32 / kQCopies, "copy stride must cover a histogram and rotate banks");
struct DescSmemQ {
    uint32_t hist[kQCopies * kQCopyStride];
    uint32_t queue[kDescWarps][64];
    uint4 queue2[kDescWarps][64];  // second-stage ring: (packed offsets, rotated gradient) of voxels above the |grad|^2 floor
    MeshConst M;
    s3d_keypoint kp;
    float red[kDescWarps + 1];
    float wt[kDescWtCap];
    int overflow;
};
// Fixed-point scale: every contribution is mag * (trilinear weight <= 1) * (barycentric <= 1), so it is
// bounded by mag.  With qscale chosen so that mag * qscale <= kQCap for every voxel, a bin — which at
// most (2*cell+2)^3 <= 28^3 voxels can touch — stays below 28^3 * kQCap < 2^32: no wrap-around.  A voxel
// that exceeds the cap raises the CTA's overflow flag and the keypoint is redone by the FP32 kernel.
constexpr float kQCap = 190000.0f;
constexpr float kQMargin = 8.0f;   // qscale = kQCap / (kQMargin * sampled max of mag)

// One CTA per surviving keypoint, 7 warps, 3 CTAs per SM.
//  * Pairs of adjacent rows (y, y+1 at fixed z) of the window are dealt round-robin to the warps,
//    one row per half-warp (static assignment => run-to-run deterministic sums).  Per row the x
//    range is clipped to the sphere chord and the rotated 4x4x4 grid (conservatively, +-1 voxel).
//  * Cheap phase: one lane per voxel applies the reference's exact inclusion tests (sphere
//    :1270, grid :1300-1302) and pushes survivors onto a per-warp ring (ballot/popc ranks).
//  * Heavy phase, whenever 32 voxels are queued (full warps): the reference's per-voxel
//    arithmetic (:1312-1327, :1450-1522) and staging of the 24 (bin, value) contributions
//    (8 trilinear cells x 3 face vertices) in shared memory, rank-compacted.
//  * Replay: the warp walks the staged voxels in order, lane l < 24 adding contribution l into
//    the warp-private histogram — plain LDS/FADD/STS, no atomics (shared FP32 atomics are CAS
//    loops on sm_100).  Warp histograms are summed in fixed order at the end.
//
// Q = true is the production variant: contributions are accumulated in FIXED POINT with native
// integer shared-memory atomics into one per-CTA histogram (order-independent, hence still
// run-to-run deterministic); the scale comes from a 900-sample estimate of the largest gradient
// magnitude in the window, and a keypoint whose contributions would exceed the overflow-safe cap is
// reported in redo_list for the FP32 variant (Q = false: the staged, ordered replay described
// above), which takes its keypoints from klist when given.  Inclusion tests, face selection and
// all per-voxel arithmetic are identical in both variants.
template <bool Q>
__device__ __forceinline__ void describe_one(const int k, const s3d_keypoint* __restrict__ extre, const int* __restrict__ surv,
                                             const LevelTable& tab, const MeshConst* __restrict__ meshp,
                                             s3d_keypoint* __restrict__ kps_out, float* __restrict__ desc_out,
                                             int* __restrict__ redo_list, int* redo_count, float qmargin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename std::conditional<Q, DescSmemQ, DescSmem>::type Smem;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    {
        const int* src = reinterpret_cast<const int*>(extre + surv[k]);
        int* dstp = reinterpret_cast<int*>(&S.kp);
        for (int i = tid; i < (int)(sizeof(s3d_keypoint) / 4); i += kDescThreads) dstp[i] = src[i];
        const int* ms = reinterpret_cast<const int*>(meshp);
        int* md = reinterpret_cast<int*>(&S.M);
        for (int i = tid; i < (int)(sizeof(MeshConst) / 4); i += kDescThreads) md[i] = ms[i];
        if constexpr (Q) {
            for (int i = tid; i < kQCopies * kQCopyStride; i += kDescThreads) S.hist[i] = 0u;
            if (tid == 0) S.overflow = 0;
        } else {
            for (int i = tid; i < kDescWarps * kHistStride; i += kDescThreads) (&S.hist[0][0])[i] = 0.0f;
        }
    }
    __syncthreads();
    const MeshConst& M = S.M;
    const s3d_keypoint& kp = S.kp;
    const int o = kp.octave, lvl = kp.level;
    const int nx = tab.dims[o][0], ny = tab.dims[o][1], nz = tab.dims[o][2];
    const float* g = tab.gss[o * tab.G + lvl];
    const float* wt_g = tab.wtab + tab.desc_off[o * tab.G + lvl];
    const int wt_n = tab.desc_n[o * tab.G + lvl];
    const bool wt_s = wt_n <= kDescWtCap;
    if (wt_s)
        for (int i = tid; i < wt_n; i += kDescThreads) S.wt[i] = wt_g[i];
    __syncthreads();
    const float u = (float)(1 << o);
    const float bary_eps = (float)(FLT_EPSILON * 1E1);          // :23
    const float sigma = kp.scale * 7.071067812f;                // desc_sig_fctr :30,:1155
    const float win_radius = 2.0f * sigma;                      // desc_rad_fctr :31,:1156
    const float desc_hw = (float)((double)win_radius / sqrt(2.0));
    const float desc_width = 2.0f * desc_hw;
    const float desc_bin_fctr = 4.0f / desc_width;
    const float r2 = win_radius * win_radius;
    const float cx = kp.x, cy = kp.y, cz = kp.z;
    // Transpose_Matrix(kp.Rotation) :1214 — R below is the transpose (App. B Q11)
    const float R0 = kp.Rotation[0], R1 = kp.Rotation[3], R2 = kp.Rotation[6];
    const float R3 = kp.Rotation[1], R4 = kp.Rotation[4], R5 = kp.Rotation[7];
    const float R6 = kp.Rotation[2], R7 = kp.Rotation[5], R8 = kp.Rotation[8];
    int xs, xe, y0, y1, z0, z1;
    window_bounds(cx, win_radius / u, nx, xs, xe);
    window_bounds(cy, win_radius / u, ny, y0, y1);
    window_bounds(cz, win_radius / u, nz, z0, z1);
    const int wyn = y1 - y0 + 1, wzn = z1 - z0 + 1;
    const int wyp = (wyn + 1) >> 1;  // row pairs per z slice
    const int npairs = (wyn > 0 && wzn > 0) ? wyp * wzn : 0;
    const ll ys = nx, zs = (ll)nx * ny;
    uint32_t* que = S.queue[wid];
    const float iu = 1.0f / u, inv_u2 = iu * iu;
    // ---- Q: fixed-point scale from a 10 x 10 x 9 lattice of window samples -----------------------
    float qscale = 0.0f;
    if constexpr (Q) {
        float gm = 0.0f;
        const int wxn_ = xe - xs + 1, wyn_ = y1 - y0 + 1, wzn_ = z1 - z0 + 1;
        if (wxn_ > 0 && wyn_ > 0 && wzn_ > 0)
            for (int sidx = tid; sidx < 900; sidx += kDescThreads) {
                const int sx = sidx % 10, sy = (sidx / 10) % 10, sz = sidx / 100;
                const int xx = xs + ((2 * sx + 1) * wxn_) / 20, yy = y0 + ((2 * sy + 1) * wyn_) / 20, zz = z0 + ((2 * sz + 1) * wzn_) / 18;
                const float dx = ((float)xx - cx) * u, dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;
                const float sq = dx * dx + dy * dy + dz * dz;
                if (sq > r2) continue;
                const int mi = (int)(sq * inv_u2);
                const float weight = wt_s ? S.wt[mi] : wt_g[mi];
                const ll i = (ll)xx + (ll)yy * ys + (ll)zz * zs;
                const float gx = 0.5f * (g[i + 1] - g[i - 1]), gy = 0.5f * (g[i + ys] - g[i - ys]), gz = 0.5f * (g[i + zs] - g[i - zs]);
                gm = fmaxf(gm, (gx * gx + gy * gy + gz * gz) * (weight * weight));
            }
        gm = warp_max(gm);
        if (lane == 0) S.red[wid] = gm;
        __syncthreads();
        gm = 0.0f;
#pragma unroll
        for (int w = 0; w < kDescWarps; ++w) gm = fmaxf(gm, S.red[w]);
        __syncthreads();
        const float gest = sqrtf(gm) * iu;  // largest sampled |gradient| * weight
        if (gest > 1e-30f) qscale = kQCap / (qmargin * gest);
        else if (tid == 0) S.overflow = 1;  // nothing sampled: let the FP32 variant decide
    }
    bool q_over = false;
    // conservative clip of a row to the rotated grid: |R_k . disp| < slab for k = 0..2, with
    // R_k . disp = a_k * (x - cx) + (R_k1*dy + R_k2*dz); the reciprocals are per keypoint
    const float slab = desc_hw * 1.001f + 1e-3f;
    const float a0 = R0 * u, a1 = R3 * u, a2 = R6 * u;
    const bool f0 = fabsf(a0) > 1e-6f, f1 = fabsf(a1) > 1e-6f, f2 = fabsf(a2) > 1e-6f;
    const float ia0 = f0 ? 1.0f / a0 : 0.0f, ia1 = f1 ? 1.0f / a1 : 0.0f, ia2 = f2 ? 1.0f / a2 : 0.0f;
    const int half = lane >> 4, hl = lane & 15;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int l24 = lane < 24 ? lane : 0;
    // Q: grid coordinates through fused multiply-adds, vb_k ~ Rf_k . disp + (hw*f - 0.5).  They differ from
    // the reference's unfused evaluation by < 1e-5; a voxel whose coordinate lies within 1e-4 of an
    // inclusion threshold (-0.5 / 3.5) is decided by the exact expression instead, so the SET of
    // contributing voxels is still the reference's.  (sq needs no such care: dx, dy, dz are integers times
    // a power of two and sq < 2^24, so it is exact in any evaluation order.)
    const float Rf0 = R0 * desc_bin_fctr, Rf1 = R1 * desc_bin_fctr, Rf2 = R2 * desc_bin_fctr;
    const float Rf3 = R3 * desc_bin_fctr, Rf4 = R4 * desc_bin_fctr, Rf5 = R5 * desc_bin_fctr;
    const float Rf6 = R6 * desc_bin_fctr, Rf7 = R7 * desc_bin_fctr, Rf8 = R8 * desc_bin_fctr;
    const float hwf = desc_hw * desc_bin_fctr - 0.5f;

    // ---- heavy phase + replay for up to 32 queued voxels (packed = dx | dy<<10 | dz<<20 offsets) ----
    auto heavy = [&](uint32_t packed, bool valid) {
        bool contrib = false;
        float vb0 = 0.f, vb1 = 0.f, vb2 = 0.f, mag = 0.f, b[3] = {0.f, 0.f, 0.f};
        int face = 0;
        if (valid) {
            const int xx = xs + (int)(packed & 1023u), yy = y0 + (int)((packed >> 10) & 1023u), zz = z0 + (int)(packed >> 20);
            const float dx = ((float)xx - cx) * u, dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;
            const float sq = dx * dx + dy * dy + dz * dz;
            if constexpr (Q) {  // cell split only (no inclusion decision here): fused form
                vb0 = __fmaf_rn(Rf0, dx, __fmaf_rn(Rf1, dy, __fmaf_rn(Rf2, dz, hwf)));
                vb1 = __fmaf_rn(Rf3, dx, __fmaf_rn(Rf4, dy, __fmaf_rn(Rf5, dz, hwf)));
                vb2 = __fmaf_rn(Rf6, dx, __fmaf_rn(Rf7, dy, __fmaf_rn(Rf8, dz, hwf)));
            } else {
                vb0 = (R0 * dx + R1 * dy + R2 * dz + desc_hw) * desc_bin_fctr;
                vb1 = (R3 * dx + R4 * dy + R5 * dz + desc_hw) * desc_bin_fctr;
                vb2 = (R6 * dx + R7 * dy + R8 * dz + desc_hw) * desc_bin_fctr;
                vb0 -= 0.5f; vb1 -= 0.5f; vb2 -= 0.5f;
            }
            const int mi = (int)(sq * inv_u2);  // sq = u*u*m exactly, see wtab_kernel
            const float weight = wt_s ? S.wt[mi] : wt_g[mi];  // == expf(-0.5f * sq / s2), :1312
            const ll i = (ll)xx + (ll)yy * ys + (ll)zz * zs;
            float gx = 0.5f * (g[i + 1] - g[i - 1]);  // == (float)(0.5 * (double)(a - b))
            float gy = 0.5f * (g[i + ys] - g[i - ys]);
            float gz = 0.5f * (g[i + zs] - g[i - zs]);
            gx *= iu; gy *= iu; gz *= iu;
            gx = gx * weight; gy = gy * weight; gz = gz * weight;  // SIFT3D_CVEC_SCALE :1322
            const float rx = R0 * gx + R1 * gy + R2 * gz;
            const float ry = R3 * gx + R4 * gy + R5 * gz;
            const float rz = R6 * gx + R7 * gy + R8 * gz;
            const float n2 = rx * rx + ry * ry + rz * rz;  // exact form: input of the reference's |grad|^2 floor
            if (!(n2 < bary_eps)) {  // Check_intersect_faces :1544
                face = find_face(M, rx, ry, rz, bary_eps, b[0], b[1], b[2]);
                if (face >= 0) {
                    contrib = true;
                    mag = sqrtf(n2);
                }
            }
        }
        if constexpr (!Q) {
        float* myh = S.hist[wid];
        uint2(*stg)[kStageCols] = S.stage[wid];
        const unsigned mc = __ballot_sync(0xffffffffu, contrib);
        if (contrib) {
            // Trilinear_interpolation_over_desc_debug :1466-1522
            const int rr = __popc(mc & lt_mask);  // rank among the contributing lanes = stage column
            const int ib0 = (int)vb0, ib1 = (int)vb1, ib2 = (int)vb2;  // truncation, Q12
            const float dv0 = vb0 - floorf(vb0), dv1 = vb1 - floorf(vb1), dv2 = vb2 - floorf(vb2);
            const double wx[2] = {1.0 - (double)dv0, (double)dv0};
            const double wy[2] = {1.0 - (double)dv1, (double)dv1};
            const double wz[2] = {1.0 - (double)dv2, (double)dv2};
            // cell (ib+dd) is inside the 4x4x4 grid iff 0 <= ib+dd <= 3, per axis
            const bool okx[2] = {ib0 >= 0 && ib0 <= 3, ib0 >= -1 && ib0 <= 2};
            const bool oky[2] = {ib1 >= 0 && ib1 <= 3, ib1 >= -1 && ib1 <= 2};
            const bool okz[2] = {ib2 >= 0 && ib2 <= 3, ib2 >= -1 && ib2 <= 2};
            const int ax[2] = {ib0 * 12, ib0 * 12 + 12}, ay[2] = {ib1 * 48, ib1 * 48 + 48}, az[2] = {ib2 * kHistZ, ib2 * kHistZ + kHistZ};
            const int i0 = M.idx[face][0], i1 = M.idx[face][1], i2 = M.idx[face][2];
            uint2* col = &stg[0][rr];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ddx = (c >> 2) & 1, ddy = (c >> 1) & 1, ddz = c & 1;
                const bool ok = okx[ddx] && oky[ddy] && okz[ddz];
                const float wt = (float)(wx[ddx] * wy[ddy] * wz[ddz]);
                const float mw = mag * wt;
                const int base = ok ? ax[ddx] + ay[ddy] + az[ddz] : kHistDump;  // dump: 12 slots, any vertex
                col[(c * 3 + 0) * kStageCols] = make_uint2(base + i0, __float_as_uint(mw * b[0]));
                col[(c * 3 + 1) * kStageCols] = make_uint2(base + i1, __float_as_uint(mw * b[1]));
                col[(c * 3 + 2) * kStageCols] = make_uint2(base + i2, __float_as_uint(mw * b[2]));
            }
        }
        __syncwarp();
        const int cnt = __popc(mc);
        if (cnt) {
            // lane l < 24 adds entry l of each staged voxel, in order; entries are prefetched two ahead
            const uint2* row = stg[l24];
            uint2 e0 = row[0], e1 = row[1];
            int j = 0;
            for (; j + 2 <= cnt; j += 2) {
                const uint2 n0 = row[j + 2], n1 = row[min(j + 3, kStageCols - 1)];
                if (lane < 24) myh[e0.x] += __uint_as_float(e0.y);
                __syncwarp();
                if (lane < 24) myh[e1.x] += __uint_as_float(e1.y);
                __syncwarp();
                e0 = n0; e1 = n1;
            }
            if (j < cnt) {
                if (lane < 24) myh[e0.x] += __uint_as_float(e0.y);
                __syncwarp();
            }
        }
        }
    };

    // ---- Q: two compaction stages.  Stage A (full warps of voxels that passed the inclusion tests): window
    // weight, gradient, rotation into the keypoint frame, the reference's |grad|^2 floor (:1544).  On smooth data a
    // large share of the voxels stops here (ncu on the 512^3 benchmark volume: 18 of 32 lanes were active in
    // everything downstream), so survivors are queued a second time and stage B — face search, barycentric
    // split, the 24 fixed-point atomics — again runs on full warps.  Per-voxel arithmetic is unchanged and
    // integer adds commute, so the descriptors are bit-identical to the single-stage kernel's.
    int q2head = 0, q2n = 0;  // warp-uniform ring state of the second stage
    auto stage_b = [&](uint4 ent, bool valid) {
        if constexpr (Q) {
            if (!valid) return;
            const uint32_t packed = ent.x;
            const float rx = __uint_as_float(ent.y), ry = __uint_as_float(ent.z), rz = __uint_as_float(ent.w);
            const int xx = xs + (int)(packed & 1023u), yy = y0 + (int)((packed >> 10) & 1023u), zz = z0 + (int)(packed >> 20);
            const float dx = ((float)xx - cx) * u, dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;
            // cell split only (no inclusion decision here): fused form
            const float vb0 = __fmaf_rn(Rf0, dx, __fmaf_rn(Rf1, dy, __fmaf_rn(Rf2, dz, hwf)));
            const float vb1 = __fmaf_rn(Rf3, dx, __fmaf_rn(Rf4, dy, __fmaf_rn(Rf5, dz, hwf)));
            const float vb2 = __fmaf_rn(Rf6, dx, __fmaf_rn(Rf7, dy, __fmaf_rn(Rf8, dz, hwf)));
            float b[3] = {0.f, 0.f, 0.f};
            const int face = find_face(M, rx, ry, rz, bary_eps, b[0], b[1], b[2]);
            if (face < 0) return;
            const float n2 = rx * rx + ry * ry + rz * rz;  // the expression stage A tested
            const float mag = sqrtf(n2);
            // Trilinear_interpolation_over_desc_debug :1466-1522, contributions in fixed point
            const float msq = mag * qscale;
            q_over |= msq > kQCap;
            const int ib0 = (int)vb0, ib1 = (int)vb1, ib2 = (int)vb2;  // truncation, Q12
            const float dv0 = vb0 - floorf(vb0), dv1 = vb1 - floorf(vb1), dv2 = vb2 - floorf(vb2);
            const float wx[2] = {1.0f - dv0, dv0}, wy[2] = {1.0f - dv1, dv1}, wz[2] = {1.0f - dv2, dv2};
            const bool okx[2] = {ib0 >= 0 && ib0 <= 3, ib0 >= -1 && ib0 <= 2};
            const bool oky[2] = {ib1 >= 0 && ib1 <= 3, ib1 >= -1 && ib1 <= 2};
            const bool okz[2] = {ib2 >= 0 && ib2 <= 3, ib2 >= -1 && ib2 <= 2};
            const int ax[2] = {ib0 * 12, ib0 * 12 + 12}, ay[2] = {ib1 * 48, ib1 * 48 + 48}, az[2] = {ib2 * kHistZ, ib2 * kHistZ + kHistZ};
            const int i0 = M.idx[face][0], i1 = M.idx[face][1], i2 = M.idx[face][2];
            const float qb0 = msq * b[0], qb1 = msq * b[1], qb2 = msq * b[2];
            uint32_t* hq = S.hist + (lane & (kQCopies - 1)) * kQCopyStride;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int ddx = (c >> 2) & 1, ddy = (c >> 1) & 1, ddz = c & 1;
                if (okx[ddx] && oky[ddy] && okz[ddz]) {
                    const float wt = wx[ddx] * wy[ddy] * wz[ddz];
                    const int base = ax[ddx] + ay[ddy] + az[ddz];
                    // round-to-nearest via the 2^23 trick (values are < 2^23): full-rate FFMA + IADD instead of F2I
                    atomicAdd(&hq[base + i0], __float_as_uint(__fmaf_rn(wt, qb0, 8388608.0f)) - 0x4B000000u);
                    atomicAdd(&hq[base + i1], __float_as_uint(__fmaf_rn(wt, qb1, 8388608.0f)) - 0x4B000000u);
                    atomicAdd(&hq[base + i2], __float_as_uint(__fmaf_rn(wt, qb2, 8388608.0f)) - 0x4B000000u);
                }
            }
        }
    };
    auto stage_a = [&](uint32_t packed, bool valid) {
        if constexpr (Q) {
            uint4* que2 = S.queue2[wid];
            bool pass = false;
            float rx = 0.f, ry = 0.f, rz = 0.f;
            if (valid) {
                const int xx = xs + (int)(packed & 1023u), yy = y0 + (int)((packed >> 10) & 1023u), zz = z0 + (int)(packed >> 20);
                const float dx = ((float)xx - cx) * u, dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;
                const float sq = dx * dx + dy * dy + dz * dz;
                const int mi = (int)(sq * inv_u2);  // sq = u*u*m exactly, see wtab_kernel
                const float weight = wt_s ? S.wt[mi] : wt_g[mi];  // == expf(-0.5f * sq / s2), :1312
                const ll i = (ll)xx + (ll)yy * ys + (ll)zz * zs;
                float gx = 0.5f * (g[i + 1] - g[i - 1]);  // == (float)(0.5 * (double)(a - b))
                float gy = 0.5f * (g[i + ys] - g[i - ys]);
                float gz = 0.5f * (g[i + zs] - g[i - zs]);
                gx *= iu; gy *= iu; gz *= iu;
                gx = gx * weight; gy = gy * weight; gz = gz * weight;  // SIFT3D_CVEC_SCALE :1322
                rx = R0 * gx + R1 * gy + R2 * gz;
                ry = R3 * gx + R4 * gy + R5 * gz;
                rz = R6 * gx + R7 * gy + R8 * gz;
                const float n2 = rx * rx + ry * ry + rz * rz;  // exact form: input of the reference's |grad|^2 floor
                pass = !(n2 < bary_eps);  // Check_intersect_faces :1544
            }
            const unsigned mp = __ballot_sync(0xffffffffu, pass);
            if (pass)
                que2[(q2head + q2n + __popc(mp & lt_mask)) & 63] =
                    make_uint4(packed, __float_as_uint(rx), __float_as_uint(ry), __float_as_uint(rz));
            q2n += __popc(mp);
            __syncwarp();
            if (q2n >= 32) {
                const uint4 ent = que2[(q2head + lane) & 63];
                __syncwarp();  // every lane holds its entry before any lane may push into the freed slots again
                stage_b(ent, true);
                q2head = (q2head + 32) & 63;
                q2n -= 32;
            }
        }
    };

    int qhead = 0, qn = 0;  // warp-uniform ring state
    // Row pairs rp = wid, wid + kDescWarps, ... belong to this warp; pair t's rows go to the two half-warps.
    // The per-row set-up (sphere chord, clip against the rotated grid, row constants: ~160 instructions) is
    // done LANE-PARALLEL for 16 pairs at a time — lane L prepares row (L & 1) of pair t0 + (L >> 1) — and
    // handed to the half-warps by shuffles; the first version had all 32 lanes repeat it for their 2 rows
    // (16 % of the kernel's instructions).  Row order, hence queue order, is unchanged.
    const int nmy = npairs > wid ? (npairs - wid + kDescWarps - 1) / kDescWarps : 0;
    for (int t0 = 0; t0 < nmy; t0 += 16) {
        int xlo_l = 1, xhi_l = 0;  // empty
        uint32_t pyz_l = 0;
        float dyz2_l = 0.0f, rb0_l = 0.0f, rb1_l = 0.0f, rb2_l = 0.0f;
        {
            const int t = t0 + (lane >> 1);
            if (t < nmy) {
                const int rp = wid + t * kDescWarps;
                const int pz = rp / wyp, py = rp - pz * wyp;
                const int yy = y0 + 2 * py + (lane & 1), zz = z0 + pz;
                const float dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;
                // fl(fl(dx^2+dy^2)+dz^2) >= fl(dy^2+dz^2) by monotonicity of rounding: safe row reject
                const float dyz2 = dy * dy + dz * dz;
                if (yy <= y1 && !(dyz2 > r2)) {
                    float lo = -(sqrtf(r2 - dyz2) * iu), hi = -lo;  // sphere chord, in voxels about cx
                    bool empty = false;
                    const float c0 = R1 * dy + R2 * dz, c1 = R4 * dy + R5 * dz, c2 = R7 * dy + R8 * dz;
#define S3D_CLIP(fk, ck, iak)                                                    \
                    if (fk) {                                                            \
                        const float t0c = (-slab - ck) * iak, t1c = (slab - ck) * iak;   \
                        lo = fmaxf(lo, fminf(t0c, t1c)); hi = fminf(hi, fmaxf(t0c, t1c)); \
                    } else if (fabsf(ck) > slab) empty = true;
                    S3D_CLIP(f0, c0, ia0) S3D_CLIP(f1, c1, ia1) S3D_CLIP(f2, c2, ia2)
#undef S3D_CLIP
                    if (!empty && !(lo > hi + 2.0f)) {
                        xlo_l = max(xs, (int)floorf(cx + lo) - 1);
                        xhi_l = min(xe, (int)ceilf(cx + hi) + 1);
                    }
                }
                pyz_l = ((uint32_t)(yy - y0) << 10) | ((uint32_t)(zz - z0) << 20);
                dyz2_l = dyz2;
                // row constants of the fused grid coordinates (Q)
                rb0_l = __fmaf_rn(Rf1, dy, __fmaf_rn(Rf2, dz, hwf));
                rb1_l = __fmaf_rn(Rf4, dy, __fmaf_rn(Rf5, dz, hwf));
                rb2_l = __fmaf_rn(Rf7, dy, __fmaf_rn(Rf8, dz, hwf));
            }
        }
        const int nb = min(16, nmy - t0);
        for (int j = 0; j < nb; ++j) {
        const int srcl = 2 * j + half;
        const int xlo = __shfl_sync(0xffffffffu, xlo_l, srcl), xhi = __shfl_sync(0xffffffffu, xhi_l, srcl);
        const uint32_t pyz = __shfl_sync(0xffffffffu, pyz_l, srcl);
        const float dyz2 = __shfl_sync(0xffffffffu, dyz2_l, srcl);
        const float rb0 = __shfl_sync(0xffffffffu, rb0_l, srcl), rb1 = __shfl_sync(0xffffffffu, rb1_l, srcl),
                    rb2 = __shfl_sync(0xffffffffu, rb2_l, srcl);
        const int len = xhi - xlo + 1;
        const int len_other = __shfl_xor_sync(0xffffffffu, len, 16);
        const int iters = (max(len, len_other) + 15) >> 4;
        const int yy = y0 + (int)((pyz >> 10) & 1023u), zz = z0 + (int)(pyz >> 20);
        const float dy = ((float)yy - cy) * u, dz = ((float)zz - cz) * u;  // exact_pass / FP32 variant
        for (int it = 0; it < iters; ++it) {
            const int xx = xlo + it * 16 + hl;
            bool pass = false;
            if (xx <= xhi) {
                const float dx = ((float)xx - cx) * u;
                auto exact_pass = [&]() -> bool {
                    float vb0 = (R0 * dx + R1 * dy + R2 * dz + desc_hw) * desc_bin_fctr;
                    float vb1 = (R3 * dx + R4 * dy + R5 * dz + desc_hw) * desc_bin_fctr;
                    float vb2 = (R6 * dx + R7 * dy + R8 * dz + desc_hw) * desc_bin_fctr;
                    vb0 -= 0.5f; vb1 -= 0.5f; vb2 -= 0.5f;
                    return !(vb0 <= -0.5f || vb1 <= -0.5f || vb2 <= -0.5f || vb0 >= 3.5f || vb1 >= 3.5f || vb2 >= 3.5f);  // :1300
                };
                if constexpr (Q) {
                    const float sq = __fmaf_rn(dx, dx, dyz2);  // exact (see above)
                    if (!(sq > r2)) {  // :1270
                        const float vb0 = __fmaf_rn(Rf0, dx, rb0), vb1 = __fmaf_rn(Rf3, dx, rb1), vb2 = __fmaf_rn(Rf6, dx, rb2);
                        // distance of the three coordinates to the interval (-0.5, 3.5): inside by more than the margin /
                        // outside by more than the margin / too close to call
                        const float lo3 = fminf(vb0, fminf(vb1, vb2)), hi3 = fmaxf(vb0, fmaxf(vb1, vb2));
                        const float inside = fminf(lo3 + 0.5f, 3.5f - hi3);
                        pass = inside > 1e-4f;
                        if (fabsf(lo3 + 0.5f) <= 1e-4f || fabsf(3.5f - hi3) <= 1e-4f) pass = exact_pass();
                    }
                } else {
                    const float sq = dx * dx + dy * dy + dz * dz;
                    if (!(sq > r2)) pass = exact_pass();  // :1270
                }
            }
            const unsigned mp = __ballot_sync(0xffffffffu, pass);
            if (pass) que[(qhead + qn + __popc(mp & lt_mask)) & 63] = (uint32_t)(xx - xs) | pyz;
            qn += __popc(mp);
            __syncwarp();
            if (qn >= 32) {
                if constexpr (Q) stage_a(que[(qhead + lane) & 63], true);
                else heavy(que[(qhead + lane) & 63], true);
                qhead = (qhead + 32) & 63;
                qn -= 32;
            }
        }
        }
    }
    if constexpr (Q) {
        if (qn > 0) stage_a(que[(qhead + lane) & 63], lane < qn);
        if (q2n > 0) stage_b(S.queue2[wid][(q2head + lane) & 63], lane < q2n);
    } else {
        if (qn > 0) heavy(que[(qhead + lane) & 63], lane < qn);
    }
    if constexpr (Q) {
        if (q_over) S.overflow = 1;
    }
    __syncthreads();
    if constexpr (Q) {
        if (S.overflow) {  // CTA-uniform: hand this keypoint to the FP32 variant
            if (tid == 0) redo_list[atomicAdd(redo_count, 1)] = k;
            return;
        }
    }
    // fixed-order sum over the per-warp histograms, then normalise / clamp / normalise (:1350-1358)
    constexpr int kPer = (S3D_DESC_LEN + kDescThreads - 1) / kDescThreads;
    float v[kPer];
    float ss = 0.0f;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const int i = tid + e * kDescThreads;        // output index (x + 4y + 16z)*12 + v
        float a = 0.0f;
        if (i < S3D_DESC_LEN) {
            const int hz = i / 192, hr = i - hz * 192;  // -> padded histogram address
            if constexpr (Q) {
                uint32_t acc = 0;  // the copies together stay below 2^32 (see kQCap)
#pragma unroll
                for (int cp = 0; cp < kQCopies; ++cp) acc += S.hist[cp * kQCopyStride + hz * kHistZ + hr];
                a = (float)acc;  // in units of 1/qscale; the normalisation below removes the scale
            } else {
#pragma unroll
                for (int w = 0; w < kDescWarps; ++w) a += S.hist[w][hz * kHistZ + hr];
            }
        }
        v[e] = a;
        ss += a * a;
    }
    auto block_sum = [&](float x) -> float {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        __syncthreads();
        if (lane == 0) S.red[wid] = x;
        __syncthreads();
        float tsum = 0.0f;
#pragma unroll
        for (int w = 0; w < kDescWarps; ++w) tsum += S.red[w];
        return tsum;
    };
    float norm = block_sum(ss);
    norm = (float)((double)sqrtf(norm) + DBL_EPSILON);
    float norm_inv = (float)(1.0 / (double)norm);
    const float trunc_thresh = (float)(0.2 * 128 / S3D_DESC_LEN);
    ss = 0.0f;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        v[e] *= norm_inv;
        v[e] = v[e] < trunc_thresh ? v[e] : trunc_thresh;
        ss += v[e] * v[e];
    }
    norm = block_sum(ss);
    norm = (float)((double)sqrtf(norm) + DBL_EPSILON);
    norm_inv = (float)(1.0 / (double)norm);
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const int i = tid + e * kDescThreads;
        if (i < S3D_DESC_LEN) desc_out[(size_t)k * S3D_DESC_LEN + i] = v[e] * norm_inv;
    }
    if (tid == 0) {
        s3d_keypoint outk = kp;
        outk.Rotation[0] = R0; outk.Rotation[1] = R1; outk.Rotation[2] = R2;
        outk.Rotation[3] = R3; outk.Rotation[4] = R4; outk.Rotation[5] = R5;
        outk.Rotation[6] = R6; outk.Rotation[7] = R7; outk.Rotation[8] = R8;
        const float coord_factor = (float)(1 << o);  // pow(2.0, octave) :1161
        outk.rx = kp.x * coord_factor; outk.ry = kp.y * coord_factor; outk.rz = kp.z * coord_factor;  // :1377-1379
        outk.desc = nullptr;
        kps_out[k] = outk;
    }
}

// The kernel proper: one CTA per keypoint slot of the launch.  The number of keypoints is produced on the device and
// never visits the host before the launch: the grid covers the CAPACITY of the keypoint buffers and the surplus CTAs
// leave at once (a few ns each).  item -> slot k = klist[item] (heavy-first order) or item; slots k >= slot_cap do not
// exist (the optimistic capacity was too small: the host repeats the stage with exact sizes, s3d_wait).
// (measured: persistent CTAs drawing keypoints from an atomic counter cost 28 bytes of spills at the 72-register budget
// and ran 2.5 % slower than the hardware's own CTA scheduler: 9.78 vs 9.53 ms on the 512^3 benchmark volume)
template <bool Q>
__global__ void __launch_bounds__(kDescThreads, Q ? kQMinCtas : 3) describe_kernel(const s3d_keypoint* __restrict__ extre,
                                                                   const int* __restrict__ surv, int nkp_max, LevelTable tab,
                                                                   const MeshConst* __restrict__ meshp,
                                                                   s3d_keypoint* __restrict__ kps_out,
                                                                   float* __restrict__ desc_out,
                                                                   const int* __restrict__ klist,
                                                                   const int* __restrict__ nkp_dev,
                                                                   int* __restrict__ redo_list, int* redo_count,
                                                                   float qmargin, int slot_cap) {
    const int item = (int)blockIdx.x;
    if (item >= (nkp_dev ? min(*nkp_dev, nkp_max) : nkp_max)) return;
    const int k = klist ? klist[item] : item;
    if (k >= slot_cap) return;
    describe_one<Q>(k, extre, surv, tab, meshp, kps_out, desc_out, redo_list, redo_count, qmargin);
}

// FP-contract self test: with -fmad=false, a*b+c must round twice.
__global__ void selftest_kernel(float a, float b, float c, float d, float* out) {
    out[0] = a * b + c;   // must equal fadd(fmul(a,b),c), not fma
    out[1] = __fadd_rn(__fmul_rn(a, b), c);
    out[2] = fmaf(a, b, c);
    out[3] = a / d;
    out[4] = __fdiv_rn(a, d);
    out[5] = sqrtf(d);
    out[6] = __fsqrt_rn(d);
    out[7] = s3d_expf_ref(-d);
}

}  // namespace s3d
