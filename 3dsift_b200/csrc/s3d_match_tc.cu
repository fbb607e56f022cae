// s3d_match_tc.cu — tensor-core candidate search for the brute-force matcher (sm_100a only).
//
// The reference's calMatches (/root/reference/3DSIFT/Src/cMatcher.cc:40-79) is a dense
// 768-deep contraction followed by a per-row best/second-best selection.  Here:
//
//   1. cvt_f16_kernel     descriptors -> fp16, scaled by 64 (all entries are >= 0 and <= 1).
//   2. tc_topk_kernel     S = Q16 . DB16^T on the 5th-gen tensor cores: tcgen05.mma
//                         (cta_group::1, kind::f16, M=128, N=256, K=16) issued by one thread,
//                         operands staged by TMA (128B-swizzled 64-wide K blocks, 4-stage
//                         mbarrier ring), FP32 accumulators double-buffered in TMEM.  The epilogue
//                         warps read the accumulators with tcgen05.ld and keep a running top-8
//                         (approximate dot, index) per query row in registers; the N x M score
//                         matrix is never materialised.
//   3. rerank_kernel      exact dots of the candidates in the reference's arithmetic (float
//                         product, sequential double sum, KP_squareSum cMatcher.cc:17-23), exact
//                         top-2 under (dot desc, index asc), and the GUARD: every non-candidate j
//                         has approx_j <= a8 (the smallest kept approximate value), hence
//                         exact_j <= a8*(1+eps)+eps_abs with eps bounding the fp16 input rounding
//                         and FP32 accumulation error for non-negative data.  If the exact
//                         second-best exceeds that bound the candidate set provably contains the
//                         true top-2; otherwise the row is flagged and recomputed by the exact
//                         CUDA-core kernel (s3d_match.cu).  Results are therefore identical to the
//                         exact path on every input.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cfloat>

#include "s3d_common.h"
#include "s3d_match_internal.h"

namespace s3d {

namespace tc {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, KD = S3D_DESC_LEN, KBLKS = KD / BK;
constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TOPK = 8;
constexpr float SCALE = 64.0f, INV_SCALE2 = 1.0f / (64.0f * 64.0f);
constexpr int kThreads = 192;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + 256;

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, FP32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cluster / 2-CTA (cta_group::2) variants ------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `local` (a shared::cta address of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on an mbarrier anywhere in the cluster (address from mapa)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load into THIS CTA's shared memory that signals an mbarrier in this CTA or its peer
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t cluster_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(cluster_bar)
        : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy encodings as used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
__device__ __forceinline__ void tma_load_2d_pair_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t cluster_bar,
                                                      uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
        "[%4], %5;" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(cluster_bar), "l"(hint)
        : "memory");
}
// commit of the pair's MMAs: arrives on the mbarrier at the same shared offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B, 64 fp16 (128 B) per row, 8-row atoms
// 1024 B apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=2 (SWIZZLE_128B) [61,64)).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;             // LBO (ignored for swizzled K-major layouts)
    d |= (uint64_t)(1024 >> 4) << 32;   // SBO
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b format F16 (0),
// a/b K-major (0), N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Epilogue of one accumulator tile, shared by both candidate kernels: fold the BN columns of this
// thread's row (TMEM lane) into the running top-8 (RAW accumulator value = 4096 x approximate dot,
// database index), kept sorted in registers.  j0 = database index of column 0.
//   * the TMEM loads are software-pipelined (chunk c+1 is in flight while chunk c is reduced);
//   * the common case costs a 32-wide max (FMNMX3) and one compare per chunk;
//   * a chunk that does hold a new candidate is handled WITHOUT per-element branches: locate the
//     maximum (select chain), insert it, knock it out, re-reduce — one round per insertion, and
//     insertions become rare as the 8th-best value rises (a warp leaves the fast path only while
//     one of its 32 rows still improves).
__device__ __forceinline__ void topk_insert(float x, int j, float (&v)[TOPK], int (&id)[TOPK]) {
    v[TOPK - 1] = x; id[TOPK - 1] = j;
#pragma unroll
    for (int i = TOPK - 1; i > 0; --i) {
        const bool sw = v[i] > v[i - 1];
        const float hi = sw ? v[i] : v[i - 1], lo = sw ? v[i - 1] : v[i];
        const int ihi = sw ? id[i] : id[i - 1], ilo = sw ? id[i - 1] : id[i];
        v[i - 1] = hi; v[i] = lo; id[i - 1] = ihi; id[i] = ilo;
    }
}

__device__ __forceinline__ void topk_reduce32(uint32_t (&r)[32], int j0, int nvalid, float (&v)[TOPK], int (&id)[TOPK]) {
    float mx = __uint_as_float(r[0]);
#pragma unroll
    for (int e = 1; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
    while (mx > v[TOPK - 1]) {
        int idx = 31;
#pragma unroll
        for (int e = 30; e >= 0; --e) idx = (__uint_as_float(r[e]) == mx) ? e : idx;
        if (idx < nvalid) topk_insert(mx, j0 + idx, v, id);
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = (e == idx) ? 0xff800000u /* -inf */ : r[e];
        mx = __uint_as_float(r[0]);
#pragma unroll
        for (int e = 1; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
    }
}

__device__ __forceinline__ void topk_scan_tile(uint32_t taddr, int j0, int nd, float (&v)[TOPK], int (&id)[TOPK]) {
    uint32_t ra[32], rb[32];
    tc_ld32(taddr, ra);
#pragma unroll 1
    for (int c = 0; c < BN / 32; c += 2) {
        tc_wait_ld();
        tc_ld32(taddr + (uint32_t)((c + 1) * 32), rb);
        topk_reduce32(ra, j0 + c * 32, nd - (j0 + c * 32), v, id);
        tc_wait_ld();
        if (c + 2 < BN / 32) tc_ld32(taddr + (uint32_t)((c + 2) * 32), ra);
        topk_reduce32(rb, j0 + (c + 1) * 32, nd - (j0 + (c + 1) * 32), v, id);
    }
}

// FP32 -> FP16 (x64) copy for the candidate pass.  The guard of the exact re-rank (rerank_kernel) bounds the error of
// the FP16 dot product RELATIVE to the dot product itself, which only holds when no term can cancel another and no
// entry leaves the range the scaling was chosen for: descriptors as the reference produces them are non-negative with
// entries <= 1 (normalise / clamp / normalise, Src/cSIFT3D.cc:1350-1358).  s3d_match accepts arbitrary floats, so the
// conversion also CHECKS the precondition: any entry that is negative, above 1 or not finite raises *bad, and the
// caller sends the whole search to the exact kernel instead (tc_search returns kTcRefused).
__global__ void cvt_f16_kernel(const float* __restrict__ src, const int* __restrict__ rows, int nrows, __half* __restrict__ dst,
                               int* __restrict__ bad) {
    // one warp per row (768 floats = 6 float4 per lane)
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nrows) return;
    const float4* s = reinterpret_cast<const float4*>(src + (size_t)(rows ? rows[w] : w) * KD);
    uint2* d = reinterpret_cast<uint2*>(dst + (size_t)w * KD);
    bool out_of_range = false;
#pragma unroll
    for (int i = 0; i < KD / 4 / 32; ++i) {
        const float4 v = s[lane + 32 * i];
        // !(0 <= x <= 1) is true for negative, > 1, NaN and +-inf entries (-0.0f passes: it is zero)
        out_of_range |= !(v.x >= 0.0f && v.x <= 1.0f) || !(v.y >= 0.0f && v.y <= 1.0f) || !(v.z >= 0.0f && v.z <= 1.0f) ||
                        !(v.w >= 0.0f && v.w <= 1.0f);
        const __half2 a = __floats2half2_rn(v.x * SCALE, v.y * SCALE), b = __floats2half2_rn(v.z * SCALE, v.w * SCALE);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned*>(&a);
        o.y = *reinterpret_cast<const unsigned*>(&b);
        d[lane + 32 * i] = o;
    }
    if (__any_sync(0xffffffffu, out_of_range) && lane == 0) *bad = 1;
}

// Work item = (query tile qt, database part p): the CTA sweeps db tiles [p*tiles_per_part, ...) for
// its 128 query rows and writes each row's top-8 of that part to cand_*[row][p*8 + k].
__global__ void __launch_bounds__(kThreads, 1) tc_topk_kernel(const __grid_constant__ CUtensorMap map_q,
                                                              const __grid_constant__ CUtensorMap map_db, int nq, int nd,
                                                              int n_qtiles, int parts, int tiles_per_part, int n_dbtiles,
                                                              float* __restrict__ cand_val, int* __restrict__ cand_idx) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    // barriers: full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base address
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
    }
    if (warp == 1) {  // 512 TMEM columns = two 128x256 FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int n_items = n_qtiles * parts;
    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int qt = item / parts, p = item - qt * parts;
                const int t0 = p * tiles_per_part, t1 = min(n_dbtiles, t0 + tiles_per_part);
                for (int t = t0; t < t1; ++t)
                    for (int kb = 0; kb < KBLKS; ++kb) {
                        mbar_wait(empty(stage), phase ^ 1);
                        mbar_expect_tx(full(stage), STAGE_BYTES);
                        const uint32_t sa = base + stage * STAGE_BYTES;
                        tma_load_2d(sa, &map_q, kb * BK, qt * BM, full(stage));
                        tma_load_2d(sa + A_BYTES, &map_db, kb * BK, t * BN, full(stage));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int qt = item / parts, p = item - qt * parts;
                const int t0 = p * tiles_per_part, t1 = min(n_dbtiles, t0 + tiles_per_part);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(tempty(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int kb = 0; kb < KBLKS; ++kb) {
                        mbar_wait(full(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = base + stage * STAGE_BYTES;
                        const uint64_t da = make_desc(sa), db = make_desc(sa + A_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < BK / 16; ++k4)  // advance 32 B (= 2 x 16 B units) per K=16 step
                            tc_mma(d_tmem, da + (uint64_t)(2 * k4), db + (uint64_t)(2 * k4), idesc, (kb | k4) ? 1u : 0u);
                        tc_commit(empty(stage));  // frees the smem stage once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull(acc));  // accumulator complete -> epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +31 (row = lane of the tile) =====
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int qt = item / parts, p = item - qt * parts;
            const int t0 = p * tiles_per_part, t1 = min(n_dbtiles, t0 + tiles_per_part);
            const int row = qt * BM + quad * 32 + lane;
            float v[TOPK];
            int id[TOPK];
#pragma unroll
            for (int i = 0; i < TOPK; ++i) { v[i] = -1.0f; id[i] = -1; }
            for (int t = t0; t < t1; ++t) {
                mbar_wait(tfull(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
                topk_scan_tile(taddr, t * BN, nd, v, id);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc));  // 4 epilogue warps -> count 4
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < nq) {
                const size_t o = ((size_t)row * parts + p) * TOPK;
#pragma unroll
                for (int i = 0; i < TOPK; ++i) { cand_val[o + i] = id[i] >= 0 ? v[i] * INV_SCALE2 : -1.0f; cand_idx[o + i] = id[i]; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair candidate kernel (cta_group::2): the query tile stays RESIDENT in shared memory.
//
// With one CTA per tile the kernel above streams 16 KB of queries + 32 KB of database rows per
// 512 tensor clocks = 96 B/clk/SM, more than twice what L2 can deliver to every SM at once
// (~6300 B/clk chip-wide = 42.6 B/clk/SM), so the tensor pipe idles ~60 % of the time.  Here two
// CTAs of one TPC form a pair: each keeps its own 128 query rows x 768 (fp16, 192 KB, 12 blocks of
// 128 B-swizzled 64-wide K) in shared memory for the whole sweep, and each stages only HALF of the
// database rows of an MMA (64 of N = 128) — tcgen05.mma.cta_group::2 (M = 256) reads B from both
// CTAs' shared memory.  Streaming traffic drops to 8 KB per 256 tensor clocks = 32 B/clk/SM.
//   ring      4 stages x 8 KB (64 database rows x 64 K), TMA with .cta_group::2 so the peer's loads
//             complete on the LEADER's full barrier; empty barriers are signalled in both CTAs by
//             tcgen05.commit ... multicast::cluster
//   TMEM      512 columns = two 128-lane x 256-column FP32 accumulators per CTA (double-buffered
//             database tiles; a tile is two N = 128 halves)
//   epilogue  4 warps per CTA, the same running top-8 as above on the CTA's own 128 rows; drained
//             accumulators are released with a remote arrive on the leader's barrier.
// Work unit = (pair of query tiles, database part), owned by ONE cluster for the whole launch.  The
// database part is swept in L2-sized chunks, chunk-major: every cluster runs all its units over
// chunk k before anyone moves to chunk k+1, so the ~150 SMs stream the SAME few tens of MB at the
// same time and the ring's TMA loads hit L2 (a database of 1 M rows is 1.5 GB in fp16; swept
// unit-major it would be re-read from HBM by every cluster, 9 TB/s of demand).  A unit's running
// top-8 is parked in the candidate arrays between chunks (same thread writes and re-reads it).
// Query tiles are loaded with an evict-first hint so they do not displace the chunk.
// ---------------------------------------------------------------------------------------------
namespace pr {
constexpr int BMC = 128, BNH = 128, BROWS = 64, BSTAGES = 4;
constexpr uint32_t A_BLK = BMC * BK * 2, A_ALL = KBLKS * A_BLK, B_STAGE = BROWS * BK * 2;
constexpr size_t SMEM_BYTES = 1024 + (size_t)A_ALL + (size_t)BSTAGES * B_STAGE + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
constexpr int kChunkTiles = 64;  // 64 tiles x 256 rows x 1536 B = 24 MB of fp16 database rows per chunk
}  // namespace pr

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    tc_pair_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, int nq, int nd,
                        int n_qpairs, int parts, int tiles_per_part, int n_dbtiles, int chunk_tiles,
                        float* __restrict__ cand_val, int* __restrict__ cand_idx) {
    using namespace pr;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_b = base + A_ALL;
    const uint32_t bars = smem_b + BSTAGES * B_STAGE;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (BSTAGES + s); };
    auto tfull = [&](int a) { return bars + 8u * (2 * BSTAGES + a); };
    auto tempty = [&](int a) { return bars + 8u * (2 * BSTAGES + 2 + a); };
    const uint32_t afull = bars + 8u * (2 * BSTAGES + 4), aempty = bars + 8u * (2 * BSTAGES + 5);
    const uint32_t tmem_slot = bars + 8u * (2 * BSTAGES + 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < BSTAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }  // 4 epilogue warps x 2 CTAs
        mbar_init(afull, 1);
        mbar_init(aempty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_db) : "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates the pair's 512 columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits and the allocation are visible to the peer before any remote arrive
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int n_units = n_qpairs * parts;
    const int n_chunks = (tiles_per_part + chunk_tiles - 1) / chunk_tiles;
    // every role walks the same (chunk, unit) sequence; empty chunk ranges are skipped by all alike
    auto for_items = [&](auto&& body) {
        for (int ck = 0; ck < n_chunks; ++ck)
            for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
                const int qp = unit / parts, p = unit - qp * parts;
                const int pt1 = min(n_dbtiles, (p + 1) * tiles_per_part);
                const int t0 = p * tiles_per_part + ck * chunk_tiles, t1 = min(pt1, t0 + chunk_tiles);
                if (t0 < t1) body(ck, qp, p, t0, t1);
            }
    };
    if (warp == 0) {
        // ===== TMA producer (one thread in EACH CTA; transaction bytes land on the leader's barriers) =====
        if (lane == 0) {
            const uint32_t afull_l = mapa(afull, 0);
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            for_items([&](int ck, int qp, int p, int t0, int t1) {
                mbar_wait(aempty, aphase ^ 1);  // the previous item's MMAs no longer read the resident tile
                if (leader) mbar_expect_tx(afull, 2u * A_ALL);
                for (int kb = 0; kb < KBLKS; ++kb)
                    tma_load_2d_pair_hint(base + kb * A_BLK, &map_q, kb * BK, qp * (2 * BMC) + (int)rank * BMC, afull_l, kEvictFirst);
                aphase ^= 1;
                for (int t = t0; t < t1; ++t)
                    for (int h = 0; h < 2; ++h)
                        for (int kb = 0; kb < KBLKS; ++kb) {
                            mbar_wait(empty(stage), phase ^ 1);
                            if (leader) mbar_expect_tx(full(stage), 2u * B_STAGE);
                            tma_load_2d_pair(smem_b + stage * B_STAGE, &map_db, kb * BK, t * BN + h * BNH + (int)rank * BROWS,
                                             mapa(full(stage), 0));
                            if (++stage == BSTAGES) { stage = 0; phase ^= 1; }
                        }
            });
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(2 * BMC, BNH);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, aphase = 0;
            for_items([&](int ck, int qp, int p, int t0, int t1) {
                mbar_wait(afull, aphase);
                aphase ^= 1;
                tc_fence_after();
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(tempty(acc), acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
                    tc_fence_after();
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN + h * BNH);
                        for (int kb = 0; kb < KBLKS; ++kb) {
                            mbar_wait(full(stage), phase);
                            tc_fence_after();
                            const uint64_t da = make_desc(base + kb * A_BLK), db = make_desc(smem_b + stage * B_STAGE);
#pragma unroll
                            for (int k4 = 0; k4 < BK / 16; ++k4)
                                tc_mma_pair(d_tmem, da + (uint64_t)(2 * k4), db + (uint64_t)(2 * k4), idesc, (kb | k4) ? 1u : 0u);
                            tc_commit_pair(empty(stage));
                            if (++stage == BSTAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                    tc_commit_pair(tfull(acc));
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                tc_commit_pair(aempty);
            });
        }
    } else {
        // ===== epilogue (both CTAs): warps 2..5 own TMEM lanes 32*(warp%4) .. +31 of their CTA =====
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t tempty_l[2] = {mapa(tempty(0), 0), mapa(tempty(1), 0)};
        for_items([&](int ck, int qp, int p, int t0, int t1) {
            const int row = qp * (2 * BMC) + (int)rank * BMC + quad * 32 + lane;
            const size_t o = ((size_t)row * parts + p) * TOPK;
            float v[TOPK];
            int id[TOPK];
            if (ck > 0 && row < nq) {  // resume the running top-8 parked after the previous chunk
#pragma unroll
                for (int i = 0; i < TOPK; ++i) {
                    id[i] = cand_idx[o + i];
                    v[i] = id[i] >= 0 ? cand_val[o + i] * (SCALE * SCALE) : -1.0f;  // back to raw accumulator units (exact)
                }
            } else {
#pragma unroll
                for (int i = 0; i < TOPK; ++i) { v[i] = -1.0f; id[i] = -1; }
            }
            for (int t = t0; t < t1; ++t) {
                mbar_wait(tfull(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
                topk_scan_tile(taddr, t * BN, nd, v, id);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_l[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < nq) {
#pragma unroll
                for (int i = 0; i < TOPK; ++i) { cand_val[o + i] = id[i] >= 0 ? v[i] * INV_SCALE2 : -1.0f; cand_idx[o + i] = id[i]; }
            }
        });
    }
    tc_fence_before();
    cluster_sync_all();  // neither CTA may leave (or free TMEM) while the other can still signal it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Exact re-rank.  Four queries per warp, eight lanes per query:
//   1. merge  the parts' top-8 lists of a query are merged to ONE top-8 by approximate value (8
//             rounds of an 8-lane arg-max), so the exact work does not grow with the number of parts.
//   2. exact  dots of the 32 (query, candidate) pairs of the warp in the reference's arithmetic
//             (float product, sequential double sum over k = 0..767, KP_squareSum cMatcher.cc:17-23).
//             Rows are read COALESCED — for each pair the warp loads 32 consecutive k of both rows
//             and parks the 32 products in a shared-memory tile; lane l then adds the 32 products of
//             pair l in order.  (One lane walking one database row touches a new 32-byte sector
//             every 8 elements: 8x the traffic, which made this kernel L2-bound.)
//   3. top-2  under (dot desc, index asc) over the query's eight lanes, and the GUARD: every row that
//             is not among the merged eight has approximate dot <= a8 (the 8th merged value, which is
//             >= each full part's own 8th), hence exact dot <= a8*(1+eps)+eps_abs; if the exact
//             second-best exceeds that bound the eight provably contain the true top-2, otherwise the
//             row goes to the exact CUDA-core kernel.
// q rows are addressed through qlist (original row index); results are written to out[orig].
// The candidate lists of query qi are `parts` (<= 8) top-8 lists at cand[qi * stride_q + p * stride_p + 0..7]: the parts
// of one tensor-core pass ([query][part][8]: stride_q = parts * 8, stride_p = 8) or, in the database-sharded search, one
// merged list per rank as they arrive from the ranks ([rank][query][8]: stride_q = 8, stride_p = rows * 8).
// MERGE_ONLY: stop after step 1 and write the merged list (value, index + idx_offset) to mrg_val / mrg_idx [query][8]
// (what a rank sends to the rank that re-ranks the query).
constexpr int kRrWarps = 8;
template <bool MERGE_ONLY>
__global__ void __launch_bounds__(kRrWarps * 32) rerank_kernel(const float* __restrict__ q, const int* __restrict__ qlist, int nql,
                                                               const float* __restrict__ db, int nd, int db_offset, int parts,
                                                               size_t stride_q, size_t stride_p,
                                                               const float* __restrict__ cand_val,
                                                               const int* __restrict__ cand_idx, Top2* __restrict__ out,
                                                               int* __restrict__ fb_list, int* fb_count,
                                                               float* __restrict__ mrg_val, int* __restrict__ mrg_idx, int idx_offset) {
    __shared__ float tile[MERGE_ONLY ? 1 : kRrWarps][32][33];
    __shared__ int pair_j[kRrWarps][32];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const int gw = blockIdx.x * kRrWarps + wid;
    if (gw * 4 >= nql) return;  // warp-uniform
    const int qi = gw * 4 + g;
    const bool qvalid = qi < nql;
    const int orig = qvalid ? (qlist ? qlist[qi] : qi) : 0;
    // ---- 1. merge ----
    float ev[8];
    int ej[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const bool ok = qvalid && p < parts;
        ej[p] = ok ? cand_idx[(size_t)qi * stride_q + p * stride_p + c] : -1;
        ev[p] = ok ? cand_val[(size_t)qi * stride_q + p * stride_p + c] : -1.0f;
    }
    int myj = -1;
    float myv = -1.0f;
    float a8 = -1.0f;
#pragma unroll 1
    for (int r = 0; r < TOPK; ++r) {
        float lv = -2.0f;
        int lp = -1;
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (ej[p] >= 0 && ev[p] > lv) { lv = ev[p]; lp = p; }
        float gv = lv;
        int gl = c;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, gv, off);
            const int ol = __shfl_xor_sync(0xffffffffu, gl, off);
            if (ov > gv || (ov == gv && ol < gl)) { gv = ov; gl = ol; }
        }
        int lj = -1;
#pragma unroll
        for (int p = 0; p < 8; ++p)
            if (p == lp) lj = ej[p];
        const int selj = __shfl_sync(0xffffffffu, lj, (lane & ~7) | gl);
        if (c == r) { myj = selj; myv = gv; }
        if (c == gl && lp >= 0) {
#pragma unroll
            for (int p = 0; p < 8; ++p)
                if (p == lp) ej[p] = -1;
        }
        if (r == TOPK - 1) a8 = selj >= 0 ? gv : -1.0f;
    }
    if constexpr (MERGE_ONLY) {
        if (qvalid) {
            mrg_val[(size_t)qi * TOPK + c] = myj >= 0 ? myv : -1.0f;
            mrg_idx[(size_t)qi * TOPK + c] = myj >= 0 ? myj + idx_offset : -1;
        }
        return;
    }
    // ---- 2. exact dots, coalesced ----
    pair_j[wid][lane] = myj;
    int qrow[4];
#pragma unroll
    for (int gg = 0; gg < 4; ++gg) qrow[gg] = __shfl_sync(0xffffffffu, orig, gg * 8);
    __syncwarp();
    double sum = 0.0;
#pragma unroll 1
    for (int kb = 0; kb < KD / 32; ++kb) {
        float av[4];
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) av[gg] = q[(size_t)qrow[gg] * KD + kb * 32 + lane];
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const int j = pair_j[wid][l];
            if (j >= 0) tile[wid][l][lane] = __fmul_rn(av[l >> 3], db[(size_t)j * KD + kb * 32 + lane]);
        }
        __syncwarp();
        if (myj >= 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) sum = __dadd_rn(sum, (double)tile[wid][lane][k]);
        }
        __syncwarp();
    }
    // ---- 3. top-2 of the query's eight lanes, guard ----
    Top2 best;
    top2_init(best);
    if (myj >= 0) top2_push(best, sum, myj + db_offset);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        Top2 o;
        o.d1 = __shfl_xor_sync(0xffffffffu, best.d1, off); o.d2 = __shfl_xor_sync(0xffffffffu, best.d2, off);
        o.i1 = __shfl_xor_sync(0xffffffffu, best.i1, off); o.i2 = __shfl_xor_sync(0xffffffffu, best.i2, off);
        top2_merge(best, o);
    }
    if (c == 0 && qvalid) {
        out[orig] = best;
        const double bound = (double)a8 * (1.0 + 1.2e-3) + 1e-6;
        const bool safe = a8 < 0.0f || (best.i2 >= 0 && best.d2 > bound);
        if (!safe) fb_list[atomicAdd(fb_count, 1)] = orig;
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    }
    return fn;
}

static int make_map(CUtensorMap* m, const __half* base, int rows, int box_rows) {
    auto enc = get_encode();
    if (!enc) return fail(S3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)KD, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)KD * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(S3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return S3D_OK;
}

}  // namespace tc

// Number of database parts per query tile: the smallest split whose last wave is at least 90 % full
// (items = units * parts over `workers` persistent CTAs / CTA pairs), else the best one seen.
static int choose_parts(int units, int n_dbtiles, int workers) {
    int best = 1;
    double best_eff = 0.0;
    const int pmax = std::max(1, std::min(8, n_dbtiles / 4));
    for (int parts = 1; parts <= pmax; ++parts) {
        const long long items = (long long)units * parts;
        const long long waves = (items + workers - 1) / workers;
        const double eff = (double)items / (double)(waves * workers);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = parts; }
        if (eff >= 0.9) break;
    }
    return best;
}

// ---- host side of the tensor-core search, in reusable steps (the single-GPU search runs them back to back; the
// database-sharded search interleaves them with the exchange of candidate lists, s3d_match.cu) ------------------------
void tc_free(TcWork& w, cudaStream_t st) {
    void* tmp[] = {w.q16, w.db16, w.cand_val, w.cand_idx, w.d_bad};
    for (void* p : tmp) if (p) cudaFreeAsync(p, st);
    w = TcWork();
}

// FP16 copies of the listed query rows and of the database rows, and the precondition flag *w.d_bad (device).
int tc_convert(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, cudaStream_t st, int variant, TcWork& w) {
    using namespace tc;
    int dev = 0, sms = 148;
    S3D_CUDA(cudaGetDevice(&dev));
    S3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        S3D_CUDA(cudaFuncSetAttribute(tc_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        S3D_CUDA(cudaFuncSetAttribute(tc_pair_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pr::SMEM_BYTES));
        attr_set[dev] = true;
    }
    // Measured on B200 (profiles/): the one-CTA-per-tile kernel keeps the tensor pipe 90 % busy; the pair
    // kernel moves a third of the L2->SM bytes (and holds higher clocks under the power cap) but its
    // 4 x 8 KB ring leaves the pipe 79 % busy, so it is selected only on request.
    w.pair = variant == 2;
    w.sms = sms;
    w.nql = nql; w.nd = nd;
    w.n_dbtiles = (nd + BN - 1) / BN;
    w.n_units = w.pair ? (nql + 2 * pr::BMC - 1) / (2 * pr::BMC) : (nql + BM - 1) / BM;
    const int workers = w.pair ? std::max(1, sms / 2) : sms;
    int parts = choose_parts(w.n_units, std::max(w.n_dbtiles, 1), workers);
    w.tiles_per_part = std::max(1, (w.n_dbtiles + parts - 1) / parts);
    w.parts = std::max(1, (w.n_dbtiles + w.tiles_per_part - 1) / w.tiles_per_part);
    w.ncand = w.parts * TOPK;
    // capacity: the FP16 copies and candidate lists are temporaries on top of the caller's FP32 sets (1 M x 1 M:
    // 2 x 1.5 GB + 64 MB * parts); refuse up front rather than fail half-way through a stream-ordered allocation
    {
        size_t free_b = 0, total_b = 0;
        const size_t need = sizeof(__half) * KD * ((size_t)nql + (size_t)nd) + (sizeof(float) + sizeof(int)) * (size_t)nql * w.ncand;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            uint64_t pooled = 0;  // bytes the stream-ordered pool already holds and can hand out again
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                uint64_t reserved = 0, used = 0;
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
                cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
                pooled = reserved > used ? reserved - used : 0;
            }
            if (need > free_b + pooled)
                return fail(S3D_ERR_CAPACITY, "tensor-core search of %d x %d rows needs %.2f GB of temporaries, %.2f GB available",
                            nql, nd, need / 1e9, (free_b + pooled) / 1e9);
        }
    }
    S3D_CUDA(cudaMallocAsync((void**)&w.q16, sizeof(__half) * KD * (size_t)std::max(nql, 1), st));
    S3D_CUDA(cudaMallocAsync((void**)&w.db16, sizeof(__half) * KD * (size_t)std::max(nd, 1), st));
    S3D_CUDA(cudaMallocAsync((void**)&w.d_bad, sizeof(int), st));
    S3D_CUDA(cudaMemsetAsync(w.d_bad, 0, sizeof(int), st));
    if (nql > 0) S3D_LAUNCH(cvt_f16_kernel, s3d_blocks((size_t)nql * 32, 256), 256, 0, st, d_q, d_qlist, nql, (__half*)w.q16, w.d_bad);
    if (nd > 0) S3D_LAUNCH(cvt_f16_kernel, s3d_blocks((size_t)nd * 32, 256), 256, 0, st, d_db, (const int*)nullptr, nd, (__half*)w.db16, w.d_bad);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// The candidate pass: w.cand_val / w.cand_idx [nql][parts * 8] (index -1 = empty slot; indices local to the database).
int tc_topk(cudaStream_t st, TcWork& w) {
    using namespace tc;
    const int nql = w.nql, nd = w.nd;
    S3D_CUDA(cudaMallocAsync((void**)&w.cand_val, sizeof(float) * (size_t)std::max(nql, 1) * w.ncand, st));
    S3D_CUDA(cudaMallocAsync((void**)&w.cand_idx, sizeof(int) * (size_t)std::max(nql, 1) * w.ncand, st));
    if (nql <= 0) return S3D_OK;
    if (nd <= 0) {  // an empty database shard: every slot empty
        S3D_CUDA(cudaMemsetAsync(w.cand_idx, 0xFF, sizeof(int) * (size_t)nql * w.ncand, st));
        S3D_CUDA(cudaMemsetAsync(w.cand_val, 0, sizeof(float) * (size_t)nql * w.ncand, st));
        return S3D_OK;
    }
    CUtensorMap mq, mdb;
    S3D_TRY(make_map(&mq, (const __half*)w.q16, nql, BM));
    S3D_TRY(make_map(&mdb, (const __half*)w.db16, nd, w.pair ? pr::BROWS : BN));
    const int workers = w.pair ? std::max(1, w.sms / 2) : w.sms;
    if (w.pair) {
        const int grid = 2 * std::min(workers, w.n_units * w.parts);
        // chunk = the slice of the database all clusters sweep together; parts are swept concurrently, so
        // together they should stay well inside L2 (126 MB, shared with the streaming query tiles)
        const int chunk_tiles = std::max(8, pr::kChunkTiles / w.parts);
        S3D_LAUNCH(tc_pair_topk_kernel, grid, kThreads, pr::SMEM_BYTES, st, mq, mdb, nql, nd, w.n_units, w.parts, w.tiles_per_part,
                   w.n_dbtiles, chunk_tiles, w.cand_val, w.cand_idx);
    } else {
        const int grid = std::min(w.sms, w.n_units * w.parts);
        S3D_LAUNCH(tc_topk_kernel, grid, kThreads, SMEM_BYTES, st, mq, mdb, nql, nd, w.n_units, w.parts, w.tiles_per_part, w.n_dbtiles,
                   w.cand_val, w.cand_idx);
    }
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// Merge the parts' lists to one top-8 per query: out_val / out_idx [nql][8], indices + idx_offset.
int tc_merge8(cudaStream_t st, const TcWork& w, int idx_offset, float* out_val, int* out_idx) {
    using namespace tc;
    if (w.nql <= 0) return S3D_OK;
    S3D_LAUNCH(rerank_kernel<true>, s3d_blocks((size_t)w.nql, 4 * kRrWarps), kRrWarps * 32, 0, st, (const float*)nullptr, (const int*)nullptr,
               w.nql, (const float*)nullptr, 0, 0, w.parts, (size_t)w.ncand, (size_t)TOPK, w.cand_val, w.cand_idx, (Top2*)nullptr,
               (int*)nullptr, (int*)nullptr, out_val, out_idx, idx_offset);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// Exact re-rank of `nql` listed queries against db from `nlists` (<= 8) top-8 lists per query (see rerank_kernel).
int tc_rerank(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, const float* cand_val,
              const int* cand_idx, int nlists, size_t stride_q, size_t stride_p, Top2* d_out, int* d_fb_list, int* d_fb_count,
              cudaStream_t st) {
    using namespace tc;
    if (nql <= 0) return S3D_OK;
    if (nlists > 8) return fail(S3D_ERR_ARG, "re-rank merges at most 8 candidate lists per query (%d given)", nlists);
    S3D_LAUNCH(rerank_kernel<false>, s3d_blocks((size_t)nql, 4 * kRrWarps), kRrWarps * 32, 0, st, d_q, d_qlist, nql, d_db, nd, db_offset,
               nlists, stride_q, stride_p, cand_val, cand_idx, d_out, d_fb_list, d_fb_count, (float*)nullptr, (int*)nullptr, 0);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// Tensor-core search of `nql` query rows (qlist maps to original rows, may be null) against db.
// out[orig] receives the exact top-2; rows whose candidate set could not be proven complete are
// appended to fb_list / *fb_count (device) for the exact kernel.
// variant: 0 = choose by size, 1 = one CTA per tile (tc_topk_kernel), 2 = CTA pairs with the query
// tile resident in shared memory (tc_pair_topk_kernel).
int tc_search(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, Top2* d_out,
              int* d_fb_list, int* d_fb_count, cudaStream_t st, int variant) {
    if (nql <= 0 || nd <= 0) return S3D_OK;
    TcWork w;
    // every exit goes through tc_free, which returns the temporaries to the pool (early returns leaked)
    auto body = [&]() -> int {
        S3D_TRY(tc_convert(d_q, d_qlist, nql, d_db, nd, st, variant, w));
        int h_bad = 0;
        S3D_CUDA(cudaMemcpyAsync(&h_bad, w.d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        S3D_CUDA(cudaStreamSynchronize(st));
        if (h_bad) return kTcRefused;  // outside the guard's precondition: the caller runs the exact kernel
        S3D_TRY(tc_topk(st, w));
        return tc_rerank(d_q, d_qlist, nql, d_db, nd, db_offset, w.cand_val, w.cand_idx, w.parts, (size_t)w.ncand, (size_t)tc::TOPK, d_out,
                         d_fb_list, d_fb_count, st);
    };
    const int rc = body();
    tc_free(w, st);
    return rc;
}

}  // namespace s3d
