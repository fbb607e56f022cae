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

__global__ void cvt_f16_kernel(const float* __restrict__ src, const int* __restrict__ rows, int nrows, __half* __restrict__ dst) {
    // one warp per row (768 floats = 6 float4 per lane)
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nrows) return;
    const float4* s = reinterpret_cast<const float4*>(src + (size_t)(rows ? rows[w] : w) * KD);
    uint2* d = reinterpret_cast<uint2*>(dst + (size_t)w * KD);
#pragma unroll
    for (int i = 0; i < KD / 4 / 32; ++i) {
        const float4 v = s[lane + 32 * i];
        const __half2 a = __floats2half2_rn(v.x * SCALE, v.y * SCALE), b = __floats2half2_rn(v.z * SCALE, v.w * SCALE);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned*>(&a);
        o.y = *reinterpret_cast<const unsigned*>(&b);
        d[lane + 32 * i] = o;
    }
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
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tc_ld32(taddr + (uint32_t)(c * 32), r);
                    tc_wait_ld();
                    float mx = __uint_as_float(r[0]);
#pragma unroll
                    for (int e = 1; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
                    if (mx * INV_SCALE2 > v[TOPK - 1]) {
                        const int j0 = t * BN + c * 32;
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float x = __uint_as_float(r[e]) * INV_SCALE2;
                            if (x > v[TOPK - 1] && j0 + e < nd) {
                                v[TOPK - 1] = x; id[TOPK - 1] = j0 + e;
#pragma unroll
                                for (int i = TOPK - 1; i > 0; --i)
                                    if (v[i] > v[i - 1]) {
                                        const float tv = v[i]; v[i] = v[i - 1]; v[i - 1] = tv;
                                        const int ti = id[i]; id[i] = id[i - 1]; id[i - 1] = ti;
                                    }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc));  // 4 epilogue warps -> count 4
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < nq) {
                const size_t o = ((size_t)row * parts + p) * TOPK;
#pragma unroll
                for (int i = 0; i < TOPK; ++i) { cand_val[o + i] = v[i]; cand_idx[o + i] = id[i]; }
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

// Exact re-rank of the candidates of one query per warp (lane = candidate slot), guard, outputs.
// q rows are addressed through qlist (original row index); results are written to out[orig].
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ q, const int* __restrict__ qlist, int nql,
                                                     const float* __restrict__ db, int nd, int db_offset, int ncand,
                                                     const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                                     Top2* __restrict__ out, int* __restrict__ fb_list, int* fb_count) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nql) return;
    const int orig = qlist ? qlist[w] : w;
    const float* a = q + (size_t)orig * KD;
    Top2 best;
    top2_init(best);
    float a_floor = -1.0f;  // largest "smallest kept approximate value" over the parts that were full
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        int j = -1;
        float av = -1.0f;
        if (c < ncand) { j = cand_idx[(size_t)w * ncand + c]; av = cand_val[(size_t)w * ncand + c]; }
        // a part whose 8th slot is filled may hide better rows below its threshold
        if (c < ncand && (c % TOPK) == TOPK - 1 && j >= 0) a_floor = fmaxf(a_floor, av);
        double s = 0.0;
        if (j >= 0) {
            const float* b = db + (size_t)j * KD;
#pragma unroll 8
            for (int k = 0; k < KD; ++k) s = __dadd_rn(s, (double)__fmul_rn(a[k], b[k]));
            top2_push(best, s, j + db_offset);
        }
    }
    // warp merge under (dot desc, index asc)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Top2 o;
        o.d1 = __shfl_xor_sync(0xffffffffu, best.d1, off); o.d2 = __shfl_xor_sync(0xffffffffu, best.d2, off);
        o.i1 = __shfl_xor_sync(0xffffffffu, best.i1, off); o.i2 = __shfl_xor_sync(0xffffffffu, best.i2, off);
        top2_merge(best, o);
        a_floor = fmaxf(a_floor, __shfl_xor_sync(0xffffffffu, a_floor, off));
    }
    if (lane == 0) {
        out[orig] = best;
        // GUARD (see file header): non-candidates satisfy exact <= a_floor*(1+eps) + eps_abs
        const double bound = (double)a_floor * (1.0 + 1.2e-3) + 1e-6;
        const bool safe = a_floor < 0.0f || (best.i2 >= 0 && best.d2 > bound);
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

// Tensor-core search of `nql` query rows (qlist maps to original rows, may be null) against db.
// out[orig] receives the exact top-2; rows whose candidate set could not be proven complete are
// appended to fb_list / *fb_count (device) for the exact kernel.
int tc_search(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, Top2* d_out,
              int* d_fb_list, int* d_fb_count, cudaStream_t st) {
    using namespace tc;
    if (nql <= 0 || nd <= 0) return S3D_OK;
    int dev = 0, sms = 148;
    S3D_CUDA(cudaGetDevice(&dev));
    S3D_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_set[64] = {false};
    if (dev < 64 && !attr_set[dev]) {
        S3D_CUDA(cudaFuncSetAttribute(tc_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set[dev] = true;
    }
    __half *q16 = nullptr, *db16 = nullptr;
    float* cand_val = nullptr;
    int* cand_idx = nullptr;
    const int n_qtiles = (nql + BM - 1) / BM, n_dbtiles = (nd + BN - 1) / BN;
    // split the database only when there are too few query tiles to fill the GPU
    int parts = std::max(1, std::min(n_dbtiles, (sms + n_qtiles - 1) / n_qtiles));
    parts = std::min(parts, 4);
    const int tiles_per_part = (n_dbtiles + parts - 1) / parts;
    parts = (n_dbtiles + tiles_per_part - 1) / tiles_per_part;
    const int ncand = parts * TOPK;
    S3D_CUDA(cudaMallocAsync((void**)&q16, sizeof(__half) * KD * (size_t)nql, st));
    S3D_CUDA(cudaMallocAsync((void**)&db16, sizeof(__half) * KD * (size_t)nd, st));
    S3D_CUDA(cudaMallocAsync((void**)&cand_val, sizeof(float) * (size_t)nql * ncand, st));
    S3D_CUDA(cudaMallocAsync((void**)&cand_idx, sizeof(int) * (size_t)nql * ncand, st));
    S3D_LAUNCH(cvt_f16_kernel, s3d_blocks((size_t)nql * 32, 256), 256, 0, st, d_q, d_qlist, nql, q16);
    S3D_LAUNCH(cvt_f16_kernel, s3d_blocks((size_t)nd * 32, 256), 256, 0, st, d_db, (const int*)nullptr, nd, db16);
    CUtensorMap mq, mdb;
    S3D_TRY(make_map(&mq, q16, nql, BM));
    S3D_TRY(make_map(&mdb, db16, nd, BN));
    const int grid = std::min(sms, n_qtiles * parts);
    S3D_LAUNCH(tc_topk_kernel, grid, kThreads, SMEM_BYTES, st, mq, mdb, nql, nd, n_qtiles, parts, tiles_per_part, n_dbtiles,
               cand_val, cand_idx);
    S3D_LAUNCH(rerank_kernel, s3d_blocks((size_t)nql * 32, 256), 256, 0, st, d_q, d_qlist, nql, d_db, nd, db_offset, ncand,
               cand_val, cand_idx, d_out, d_fb_list, d_fb_count);
    S3D_CUDA(cudaGetLastError());
    void* tmp[] = {q16, db16, cand_val, cand_idx};
    for (void* p : tmp) S3D_CUDA(cudaFreeAsync(p, st));
    return S3D_OK;
}

}  // namespace s3d
