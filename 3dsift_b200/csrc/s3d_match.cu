// s3d_match.cu — brute-force descriptor matching (muBruteMatcher, /root/reference/3DSIFT/Src/cMatcher.cc).
//
// Exact path: for every query the best / second-best dot product over the database, where a dot
// product is the reference's own  sum_{i=0..767} (double)(float)(a_i * b_i)  accumulated
// sequentially in double (KP_squareSum, Src/cMatcher.cc:17-23).  The running top-2 uses the total
// order (dot desc, index asc), which equals the reference's ascending-j strict-'>' scan
// (Src/cMatcher.cc:54-70).
#include <cfloat>
#include <cstring>
#include <vector>

#include <string>
#include <thread>

#include "s3d_comm.h"
#include "s3d_common.h"
#include "s3d_match_internal.h"

namespace s3d {

constexpr int kD = S3D_DESC_LEN;

// Exact tile kernel: CTA = 64 queries x 64 database rows per step, 256 threads, each thread owns a
// 4x4 block of (query, db) pairs (db rows interleaved by 16 to keep shared reads conflict-free) and walks k = 0..767 in order with a k-chunked shared-memory
// stage, so every pair sees the reference's accumulation order.  Database rows are swept in
// `db_per_cta`-sized slices (gridDim.y) and merged by top2_merge_kernel.
constexpr int kTQ = 64, kTD = 64, kTK = 32;

// qlist (may be null) maps the nq listed queries to rows of q.
__global__ void __launch_bounds__(256) top2_exact_kernel(const float* __restrict__ q, const int* __restrict__ qlist, int nq,
                                                         const float* __restrict__ db, int nd, int db_offset,
                                                         int db_per_cta, Top2* __restrict__ part) {
    __shared__ float sq[kTK][kTQ + 1];
    __shared__ float sd[kTK][kTD + 1];
    __shared__ Top2 red[kTQ][16 + 1];
    const int tid = threadIdx.x;
    const int tq = tid / 16, td = tid % 16;  // thread owns queries tq*4.., db rows td*4..
    const int q0 = blockIdx.x * kTQ;
    const int dbeg = blockIdx.y * db_per_cta, dend = min(nd, dbeg + db_per_cta);
    Top2 best[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) top2_init(best[a]);
    for (int d0 = dbeg; d0 < dend; d0 += kTD) {
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
        for (int k0 = 0; k0 < kD; k0 += kTK) {
            __syncthreads();
            // stage kTK columns of 64 query rows and 64 db rows (transposed: [k][row])
            for (int e = tid; e < kTQ * kTK; e += 256) {
                const int row = e / kTK, kk = e % kTK;
                const int qi = q0 + row, di = d0 + row;
                sq[kk][row] = qi < nq ? q[(size_t)(qlist ? qlist[qi] : qi) * kD + k0 + kk] : 0.0f;
                sd[kk][row] = di < dend ? db[(size_t)di * kD + k0 + kk] : 0.0f;
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < kTK; ++kk) {
                float av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) av[a] = sq[kk][tq * 4 + a];
#pragma unroll
                for (int b = 0; b < 4; ++b) bv[b] = sd[kk][b * 16 + td];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b] = __dadd_rn(acc[a][b], (double)__fmul_rn(av[a], bv[b]));
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int di = d0 + b * 16 + td;
                if (di < dend) top2_push(best[a], acc[a][b], di + db_offset);
            }
    }
    // merge the 16 db-lanes of every query
#pragma unroll
    for (int a = 0; a < 4; ++a) red[tq * 4 + a][td] = best[a];
    __syncthreads();
    if (tid < kTQ) {
        Top2 r = red[tid][0];
        for (int j = 1; j < 16; ++j) top2_merge(r, red[tid][j]);
        const int qi = q0 + tid;
        if (qi < nq) part[(size_t)blockIdx.y * nq + qi] = r;
    }
}

// Merge `parts` partial lists of the listed queries into out[orig] (merging INTO the existing
// entry when `accumulate`, which the tensor-core fallback rows use to overwrite theirs).
__global__ void __launch_bounds__(256) top2_merge_kernel(int parts, int nq, const int* __restrict__ qlist,
                                                         const Top2* __restrict__ part, Top2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    Top2 r = part[i];
    for (int p = 1; p < parts; ++p) top2_merge(r, part[(size_t)p * nq + i]);
    out[qlist ? qlist[i] : i] = r;
}

__global__ void top2_fill_kernel(Top2* t, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) top2_init(t[i]);
}

// calMatches' outputs (Src/cMatcher.cc:71-77); masked-out queries get gIdx = -1 and nothing else
// (:48-52).  Optionally also the raw (double dot, index) pairs for multi-GPU merging.
__global__ void __launch_bounds__(256) top2_finalize_kernel(int nq, const Top2* __restrict__ top, const int* __restrict__ mask,
                                                            double* d1o, int* i1o, double* d2o, int* i2o, float* gDist,
                                                            int* gIdx, float* sDist, int* sIdx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    if (mask && mask[i] == 0) {
        if (gIdx) gIdx[i] = -1;
        if (i1o) { i1o[i] = -1; i2o[i] = -1; d1o[i] = (double)FLT_MIN; d2o[i] = (double)FLT_MIN; }
        return;
    }
    const Top2 r = top[i];
    if (i1o) { d1o[i] = r.d1; i1o[i] = r.i1; d2o[i] = r.d2; i2o[i] = r.i2; }
    if (gIdx) {
        gDist[i] = (float)(2 - 2 * r.d1);
        sDist[i] = (float)(2 - 2 * r.d2);
        gIdx[i] = r.i1;
        sIdx[i] = r.i2;
    }
}

// ordered list of the queries with mask != 0 — one CTA
__global__ void __launch_bounds__(1024) mask_list_kernel(const int* __restrict__ mask, int n, int* list, int* count) {
    __shared__ int part[1024];
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += mask[i] != 0;
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
        if (mask[i] != 0) list[run++] = i;
    if (threadIdx.x == 1023) *count = part[1023];
}

// Merge across shards given as separate arrays laid out [part][nq] (multi-GPU gather).
__global__ void __launch_bounds__(256) top2_merge_arrays_kernel(int parts, int nq, const double* __restrict__ d1,
                                                                const int* __restrict__ i1, const double* __restrict__ d2,
                                                                const int* __restrict__ i2, const int* __restrict__ mask,
                                                                float* gDist, int* gIdx, float* sDist, int* sIdx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    if (mask && mask[i] == 0) {
        gIdx[i] = -1;
        return;
    }
    Top2 r;
    top2_init(r);
    for (int p = 0; p < parts; ++p) {
        const size_t o = (size_t)p * nq + i;
        if (i1[o] >= 0) top2_push(r, d1[o], i1[o]);
        if (i2[o] >= 0) top2_push(r, d2[o], i2[o]);
    }
    gDist[i] = (float)(2 - 2 * r.d1);
    sDist[i] = (float)(2 - 2 * r.d2);
    gIdx[i] = r.i1;
    sIdx[i] = r.i2;
}

// filter, Src/cMatcher.cc:81-97: float quotient of squared distances vs double thr^2; idx *= -1
// (so index 0 cannot be rejected, App. B Q19; NaN keeps the match, Q20).
__global__ void ratio_filter_kernel(int* gIdx, const float* __restrict__ gDist, const float* __restrict__ sDist, int n,
                                    double thresSquare) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (gIdx[i] < 0) return;
    const float d1 = gDist[i], d2 = sDist[i];
    if ((double)__fdiv_rn(d1, d2) >= thresSquare) gIdx[i] *= -1;
}

// countMatched, Src/cMatcher.cc:114-120
__global__ void count_kernel(const int* __restrict__ gIdx, int n_ref, int* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const int idx = gIdx[i];
    if (idx >= 0) atomicAdd(&counts[idx], 1);
}
// toMask, Src/cMatcher.cc:122-131
__global__ void mask_kernel(int* counts, int n_tar, int thres) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tar) return;
    counts[i] = counts[i] > thres ? 1 : 0;
}
// bijectFilter, Src/cMatcher.cc:133-144
__global__ void biject_kernel(int* gIdx, int n_ref, const int* __restrict__ mask, const int* __restrict__ gIdx2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ref) return;
    const int m = gIdx[i];
    if (m < 0 || mask[m] == 0) return;
    if (gIdx2[m] != i) gIdx[i] *= -1;
}
// toCvec, Src/cMatcher.cc:99-112: ordered compaction of (i, gIdx[i]) for gIdx[i] >= 0 — one CTA.
__global__ void __launch_bounds__(1024) pairs_kernel(const int* __restrict__ gIdx, int n, int* pair_ref, int* pair_tar,
                                                     int* n_pairs) {
    __shared__ int part[1024];
    const int per = (n + 1023) / 1024;
    const int b = threadIdx.x * per, e = min(n, b + per);
    int s = 0;
    for (int i = b; i < e; ++i) s += gIdx[i] >= 0;
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
        if (gIdx[i] >= 0) {
            pair_ref[run] = i;
            pair_tar[run] = gIdx[i];
            ++run;
        }
    if (threadIdx.x == 1023) *n_pairs = part[1023];
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// 0 = auto (tensor cores for large searches), 1 = exact CUDA-core kernel only, 2 = tensor cores always
// (kernel variant by size), 3 = tensor cores, one CTA per tile, 4 = tensor cores, CTA pairs with the
// query tile resident in shared memory  (initial value from the environment: S3D_MATCH_PATH=0..4)
static int initial_match_path() {
    const char* e = getenv("S3D_MATCH_PATH");
    return (e && e[0] >= '0' && e[0] <= '4' && !e[1]) ? e[0] - '0' : 0;
}
static std::atomic<int> g_match_path{initial_match_path()};
static std::atomic<unsigned long long> g_tc_rows{0}, g_fb_rows{0}, g_refused_searches{0};

// Exact CUDA-core search of the listed queries; results go to out[orig].
static int exact_search(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, Top2* d_out,
                        cudaStream_t st) {
    if (nql <= 0) return S3D_OK;
    // split the database so the grid fills the GPU: ~148*4 CTAs
    const int qtiles = (nql + kTQ - 1) / kTQ;
    int parts = std::max(1, std::min((nd + kTD - 1) / kTD, (148 * 4 + qtiles - 1) / qtiles));
    int db_per_cta = ((nd + parts - 1) / parts + kTD - 1) / kTD * kTD;
    if (db_per_cta < kTD) db_per_cta = kTD;
    parts = std::max(1, (nd + db_per_cta - 1) / db_per_cta);
    Top2* d_part = nullptr;
    S3D_CUDA(cudaMallocAsync((void**)&d_part, sizeof(Top2) * (size_t)parts * nql, st));
    dim3 grid(qtiles, parts);
    S3D_LAUNCH(top2_exact_kernel, grid, 256, 0, st, d_q, d_qlist, nql, d_db, nd, db_offset, db_per_cta, d_part);
    S3D_LAUNCH(top2_merge_kernel, s3d_blocks(nql, 256), 256, 0, st, parts, nql, d_qlist, d_part, d_out);
    S3D_CUDA(cudaGetLastError());
    S3D_CUDA(cudaFreeAsync(d_part, st));
    return S3D_OK;
}

// One search direction (calMatches, Src/cMatcher.cc:40-79) on device arrays.
static int search_device(const float* d_q, int nq, const float* d_db, int nd, int db_offset, const int* d_mask,
                         double* d1, int* i1, double* d2, int* i2, float* gDist, int* gIdx, float* sDist, int* sIdx,
                         cudaStream_t st) {
    if (nq <= 0) return S3D_OK;
    Top2* d_top = nullptr;
    int *d_list = nullptr, *d_cnt = nullptr, *d_fb = nullptr;
    S3D_CUDA(cudaMallocAsync((void**)&d_top, sizeof(Top2) * (size_t)nq, st));
    S3D_CUDA(cudaMallocAsync((void**)&d_cnt, sizeof(int) * 2, st));
    S3D_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * 2, st));
    S3D_LAUNCH(top2_fill_kernel, s3d_blocks(nq, 256), 256, 0, st, d_top, nq);
    int n_act = nq;
    if (d_mask) {  // only the unmasked queries are searched (Src/cMatcher.cc:48-52)
        S3D_CUDA(cudaMallocAsync((void**)&d_list, sizeof(int) * (size_t)nq, st));
        S3D_LAUNCH(mask_list_kernel, 1, 1024, 0, st, d_mask, nq, d_list, d_cnt);
        S3D_CUDA(cudaMemcpyAsync(&n_act, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
        S3D_CUDA(cudaStreamSynchronize(st));
    }
    const int path = g_match_path.load();
    const bool use_tc = nd > 0 && n_act > 0 && (path >= 2 || (path == 0 && (double)n_act * nd >= 4.0e6 && nd >= 512));
    if (use_tc) {
        S3D_CUDA(cudaMallocAsync((void**)&d_fb, sizeof(int) * (size_t)n_act, st));
        const int rc = tc_search(d_q, d_list, n_act, d_db, nd, db_offset, d_top, d_fb, d_cnt + 1, st, path == 3 ? 1 : (path == 4 ? 2 : 0));
        if (rc == kTcRefused) {
            // descriptors outside the tensor-core guard's precondition (negative / > 1 / non-finite entries):
            // the exact kernel is right for any input
            g_refused_searches++;
            S3D_TRY(exact_search(d_q, d_list, n_act, d_db, nd, db_offset, d_top, st));
        } else {
            S3D_TRY(rc);
            int n_fb = 0;
            S3D_CUDA(cudaMemcpyAsync(&n_fb, d_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            S3D_CUDA(cudaStreamSynchronize(st));
            g_tc_rows += (unsigned long long)n_act;
            g_fb_rows += (unsigned long long)n_fb;
            if (n_fb > 0) S3D_TRY(exact_search(d_q, d_fb, n_fb, d_db, nd, db_offset, d_top, st));
        }
    } else if (nd > 0) {
        S3D_TRY(exact_search(d_q, d_list, n_act, d_db, nd, db_offset, d_top, st));
    }
    S3D_LAUNCH(top2_finalize_kernel, s3d_blocks(nq, 256), 256, 0, st, nq, d_top, d_mask, d1, i1, d2, i2, gDist, gIdx, sDist, sIdx);
    S3D_CUDA(cudaGetLastError());
    void* tmp[] = {d_top, d_list, d_cnt, d_fb};
    for (void* p : tmp) if (p) S3D_CUDA(cudaFreeAsync(p, st));
    return S3D_OK;
}


// ---- database-sharded search (the inner database loop of calMatches, Src/cMatcher.cc:58) ---------
// Every rank holds the full query and database sets (replicated) and SEARCHES database rows [lo_r, hi_r) for all
// queries on tensor cores.  The approximate top-8 lists are then re-distributed so that the EXACT re-rank is sharded
// by query: rank b receives every rank's list for its block of queries, merges them (the guard argument holds for the
// merged list: its 8th value is >= every full list's 8th) and recomputes the 8 candidates in the reference's arithmetic
// against the full database.  Exact top-2 blocks are all-gathered; the cheap filters run replicated.  Per-rank work is
// ~1/world of every heavy step, so the path scales until the collectives (64 B/query each way) matter.
__global__ void top2_pack_kernel(const Top2* __restrict__ top, const int* __restrict__ list, int first, int n, Top2* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    out[j] = top[list ? list[first + j] : first + j];
}

__global__ void top2_unpack_kernel(const Top2* __restrict__ all, int maxb, int world, int n_act, const int* __restrict__ list,
                                   Top2* __restrict__ top) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // position in the list of active queries
    if (i >= n_act) return;
    // rank r owns positions [lo(r), lo(r+1)) — shard_lo, inlined
    const int base = n_act / world, rem = n_act % world;
    int r = i / (base + 1);
    if (r >= rem) r = base > 0 ? rem + (i - rem * (base + 1)) / base : world - 1;
    const int lo = base * r + (rem < r ? rem : r);
    top[list ? list[i] : i] = all[(size_t)r * maxb + (i - lo)];
}

static int search_sharded(Comm& cm, const float* d_q, int nq, const float* d_db, int nd, const int* d_mask, float* gDist, int* gIdx,
                          float* sDist, int* sIdx, cudaStream_t st) {
    if (nq <= 0) return S3D_OK;
    const int W = cm.world, R = cm.rank;
    Top2 *d_top = nullptr, *d_send = nullptr, *d_all = nullptr;
    int *d_list = nullptr, *d_cnt = nullptr, *d_fb = nullptr;
    float *mv = nullptr, *rv = nullptr;
    int *mi = nullptr, *ri = nullptr;
    TcWork w;
    auto body = [&]() -> int {
        S3D_CUDA(cudaMallocAsync((void**)&d_top, sizeof(Top2) * (size_t)nq, st));
        S3D_CUDA(cudaMallocAsync((void**)&d_cnt, sizeof(int) * 2, st));
        S3D_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * 2, st));
        S3D_LAUNCH(top2_fill_kernel, s3d_blocks(nq, 256), 256, 0, st, d_top, nq);
        int n_act = nq;
        if (d_mask) {  // only the unmasked queries are searched (Src/cMatcher.cc:48-52); the mask is replicated
            S3D_CUDA(cudaMallocAsync((void**)&d_list, sizeof(int) * (size_t)nq, st));
            S3D_LAUNCH(mask_list_kernel, 1, 1024, 0, st, d_mask, nq, d_list, d_cnt);
            S3D_CUDA(cudaMemcpyAsync(&n_act, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
            S3D_CUDA(cudaStreamSynchronize(st));
        }
        const int lo = shard_lo(nd, W, R), hi = shard_lo(nd, W, R + 1), nds = hi - lo;
        const float* db_s = d_db + (size_t)lo * kD;
        const int path = g_match_path.load();
        bool use_tc = W <= 8 && nd > 0 && n_act > 0 && (path >= 2 || (path == 0 && (double)n_act * nd >= 4.0e6 && nd >= 512));
        if (use_tc) {
            S3D_TRY(tc_convert(d_q, d_list, n_act, db_s, nds, st, path == 3 ? 1 : (path == 4 ? 2 : 0), w));
            S3D_TRY(cm.allreduce_max_u32((unsigned*)w.d_bad, 1, st));  // every rank takes the same path
            int h_bad = 0;
            S3D_CUDA(cudaMemcpyAsync(&h_bad, w.d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
            S3D_CUDA(cudaStreamSynchronize(st));
            if (h_bad) { use_tc = false; g_refused_searches++; }
        }
        if (use_tc) {
            S3D_TRY(tc_topk(st, w));
            S3D_CUDA(cudaMallocAsync((void**)&mv, sizeof(float) * 8 * (size_t)n_act, st));
            S3D_CUDA(cudaMallocAsync((void**)&mi, sizeof(int) * 8 * (size_t)n_act, st));
            S3D_TRY(tc_merge8(st, w, lo, mv, mi));
            const int q0 = shard_lo(n_act, W, R), bs = shard_lo(n_act, W, R + 1) - q0;
            const int maxb = (n_act + W - 1) / W;
            S3D_CUDA(cudaMallocAsync((void**)&rv, sizeof(float) * 8 * (size_t)std::max(bs, 1) * W, st));
            S3D_CUDA(cudaMallocAsync((void**)&ri, sizeof(int) * 8 * (size_t)std::max(bs, 1) * W, st));
            std::vector<Xfer> sends, recvs;
            for (int b = 0; b < W; ++b) {
                if (b == R) continue;
                const int b0 = shard_lo(n_act, W, b), bn = shard_lo(n_act, W, b + 1) - b0;
                if (bn > 0) {
                    sends.push_back({b, mv + (size_t)b0 * 8, sizeof(float) * 8 * (size_t)bn});
                    sends.push_back({b, mi + (size_t)b0 * 8, sizeof(int) * 8 * (size_t)bn});
                }
                if (bs > 0) {
                    recvs.push_back({b, rv + (size_t)b * bs * 8, sizeof(float) * 8 * (size_t)bs});
                    recvs.push_back({b, ri + (size_t)b * bs * 8, sizeof(int) * 8 * (size_t)bs});
                }
            }
            if (bs > 0) {
                S3D_CUDA(cudaMemcpyAsync(rv + (size_t)R * bs * 8, mv + (size_t)q0 * 8, sizeof(float) * 8 * (size_t)bs, cudaMemcpyDeviceToDevice, st));
                S3D_CUDA(cudaMemcpyAsync(ri + (size_t)R * bs * 8, mi + (size_t)q0 * 8, sizeof(int) * 8 * (size_t)bs, cudaMemcpyDeviceToDevice, st));
            }
            S3D_TRY(cm.p2p(sends, recvs, st));
            // exact re-rank of my block of queries against the FULL database
            const float* qb = d_list ? d_q : d_q + (size_t)q0 * kD;
            const int* lb = d_list ? d_list + q0 : nullptr;
            Top2* ob = d_list ? d_top : d_top + q0;
            S3D_CUDA(cudaMallocAsync((void**)&d_fb, sizeof(int) * (size_t)std::max(bs, 1), st));
            S3D_TRY(tc_rerank(qb, lb, bs, d_db, nd, 0, rv, ri, W, (size_t)8, (size_t)bs * 8, ob, d_fb, d_cnt + 1, st));
            int n_fb = 0;
            S3D_CUDA(cudaMemcpyAsync(&n_fb, d_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            S3D_CUDA(cudaStreamSynchronize(st));
            g_tc_rows += (unsigned long long)bs;
            g_fb_rows += (unsigned long long)n_fb;
            if (n_fb > 0) S3D_TRY(exact_search(qb, d_fb, n_fb, d_db, nd, 0, ob, st));
            // all-gather the blocks' exact top-2
            S3D_CUDA(cudaMallocAsync((void**)&d_send, sizeof(Top2) * (size_t)maxb, st));
            S3D_CUDA(cudaMallocAsync((void**)&d_all, sizeof(Top2) * (size_t)maxb * W, st));
            S3D_CUDA(cudaMemsetAsync(d_send, 0, sizeof(Top2) * (size_t)maxb, st));
            if (bs > 0) S3D_LAUNCH(top2_pack_kernel, s3d_blocks(bs, 256), 256, 0, st, d_top, d_list, q0, bs, d_send);
            S3D_TRY(cm.allgather(d_send, d_all, sizeof(Top2) * (size_t)maxb, st));
            S3D_LAUNCH(top2_unpack_kernel, s3d_blocks(n_act, 256), 256, 0, st, d_all, maxb, W, n_act, d_list, d_top);
        } else if (nd > 0 && n_act > 0) {
            // exact CUDA-core kernel over this rank's database rows, then merge under (dot desc, index asc)
            S3D_CUDA(cudaMallocAsync((void**)&d_send, sizeof(Top2) * (size_t)nq, st));
            S3D_CUDA(cudaMallocAsync((void**)&d_all, sizeof(Top2) * (size_t)nq * W, st));
            S3D_LAUNCH(top2_fill_kernel, s3d_blocks(nq, 256), 256, 0, st, d_send, nq);
            if (nds > 0) S3D_TRY(exact_search(d_q, d_list, n_act, db_s, nds, lo, d_send, st));
            S3D_TRY(cm.allgather(d_send, d_all, sizeof(Top2) * (size_t)nq, st));
            S3D_LAUNCH(top2_merge_kernel, s3d_blocks(nq, 256), 256, 0, st, W, nq, (const int*)nullptr, d_all, d_top);
        }
        S3D_LAUNCH(top2_finalize_kernel, s3d_blocks(nq, 256), 256, 0, st, nq, d_top, d_mask, (double*)nullptr, (int*)nullptr,
                   (double*)nullptr, (int*)nullptr, gDist, gIdx, sDist, sIdx);
        S3D_CUDA(cudaGetLastError());
        return cm.finish(st);  // the peers have read what they were sent: the temporaries may be freed in stream order
    };
    const int rc = body();
    tc_free(w, st);
    void* tmp[] = {d_top, d_send, d_all, d_list, d_cnt, d_fb, mv, rv, mi, ri};
    for (void* p : tmp) if (p) cudaFreeAsync(p, st);
    return rc;
}

// bijectMatchBase (Src/cMatcher.cc:146-215) with both searches sharded over the communicator; outputs are complete
// on every rank.
static int match_sharded_device(Comm& cm, int type, const float* d_ref, int n_ref, const float* d_tar, int n_tar, double thr, int* d_gIdx,
                                float* d_gDist, int* d_sIdx, float* d_sDist, int* d_gIdx2, float* d_gDist2, int* d_sIdx2, float* d_sDist2,
                                int* d_pair_ref, int* d_pair_tar, int* d_n_pairs, cudaStream_t st) {
    if (type < 1 || type > 3) return fail(S3D_ERR_ARG, "match type %d (1 inject, 2 biject, 3 enhanced)", type);
    if (n_ref < 0 || n_tar < 0) return fail(S3D_ERR_ARG, "negative size");
    if (!d_gIdx || !d_gDist || !d_sIdx || !d_sDist) return fail(S3D_ERR_ARG, "forward outputs are required");
    if (n_ref > 0) {
        S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_ref, 256), 256, 0, st, d_gIdx, n_ref, -1);
        S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_ref, 256), 256, 0, st, d_sIdx, n_ref, -1);
        S3D_CUDA(cudaMemsetAsync(d_gDist, 0, sizeof(float) * n_ref, st));
        S3D_CUDA(cudaMemsetAsync(d_sDist, 0, sizeof(float) * n_ref, st));
    }
    S3D_TRY(search_sharded(cm, d_ref, n_ref, d_tar, n_tar, nullptr, d_gDist, d_gIdx, d_sDist, d_sIdx, st));
    S3D_TRY(s3d_ratio_filter_device(d_gIdx, d_gDist, d_sDist, n_ref, thr, st));
    if (type != 1) {
        if (!d_gIdx2 || !d_gDist2 || !d_sIdx2 || !d_sDist2) return fail(S3D_ERR_ARG, "reverse outputs are required");
        int* d_mask = nullptr;
        S3D_CUDA(cudaMallocAsync((void**)&d_mask, sizeof(int) * std::max(n_tar, 1), st));
        if (n_tar > 0) {
            S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_tar, 256), 256, 0, st, d_gIdx2, n_tar, -1);
            S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_tar, 256), 256, 0, st, d_sIdx2, n_tar, -1);
            S3D_CUDA(cudaMemsetAsync(d_gDist2, 0, sizeof(float) * n_tar, st));
            S3D_CUDA(cudaMemsetAsync(d_sDist2, 0, sizeof(float) * n_tar, st));
        }
        S3D_TRY(s3d_count_mask_device(d_gIdx, n_ref, d_mask, n_tar, type == 2 ? 0 : 1, st));
        const int rc = search_sharded(cm, d_tar, n_tar, d_ref, n_ref, d_mask, d_gDist2, d_gIdx2, d_sDist2, d_sIdx2, st);
        if (rc != S3D_OK) { cudaFreeAsync(d_mask, st); return rc; }
        S3D_TRY(s3d_ratio_filter_device(d_gIdx2, d_gDist2, d_sDist2, n_tar, thr, st));
        S3D_TRY(s3d_biject_filter_device(d_gIdx, n_ref, d_mask, d_gIdx2, st));
        S3D_CUDA(cudaFreeAsync(d_mask, st));
    }
    if (d_pair_ref && d_pair_tar && d_n_pairs) S3D_TRY(s3d_pairs_device(d_gIdx, n_ref, d_pair_ref, d_pair_tar, d_n_pairs, st));
    return S3D_OK;
}

}  // namespace s3d

using namespace s3d;

extern "C" {

int s3d_set_match_path(int path) {
    if (path < 0 || path > 4)
        return fail(S3D_ERR_ARG, "match path %d (0 auto, 1 exact, 2 tensor-core, 3 tensor-core single-CTA, 4 tensor-core CTA-pair)", path);
    g_match_path = path;
    return S3D_OK;
}

void s3d_match_stats(unsigned long long* tc_rows, unsigned long long* fallback_rows, int reset) {
    if (tc_rows) *tc_rows = g_tc_rows.load();
    if (fallback_rows) *fallback_rows = g_fb_rows.load();
    if (reset) { g_tc_rows = 0; g_fb_rows = 0; }
}

int s3d_top2_device(const float* d_q, int n_q, const float* d_db, int n_db, int db_offset, const int* d_mask,
                    double* d_dot1, int* d_idx1, double* d_dot2, int* d_idx2, void* stream) {
    clear_error();
    if (!d_q || !d_dot1 || !d_idx1 || !d_dot2 || !d_idx2 || n_q < 0 || n_db < 0) return fail(S3D_ERR_ARG, "bad argument");
    if (n_db > 0 && !d_db) return fail(S3D_ERR_ARG, "null database");
    int dev;
    S3D_TRY(use_device(-1, &dev));
    return search_device(d_q, n_q, d_db, n_db, db_offset, d_mask, d_dot1, d_idx1, d_dot2, d_idx2, nullptr, nullptr, nullptr,
                         nullptr, (cudaStream_t)stream);
}

int s3d_top2_merge_device(int parts, int n_q, const double* d_dot1, const int* d_idx1, const double* d_dot2,
                          const int* d_idx2, const int* d_mask, float* d_gDist, int* d_gIdx, float* d_sDist, int* d_sIdx,
                          void* stream) {
    clear_error();
    if (parts < 1 || n_q < 0 || !d_dot1 || !d_idx1 || !d_dot2 || !d_idx2 || !d_gDist || !d_gIdx || !d_sDist || !d_sIdx)
        return fail(S3D_ERR_ARG, "bad argument");
    if (n_q == 0) return S3D_OK;
    S3D_LAUNCH(top2_merge_arrays_kernel, s3d_blocks(n_q, 256), 256, 0, (cudaStream_t)stream, parts, n_q, d_dot1, d_idx1,
               d_dot2, d_idx2, d_mask, d_gDist, d_gIdx, d_sDist, d_sIdx);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

int s3d_ratio_filter_device(int* d_gIdx, const float* d_gDist, const float* d_sDist, int n, double thr, void* stream) {
    clear_error();
    if (n <= 0) return S3D_OK;
    S3D_LAUNCH(ratio_filter_kernel, s3d_blocks(n, 256), 256, 0, (cudaStream_t)stream, d_gIdx, d_gDist, d_sDist, n, thr * thr);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

int s3d_count_mask_device(const int* d_gIdx, int n_ref, int* d_mask, int n_tar, int count_thres, void* stream) {
    clear_error();
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tar <= 0) return S3D_OK;
    S3D_CUDA(cudaMemsetAsync(d_mask, 0, sizeof(int) * n_tar, st));
    if (n_ref > 0) S3D_LAUNCH(count_kernel, s3d_blocks(n_ref, 256), 256, 0, st, d_gIdx, n_ref, d_mask);
    S3D_LAUNCH(mask_kernel, s3d_blocks(n_tar, 256), 256, 0, st, d_mask, n_tar, count_thres);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

int s3d_biject_filter_device(int* d_gIdx, int n_ref, const int* d_mask, const int* d_gIdx2, void* stream) {
    clear_error();
    if (n_ref <= 0) return S3D_OK;
    S3D_LAUNCH(biject_kernel, s3d_blocks(n_ref, 256), 256, 0, (cudaStream_t)stream, d_gIdx, n_ref, d_mask, d_gIdx2);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

int s3d_pairs_device(const int* d_gIdx, int n_ref, int* d_pair_ref, int* d_pair_tar, int* d_n_pairs, void* stream) {
    clear_error();
    S3D_LAUNCH(pairs_kernel, 1, 1024, 0, (cudaStream_t)stream, d_gIdx, n_ref, d_pair_ref, d_pair_tar, d_n_pairs);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// bijectMatchBase, Src/cMatcher.cc:146-215, on device arrays.
int s3d_match_device(int type, const float* d_ref, int n_ref, const float* d_tar, int n_tar, double thr, int* d_gIdx,
                     float* d_gDist, int* d_sIdx, float* d_sDist, int* d_gIdx2, float* d_gDist2, int* d_sIdx2,
                     float* d_sDist2, int* d_pair_ref, int* d_pair_tar, int* d_n_pairs, void* stream) {
    clear_error();
    if (type < 1 || type > 3) return fail(S3D_ERR_ARG, "match type %d (1 inject, 2 biject, 3 enhanced)", type);
    if (n_ref < 0 || n_tar < 0) return fail(S3D_ERR_ARG, "negative size");
    if (!d_gIdx || !d_gDist || !d_sIdx || !d_sDist) return fail(S3D_ERR_ARG, "forward outputs are required");
    int dev;
    S3D_TRY(use_device(-1, &dev));
    cudaStream_t st = (cudaStream_t)stream;
    // glodenIdx/silverIdx start at -1 (:154-155)
    if (n_ref > 0) {
        S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_ref, 256), 256, 0, st, d_gIdx, n_ref, -1);
        S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_ref, 256), 256, 0, st, d_sIdx, n_ref, -1);
        S3D_CUDA(cudaMemsetAsync(d_gDist, 0, sizeof(float) * n_ref, st));
        S3D_CUDA(cudaMemsetAsync(d_sDist, 0, sizeof(float) * n_ref, st));
    }
    S3D_TRY(search_device(d_ref, n_ref, d_tar, n_tar, 0, nullptr, nullptr, nullptr, nullptr, nullptr, d_gDist, d_gIdx,
                          d_sDist, d_sIdx, st));
    S3D_TRY(s3d_ratio_filter_device(d_gIdx, d_gDist, d_sDist, n_ref, thr, st));
    if (type != 1) {
        if (!d_gIdx2 || !d_gDist2 || !d_sIdx2 || !d_sDist2) return fail(S3D_ERR_ARG, "reverse outputs are required");
        int* d_mask = nullptr;
        S3D_CUDA(cudaMallocAsync((void**)&d_mask, sizeof(int) * std::max(n_tar, 1), st));
        if (n_tar > 0) {
            S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_tar, 256), 256, 0, st, d_gIdx2, n_tar, -1);
            S3D_LAUNCH(fill_int_kernel, s3d_blocks(n_tar, 256), 256, 0, st, d_sIdx2, n_tar, -1);
            S3D_CUDA(cudaMemsetAsync(d_gDist2, 0, sizeof(float) * n_tar, st));
            S3D_CUDA(cudaMemsetAsync(d_sDist2, 0, sizeof(float) * n_tar, st));
        }
        S3D_TRY(s3d_count_mask_device(d_gIdx, n_ref, d_mask, n_tar, type == 2 ? 0 : 1, st));
        S3D_TRY(search_device(d_tar, n_tar, d_ref, n_ref, 0, d_mask, nullptr, nullptr, nullptr, nullptr, d_gDist2, d_gIdx2,
                              d_sDist2, d_sIdx2, st));
        S3D_TRY(s3d_ratio_filter_device(d_gIdx2, d_gDist2, d_sDist2, n_tar, thr, st));
        S3D_TRY(s3d_biject_filter_device(d_gIdx, n_ref, d_mask, d_gIdx2, st));
        S3D_CUDA(cudaFreeAsync(d_mask, st));
    }
    if (d_pair_ref && d_pair_tar && d_n_pairs) S3D_TRY(s3d_pairs_device(d_gIdx, n_ref, d_pair_ref, d_pair_tar, d_n_pairs, st));
    return S3D_OK;
}

int s3d_match(int type, const float* ref_desc, int n_ref, const float* tar_desc, int n_tar, double thr, int* gIdx,
              float* gDist, int* sIdx, float* sDist, int* gIdx2, float* gDist2, int* sIdx2, float* sDist2, int* pair_ref,
              int* pair_tar, int* n_pairs, double* times3) {
    return s3d_match_ex(type, ref_desc, n_ref, 0, tar_desc, n_tar, 0, thr, gIdx, gDist, sIdx, sDist, gIdx2, gDist2, sIdx2,
                        sDist2, pair_ref, pair_tar, n_pairs, times3);
}

int s3d_match_ex(int type, const float* ref_desc, int n_ref, int ref_on_device, const float* tar_desc, int n_tar,
                 int tar_on_device, double thr, int* gIdx, float* gDist, int* sIdx, float* sDist, int* gIdx2,
                 float* gDist2, int* sIdx2, float* sDist2, int* pair_ref, int* pair_tar, int* n_pairs, double* times3) {
    clear_error();
    if (type < 1 || type > 3) return fail(S3D_ERR_ARG, "match type %d (1 inject, 2 biject, 3 enhanced)", type);
    if (n_ref < 0 || n_tar < 0 || (n_ref > 0 && !ref_desc) || (n_tar > 0 && !tar_desc)) return fail(S3D_ERR_ARG, "bad argument");
    int dev;
    S3D_TRY(use_device(-1, &dev));
    cudaStream_t st;
    S3D_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t nr = std::max(n_ref, 1), nt = std::max(n_tar, 1);
    float *d_ref = nullptr, *d_tar = nullptr, *d_f = nullptr;
    int* d_i = nullptr;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int rc = S3D_OK;
    auto body = [&]() -> int {
        if (!(ref_on_device && n_ref)) S3D_CUDA(cudaMallocAsync((void**)&d_ref, sizeof(float) * kD * nr, st));
        if (!(tar_on_device && n_tar)) S3D_CUDA(cudaMallocAsync((void**)&d_tar, sizeof(float) * kD * nt, st));
        S3D_CUDA(cudaMallocAsync((void**)&d_f, sizeof(float) * 2 * (nr + nt), st));
        S3D_CUDA(cudaMallocAsync((void**)&d_i, sizeof(int) * (4 * nr + 2 * nt + 4), st));
        const float* q_ref = d_ref;
        const float* q_tar = d_tar;
        if (ref_on_device && n_ref) q_ref = ref_desc;
        else if (n_ref) S3D_CUDA(cudaMemcpyAsync(d_ref, ref_desc, sizeof(float) * kD * (size_t)n_ref, cudaMemcpyHostToDevice, st));
        if (tar_on_device && n_tar) q_tar = tar_desc;
        else if (n_tar) S3D_CUDA(cudaMemcpyAsync(d_tar, tar_desc, sizeof(float) * kD * (size_t)n_tar, cudaMemcpyHostToDevice, st));
        float *dg = d_f, *ds = d_f + nr, *dg2 = d_f + 2 * nr, *ds2 = d_f + 2 * nr + nt;
        int *ig = d_i, *is = d_i + nr, *pr = d_i + 2 * nr, *pt = d_i + 3 * nr, *ig2 = d_i + 4 * nr, *is2 = d_i + 4 * nr + nt,
            *np = d_i + 4 * nr + 2 * nt;
        S3D_CUDA(cudaMemsetAsync(np, 0, sizeof(int), st));
        S3D_CUDA(cudaMemsetAsync(d_f, 0, sizeof(float) * 2 * (nr + nt), st));
        S3D_CUDA(cudaEventRecord(e0, st));
        S3D_TRY(s3d_match_device(type, q_ref, n_ref, q_tar, n_tar, thr, ig, dg, is, ds, ig2, dg2, is2, ds2, pr, pt, np, st));
        S3D_CUDA(cudaEventRecord(e1, st));
        int h_np = 0;
        S3D_CUDA(cudaMemcpyAsync(&h_np, np, sizeof(int), cudaMemcpyDeviceToHost, st));
        auto back = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
            return (h && bytes) ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
        };
        S3D_CUDA(back(gIdx, ig, sizeof(int) * n_ref));
        S3D_CUDA(back(sIdx, is, sizeof(int) * n_ref));
        S3D_CUDA(back(gDist, dg, sizeof(float) * n_ref));
        S3D_CUDA(back(sDist, ds, sizeof(float) * n_ref));
        if (type != 1) {
            S3D_CUDA(back(gIdx2, ig2, sizeof(int) * n_tar));
            S3D_CUDA(back(sIdx2, is2, sizeof(int) * n_tar));
            S3D_CUDA(back(gDist2, dg2, sizeof(float) * n_tar));
            S3D_CUDA(back(sDist2, ds2, sizeof(float) * n_tar));
        }
        S3D_CUDA(back(pair_ref, pr, sizeof(int) * n_ref));
        S3D_CUDA(back(pair_tar, pt, sizeof(int) * n_ref));
        S3D_CUDA(cudaStreamSynchronize(st));
        if (n_pairs) *n_pairs = h_np;
        if (times3) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            times3[0] = times3[2] = ms * 1e-3;
            times3[1] = 0;
        }
        return S3D_OK;
    };
    rc = body();
    void* ptrs[] = {d_ref, d_tar, d_f, d_i};
    for (void* p : ptrs) if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return rc;
}

// muBruteMatcher over the ranks of a communicator (collective): the searched set of each direction is sharded over
// the ranks, the exact re-rank over the queries; every rank passes the FULL sets (device memory, replicated) and
// receives the complete outputs (device memory).  Results are bit-identical to s3d_match_device.
int s3d_match_sharded(s3d_comm_t comm, int type, const float* d_ref, int n_ref, const float* d_tar, int n_tar, double thr,
                      int* d_gIdx, float* d_gDist, int* d_sIdx, float* d_sDist, int* d_gIdx2, float* d_gDist2, int* d_sIdx2,
                      float* d_sDist2, int* d_pair_ref, int* d_pair_tar, int* d_n_pairs, void* stream) {
    clear_error();
    if (!comm || !comm->impl) return fail(S3D_ERR_ARG, "null communicator");
    int dev;
    S3D_TRY(use_device(comm->impl->device, &dev));
    return match_sharded_device(*comm->impl, type, d_ref, n_ref, d_tar, n_tar, thr, d_gIdx, d_gDist, d_sIdx, d_sDist, d_gIdx2, d_gDist2,
                                d_sIdx2, d_sDist2, d_pair_ref, d_pair_tar, d_n_pairs, (cudaStream_t)stream);
}

// The same in ONE process over `ndev` devices (host threads + peer copies): host descriptor sets in, host outputs out
// (as s3d_match).  devices == NULL: `ndev` logical shards on the current device.
int s3d_match_multi(int type, const float* ref_desc, int n_ref, const float* tar_desc, int n_tar, double thr, const int* devices,
                    int ndev, int* gIdx, float* gDist, int* sIdx, float* sDist, int* gIdx2, float* gDist2, int* sIdx2, float* sDist2,
                    int* pair_ref, int* pair_tar, int* n_pairs, double* times3) {
    clear_error();
    if (type < 1 || type > 3) return fail(S3D_ERR_ARG, "match type %d (1 inject, 2 biject, 3 enhanced)", type);
    if (n_ref < 0 || n_tar < 0 || (n_ref > 0 && !ref_desc) || (n_tar > 0 && !tar_desc) || ndev < 1 || ndev > 64) return fail(S3D_ERR_ARG, "bad argument");
    int cur = 0;
    S3D_TRY(use_device(-1, &cur));
    LocalGroup* group = local_group_create(ndev);
    std::vector<int> rc(ndev, S3D_OK);
    std::vector<std::string> msg(ndev);
    const size_t nr = std::max(n_ref, 1), nt = std::max(n_tar, 1);
    auto work = [&](int g) {
        int dev = 0;
        int r = use_device(devices ? devices[g] : cur, &dev);
        Comm* cm = r == S3D_OK ? local_comm_create(group, g, dev) : nullptr;
        if (r == S3D_OK && !cm) r = S3D_ERR_CUDA;
        cudaStream_t st = nullptr;
        float *d_ref = nullptr, *d_tar = nullptr, *d_f = nullptr;
        int* d_i = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        auto body = [&]() -> int {
            if (devices)
                for (int k = 0; k < ndev; ++k) {
                    int can = 0;
                    if (devices[k] != dev && cudaDeviceCanAccessPeer(&can, dev, devices[k]) == cudaSuccess && can)
                        if (cudaDeviceEnablePeerAccess(devices[k], 0) != cudaSuccess) cudaGetLastError();
                }
            S3D_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            S3D_CUDA(cudaEventCreate(&e0));
            S3D_CUDA(cudaEventCreate(&e1));
            S3D_CUDA(cudaMallocAsync((void**)&d_ref, sizeof(float) * kD * nr, st));
            S3D_CUDA(cudaMallocAsync((void**)&d_tar, sizeof(float) * kD * nt, st));
            S3D_CUDA(cudaMallocAsync((void**)&d_f, sizeof(float) * 2 * (nr + nt), st));
            S3D_CUDA(cudaMallocAsync((void**)&d_i, sizeof(int) * (4 * nr + 2 * nt + 4), st));
            if (n_ref) S3D_CUDA(cudaMemcpyAsync(d_ref, ref_desc, sizeof(float) * kD * (size_t)n_ref, cudaMemcpyHostToDevice, st));
            if (n_tar) S3D_CUDA(cudaMemcpyAsync(d_tar, tar_desc, sizeof(float) * kD * (size_t)n_tar, cudaMemcpyHostToDevice, st));
            float *dg = d_f, *ds = d_f + nr, *dg2 = d_f + 2 * nr, *ds2 = d_f + 2 * nr + nt;
            int *ig = d_i, *is = d_i + nr, *pr = d_i + 2 * nr, *pt = d_i + 3 * nr, *ig2 = d_i + 4 * nr, *is2 = d_i + 4 * nr + nt,
                *np = d_i + 4 * nr + 2 * nt;
            S3D_CUDA(cudaMemsetAsync(np, 0, sizeof(int), st));
            S3D_CUDA(cudaMemsetAsync(d_f, 0, sizeof(float) * 2 * (nr + nt), st));
            S3D_CUDA(cudaEventRecord(e0, st));
            S3D_TRY(match_sharded_device(*cm, type, d_ref, n_ref, d_tar, n_tar, thr, ig, dg, is, ds, ig2, dg2, is2, ds2, pr, pt, np, st));
            S3D_CUDA(cudaEventRecord(e1, st));
            if (g == 0) {
                int h_np = 0;
                S3D_CUDA(cudaMemcpyAsync(&h_np, np, sizeof(int), cudaMemcpyDeviceToHost, st));
                auto back = [&](void* h, const void* d, size_t bytes) -> cudaError_t {
                    return (h && bytes) ? cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
                };
                S3D_CUDA(back(gIdx, ig, sizeof(int) * n_ref));
                S3D_CUDA(back(sIdx, is, sizeof(int) * n_ref));
                S3D_CUDA(back(gDist, dg, sizeof(float) * n_ref));
                S3D_CUDA(back(sDist, ds, sizeof(float) * n_ref));
                if (type != 1) {
                    S3D_CUDA(back(gIdx2, ig2, sizeof(int) * n_tar));
                    S3D_CUDA(back(sIdx2, is2, sizeof(int) * n_tar));
                    S3D_CUDA(back(gDist2, dg2, sizeof(float) * n_tar));
                    S3D_CUDA(back(sDist2, ds2, sizeof(float) * n_tar));
                }
                S3D_CUDA(back(pair_ref, pr, sizeof(int) * n_ref));
                S3D_CUDA(back(pair_tar, pt, sizeof(int) * n_ref));
                S3D_CUDA(cudaStreamSynchronize(st));
                if (n_pairs) *n_pairs = h_np;
                if (times3) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    times3[0] = times3[2] = ms * 1e-3;
                    times3[1] = 0;
                }
            }
            S3D_CUDA(cudaStreamSynchronize(st));
            return S3D_OK;
        };
        if (r == S3D_OK) r = body();
        if (r != S3D_OK) { msg[g] = s3d_last_error(); local_group_fail(group); }
        void* ptrs[] = {d_ref, d_tar, d_f, d_i};
        if (st) {
            for (void* p : ptrs) if (p) cudaFreeAsync(p, st);
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
        }
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        delete cm;
        rc[g] = r;
    };
    if (ndev == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < ndev; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    local_group_destroy(group);
    cudaSetDevice(cur);
    for (int g = 0; g < ndev; ++g)
        if (rc[g] != S3D_OK) {
            int first = g;
            for (int k = 0; k < ndev; ++k)
                if (rc[k] != S3D_OK && msg[k].find("another shard") == std::string::npos) { first = k; break; }
            return fail(rc[first], "shard %d: %s", first, msg[first].c_str());
        }
    return S3D_OK;
}

}  // extern "C"
