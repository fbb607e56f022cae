// s3d_match_internal.h — pieces shared by the exact (s3d_match.cu) and tensor-core
// (s3d_match_tc.cu) matcher units.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>

namespace s3d {

// Running best / second-best of one query: double dots (KP_squareSum, Src/cMatcher.cc:17-23) and
// GLOBAL database indices; index -1 = empty slot.
struct Top2 {
    double d1, d2;
    int i1, i2;
};

__device__ __forceinline__ void top2_init(Top2& t) {
    t.d1 = (double)FLT_MIN; t.d2 = (double)FLT_MIN;  // Src/cMatcher.cc:54-55
    t.i1 = -1; t.i2 = -1;
}

// is (da, ia) ahead of (db, ib) in (dot desc, index asc)?  index -1 = empty slot (never ahead)
__device__ __forceinline__ bool ahead(double da, int ia, double db, int ib) {
    if (ia < 0) return false;
    if (ib < 0) return true;
    return da > db || (da == db && ia < ib);
}

// Insert one candidate (s, j).  Candidates with s <= FLT_MIN never enter (strict '>' against the
// FLT_MIN initial values, Src/cMatcher.cc:60,66).  The total order (dot desc, index asc) equals
// the reference's ascending-j strict-'>' scan.
__device__ __forceinline__ void top2_push(Top2& t, double s, int j) {
    if (!(s > (double)FLT_MIN)) return;
    if (j == t.i1 || j == t.i2) return;  // the same database row offered twice (merge paths)
    if (ahead(s, j, t.d1, t.i1)) {
        t.d2 = t.d1; t.i2 = t.i1;
        t.d1 = s; t.i1 = j;
    } else if (ahead(s, j, t.d2, t.i2)) {
        t.d2 = s; t.i2 = j;
    }
}

__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) {
    if (b.i1 >= 0) top2_push(a, b.d1, b.i1);
    if (b.i2 >= 0) top2_push(a, b.d2, b.i2);
}

// s3d_match_tc.cu.  Returns S3D_OK, an error code, or kTcRefused when the sets violate the precondition of the
// tensor-core guard (an entry that is negative, above 1 or not finite): nothing was written to d_out / d_fb_list and the
// caller must run the exact kernel for these rows.
constexpr int kTcRefused = -1000;
// Working set of one tensor-core search (device temporaries from the stream-ordered pool; release with tc_free).
struct TcWork {
    void* q16 = nullptr;      // __half [nql][768], x64
    void* db16 = nullptr;     // __half [nd][768], x64
    float* cand_val = nullptr;
    int* cand_idx = nullptr;  // [nql][parts * 8]
    int* d_bad = nullptr;     // device flag: an entry outside [0, 1] or not finite was seen
    int nql = 0, nd = 0, parts = 0, ncand = 0, tiles_per_part = 0, n_dbtiles = 0, n_units = 0, sms = 148;
    bool pair = false;
};
int tc_convert(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, cudaStream_t st, int variant, TcWork& w);
int tc_topk(cudaStream_t st, TcWork& w);
int tc_merge8(cudaStream_t st, const TcWork& w, int idx_offset, float* out_val, int* out_idx);
int tc_rerank(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, const float* cand_val,
              const int* cand_idx, int nlists, size_t stride_q, size_t stride_p, Top2* d_out, int* d_fb_list, int* d_fb_count,
              cudaStream_t st);
void tc_free(TcWork& w, cudaStream_t st);
int tc_search(const float* d_q, const int* d_qlist, int nql, const float* d_db, int nd, int db_offset, Top2* d_out,
              int* d_fb_list, int* d_fb_count, cudaStream_t st, int variant);

}  // namespace s3d
