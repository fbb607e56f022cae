// s3d_ctx.h — the extraction handle and its stage functions, shared by s3d_extract.cu (the stages and every kernel
// launch of the extraction path) and s3d_slab.cu (the multi-GPU orchestration of those stages).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "s3d_common.h"

namespace s3d {

constexpr int kCtxMaxOct = 16;   // == kMaxOct of s3d_kernels.cuh
constexpr int kCtxMaxG = 12;     // == kMaxG
constexpr int kCtxMaxHW = 16;    // == kMaxHW

// mirror of s3d_kernels.cuh::Taps (kept layout-identical; s3d_extract.cu static_asserts it)
struct TapsH {
    int hw;
    float w[2 * kCtxMaxHW + 1];
    int ext_il[kCtxMaxHW + 1];
    float ext_frac[kCtxMaxHW + 1];
};

// ---- per-kernel-class device timing (params.profile) ---------------------------------------
enum KCls { K_MAXABS, K_NORMALIZE, K_BLUR_X, K_BLUR_Y, K_BLUR_XY, K_BLUR_Z_DOG, K_BLUR_GENERIC, K_DOWNSAMPLE, K_DETECT, K_COMPACT,
            K_ORIENT, K_ORIENT_EXACT, K_SURVIVORS, K_DESCRIBE, K_DESCRIBE_REDO, K_BLUR_XYZ, K_NCLS };

// Timing events are recycled through a per-device free list: a 512^3 step brackets ~120 launches, and creating and
// destroying 240 events per step cost more host time than the small octaves' kernels take.
struct EventPool {
    std::mutex mu;
    std::vector<cudaEvent_t> free_[64];
    cudaEvent_t get(int dev) {
        {
            std::lock_guard<std::mutex> lk(mu);
            auto& f = free_[dev & 63];
            if (!f.empty()) { cudaEvent_t e = f.back(); f.pop_back(); return e; }
        }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void put(int dev, cudaEvent_t e) {
        std::lock_guard<std::mutex> lk(mu);
        free_[dev & 63].push_back(e);
    }
};
extern EventPool g_event_pool;

struct Prof {
    bool on = false;
    int dev = 0;
    cudaStream_t st = nullptr;
    struct Rec { int cls; cudaEvent_t a, b; double bytes; };
    std::vector<Rec> recs;
    double ms[K_NCLS] = {0};
    long long cnt[K_NCLS] = {0};
    double bytes[K_NCLS] = {0};
    // (measured: letting back-to-back scopes share a boundary event saves 0.2 ms of a 19 ms step but charges the
    // inter-kernel gaps to the next class - blur_xy +7 % - so every scope keeps its own two events)
    void begin(int cls, double by) {
        if (!on) return;
        Rec r; r.cls = cls; r.bytes = by;
        r.a = g_event_pool.get(dev); r.b = g_event_pool.get(dev);
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end() {
        if (!on) return;
        cudaEventRecord(recs.back().b, st);
    }
    void resolve() {  // after the stream has been synchronised
        for (auto& r : recs) {
            float t = 0;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.cls] += t; cnt[r.cls]++; bytes[r.cls] += r.bytes; }
            g_event_pool.put(dev, r.a); g_event_pool.put(dev, r.b);
        }
        recs.clear();
    }
};
struct ProfScope {
    Prof* p;
    ProfScope(Prof* p_, int cls, double by) : p(p_) { if (p) p->begin(cls, by); }
    ~ProfScope() { if (p) p->end(); }
};

}  // namespace s3d

struct MeshConstOpaque;

// ---------------------------------------------------------------------------------------------
struct s3d_ctx {
    s3d_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    s3d::Prof prof;
    int nx = 0, ny = 0, nz = 0;
    size_t n0 = 0;
    int noct = 0, G = 0, D = 0, L = 0;
    int dims[s3d::kCtxMaxOct][3];
    size_t nvox[s3d::kCtxMaxOct];
    float sig[s3d::kCtxMaxG];
    s3d::TapsH taps[s3d::kCtxMaxG];
    float* d_input = nullptr;       // normalised input (Host_Im); a shard: local planes [za[0], zb[0])
    std::vector<float*> gss, dog;   // device levels
    float* d_tmp[2] = {nullptr, nullptr};
    unsigned* d_slots = nullptr;    // [0] input max|v|, [1 + o*D + i] max|DoG(o,i)|
    float* d_thres = nullptr;       // noct * L thresholds
    void* d_mesh = nullptr;         // MeshConst
    // sparse stage
    int n_extre = 0, n_kps = 0;
    s3d_keypoint* d_extre = nullptr;
    int* d_codes = nullptr;
    int* d_xyz5 = nullptr;
    s3d_keypoint* d_kps = nullptr;
    float* d_desc = nullptr;
    int n_rechecked = 0, n_flipped = 0;
    int* d_redo = nullptr;          // keypoint indices for the FP32 descriptor kernel (freed in s3d_wait)
    int n_desc_redo = 0;            // keypoints the fixed-point descriptor kernel handed to the FP32 one
    int* d_counts = nullptr;        // the run's device-side counters (read once, in s3d_wait)
    int cap_extre = 0, cap_kps = 0; // capacities of the detection / keypoint buffers (0: choose in stage_sparse)
    size_t own_total = 0;           // owned voxels over all octaves (the capacities' yardstick)
    int n_resized = 0;              // times the sparse stage was repeated with larger buffers
    bool ran = false, levels_alive = false, queued = false, h2d_pending = false, d2h_pending = false;
    // z-slab sharding.  Unsharded: slab = false, za = p0 = 0, zb = p1 = nz_o.
    // A shard OWNS global planes [p0[o], p1[o]) of octave o (plane k of octave o belongs to the owner of octave-0 plane
    // k * 2^o) and keeps local buffers for planes [za[o], zb[o]) (owned + halo).  nz / dims[][2] / nvox[] stay the GLOBAL
    // sizes.  An octave with full[o] set is REPLICATED: every shard holds and computes all of its planes (za = 0,
    // zb = nz_o) and only detection / orientation / description are restricted to the owned planes.
    bool slab = false;
    int own0 = 0, own1 = 0, halo = 0;      // owned octave-0 planes, halo depth (planes, every octave)
    int za[s3d::kCtxMaxOct], zb[s3d::kCtxMaxOct], p0[s3d::kCtxMaxOct], p1[s3d::kCtxMaxOct];
    bool full[s3d::kCtxMaxOct];
    int first_full = 1 << 30;               // first replicated octave (>= noct: none)
    int stage = 0;                          // 0 created, 1 initialised, 100 sparse done
    size_t lvox(int o) const { return (size_t)dims[o][0] * dims[o][1] * (size_t)(zb[o] - za[o]); }
    size_t plane(int o) const { return (size_t)dims[o][0] * dims[o][1]; }
    cudaEvent_t ev[8];
    bool ev_ok = false;
    double timers[10] = {0};
    // phase boundaries of a sharded run (s3d_slab.cu) and their device times in ms (s3d_slab_phases)
    cudaEvent_t ph_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double ph_ms[8] = {0};
};

namespace s3d {

// ---- stages (s3d_extract.cu) --------------------------------------------------------------------
int ctx_common_init(s3d_ctx* c, int nx, int ny, int nz, const s3d_params* p);
// planes a shard needs beyond its owned range for the descriptor windows of keypoint level `lvl` (1..L) — in planes of
// the level's own octave, the same for every octave
int slab_window_halo(const s3d_ctx* c, int lvl);
int slab_halo_from_params(const s3d_params* p, int* halo);
// octave count, taps, per-octave extents (uses c->slab / own0 / own1 / halo / first_full), level allocation
int stage_init(s3d_ctx* c);
// max|v| of `n` floats into c->d_slots[0] (atomic max: call it on disjoint pieces) / v = v / slots[0] in place
int stage_input_max(s3d_ctx* c, const float* d, size_t n);
int stage_input_normalize(s3d_ctx* c, const float* src, float* dst, size_t n);
// octave seed: decimate planes [k0, k1) of octave o from level L of octave o-1 (DownSample_3D)
int stage_seed(s3d_ctx* c, int o, int k0, int k1);
// Gaussian level i of octave o on output planes [zlo, zhi) (+ DoG i-1 and its max for i >= 1); needs the source level
// valid on [zlo - hw_i, zhi + hw_i) clipped to the level
int stage_level(s3d_ctx* c, int o, int i, int zlo, int zhi);
int stage_sparse(s3d_ctx* c);
void free_levels(s3d_ctx* c);

}  // namespace s3d
