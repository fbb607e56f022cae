// s3d_extract.cu — host orchestration + C ABI of the extraction path.
// Compiled with -fmad=false (see s3d_kernels.cuh).  Mirrors, stage by stage,
// CSIFT3D::KpSiftAlgorithm (/root/reference/3DSIFT/Src/cSIFT3D.cc:165-235).
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "s3d_common.h"
#include "s3d_devcache.h"
#include "s3d_kernels.cuh"

namespace s3d {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void clear_error() { g_err[0] = 0; }

int use_device(int device, int* resolved) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(S3D_ERR_CUDA, "no CUDA device available (%s); libsift3d_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) S3D_CUDA(cudaGetDevice(&device));
    if (device >= n) return fail(S3D_ERR_ARG, "device %d out of range (%d devices)", device, n);
    S3D_CUDA(cudaSetDevice(device));
    static std::mutex mu;
    static bool pool_set[64] = {false};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (device < 64 && !pool_set[device]) {
            int major = 0;
            S3D_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
            if (major != 10)
                return fail(S3D_ERR_CUDA, "device %d has compute capability %d.x; this library is built for sm_100a only",
                            device, major);
            // keep freed pyramid memory in the stream-ordered pool: the next volume reuses it
            cudaMemPool_t pool;
            S3D_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t thr = UINT64_MAX;
            S3D_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
            pool_set[device] = true;
        }
    }
    if (resolved) *resolved = device;
    return S3D_OK;
}

// ---- host-side constants, computed exactly as the reference does (App. A.1, Q10) -------------

// Src/cSIFT3D.cc:270-287,299
static void host_sigmas(int L, float sigma_default, float sigma_n, float* sig) {
    float k = (float)pow(2.0, 1.0 / L);
    float base = (float)(sigma_default * pow(2.0, -1.0 / 3.0));
    sig[0] = sqrtf(base * base - sigma_n * sigma_n);
    for (int i = 1; i < L + 3; i++) {
        float sig_prev = (float)(pow((double)k, (double)(i - 1)) * base);
        float sig_total = sig_prev * k;
        sig[i] = sqrtf(sig_total * sig_total - sig_prev * sig_prev);
    }
}

// Src/cSIFT3D.cc:541-572.  Returns hw (or -1 if the kernel is wider than kMaxHW).
static int host_taps(float sigma, Taps* t) {
    sigma = sigma > 0 ? sigma : 0;
    int c = (int)ceil(sigma * 3.0);
    const int hw = sigma > 0 ? (c > 1 ? c : 1) : 1;
    if (hw > kMaxHW) return -1;
    const int width = 2 * hw + 1;
    float acc = 0;
    for (int i = 0; i < width; i++) {
        float x = (float)(i - hw);
        x = (float)((double)x / ((double)sigma + DBL_EPSILON));
        t->w[i] = (float)exp(-0.5 * (double)x * (double)x);
        acc += t->w[i];
    }
    for (int i = 0; i < width; i++) t->w[i] /= acc;
    for (int i = width; i < 2 * kMaxHW + 1; i++) t->w[i] = 0.0f;
    t->hw = hw;
    return hw;
}

// Src/cUtil.cc:182,209-210
static float host_level_scale(int o, int s, int L, float sigma_default) {
    double sigma0 = sigma_default * pow(2.0, -1.0 / 3.0);
    double scale_factor = pow(2.0, o + (double)s / L);
    return (float)(scale_factor * sigma0);
}

// Icosahedron, Src/cUtil.cc:19-55 + Initialize_geometry :113-175, and the face-only terms of
// cart2bary (Src/cSIFT3D.cc:1600-1619,1634) in the same FP32 arithmetic.  `volatile` keeps the
// host compiler from contracting or reassociating anything.
static void host_mesh(MeshConst* M) {
    const double gr = 1.6180339887;
    const double vert[36] = {0, 1, gr, 0, -1, gr, 0, 1, -gr, 0, -1, -gr, 1, gr, 0, -1, gr, 0,
                             1, -gr, 0, -1, -gr, 0, gr, 0, 1, -gr, 0, 1, gr, 0, -1, -gr, 0, -1};
    const int faces[60] = {0, 1, 8, 0, 8, 4, 0, 4, 5, 0, 5, 9, 0, 9, 1, 1, 6, 8, 8, 6, 10, 8, 10, 4, 4, 10, 2,
                           4, 2, 5, 5, 2, 11, 5, 11, 9, 9, 11, 7, 9, 7, 1, 1, 7, 6, 3, 6, 7, 3, 7, 11, 3, 11, 2,
                           3, 2, 10, 3, 10, 6};
    auto mul = [](float a, float b) { volatile float r = a * b; return (float)r; };
    auto add = [](float a, float b) { volatile float r = a + b; return (float)r; };
    auto sub = [](float a, float b) { volatile float r = a - b; return (float)r; };
    float cen[20][3];
    for (int i = 0; i < 20; i++) {
        float v[3][3];
        for (int j = 0; j < 3; j++) {
            M->idx[i][j] = faces[i * 3 + j];
            for (int c = 0; c < 3; c++) v[j][c] = (float)vert[faces[i * 3 + j] * 3 + c];
            float n2 = add(add(mul(v[j][0], v[j][0]), mul(v[j][1], v[j][1])), mul(v[j][2], v[j][2]));
            double mag = (double)sqrtf(n2);
            double s = 1.0 / mag;
            for (int c = 0; c < 3; c++) v[j][c] = (float)((double)v[j][c] * s);
        }
        float a[3], b[3], n[3];
        for (int c = 0; c < 3; c++) { a[c] = sub(v[2][c], v[1][c]); b[c] = sub(v[1][c], v[0][c]); }
        n[0] = sub(mul(a[1], b[2]), mul(a[2], b[1]));
        n[1] = sub(mul(a[2], b[0]), mul(a[0], b[2]));
        n[2] = sub(mul(a[0], b[1]), mul(a[1], b[0]));
        float dot = add(add(mul(n[0], v[0][0]), mul(n[1], v[0][1])), mul(n[2], v[0][2]));
        if (dot < 0)
            for (int c = 0; c < 3; c++) { float tmp = v[0][c]; v[0][c] = v[1][c]; v[1][c] = tmp; }
        float* e1 = M->e1[i]; float* e2 = M->e2[i]; float* t = M->t[i]; float* q = M->q[i];
        for (int c = 0; c < 3; c++) {
            e1[c] = sub(v[1][c], v[0][c]);
            e2[c] = sub(v[2][c], v[0][c]);
            t[c] = (float)((double)v[0][c] * (-1.0));
        }
        q[0] = sub(mul(t[1], e1[2]), mul(t[2], e1[1]));
        q[1] = sub(mul(t[2], e1[0]), mul(t[0], e1[2]));
        q[2] = sub(mul(t[0], e1[1]), mul(t[1], e1[0]));
        M->qe2[i] = add(add(mul(q[0], e2[0]), mul(q[1], e2[1])), mul(q[2], e2[2]));
        for (int c = 0; c < 3; c++) cen[i][c] = v[0][c] + v[1][c] + v[2][c];
    }
    // antipodal face pairs (the icosahedron is centrally symmetric): 10 centroids decide 20 faces
    bool used[20] = {false};
    int np = 0;
    for (int i = 0; i < 20; i++) {
        if (used[i]) continue;
        int opp = -1;
        for (int j = i + 1; j < 20; j++)
            if (!used[j] && fabsf(cen[i][0] + cen[j][0]) + fabsf(cen[i][1] + cen[j][1]) + fabsf(cen[i][2] + cen[j][2]) < 1e-4f) opp = j;
        used[i] = true;
        if (opp >= 0) used[opp] = true;
        for (int c = 0; c < 3; c++) M->cen10[np][c] = cen[i][c];
        M->pos[np] = i;
        M->neg[np] = opp >= 0 ? opp : i;
        np++;
    }
    // Octant pre-selection tables, derived from the mesh itself (vertex positions V0 = -t, V1 = V0+e1, V2 = V0+e2).
    auto vtx = [&](int f, int j, double* out) {
        for (int c = 0; c < 3; c++) out[c] = -(double)M->t[f][c] + (j == 1 ? (double)M->e1[f][c] : (j == 2 ? (double)M->e2[f][c] : 0.0));
    };
    auto same = [](const double* a, const double* b) { return fabs(a[0] - b[0]) + fabs(a[1] - b[1]) + fabs(a[2] - b[2]) < 1e-4; };
    for (int o = 0; o < 8; o++) {
        const double sg[3] = {(o & 1) ? -1.0 : 1.0, (o & 2) ? -1.0 : 1.0, (o & 4) ? -1.0 : 1.0};
        // the octant face: all three vertices inside the closed octant
        int fo = 0;
        for (int f = 0; f < 20; f++) {
            bool in = true;
            for (int j = 0; j < 3 && in; j++) {
                double v[3];
                vtx(f, j, v);
                for (int c = 0; c < 3; c++) in = in && v[c] * sg[c] > -1e-6;
            }
            if (in) fo = f;
        }
        M->octface[o][0] = fo;
        for (int k = 0; k < 3; k++) {  // edge k joins vertices k and (k+1)%3; the third vertex is inside
            double a[3], b[3], cth[3];
            vtx(fo, k, a); vtx(fo, (k + 1) % 3, b); vtx(fo, (k + 2) % 3, cth);
            double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
            if (n[0] * cth[0] + n[1] * cth[1] + n[2] * cth[2] < 0) for (int c = 0; c < 3; c++) n[c] = -n[c];
            for (int c = 0; c < 3; c++) M->octn[o][k][c] = (float)n[c];
            int nb = fo;
            for (int f = 0; f < 20; f++) {  // the other face that contains both a and b
                if (f == fo) continue;
                bool ha = false, hb = false;
                for (int j = 0; j < 3; j++) {
                    double v[3];
                    vtx(f, j, v);
                    ha = ha || same(v, a);
                    hb = hb || same(v, b);
                }
                if (ha && hb) nb = f;
            }
            M->octface[o][1 + k] = nb;
        }
    }
}

// Extended-line table of a pass along an axis of length n (Src/cSIFT3D.cc:751-760): for tap
// coordinate c = q = n-1+e the reference samples at c' = 2(n-1) - c - 0.1f, lo = (int)c',
// frac = c' - lo.  Same FP32 operations as the reference (volatile: no wider evaluation).
// A z-slab shard passes the GLOBAL line length and the global coordinate of its first local plane:
// frac is "whatever FP32 gives" for c' (App. A.2), so it depends on the magnitude of n — a shard that
// holds the top of the volume must blend with the global c' (and address il relative to its buffer).
static Taps with_ext(const Taps& t0, int n, int n_glob = -1, int off = 0) {
    Taps t = t0;
    const bool top_is_global = n_glob > 0 && off + n == n_glob;
    const int ng = top_is_global ? n_glob : n;
    for (int e = 0; e <= kMaxHW; e++) {
        volatile float c = (float)(2 * (ng - 1));
        c = c - (float)(ng - 1 + e);
        c = c - 0.1f;
        int il = (int)c;
        volatile float fr = c - (float)il;
        t.ext_il[e] = top_is_global ? il - off : il;
        t.ext_frac[e] = fr;
    }
    return t;
}

// Descriptor accumulation: 0 = fixed-point shared-memory atomics with FP32 redo of overflow-flagged
// keypoints (default), 1 = FP32 staged/ordered accumulation for every keypoint, 2 = as 0 but with a
// deliberately tiny scale margin so that (nearly) every keypoint takes the redo path (tests).
// Initial value from the environment: S3D_DESC_PATH=0|1|2.
static int initial_describe_path() {
    const char* e = getenv("S3D_DESC_PATH");
    return (e && e[0] >= '0' && e[0] <= '2' && !e[1]) ? e[0] - '0' : 0;
}
static std::atomic<int> g_describe_path{initial_describe_path()};

static bool desc_order_enabled() {
    static const bool on = !(getenv("S3D_DESC_ORDER") && getenv("S3D_DESC_ORDER")[0] == '0');
    return on;
}

static bool supported_fast_hw(int hw) { return hw == 2 || hw == 3 || hw == 4 || hw == 5 || hw == 6 || hw == 8; }

template <int HW>
static void launch_x(const float* src, float* dst, int nx, ll nrows, const Taps& t, ll total, cudaStream_t st) {
    ll threads = nrows * (nx >> 2);
    auto kfn = blur_x_kernel<HW>;
    S3D_LAUNCH(kfn, s3d_blocks((size_t)threads, 256), 256, 0, st, src, dst, nx, (unsigned)threads, t);
}

static int pick_seg(int nx4, int n_other, int n) {
    // longest segment that still yields >= ~256k threads, so small octaves keep the GPU busy
    static const int seg_env = [] { const char* e = getenv("S3D_MARCH_SEG"); return e ? atoi(e) : 0; }();  // experiments
    int seg = seg_env > 0 ? seg_env : 64;
    while (seg > 8 && (ll)nx4 * n_other * ((n + seg - 1) / seg) < 262144) seg >>= 1;
    return seg;
}

template <int HW>
static void launch_march(const float* src, float* dst, int nx, int n, ll st_m, int n_other, ll st_other, const Taps& t,
                         ll total, const float* prev, float* dog, unsigned* slot, cudaStream_t st) {
    const int nx4 = nx >> 2;
    const int seg = pick_seg(nx4, n_other, n);
    const int nseg = (n + seg - 1) / seg;
    ll threads = (ll)nx4 * n_other * nseg;
    // S3D_ZVAR: 0 = register-ring prefetch (blur_march_kernel), 1 = cp.async ring (blur_marchc_kernel, default)
    static const int zvar = [] { const char* e = getenv("S3D_ZVAR"); return (e && e[0] >= '0' && e[0] <= '1' && !e[1]) ? e[0] - '0' : 1; }();
    const unsigned blocks = s3d_blocks((size_t)threads, 128);
    if (dog) {
        if (zvar == 0) {
            auto kfn = blur_march_kernel<HW, true>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t, prev, dog, slot);
        } else {
            auto kfn = blur_marchc_kernel<HW, true>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t, prev, dog, slot);
        }
    } else {
        if (zvar == 0) {
            auto kfn = blur_march_kernel<HW, false>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t,
                       (const float*)nullptr, (float*)nullptr, (unsigned*)nullptr);
        } else {
            auto kfn = blur_marchc_kernel<HW, false>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t,
                       (const float*)nullptr, (float*)nullptr, (unsigned*)nullptr);
        }
    }
}

// ---- per-kernel-class device timing (params.profile) ---------------------------------------
enum KCls { K_MAXABS, K_NORMALIZE, K_BLUR_X, K_BLUR_Y, K_BLUR_XY, K_BLUR_Z_DOG, K_BLUR_GENERIC, K_DOWNSAMPLE, K_DETECT, K_COMPACT,
            K_ORIENT, K_ORIENT_EXACT, K_SURVIVORS, K_DESCRIBE, K_DESCRIBE_REDO, K_NCLS };
static const char* kClsName[K_NCLS] = {"maxabs", "normalize", "blur_x", "blur_y", "blur_xy", "blur_z_dog", "blur_generic", "downsample",
                                       "detect", "compact", "orient", "orient_exact", "survivors", "describe", "describe_redo"};
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
static EventPool g_event_pool;

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

#define S3D_HW_SWITCH(hw, CALL)            \
    switch (hw) {                          \
        case 2: CALL(2); break;            \
        case 3: CALL(3); break;            \
        case 4: CALL(4); break;            \
        case 5: CALL(5); break;            \
        case 6: CALL(6); break;            \
        case 8: CALL(8); break;            \
        default: break;                    \
    }

// One separable pass.  variant 0 = generic kernel; 1 = fast kernels when eligible.
// prev/dog/slot non-null fuses the DoG subtraction + max|DoG| (only meaningful on the last pass).
// nz_glob / z_off (z passes of a z-slab shard): global plane count and global index of local plane 0.
static void blur_pass(const float* src, float* dst, int nx, int ny, int nz, int axis, const Taps& t0, int variant,
                      const float* prev, float* dog, unsigned* slot, cudaStream_t st, Prof* prof = nullptr, int nz_glob = -1,
                      int z_off = 0) {
    const ll total = (ll)nx * ny * nz;
    const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const bool partial = axis == 2 && nz_glob > 0 && (z_off != 0 || nz != nz_glob);
    const Taps t = partial ? with_ext(t0, n, nz_glob, z_off) : with_ext(t0, n);
    const bool fast = variant == 1 && (nx % 4 == 0) && supported_fast_hw(t.hw) && n >= 2 * t.hw + 2 && total >= 4096 &&
                      !(axis == 0 && dog);
    // algorithmic bytes: read src + write dst (+ read prev + write dog on the fused pass)
    ProfScope ps(prof, !fast ? K_BLUR_GENERIC : (axis == 0 ? K_BLUR_X : (axis == 1 ? K_BLUR_Y : K_BLUR_Z_DOG)),
                 (dog ? 16.0 : 8.0) * (double)total);
    if (!fast) {
        S3D_LAUNCH(blur_generic_kernel, s3d_blocks((size_t)total, 256), 256, 0, st, src, dst, nx, ny, nz, axis, t, prev,
                   dog, slot, partial ? nz_glob : 0, partial ? z_off : 0);
        return;
    }
    if (axis == 0) {
#define CALLX(H) launch_x<H>(src, dst, nx, (ll)ny * nz, t, total, st)
        S3D_HW_SWITCH(t.hw, CALLX)
#undef CALLX
    } else if (axis == 1) {
#define CALLY(H) launch_march<H>(src, dst, nx, ny, (ll)nx, nz, (ll)nx * ny, t, total, prev, dog, slot, st)
        S3D_HW_SWITCH(t.hw, CALLY)
#undef CALLY
    } else {
#define CALLZ(H) launch_march<H>(src, dst, nx, nz, (ll)nx * ny, ny, (ll)nx, t, total, prev, dog, slot, st)
        S3D_HW_SWITCH(t.hw, CALLZ)
#undef CALLZ
    }
}

// Fused X + Y pass (blur_xy_kernel) when the shape allows it; returns false when the caller must run
// the two separate passes.  S3D_BLUR_XY=0 in the environment disables the fused kernel.
static int xy_seg_override() {
    const char* e = getenv("S3D_BLUR_XY_SEG");
    return e ? atoi(e) : 0;
}
static int blur_xy_mode() {
    // S3D_BLUR_XY: 0 = separate X and Y passes, 1 = blur_xy_kernel for hw <= 6 (separate passes at hw 8),
    // 2 = blur_xy_kernel for hw <= 3 and the compact-code blur_xyc_kernel for hw >= 4 (default),
    // 3 = blur_xyc_kernel for every hw
    const char* e = getenv("S3D_BLUR_XY");
    return (e && e[0] >= '0' && e[0] <= '3' && !e[1]) ? e[0] - '0' : 2;
}
static bool blur_xy_pass(const float* src, float* dst, int nx, int ny, int nz, const Taps& t0, cudaStream_t st, Prof* prof) {
    static const int mode = blur_xy_mode();
    static const int seg_env = xy_seg_override();
    const int hw = t0.hw, T = nx >> 2;
    // measured on B200 (profiles/r01_ncu_blur_xy.txt): blur_xy_kernel beats X + Y for hw <= 6 (octave 0:
    // 191/241/339/385/542 us against 412/434/460/515/645) but stalls on instruction fetch from hw 4 up (its
    // unrolled ring replicates the X body); blur_xyc_kernel is the same arithmetic in compact code
    const bool compact = mode == 3 || (mode == 2 && hw >= 4);
    if (mode == 0 || (mode == 1 && hw > 6) || nx % 4 != 0 || T < 1 || T > 128 || 128 % T != 0 || T <= hw + 1 ||
        !supported_fast_hw(hw) || nx < 2 * hw + 2 || ny < 2 * hw + 2 || (ll)nx * ny * nz < 4096)
        return false;
    const Taps tx = with_ext(t0, nx), ty = with_ext(t0, ny);
    const int lpc = 128 / T, zgroups = (nz + lpc - 1) / lpc;
    // y segment: long enough to amortise the 2*hw-row prologue, short enough for >= ~2 waves of CTAs
    int seg = seg_env > 0 ? seg_env : 128;
    if (seg_env <= 0)
        while (seg > 32 && (ll)zgroups * ((ny + seg - 1) / seg) < 148 * 4 * 2) seg >>= 1;
    const int nseg = (ny + seg - 1) / seg;
    const size_t smem = (size_t)lpc * kXYBufs * (nx + 2 * ((hw + 3) / 4 * 4)) * sizeof(float);
    ProfScope ps(prof, K_BLUR_XY, 8.0 * (double)nx * ny * nz);
#define CALLXY(H) S3D_LAUNCH(blur_xy_kernel<H>, (unsigned)(zgroups * nseg), 128, smem, st, src, dst, nx, ny, nz, seg, tx, ty)
#define CALLXYC(H) S3D_LAUNCH(blur_xyc_kernel<H>, (unsigned)(zgroups * nseg), 128, smem, st, src, dst, nx, ny, nz, seg, tx, ty)
    if (compact) {
        S3D_HW_SWITCH(hw, CALLXYC)
    } else {
        S3D_HW_SWITCH(hw, CALLXY)
    }
#undef CALLXY
#undef CALLXYC
    return true;
}

}  // namespace s3d

using namespace s3d;

// ---------------------------------------------------------------------------------------------
struct s3d_ctx {
    s3d_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    Prof prof;
    int nx = 0, ny = 0, nz = 0;
    size_t n0 = 0;
    int noct = 0, G = 0, D = 0, L = 0;
    int dims[kMaxOct][3];
    size_t nvox[kMaxOct];
    float sig[kMaxG];
    Taps taps[kMaxG];
    float* d_input = nullptr;       // normalised input (Host_Im)
    std::vector<float*> gss, dog;   // device levels
    float* d_tmp[2] = {nullptr, nullptr};
    unsigned* d_slots = nullptr;    // [0] input max|v|, [1 + o*D + i] max|DoG(o,i)|
    float* d_thres = nullptr;       // noct * L thresholds
    MeshConst* d_mesh = nullptr;
    // sparse stage
    int n_extre = 0, n_kps = 0;
    s3d_keypoint* d_extre = nullptr;
    int* d_codes = nullptr;
    int* d_xyz5 = nullptr;
    s3d_keypoint* d_kps = nullptr;
    float* d_desc = nullptr;
    int n_rechecked = 0, n_flipped = 0;
    int* d_redo = nullptr;          // [0] = count, [1..] = keypoint indices (freed in s3d_wait)
    int n_desc_redo = 0;            // keypoints the fixed-point descriptor kernel handed to the FP32 one
    bool ran = false, levels_alive = false, queued = false, h2d_pending = false, d2h_pending = false;
    // z-slab sharding.  Unsharded: slab = false, za = p0 = 0, zb = p1 = nz_o.
    // A shard OWNS global planes [p0[o], p1[o]) of octave o and keeps local buffers for planes
    // [za[o], zb[o]) (owned + halo).  nz / dims[][2] / nvox[] stay the GLOBAL sizes.
    bool slab = false;
    int own0 = 0, own1 = 0, halo = 0;      // owned octave-0 planes, halo depth (planes, every octave)
    int za[kMaxOct], zb[kMaxOct], p0[kMaxOct], p1[kMaxOct];
    int stage = 0;                          // 0 created, 1 initialised, 2 + o = octave o done, 100 sparse done
    int next_octave = 0;
    size_t lvox(int o) const { return (size_t)dims[o][0] * dims[o][1] * (size_t)(zb[o] - za[o]); }
    size_t plane(int o) const { return (size_t)dims[o][0] * dims[o][1]; }
    cudaEvent_t ev[8];
    bool ev_ok = false;
    double timers[10] = {0};
};

static void free_levels(s3d_ctx* c) {
    for (auto& p : c->gss) if (p) { s3d::dev_free(p, c->stream); p = nullptr; }
    for (auto& p : c->dog) if (p) { s3d::dev_free(p, c->stream); p = nullptr; }
    for (int i = 0; i < 2; i++) if (c->d_tmp[i]) { s3d::dev_free(c->d_tmp[i], c->stream); c->d_tmp[i] = nullptr; }
    c->levels_alive = false;
}

static int ctx_common_init(s3d_ctx* c, int nx, int ny, int nz, const s3d_params* p) {
    s3d_params def;
    s3d_default_params(&def);
    c->prm = p ? *p : def;
    if (nx < 8 || ny < 8 || nz < 8) return fail(S3D_ERR_ARG, "volume %dx%dx%d too small (min 8 per axis)", nx, ny, nz);
    if ((double)nx * ny * nz >= 4294967295.0) return fail(S3D_ERR_ARG, "volume too large for 32-bit voxel keys");
    c->L = c->prm.num_kp_levels;
    if (c->L < 1 || c->L + 3 > kMaxG) return fail(S3D_ERR_ARG, "num_kp_levels %d unsupported (1..%d)", c->L, kMaxG - 3);
    c->G = c->L + 3;
    c->D = c->L + 2;
    S3D_TRY(use_device(c->prm.device, &c->device));
    c->nx = nx; c->ny = ny; c->nz = nz;
    c->n0 = (size_t)nx * ny * nz;
    if (c->prm.stream) {
        c->stream = (cudaStream_t)c->prm.stream;
    } else {
        S3D_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    c->prof.on = c->prm.profile != 0;
    c->prof.st = c->stream;
    c->prof.dev = c->device;
    for (int i = 0; i < 8; i++) S3D_CUDA(cudaEventCreate(&c->ev[i]));
    c->ev_ok = true;
    return S3D_OK;
}

static int ctx_normalize(s3d_ctx* c, const float* d_raw) {
    // ctor: data_scale (Src/cUtil.cc:536-564)
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_input, c->n0 * sizeof(float), c->stream));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_slots, 256 * sizeof(unsigned), c->stream));
    S3D_CUDA(cudaMemsetAsync(c->d_slots, 0, 256 * sizeof(unsigned), c->stream));
    const unsigned grid = (unsigned)std::min<size_t>(s3d_blocks(c->n0 / 4 + 1, 256), 148 * 16);
    {
        ProfScope ps(&c->prof, K_MAXABS, 4.0 * c->n0);
        S3D_LAUNCH(maxabs_kernel, grid, 256, 0, c->stream, d_raw, c->n0, c->d_slots);
    }
    {
        ProfScope ps(&c->prof, K_NORMALIZE, 8.0 * c->n0);
        S3D_LAUNCH(normalize_kernel, grid, 256, 0, c->stream, d_raw, c->d_input, c->n0, c->d_slots);
    }
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

extern "C" {

int s3d_version(void) { return 100; }
const char* s3d_last_error(void) { return g_err; }
uint64_t s3d_launch_count(void) { return g_launches.load(); }

int s3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

void s3d_default_params(s3d_params* p) {
    p->num_kp_levels = 3;       // NUM_KP_LEVELS   Include/cSIFT3D.h:15
    p->sigma_default = 1.6f;    // SIGMA_DEFAULT   :13
    p->sigma_n_default = 1.15f; // SIGMA_N_DEFAULT :14
    p->peak_thresh = 0.1f;      // PEAK_THRESH     :18
    p->max_eig_thres = 0.9f;    // EIG_THRES       :19
    p->corner_thresh = 0.4f;    // CORNER_THRESH   :20
    p->device = -1;
    p->keep_levels = 0;
    p->exact_recheck = 1;
    p->profile = 0;
    p->stream = nullptr;
}

int s3d_selftest(int device) {
    clear_error();
    int dev;
    S3D_TRY(use_device(device, &dev));
    float* d;
    S3D_CUDA(cudaMalloc((void**)&d, 8 * sizeof(float)));
    // a*b+c differs between fused and unfused evaluation for these values
    const float a = 1.0f + 0x1p-12f, b = 1.0f + 0x1p-12f, c = -1.0f, dd = 3.0f;
    S3D_LAUNCH(selftest_kernel, 1, 1, 0, 0, a, b, c, dd, d);
    float h[8];
    S3D_CUDA(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d);
    if (h[1] == h[2]) return fail(S3D_ERR_STATE, "selftest values do not separate fma from mul+add");
    if (h[0] != h[1]) return fail(S3D_ERR_STATE, "FP32 a*b+c was contracted to FMA: build s3d_extract.cu with -fmad=false");
    if (h[3] != h[4]) return fail(S3D_ERR_STATE, "FP32 division is not IEEE (-prec-div=true required)");
    if (h[5] != h[6]) return fail(S3D_ERR_STATE, "FP32 sqrt is not IEEE (-prec-sqrt=true required)");
    if (h[7] != s3d_expf_ref(-dd)) return fail(S3D_ERR_STATE, "device expf_ref differs from host expf_ref");
    return S3D_OK;
}

static int create_host(const float* vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out, bool sync) {
    clear_error();
    if (!vol || !out) return fail(S3D_ERR_ARG, "null argument");
    s3d_ctx* c = new s3d_ctx();
    int r = ctx_common_init(c, nx, ny, nz, p);
    if (r != S3D_OK) { s3d_destroy(c); return r; }
    float* d_raw = nullptr;
    cudaEventRecord(c->ev[6], c->stream);
    if (s3d::dev_alloc((void**)&d_raw, c->n0 * sizeof(float), c->stream) != cudaSuccess ||
        cudaMemcpyAsync(d_raw, vol, c->n0 * sizeof(float), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        r = fail(S3D_ERR_CUDA, "H2D copy of the volume failed: %s", cudaGetErrorString(cudaGetLastError()));
        s3d_destroy(c);
        return r;
    }
    cudaEventRecord(c->ev[7], c->stream);
    r = ctx_normalize(c, d_raw);
    s3d::dev_free(d_raw, c->stream);
    if (r != S3D_OK) { s3d_destroy(c); return r; }
    c->h2d_pending = true;
    if (sync) {
        // the caller may free/reuse `vol` as soon as we return (the reference's ctor memcpy's it)
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
            r = fail(S3D_ERR_CUDA, "create: %s", cudaGetErrorString(cudaGetLastError()));
            s3d_destroy(c);
            return r;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]);
        c->timers[8] = ms * 1e-3;
        c->h2d_pending = false;
    }
    *out = c;
    return S3D_OK;
}

int s3d_create(const float* vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out) {
    return create_host(vol, nx, ny, nz, p, out, true);
}

int s3d_create_async(const float* vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out) {
    return create_host(vol, nx, ny, nz, p, out, false);
}

int s3d_create_device(const float* d_vol, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out) {
    clear_error();
    if (!d_vol || !out) return fail(S3D_ERR_ARG, "null argument");
    s3d_ctx* c = new s3d_ctx();
    int r = ctx_common_init(c, nx, ny, nz, p);
    if (r == S3D_OK) r = ctx_normalize(c, d_vol);
    if (r == S3D_OK && cudaStreamSynchronize(c->stream) != cudaSuccess)
        r = fail(S3D_ERR_CUDA, "create_device: %s", cudaGetErrorString(cudaGetLastError()));
    if (r != S3D_OK) { s3d_destroy(c); return r; }
    *out = c;
    return S3D_OK;
}

void s3d_destroy(s3d_handle c) {
    if (!c) return;
    if (c->stream) {
        cudaSetDevice(c->device);
        free_levels(c);
        void* ptrs[] = {c->d_input, c->d_slots, c->d_thres, c->d_mesh, c->d_extre, c->d_codes, c->d_xyz5, c->d_kps, c->d_desc, c->d_redo};
        for (void* q : ptrs) if (q) s3d::dev_free(q, c->stream);
        cudaStreamSynchronize(c->stream);
        c->prof.resolve();
        if (c->ev_ok) for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev[i]);
        if (c->own_stream) cudaStreamDestroy(c->stream);
    }
    delete c;
}

// Initialize (Src/cSIFT3D.cc:237-266) + Build_Gaussian_Scale_Space (:268-319) +
// Build_DOG_Scale_Space (:346-360; fused into the Z pass) + Detect_KeyPoints (:362-425) +
// Assign_Orientation (:427-482) + Extract_Description (:484-502).
// ---- the pipeline in stages (shared by the unsharded run and the z-slab shards) ----------------
//
// Initialize (Src/cSIFT3D.cc:237-266) -> per octave Build_Gaussian_Scale_Space (:268-319) with
// Build_DOG_Scale_Space (:346-360) fused into the Z pass -> Detect_KeyPoints (:362-425) ->
// Assign_Orientation (:427-482) -> Extract_Description (:484-502).

static int halo_for(const s3d_ctx* c) {
    // planes a shard needs beyond its owned range: the blur chain must leave every DoG level valid
    // on owned +-1 (detection neighbours), and the descriptor window reaches ceil(r/u)+1 planes
    // (r = 2*7.0711*scale, Src/cSIFT3D.cc:1155-1156, clamped windows :1182-1198, +1 for the gradient)
    int cum = 0;
    for (int i = 0; i < c->G; i++) cum += c->taps[i].hw;
    const int blur_need = 1 + cum + c->G;  // low side needs 1 + sum(hw); high side one more plane per level
    const float ratio = host_level_scale(0, c->L, c->L, c->prm.sigma_default);
    const int desc_need = (int)ceilf(2.0f * 7.071067812f * ratio) + 2;
    return std::max(blur_need, desc_need);
}

static int stage_init(s3d_ctx* c) {
    if (c->stage != 0) return fail(S3D_ERR_STATE, "already initialised (KpSiftAlgorithm is single-shot)");
    cudaStream_t st = c->stream;
    S3D_CUDA(cudaSetDevice(c->device));
    const int L = c->L, G = c->G, D = c->D;
    if (c->h2d_pending) {  // s3d_create_async: the copy's events are about to be reused
        S3D_CUDA(cudaEventSynchronize(c->ev[7]));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]);
        c->timers[8] = ms * 1e-3;
        c->h2d_pending = false;
    }
    S3D_CUDA(cudaEventRecord(c->ev[0], st));
    int mn = std::min(c->nx, std::min(c->ny, c->nz));
    c->noct = (int)log2f((float)mn) - 3 + 1;  // :254-255
    if (c->noct < 1 || c->noct > kMaxOct) return fail(S3D_ERR_ARG, "octave count %d out of range", c->noct);
    if (1 + c->noct * D > 255) return fail(S3D_ERR_ARG, "too many levels");
    {
        int nx = c->nx, ny = c->ny, nz = c->nz;
        for (int o = 0; o < c->noct; o++) {
            c->dims[o][0] = nx; c->dims[o][1] = ny; c->dims[o][2] = nz;
            c->nvox[o] = (size_t)nx * ny * nz;
            nx /= 2; ny /= 2; nz /= 2;  // Src/cUtil.cc:219-221
        }
    }
    host_sigmas(L, c->prm.sigma_default, c->prm.sigma_n_default, c->sig);
    for (int i = 0; i < G; i++)
        if (host_taps(c->sig[i], &c->taps[i]) < 0)
            return fail(S3D_ERR_ARG, "sigma %g needs more than %d taps per side", c->sig[i], kMaxHW);
    for (int o = 0; o < c->noct; o++) {
        const int nzo = c->dims[o][2];
        if (!c->slab) {
            c->p0[o] = c->za[o] = 0; c->p1[o] = c->zb[o] = nzo;
        } else {
            // plane k of octave o is owned iff own0 <= k * 2^o < own1 (== the owner of plane 2k one octave down)
            const int sh = 1 << o;
            c->p0[o] = std::min(nzo, (c->own0 + sh - 1) >> o);
            c->p1[o] = std::min(nzo, (c->own1 + sh - 1) >> o);
            if (o > 0) {  // nz/2 truncation can drop the last plane of an odd level
                c->p0[o] = std::min(c->p0[o], nzo);
                c->p1[o] = std::min(c->p1[o], nzo);
            }
            c->za[o] = std::max(0, c->p0[o] - c->halo);
            c->zb[o] = std::min(nzo, c->p1[o] + c->halo);
            if (c->p1[o] <= c->p0[o]) { c->za[o] = c->zb[o] = c->p0[o]; }  // owns nothing here
        }
    }
    c->gss.assign((size_t)c->noct * G, nullptr);
    c->dog.assign((size_t)c->noct * D, nullptr);
    size_t tmp_elems = 4;
    for (int o = 0; o < c->noct; o++) {
        const size_t bytes = std::max<size_t>(c->lvox(o), 4) * sizeof(float);
        tmp_elems = std::max(tmp_elems, c->lvox(o));
        for (int i = 0; i < G; i++) S3D_CUDA(s3d::dev_alloc((void**)&c->gss[o * G + i], bytes, st));
        for (int i = 0; i < D; i++) S3D_CUDA(s3d::dev_alloc((void**)&c->dog[o * D + i], bytes, st));
    }
    for (int i = 0; i < 2; i++) S3D_CUDA(s3d::dev_alloc((void**)&c->d_tmp[i], tmp_elems * sizeof(float), st));
    if (c->slab && getenv("S3D_SLAB_POISON")) {  // tests: a halo plane that was never filled must show up as NaN
        for (int o = 0; o < c->noct; o++) {
            for (int i = 0; i < G; i++) S3D_CUDA(cudaMemsetAsync(c->gss[o * G + i], 0xFF, c->lvox(o) * sizeof(float), st));
            for (int i = 0; i < D; i++) S3D_CUDA(cudaMemsetAsync(c->dog[o * D + i], 0xFF, c->lvox(o) * sizeof(float), st));
        }
    }
    c->levels_alive = true;
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_thres, sizeof(float) * c->noct * L, st));
    {
        MeshConst hm;
        host_mesh(&hm);
        S3D_CUDA(s3d::dev_alloc((void**)&c->d_mesh, sizeof(MeshConst), st));
        S3D_CUDA(cudaMemcpyAsync(c->d_mesh, &hm, sizeof(MeshConst), cudaMemcpyHostToDevice, st));
        S3D_CUDA(cudaStreamSynchronize(st));  // hm is a stack object
    }
    S3D_CUDA(cudaEventRecord(c->ev[1], st));
    c->stage = 1;
    c->next_octave = 0;
    return S3D_OK;
}

// Octave seed = even-index decimation of level L of the previous octave (:311), owned planes only
// (a shard's halo planes of the seed come from their owners, s3d_slab_level_buffer + exchange).
static int stage_seed(s3d_ctx* c, int o) {
    cudaStream_t st = c->stream;
    const int L = c->L, G = c->G;
    const int nown = c->p1[o] - c->p0[o];
    if (o < 1 || nown <= 0) return S3D_OK;
    const int nx = c->dims[o][0], ny = c->dims[o][1];
    const float* src = c->gss[(o - 1) * G + L] + (size_t)(2 * c->p0[o] - c->za[o - 1]) * c->plane(o - 1);
    float* dst = c->gss[o * G] + (size_t)(c->p0[o] - c->za[o]) * c->plane(o);
    const size_t nout = (size_t)nown * c->plane(o);
    ProfScope ps(&c->prof, K_DOWNSAMPLE, 8.0 * nout);
    S3D_LAUNCH(downsample_kernel, s3d_blocks(nout, 256), 256, 0, st, src, c->dims[o - 1][0], c->dims[o - 1][1], dst, nx, ny, nown);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// Levels of one octave over the local planes [za, zb): X -> Y -> Z (:609-617); the Z pass of level
// i >= 1 also emits DoG[i-1] = G[i-1] - G[i] and (unsharded) folds max|DoG| into its slot.  A shard
// takes the maxima over its OWNED planes only (halo planes hold partial results) with maxabs_kernel.
static int stage_octave(s3d_ctx* c, int o) {
    cudaStream_t st = c->stream;
    const int L = c->L, G = c->G, D = c->D;
    (void)L;
    const int nx = c->dims[o][0], ny = c->dims[o][1], nz = c->zb[o] - c->za[o];
    if (nz <= 0) return S3D_OK;
    unsigned* scratch_slot = c->d_slots + 255;  // sink for the fused max of a shard's Z passes
    for (int i = 0; i < G; i++) {
        if (i == 0 && o > 0) continue;  // seed already in place
        float* dst = c->gss[o * G + i];
        const float* src = (o == 0 && i == 0) ? c->d_input : c->gss[o * G + i - 1];
        const Taps& t = c->taps[i];
        if (!blur_xy_pass(src, c->d_tmp[1], nx, ny, nz, t, st, &c->prof)) {
            blur_pass(src, c->d_tmp[0], nx, ny, nz, 0, t, 1, nullptr, nullptr, nullptr, st, &c->prof);
            blur_pass(c->d_tmp[0], c->d_tmp[1], nx, ny, nz, 1, t, 1, nullptr, nullptr, nullptr, st, &c->prof);
        }
        if (i >= 1) {
            unsigned* slot = c->d_slots + 1 + o * D + i - 1;
            blur_pass(c->d_tmp[1], dst, nx, ny, nz, 2, t, 1, src, c->dog[o * D + i - 1], c->slab ? scratch_slot : slot, st,
                      &c->prof, c->dims[o][2], c->za[o]);
            if (c->slab && c->p1[o] > c->p0[o]) {
                const size_t n = (size_t)(c->p1[o] - c->p0[o]) * c->plane(o);
                const float* own = c->dog[o * D + i - 1] + (size_t)(c->p0[o] - c->za[o]) * c->plane(o);
                ProfScope ps(&c->prof, K_MAXABS, 4.0 * n);
                S3D_LAUNCH(maxabs_kernel, (unsigned)std::min<size_t>(s3d_blocks(n / 4 + 1, 256), 148 * 16), 256, 0, st, own, n, slot);
            }
        } else {
            blur_pass(c->d_tmp[1], dst, nx, ny, nz, 2, t, 1, nullptr, nullptr, nullptr, st, &c->prof, c->dims[o][2], c->za[o]);
        }
    }
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

static int stage_sparse(s3d_ctx* c) {
    cudaStream_t st = c->stream;
    const int L = c->L, G = c->G, D = c->D;
    S3D_CUDA(cudaEventRecord(c->ev[2], st));
    // level pointers as seen with GLOBAL voxel indices (virtual origin for a shard's local planes)
    auto vdog = [&](int o, int i) { return c->dog[o * D + i] - (ptrdiff_t)c->za[o] * (ptrdiff_t)c->plane(o); };
    auto vgss = [&](int o, int i) { return c->gss[o * G + i] - (ptrdiff_t)c->za[o] * (ptrdiff_t)c->plane(o); };

    // ---- Detection -------------------------------------------------------------------------------
    std::vector<uint32_t> gb_base((size_t)c->noct * L + 1, 0);
    size_t own_total = 0;
    for (int o = 0; o < c->noct; o++) {
        const size_t nown = (size_t)std::max(0, c->p1[o] - c->p0[o]) * c->plane(o);
        own_total += nown;
        for (int j = 0; j < L; j++) gb_base[o * L + j + 1] = gb_base[o * L + j] + s3d_blocks(nown, kDetectChunk);
    }
    const int nblk = std::max<int>(1, (int)gb_base[(size_t)c->noct * L]);
    const unsigned stage_cap = (unsigned)std::max<size_t>(65536, own_total / 28);
    int *d_blk_cnt = nullptr, *d_blk_off = nullptr, *d_total = nullptr;
    StageEntry* d_stage = nullptr;
    unsigned* d_stage_count = nullptr;
    Cand* d_cand = nullptr;
    S3D_CUDA(s3d::dev_alloc((void**)&d_blk_cnt, sizeof(int) * nblk, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_blk_off, sizeof(int) * nblk, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_total, sizeof(int) * 4, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_stage, sizeof(StageEntry) * (size_t)stage_cap, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_stage_count, sizeof(unsigned), st));
    S3D_CUDA(cudaMemsetAsync(d_stage_count, 0, sizeof(unsigned), st));
    S3D_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int) * 4, st));
    S3D_CUDA(cudaMemsetAsync(d_blk_cnt, 0, sizeof(int) * nblk, st));
    for (int o = 0; o < c->noct; o++) {
        const ll vb = (ll)c->p0[o] * (ll)c->plane(o), ve = (ll)c->p1[o] * (ll)c->plane(o);
        if (ve <= vb) continue;
        for (int j = 1; j <= L; j++) {
            const int unit = o * L + (j - 1);
            ProfScope ps(&c->prof, K_DETECT, 4.0 * (double)(ve - vb));
            S3D_LAUNCH(detect_kernel, s3d_blocks((size_t)(ve - vb), kDetectChunk), 256, 0, st, vdog(o, j - 1), vdog(o, j),
                       vdog(o, j + 1), c->dims[o][0], c->dims[o][1], c->dims[o][2], c->d_slots + 1 + o * D + j,
                       c->prm.peak_thresh, (uint32_t)unit, gb_base[unit], d_blk_cnt, d_stage, d_stage_count, stage_cap,
                       c->d_thres + unit, vb, ve);
        }
    }
    {
        ProfScope ps(&c->prof, K_COMPACT, 8.0 * nblk);
        S3D_LAUNCH(scan_kernel, 1, 1024, 0, st, d_blk_cnt, d_blk_off, nblk, d_total);
    }
    unsigned h_stage = 0;
    int h_total[4] = {0, 0, 0, 0};
    S3D_CUDA(cudaMemcpyAsync(&h_stage, d_stage_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaMemcpyAsync(h_total, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    if (h_stage > stage_cap)
        return fail(S3D_ERR_CAPACITY, "detection staged %u candidates, capacity %u", h_stage, stage_cap);
    c->n_extre = h_total[0];
    const int ne = c->n_extre;
    if (ne > 0) {
        S3D_CUDA(s3d::dev_alloc((void**)&d_cand, sizeof(Cand) * (size_t)ne, st));
        ProfScope ps(&c->prof, K_COMPACT, 24.0 * ne);
        S3D_LAUNCH(scatter_kernel, s3d_blocks(ne, 256), 256, 0, st, d_stage, d_stage_count, stage_cap, d_blk_off, d_cand);
    }
    S3D_CUDA(cudaEventRecord(c->ev[3], st));

    // ---- Orientation -----------------------------------------------------------------------------
    LevelTable tab;
    memset(&tab, 0, sizeof(tab));
    tab.L = L; tab.G = G;
    for (int o = 0; o < c->noct; o++) {
        for (int k = 0; k < 3; k++) tab.dims[o][k] = c->dims[o][k];
        for (int i = 0; i < G; i++) {
            tab.gss[o * G + i] = vgss(o, i);
            tab.scale[o * G + i] = host_level_scale(o, i, L, c->prm.sigma_default);
        }
    }
    // Gaussian window weight tables (wtab_kernel): one per (octave, keypoint level) and stage
    float* d_wtab = nullptr;
    {
        int off = 0;
        for (int o = 0; o < c->noct; o++)
            for (int i = 1; i <= L; i++) {
                const int lv = o * G + i;
                const float u = (float)(1 << o), scale = tab.scale[lv];
                const float so = 1.5f * scale, ro = so * 3.0f;                       // :27-28,:442,:915
                const float sd = scale * 7.071067812f, rd = 2.0f * sd;               // :30-31,:1155-1156
                tab.ori_off[lv] = off; tab.ori_n[lv] = (int)(ro * ro / (u * u)) + 2; off += tab.ori_n[lv];
                tab.desc_off[lv] = off; tab.desc_n[lv] = (int)(rd * rd / (u * u)) + 2; off += tab.desc_n[lv];
            }
        S3D_CUDA(s3d::dev_alloc((void**)&d_wtab, sizeof(float) * std::max(off, 1), st));
        tab.wtab = d_wtab;
        if (ne > 0) S3D_LAUNCH(wtab_kernel, dim3(2, c->noct * G), 256, 0, st, tab, c->noct, d_wtab);
    }
    int *d_recheck = nullptr;
    int* d_surv = nullptr;
    int* d_order = nullptr;
    const size_t nea = std::max(ne, 1);
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_extre, sizeof(s3d_keypoint) * nea, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_codes, sizeof(int) * nea, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_xyz5, sizeof(int) * 5 * nea, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_recheck, sizeof(int) * nea, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_surv, sizeof(int) * nea, st));
    if (ne > 0) {
        const unsigned grid = (unsigned)std::min<size_t>(s3d_blocks((size_t)ne * 32, 256), 148 * 32);
        {
            ProfScope ps(&c->prof, K_ORIENT, 200.0 * ne);
            S3D_LAUNCH(orient_kernel, grid, 256, 0, st, d_cand, ne, tab, c->d_extre, c->d_codes, c->d_xyz5,
                       c->prm.max_eig_thres, c->prm.corner_thresh, 2e-3f, c->prm.exact_recheck ? d_recheck : (int*)nullptr,
                       d_total + 1);
        }
        ProfScope ps(&c->prof, K_ORIENT_EXACT, 0.0);
        if (c->prm.exact_recheck)
            S3D_LAUNCH(orient_exact_kernel, (unsigned)std::min<size_t>((size_t)ne, 148 * 6),
                       kExactWarps * 32, 0, st, d_cand, tab, c->d_extre, c->d_codes, d_recheck, d_total + 1,
                       c->prm.max_eig_thres, c->prm.corner_thresh, d_total + 2);
    }
    {
        ProfScope ps(&c->prof, K_SURVIVORS, 8.0 * ne);
        S3D_LAUNCH(survivors_kernel, 1, 1024, 0, st, c->d_codes, ne, d_surv, d_total + 3);
        // heavy-first launch order of the descriptor CTAs (S3D_DESC_ORDER=0: list order)
        if (desc_order_enabled() && ne > 0) {
            S3D_CUDA(s3d::dev_alloc((void**)&d_order, sizeof(int) * nea, st));
            S3D_LAUNCH(desc_order_kernel, 1, 1024, 0, st, c->d_extre, d_surv, d_total + 3, G - 1, d_order);
        }
    }
    S3D_CUDA(cudaMemcpyAsync(h_total, d_total, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    c->n_rechecked = h_total[1];
    c->n_flipped = h_total[2];
    c->n_kps = h_total[3];
    S3D_CUDA(cudaEventRecord(c->ev[4], st));

    // ---- Description -----------------------------------------------------------------------------
    const size_t nka = std::max(c->n_kps, 1);
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_kps, sizeof(s3d_keypoint) * nka, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_desc, sizeof(float) * S3D_DESC_LEN * nka, st));
    int* d_redo = nullptr;
    if (c->n_kps > 0) {
        static bool attr_set[64] = {false};
        if (c->device < 64 && !attr_set[c->device]) {
            S3D_CUDA(cudaFuncSetAttribute(describe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescSmem)));
            S3D_CUDA(cudaFuncSetAttribute(describe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescSmemQ)));
            attr_set[c->device] = true;
        }
        const int path = g_describe_path.load();
        if (path == 1) {  // FP32 staged/ordered accumulation for every keypoint
            ProfScope ps(&c->prof, K_DESCRIBE, (176.0 + 3072.0) * c->n_kps);
            S3D_LAUNCH(describe_kernel<false>, c->n_kps, kDescWarps * 32, sizeof(DescSmem), st, c->d_extre, d_surv, c->n_kps, tab,
                       c->d_mesh, c->d_kps, c->d_desc, (const int*)nullptr, (const int*)nullptr, (int*)nullptr, (int*)nullptr, 0.0f);
        } else {
            // fixed-point atomics for all; the (normally empty) list of keypoints whose scale estimate was
            // too low is redone in FP32 — its grid is sized for the worst case and reads the count on the device
            S3D_CUDA(s3d::dev_alloc((void**)&d_redo, sizeof(int) * ((size_t)c->n_kps + 1), st));
            S3D_CUDA(cudaMemsetAsync(d_redo, 0, sizeof(int), st));
            {
                ProfScope ps(&c->prof, K_DESCRIBE, (176.0 + 3072.0) * c->n_kps);
                S3D_LAUNCH(describe_kernel<true>, c->n_kps, kDescWarps * 32, sizeof(DescSmemQ), st, c->d_extre, d_surv, c->n_kps, tab,
                           c->d_mesh, c->d_kps, c->d_desc, (const int*)d_order, (const int*)nullptr, d_redo + 1, d_redo,
                           path == 2 ? 0.02f : kQMargin);
            }
            ProfScope ps(&c->prof, K_DESCRIBE_REDO, 0.0);
            S3D_LAUNCH(describe_kernel<false>, c->n_kps, kDescWarps * 32, sizeof(DescSmem), st, c->d_extre, d_surv, c->n_kps, tab,
                       c->d_mesh, c->d_kps, c->d_desc, (const int*)(d_redo + 1), (const int*)d_redo, (int*)nullptr, (int*)nullptr, 0.0f);
            c->d_redo = d_redo;  // the count is read in s3d_wait
            d_redo = nullptr;
        }
    }
    S3D_CUDA(cudaGetLastError());
    S3D_CUDA(cudaEventRecord(c->ev[5], st));

    // ---- Release_SIFT (:1659-1678) unless the caller asked to keep the pyramids ---------------
    if (!c->prm.keep_levels) free_levels(c);
    void* tmp[] = {d_blk_cnt, d_blk_off, d_total, d_stage, d_stage_count, d_cand, d_recheck, d_surv, d_order, d_wtab, d_redo};
    for (void* q : tmp) if (q) s3d::dev_free(q, st);
    S3D_CUDA(cudaEventRecord(c->ev[6], st));
    c->queued = true;
    c->stage = 100;
    return S3D_OK;
}

static int run_impl(s3d_ctx* c) {
    if (c->ran || c->stage != 0) return fail(S3D_ERR_STATE, "s3d_run called twice on one handle (KpSiftAlgorithm is single-shot)");
    if (c->slab) return fail(S3D_ERR_STATE, "a z-slab shard is driven stage by stage (s3d_slab_*), not by s3d_run");
    S3D_TRY(stage_init(c));
    for (int o = 0; o < c->noct; o++) {
        S3D_TRY(stage_seed(c, o));
        S3D_TRY(stage_octave(c, o));
    }
    return stage_sparse(c);
}

int s3d_run_async(s3d_handle c) {
    clear_error();
    if (!c) return fail(S3D_ERR_ARG, "null handle");
    return run_impl(c);
}

int s3d_wait(s3d_handle c) {
    clear_error();
    if (!c) return fail(S3D_ERR_ARG, "null handle");
    if (!c->queued) return fail(S3D_ERR_STATE, "s3d_wait before s3d_run_async");
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_redo) {
        S3D_CUDA(cudaMemcpy(&c->n_desc_redo, c->d_redo, sizeof(int), cudaMemcpyDeviceToHost));
        s3d::dev_free(c->d_redo, c->stream);
        c->d_redo = nullptr;
    }
    if (!c->ran) {
        float ms;
        for (int i = 0; i < 6; i++) {
            cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]);
            // alloc, gss(+dog fused), detect, orient, describe, release
            static const int slot[6] = {0, 1, 3, 4, 5, 6};
            c->timers[slot[i]] = ms * 1e-3;
        }
        c->timers[2] = 0.0;  // DoG is fused into the Z pass
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[6]);
        c->timers[7] = ms * 1e-3;
        c->prof.resolve();
        c->ran = true;
    }
    return S3D_OK;
}

int s3d_run(s3d_handle c) {
    int r = s3d_run_async(c);
    if (r != S3D_OK) return r;
    return s3d_wait(c);
}

// ---- z-slab shards ---------------------------------------------------------
// One handle per shard (one process per GPU, or several logical shards on one device).  The caller
// drives the stages in lockstep over all shards and moves planes between the shards' level buffers
// (s3d_slab_level_buffer) and scalars (maxima) between the stages: NCCL send/recv + all-reduce in
// 3dsift_b200/dist.py.  Every stage only enqueues work on the handle's stream unless noted.

static int slab_halo_from_params(const s3d_params* p, int* halo) {
    s3d_ctx tmp;
    s3d_params def;
    s3d_default_params(&def);
    tmp.prm = p ? *p : def;
    tmp.L = tmp.prm.num_kp_levels;
    if (tmp.L < 1 || tmp.L + 3 > kMaxG) return fail(S3D_ERR_ARG, "num_kp_levels %d unsupported", tmp.L);
    tmp.G = tmp.L + 3;
    host_sigmas(tmp.L, tmp.prm.sigma_default, tmp.prm.sigma_n_default, tmp.sig);
    for (int i = 0; i < tmp.G; i++)
        if (host_taps(tmp.sig[i], &tmp.taps[i]) < 0) return fail(S3D_ERR_ARG, "sigma %g too wide", tmp.sig[i]);
    *halo = halo_for(&tmp);
    return S3D_OK;
}

int s3d_slab_extent(int nz, int own0, int own1, const s3d_params* p, int octave, int* out4) {
    clear_error();
    if (!out4 || nz < 1 || own0 < 0 || own1 < own0 || own1 > nz || octave < 0 || octave >= kMaxOct) return fail(S3D_ERR_ARG, "bad argument");
    int halo = 0;
    S3D_TRY(slab_halo_from_params(p, &halo));
    int nzo = nz;
    for (int o = 0; o < octave; o++) nzo /= 2;
    const int sh = 1 << octave;
    int p0 = std::min(nzo, (own0 + sh - 1) >> octave), p1 = std::min(nzo, (own1 + sh - 1) >> octave);
    int za = std::max(0, p0 - halo), zb = std::min(nzo, p1 + halo);
    if (p1 <= p0) za = zb = p0;
    out4[0] = za; out4[1] = zb; out4[2] = p0; out4[3] = p1;
    return S3D_OK;
}

int s3d_slab_create(const float* vol_ext, int on_device, int nx, int ny, int nz, int own0, int own1, const s3d_params* p,
                    s3d_handle* out) {
    clear_error();
    if (!vol_ext || !out || own0 < 0 || own1 <= own0 || own1 > nz) return fail(S3D_ERR_ARG, "bad slab arguments");
    s3d_ctx* c = new s3d_ctx();
    int r = ctx_common_init(c, nx, ny, nz, p);
    if (r == S3D_OK) r = slab_halo_from_params(&c->prm, &c->halo);
    if (r != S3D_OK) { s3d_destroy(c); return r; }
    c->slab = true;
    c->own0 = own0; c->own1 = own1;
    const int za = std::max(0, own0 - c->halo), zb = std::min(nz, own1 + c->halo);
    const size_t plane = (size_t)nx * ny, nloc = plane * (size_t)(zb - za);
    auto body = [&]() -> int {
        // d_input first holds the raw local planes; s3d_slab_begin normalises them in place
        S3D_CUDA(s3d::dev_alloc((void**)&c->d_input, nloc * sizeof(float), c->stream));
        S3D_CUDA(s3d::dev_alloc((void**)&c->d_slots, 256 * sizeof(unsigned), c->stream));
        S3D_CUDA(cudaMemsetAsync(c->d_slots, 0, 256 * sizeof(unsigned), c->stream));
        S3D_CUDA(cudaMemcpyAsync(c->d_input, vol_ext, nloc * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                 c->stream));
        const size_t nown = plane * (size_t)(own1 - own0);
        ProfScope ps(&c->prof, K_MAXABS, 4.0 * nown);
        S3D_LAUNCH(maxabs_kernel, (unsigned)std::min<size_t>(s3d_blocks(nown / 4 + 1, 256), 148 * 16), 256, 0, c->stream,
                   c->d_input + plane * (size_t)(own0 - za), nown, c->d_slots);
        S3D_CUDA(cudaGetLastError());
        S3D_CUDA(cudaStreamSynchronize(c->stream));  // `vol_ext` may be released by the caller
        return S3D_OK;
    };
    r = body();
    if (r != S3D_OK) { s3d_destroy(c); return r; }
    *out = c;
    return S3D_OK;
}

int s3d_slab_local_max(s3d_handle c, float* mx) {
    clear_error();
    if (!c || !mx || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaMemcpyAsync(mx, c->d_slots, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_slab_begin(s3d_handle c, float global_max) {
    clear_error();
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage != 0) return fail(S3D_ERR_STATE, "s3d_slab_begin called twice");
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaMemcpyAsync(c->d_slots, &global_max, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));  // global_max is a stack value
    const int za = std::max(0, c->own0 - c->halo), zb = std::min(c->nz, c->own1 + c->halo);
    const size_t nloc = (size_t)c->nx * c->ny * (size_t)(zb - za);
    {
        ProfScope ps(&c->prof, K_NORMALIZE, 8.0 * nloc);
        S3D_LAUNCH(normalize_kernel, (unsigned)std::min<size_t>(s3d_blocks(nloc / 4 + 1, 256), 148 * 16), 256, 0, c->stream,
                   c->d_input, c->d_input, nloc, c->d_slots);
    }
    S3D_TRY(stage_init(c));
    return stage_octave(c, 0);
}

int s3d_slab_info(s3d_handle c, int* noct, int* halo, int* levels_per_octave) {
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage < 1) return fail(S3D_ERR_STATE, "s3d_slab_begin first");
    if (noct) *noct = c->noct;
    if (halo) *halo = c->halo;
    if (levels_per_octave) *levels_per_octave = c->G;
    return S3D_OK;
}

int s3d_slab_seed(s3d_handle c, int octave) {
    clear_error();
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage < 1 || octave < 1 || octave >= c->noct) return fail(S3D_ERR_STATE, "bad octave / order");
    S3D_CUDA(cudaSetDevice(c->device));
    return stage_seed(c, octave);
}

int s3d_slab_octave(s3d_handle c, int octave) {
    clear_error();
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage < 1 || octave < 1 || octave >= c->noct) return fail(S3D_ERR_STATE, "bad octave / order");
    S3D_CUDA(cudaSetDevice(c->device));
    return stage_octave(c, octave);
}

int s3d_slab_level_buffer(s3d_handle c, int which, int idx, float** d_ptr, int* ext4) {
    if (!c || !c->slab || !d_ptr || !ext4) return fail(S3D_ERR_ARG, "bad argument");
    if (c->stage < 1 || !c->levels_alive) return fail(S3D_ERR_STATE, "levels not allocated");
    const int per = which == 0 ? c->G : c->D;
    if (which < 0 || which > 1 || idx < 0 || idx >= c->noct * per) return fail(S3D_ERR_ARG, "level index out of range");
    const int o = idx / per;
    *d_ptr = which == 0 ? c->gss[idx] : c->dog[idx];
    ext4[0] = c->za[o]; ext4[1] = c->zb[o]; ext4[2] = c->p0[o]; ext4[3] = c->p1[o];
    return S3D_OK;
}

int s3d_slab_get_maxima(s3d_handle c, float* out, int n) {
    clear_error();
    if (!c || !c->slab || !out) return fail(S3D_ERR_ARG, "bad argument");
    if (c->stage < 1) return fail(S3D_ERR_STATE, "s3d_slab_begin first");
    n = std::min(n, c->noct * c->D);
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaMemcpyAsync(out, c->d_slots + 1, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_slab_set_maxima(s3d_handle c, const float* in, int n) {
    clear_error();
    if (!c || !c->slab || !in) return fail(S3D_ERR_ARG, "bad argument");
    if (c->stage < 1) return fail(S3D_ERR_STATE, "s3d_slab_begin first");
    n = std::min(n, c->noct * c->D);
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaMemcpyAsync(c->d_slots + 1, in, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_slab_finish(s3d_handle c) {
    clear_error();
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage != 1) return fail(S3D_ERR_STATE, "s3d_slab_finish out of order");
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_TRY(stage_sparse(c));
    return s3d_wait(c);
}

int s3d_num_octaves(s3d_handle c, int* n) {
    if (!c || !n) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    *n = c->noct;
    return S3D_OK;
}

int s3d_level_dims(s3d_handle c, int o, int* d) {
    if (!c || !d) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    if (o < 0 || o >= c->noct) return fail(S3D_ERR_ARG, "octave %d out of range", o);
    d[0] = c->dims[o][0]; d[1] = c->dims[o][1]; d[2] = c->dims[o][2];
    return S3D_OK;
}

int s3d_num_keypoints(s3d_handle c, int* n) {
    if (!c || !n) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    *n = c->n_kps;
    return S3D_OK;
}

int s3d_get_keypoints_async(s3d_handle c, s3d_keypoint* kp, float* desc) {
    clear_error();
    if (!c) return fail(S3D_ERR_ARG, "null handle");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaEventRecord(c->ev[6], c->stream));
    if (c->n_kps > 0) {
        if (kp) S3D_CUDA(cudaMemcpyAsync(kp, c->d_kps, sizeof(s3d_keypoint) * c->n_kps, cudaMemcpyDeviceToHost, c->stream));
        if (desc)
            S3D_CUDA(cudaMemcpyAsync(desc, c->d_desc, sizeof(float) * S3D_DESC_LEN * (size_t)c->n_kps,
                                     cudaMemcpyDeviceToHost, c->stream));
    }
    S3D_CUDA(cudaEventRecord(c->ev[7], c->stream));
    c->d2h_pending = true;
    return S3D_OK;
}

int s3d_sync(s3d_handle c) {
    clear_error();
    if (!c) return fail(S3D_ERR_ARG, "null handle");
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d2h_pending) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]);
        c->timers[9] = ms * 1e-3;
        c->d2h_pending = false;
    }
    return S3D_OK;
}

int s3d_trim_cache(int device, unsigned long long* cached_bytes) {
    clear_error();
    int cur = 0;
    S3D_CUDA(cudaGetDevice(&cur));
    const int dev = device < 0 ? cur : device;
    if (cached_bytes) *cached_bytes = (unsigned long long)s3d::DevCache::get().cached_bytes(dev);
    if (dev != cur) S3D_CUDA(cudaSetDevice(dev));
    s3d::DevCache::get().trim(dev);
    if (dev != cur) S3D_CUDA(cudaSetDevice(cur));
    return S3D_OK;
}

int s3d_get_keypoints(s3d_handle c, s3d_keypoint* kp, float* desc) {
    const int r = s3d_get_keypoints_async(c, kp, desc);
    return r != S3D_OK ? r : s3d_sync(c);
}

int s3d_num_extrema(s3d_handle c, int* n) {
    if (!c || !n) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    *n = c->n_extre;
    return S3D_OK;
}

int s3d_get_extrema(s3d_handle c, s3d_keypoint* kp, int* codes, int* xyz5) {
    clear_error();
    if (!c) return fail(S3D_ERR_ARG, "null handle");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    S3D_CUDA(cudaSetDevice(c->device));
    if (c->n_extre > 0) {
        if (kp) S3D_CUDA(cudaMemcpyAsync(kp, c->d_extre, sizeof(s3d_keypoint) * c->n_extre, cudaMemcpyDeviceToHost, c->stream));
        if (codes) S3D_CUDA(cudaMemcpyAsync(codes, c->d_codes, sizeof(int) * c->n_extre, cudaMemcpyDeviceToHost, c->stream));
        if (xyz5) S3D_CUDA(cudaMemcpyAsync(xyz5, c->d_xyz5, sizeof(int) * 5 * c->n_extre, cudaMemcpyDeviceToHost, c->stream));
    }
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_get_level(s3d_handle c, int which, int idx, float* out) {
    clear_error();
    if (!c || !out) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    if (!c->levels_alive) return fail(S3D_ERR_STATE, "pyramids were released; create the handle with keep_levels=1");
    const int per = which == 0 ? c->G : c->D;
    if (which < 0 || which > 1 || idx < 0 || idx >= c->noct * per) return fail(S3D_ERR_ARG, "level index out of range");
    const float* src = which == 0 ? c->gss[idx] : c->dog[idx];
    S3D_CUDA(cudaSetDevice(c->device));
    // (a z-slab shard returns its local planes [za, zb) of the level, see s3d_slab_extent)
    S3D_CUDA(cudaMemcpyAsync(out, src, c->lvox(idx / per) * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_level_info(s3d_handle c, int which, int idx, int* dims3, float* meta4) {
    if (!c || !dims3 || !meta4) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    const int per = which == 0 ? c->G : c->D;
    if (which < 0 || which > 1 || idx < 0 || idx >= c->noct * per) return fail(S3D_ERR_ARG, "level index out of range");
    const int o = idx / per, s = idx % per;
    for (int k = 0; k < 3; k++) dims3[k] = c->dims[o][k];
    meta4[0] = host_level_scale(o, s, c->L, c->prm.sigma_default);
    meta4[1] = meta4[2] = meta4[3] = (float)(1 << o);
    return S3D_OK;
}

int s3d_device_descriptors(s3d_handle c, const float** d_desc, int* n) {
    if (!c || !d_desc || !n) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    *d_desc = c->d_desc;
    *n = c->n_kps;
    return S3D_OK;
}

int s3d_device_results(s3d_handle c, const void** d_ptrs5, int* n_kps, int* n_extre) {
    if (!c || !d_ptrs5 || !n_kps || !n_extre) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    d_ptrs5[0] = c->d_kps; d_ptrs5[1] = c->d_desc; d_ptrs5[2] = c->d_extre; d_ptrs5[3] = c->d_codes; d_ptrs5[4] = c->d_xyz5;
    *n_kps = c->n_kps;
    *n_extre = c->n_extre;
    return S3D_OK;
}

int s3d_get_input(s3d_handle c, float* out) {
    clear_error();
    if (!c || !out) return fail(S3D_ERR_ARG, "null argument");
    S3D_CUDA(cudaSetDevice(c->device));
    const size_t nin = c->slab ? (size_t)c->nx * c->ny * (size_t)(std::min(c->nz, c->own1 + c->halo) - std::max(0, c->own0 - c->halo)) : c->n0;
    S3D_CUDA(cudaMemcpyAsync(out, c->d_input, nin * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_get_thresholds(s3d_handle c, float* out, int n) {
    clear_error();
    if (!c || !out) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "not run yet");
    n = std::min(n, c->noct * c->L);
    S3D_CUDA(cudaSetDevice(c->device));
    S3D_CUDA(cudaMemcpyAsync(out, c->d_thres, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    S3D_CUDA(cudaStreamSynchronize(c->stream));
    return S3D_OK;
}

int s3d_get_kernel_stats(s3d_handle c, int cap, int* n_classes, double* ms, long long* launches, double* alg_bytes) {
    if (!c || !n_classes || !ms || !launches || !alg_bytes) return fail(S3D_ERR_ARG, "null argument");
    const int n = std::min(cap, (int)K_NCLS);
    for (int i = 0; i < n; i++) { ms[i] = c->prof.ms[i]; launches[i] = c->prof.cnt[i]; alg_bytes[i] = c->prof.bytes[i]; }
    *n_classes = n;
    return S3D_OK;
}

const char* s3d_kernel_class_name(int cls) { return cls >= 0 && cls < K_NCLS ? kClsName[cls] : ""; }

int s3d_set_describe_path(int path) {
    if (path < 0 || path > 2) return fail(S3D_ERR_ARG, "describe path %d (0 fixed-point + FP32 redo, 1 FP32, 2 forced redo)", path);
    g_describe_path = path;
    return S3D_OK;
}

int s3d_get_counters(s3d_handle c, int* out4) {
    clear_error();
    if (!c || !out4) return fail(S3D_ERR_ARG, "null argument");
    if (!c->ran) return fail(S3D_ERR_STATE, "s3d_get_counters before s3d_run");
    out4[0] = c->n_rechecked; out4[1] = c->n_flipped; out4[2] = c->n_desc_redo; out4[3] = 0;
    return S3D_OK;
}

int s3d_get_timers(s3d_handle c, double* t) {
    if (!c || !t) return fail(S3D_ERR_ARG, "null argument");
    for (int i = 0; i < 10; i++) t[i] = c->timers[i];
    return S3D_OK;
}

// ---- free kernels ------------------------------------------------------------------------------

static int blur_host_buffers(const float* src, int nx, int ny, int nz, const Taps* taps3, const int* axes, int npass,
                             int variant, float* dst) {
    int dev;
    S3D_TRY(use_device(-1, &dev));
    const size_t n = (size_t)nx * ny * nz;
    float *a = nullptr, *b = nullptr;
    S3D_CUDA(cudaMalloc((void**)&a, std::max<size_t>(n, 4) * sizeof(float)));
    S3D_CUDA(cudaMalloc((void**)&b, std::max<size_t>(n, 4) * sizeof(float)));
    S3D_CUDA(cudaMemcpy(a, src, n * sizeof(float), cudaMemcpyHostToDevice));
    for (int p = 0; p < npass; p++) {
        blur_pass(a, b, nx, ny, nz, axes[p], taps3[p], variant, nullptr, nullptr, nullptr, 0);
        std::swap(a, b);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(dst, a, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(a);
    cudaFree(b);
    if (e != cudaSuccess) return fail(S3D_ERR_CUDA, "blur: %s", cudaGetErrorString(e));
    return S3D_OK;
}

int s3d_gaussian_smooth(const float* src, int nx, int ny, int nz, float sigma, float* dst) {
    clear_error();
    if (!src || !dst || nx < 1 || ny < 1 || nz < 1) return fail(S3D_ERR_ARG, "bad argument");
    Taps t[3];
    if (host_taps(sigma, &t[0]) < 0) return fail(S3D_ERR_ARG, "sigma %g needs more than %d taps per side", sigma, kMaxHW);
    t[1] = t[0]; t[2] = t[0];
    const int axes[3] = {0, 1, 2};
    return blur_host_buffers(src, nx, ny, nz, t, axes, 3, 1, dst);
}

int s3d_blur_axis(const float* src, int nx, int ny, int nz, int axis, const float* w, int hw, int variant, float* dst) {
    clear_error();
    if (!src || !dst || !w || nx < 1 || ny < 1 || nz < 1 || axis < 0 || axis > 2) return fail(S3D_ERR_ARG, "bad argument");
    if (hw < 1 || hw > kMaxHW) return fail(S3D_ERR_ARG, "hw %d out of range (1..%d)", hw, kMaxHW);
    Taps t;
    memset(&t, 0, sizeof(t));
    t.hw = hw;
    for (int i = 0; i < 2 * hw + 1; i++) t.w[i] = w[i];
    return blur_host_buffers(src, nx, ny, nz, &t, &axis, 1, variant, dst);
}

int s3d_downsample(const float* src, int nx, int ny, int nz, float* dst) {
    clear_error();
    if (!src || !dst || nx < 2 || ny < 2 || nz < 2) return fail(S3D_ERR_ARG, "bad argument");
    int dev;
    S3D_TRY(use_device(-1, &dev));
    const int dx = nx / 2, dy = ny / 2, dz = nz / 2;
    const size_t n = (size_t)nx * ny * nz, m = (size_t)dx * dy * dz;
    float *a = nullptr, *b = nullptr;
    S3D_CUDA(cudaMalloc((void**)&a, n * sizeof(float)));
    S3D_CUDA(cudaMalloc((void**)&b, m * sizeof(float)));
    S3D_CUDA(cudaMemcpy(a, src, n * sizeof(float), cudaMemcpyHostToDevice));
    S3D_LAUNCH(downsample_kernel, s3d_blocks(m, 256), 256, 0, 0, a, nx, ny, b, dx, dy, dz);
    cudaError_t e = cudaMemcpy(dst, b, m * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(a);
    cudaFree(b);
    if (e != cudaSuccess) return fail(S3D_ERR_CUDA, "downsample: %s", cudaGetErrorString(e));
    return S3D_OK;
}

}  // extern "C"
