// s3d_extract.cu — host orchestration + C ABI of the extraction path.
// Compiled with -fmad=false (see s3d_kernels.cuh).  Mirrors, stage by stage,
// CSIFT3D::KpSiftAlgorithm (/root/reference/3DSIFT/Src/cSIFT3D.cc:165-235).
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "s3d_common.h"
#include "s3d_ctx.h"
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
// frac is "whatever FP32 gives" for c' (App. A.2), so it depends on the magnitude of n: a z-slab shard always
// passes the GLOBAL line length (its kernels address planes through a virtual origin, see blur_z).
static Taps with_ext(const Taps& t0, int n) {
    Taps t = t0;
    for (int e = 0; e <= kMaxHW; e++) {
        volatile float c = (float)(2 * (n - 1));
        c = c - (float)(n - 1 + e);
        c = c - 0.1f;
        int il = (int)c;
        volatile float fr = c - (float)il;
        t.ext_il[e] = il;
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

// March along an axis of (global) length n over the output positions [zlo, zhi); src/dst/prev/dog address position 0
// of every line (a z-slab shard passes its local buffers offset to that virtual origin).
template <int HW>
static void launch_march(const float* src, float* dst, int nx, int n, ll st_m, int n_other, ll st_other, const Taps& t,
                         const float* prev, float* dog, unsigned* slot, cudaStream_t st, int zlo, int zhi) {
    const int nx4 = nx >> 2;
    const int seg = pick_seg(nx4, n_other, zhi - zlo);
    const int nseg = (zhi - zlo + seg - 1) / seg;
    ll threads = (ll)nx4 * n_other * nseg;
    // S3D_ZVAR: 0 = register-ring prefetch (blur_march_kernel), 1 = cp.async ring (blur_marchc_kernel, default)
    static const int zvar = [] { const char* e = getenv("S3D_ZVAR"); return (e && e[0] >= '0' && e[0] <= '1' && !e[1]) ? e[0] - '0' : 1; }();
    const unsigned blocks = s3d_blocks((size_t)threads, 128);
    if (dog) {
        if (zvar == 0) {
            auto kfn = blur_march_kernel<HW, true>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t, prev, dog, slot, zlo, zhi);
        } else {
            auto kfn = blur_marchc_kernel<HW, true>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t, prev, dog, slot, zlo, zhi);
        }
    } else {
        if (zvar == 0) {
            auto kfn = blur_march_kernel<HW, false>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t,
                       (const float*)nullptr, (float*)nullptr, (unsigned*)nullptr, zlo, zhi);
        } else {
            auto kfn = blur_marchc_kernel<HW, false>;
            S3D_LAUNCH(kfn, blocks, 128, 0, st, src, dst, nx, n, st_m, n_other, st_other, seg, t,
                       (const float*)nullptr, (float*)nullptr, (unsigned*)nullptr, zlo, zhi);
        }
    }
}

// ---- per-kernel-class device timing (params.profile): s3d_ctx.h -------------------------------
static const char* kClsName[K_NCLS] = {"maxabs", "normalize", "blur_x", "blur_y", "blur_xy", "blur_z_dog", "blur_generic", "downsample",
                                       "detect", "compact", "orient", "orient_exact", "survivors", "describe", "describe_redo", "blur_xyz"};
EventPool g_event_pool;
static_assert(sizeof(TapsH) == sizeof(Taps) && kCtxMaxOct == kMaxOct && kCtxMaxG == kMaxG && kCtxMaxHW == kMaxHW, "s3d_ctx.h mirrors");
static inline const Taps& taps_of(const s3d_ctx* c, int i) { return reinterpret_cast<const Taps&>(c->taps[i]); }

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

// One separable pass over a whole buffer of nx x ny x nz voxels.  variant 0 = generic kernel; 1 = fast kernels when
// eligible.  prev/dog/slot non-null fuses the DoG subtraction + max|DoG| (only meaningful on the last pass).
static void blur_pass(const float* src, float* dst, int nx, int ny, int nz, int axis, const Taps& t0, int variant,
                      const float* prev, float* dog, unsigned* slot, cudaStream_t st, Prof* prof = nullptr) {
    const ll total = (ll)nx * ny * nz;
    const int n = axis == 0 ? nx : (axis == 1 ? ny : nz);
    const Taps t = with_ext(t0, n);
    const bool fast = variant == 1 && (nx % 4 == 0) && supported_fast_hw(t.hw) && n >= 2 * t.hw + 2 && total >= 4096 &&
                      !(axis == 0 && dog);
    // algorithmic bytes: read src + write dst (+ read prev + write dog on the fused pass)
    ProfScope ps(prof, !fast ? K_BLUR_GENERIC : (axis == 0 ? K_BLUR_X : (axis == 1 ? K_BLUR_Y : K_BLUR_Z_DOG)),
                 (dog ? 16.0 : 8.0) * (double)total);
    if (!fast) {
        S3D_LAUNCH(blur_generic_kernel, s3d_blocks((size_t)total, 256), 256, 0, st, src, dst, nx, ny, nz, axis, t, prev,
                   dog, slot, 0, 0, 0, 0);
        return;
    }
    if (axis == 0) {
#define CALLX(H) launch_x<H>(src, dst, nx, (ll)ny * nz, t, total, st)
        S3D_HW_SWITCH(t.hw, CALLX)
#undef CALLX
    } else if (axis == 1) {
#define CALLY(H) launch_march<H>(src, dst, nx, ny, (ll)nx, nz, (ll)nx * ny, t, prev, dog, slot, st, 0, ny)
        S3D_HW_SWITCH(t.hw, CALLY)
#undef CALLY
    } else {
#define CALLZ(H) launch_march<H>(src, dst, nx, nz, (ll)nx * ny, ny, (ll)nx, t, prev, dog, slot, st, 0, nz)
        S3D_HW_SWITCH(t.hw, CALLZ)
#undef CALLZ
    }
}

// Z pass of a level whose LOCAL buffers (src = the X/Y-blurred planes, dst, prev, dog) hold global planes
// [za, za + nzl) of a line of nz_glob planes, for the output planes [zlo, zhi).  The unsharded run has za = 0,
// nzl = nz_glob, [zlo, zhi) = [0, nz_glob) and takes exactly the launches of blur_pass(axis 2).  A z-slab shard
// needs src valid on [zlo - hw, zhi + hw) (clipped to the line; the mirror / blend of the reference's boundary rule
// only reaches planes next to the line's own ends, which the shard at that end holds): the fast kernels address
// the buffers through a virtual origin (global plane 0), the generic kernel through (gn, goff) with clamped fetches.
static void blur_z(const float* src, float* dst, const float* prev, float* dog, unsigned* slot, int nx, int ny, int nz_glob,
                   int za, int nzl, int zlo, int zhi, const Taps& t0, cudaStream_t st, Prof* prof) {
    if (zhi <= zlo) return;
    if (za == 0 && nzl == nz_glob && zlo == 0 && zhi == nz_glob) {
        blur_pass(src, dst, nx, ny, nz_glob, 2, t0, 1, prev, dog, slot, st, prof);
        return;
    }
    const size_t plane = (size_t)nx * ny;
    const ll total = (ll)plane * (zhi - zlo);
    const Taps t = with_ext(t0, nz_glob);
    const bool fast = (nx % 4 == 0) && supported_fast_hw(t.hw) && nz_glob >= 2 * t.hw + 2 && (ll)plane * nz_glob >= 4096;
    ProfScope ps(prof, fast ? K_BLUR_Z_DOG : K_BLUR_GENERIC, (dog ? 16.0 : 8.0) * (double)total);
    if (fast) {
        const ptrdiff_t off = (ptrdiff_t)za * (ptrdiff_t)plane;
        const float* vs = src - off; float* vd = dst - off;
        const float* vp = prev ? prev - off : nullptr; float* vg = dog ? dog - off : nullptr;
#define CALLZ(H) launch_march<H>(vs, vd, nx, nz_glob, (ll)plane, ny, (ll)nx, t, vp, vg, slot, st, zlo, zhi)
        S3D_HW_SWITCH(t.hw, CALLZ)
#undef CALLZ
    } else {
        const int b0 = std::max(za, zlo - t.hw), b1 = std::min(za + nzl, zhi + t.hw);  // planes the launch touches
        const size_t o = (size_t)(b0 - za) * plane;
        S3D_LAUNCH(blur_generic_kernel, s3d_blocks(plane * (size_t)(b1 - b0), 256), 256, 0, st, src + o, dst + o, nx, ny, b1 - b0, 2, t,
                   prev ? prev + o : nullptr, dog ? dog + o : nullptr, slot, nz_glob, b0, zlo, zhi);
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


// ---- context plumbing ---------------------------------------------------------------------------
void free_levels(s3d_ctx* c) {
    for (auto& p : c->gss) if (p) { s3d::dev_free(p, c->stream); p = nullptr; }
    for (auto& p : c->dog) if (p) { s3d::dev_free(p, c->stream); p = nullptr; }
    for (int i = 0; i < 2; i++) if (c->d_tmp[i]) { s3d::dev_free(c->d_tmp[i], c->stream); c->d_tmp[i] = nullptr; }
    c->levels_alive = false;
}

int ctx_common_init(s3d_ctx* c, int nx, int ny, int nz, const s3d_params* p) {
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
    host_sigmas(c->L, c->prm.sigma_default, c->prm.sigma_n_default, c->sig);
    for (int i = 0; i < c->G; i++)
        if (host_taps(c->sig[i], reinterpret_cast<Taps*>(&c->taps[i])) < 0)
            return fail(S3D_ERR_ARG, "sigma %g needs more than %d taps per side", c->sig[i], kMaxHW);
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_slots, 256 * sizeof(unsigned), c->stream));
    S3D_CUDA(cudaMemsetAsync(c->d_slots, 0, 256 * sizeof(unsigned), c->stream));
    return S3D_OK;
}

// data_scale (Src/cUtil.cc:536-564), first sweep: max|v| folded into d_slots[0]
int stage_input_max(s3d_ctx* c, const float* d, size_t n) {
    if (n == 0) return S3D_OK;
    const unsigned grid = (unsigned)std::min<size_t>(s3d_blocks(n / 4 + 1, 256), 148 * 16);
    ProfScope ps(&c->prof, K_MAXABS, 4.0 * n);
    S3D_LAUNCH(maxabs_kernel, grid, 256, 0, c->stream, d, n, c->d_slots);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// second sweep: v / max (IEEE division per voxel)
int stage_input_normalize(s3d_ctx* c, const float* src, float* dst, size_t n) {
    if (n == 0) return S3D_OK;
    const unsigned grid = (unsigned)std::min<size_t>(s3d_blocks(n / 4 + 1, 256), 148 * 16);
    ProfScope ps(&c->prof, K_NORMALIZE, 8.0 * n);
    S3D_LAUNCH(normalize_kernel, grid, 256, 0, c->stream, src, dst, n, c->d_slots);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

static int ctx_normalize(s3d_ctx* c, const float* d_raw) {
    // ctor: data_scale (Src/cUtil.cc:536-564)
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_input, c->n0 * sizeof(float), c->stream));
    S3D_TRY(stage_input_max(c, d_raw, c->n0));
    return stage_input_normalize(c, d_raw, c->d_input, c->n0);
}

}  // namespace s3d

using namespace s3d;

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
        void* ptrs[] = {c->d_input, c->d_slots, c->d_thres, c->d_extre, c->d_codes, c->d_xyz5, c->d_kps, c->d_desc, c->d_redo, c->d_counts};
        for (void* q : ptrs) if (q) s3d::dev_free(q, c->stream);
        cudaStreamSynchronize(c->stream);
        c->prof.resolve();
        if (c->ev_ok) for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev[i]);
        for (auto& e : c->ph_ev) if (e) cudaEventDestroy(e);
        if (c->own_stream) cudaStreamDestroy(c->stream);
    }
    delete c;
}

}  // extern "C"

namespace s3d {

// ---- the pipeline in stages (shared by the unsharded run and the z-slab shards) ----------------
//
// Initialize (Src/cSIFT3D.cc:237-266) -> per octave Build_Gaussian_Scale_Space (:268-319) with
// Build_DOG_Scale_Space (:346-360) fused into the Z pass -> Detect_KeyPoints (:362-425) ->
// Assign_Orientation (:427-482) -> Extract_Description (:484-502).

// Planes of its own octave that the windows of keypoint level `lvl` reach beyond the keypoint's plane:
// r = 2*7.0711*scale (Src/cSIFT3D.cc:1155-1156), clamped windows :1182-1198, +1 for the central difference, +1 for ceil
// (the orientation window, r = 4.5*scale :27-28,:915, is smaller).  scale/unit does not depend on the octave.
int slab_window_halo(const s3d_ctx* c, int lvl) {
    const float ratio = host_level_scale(0, lvl, c->L, c->prm.sigma_default);
    return (int)ceilf(2.0f * 7.071067812f * ratio) + 2;
}

static int halo_for(const s3d_ctx* c) {
    // local planes a shard keeps beyond its owned range: the source halo of the widest blur (hw + 1: every level is
    // produced on owned +-1 so that the DoG neighbours of detection are local) and the widest descriptor window
    // (s3d_slab.cu builds the levels in groups with one exchange per group: the deepest source halo is 1 + sum(hw))
    int hwsum = 0;
    for (int i = 0; i < c->G; i++) hwsum += c->taps[i].hw;
    return std::max(hwsum + 2, slab_window_halo(c, c->L));
}

int slab_halo_from_params(const s3d_params* p, int* halo) {
    s3d_ctx tmp;
    s3d_params def;
    s3d_default_params(&def);
    tmp.prm = p ? *p : def;
    tmp.L = tmp.prm.num_kp_levels;
    if (tmp.L < 1 || tmp.L + 3 > kMaxG) return fail(S3D_ERR_ARG, "num_kp_levels %d unsupported", tmp.L);
    tmp.G = tmp.L + 3;
    host_sigmas(tmp.L, tmp.prm.sigma_default, tmp.prm.sigma_n_default, tmp.sig);
    for (int i = 0; i < tmp.G; i++)
        if (host_taps(tmp.sig[i], reinterpret_cast<Taps*>(&tmp.taps[i])) < 0) return fail(S3D_ERR_ARG, "sigma %g too wide", tmp.sig[i]);
    *halo = halo_for(&tmp);
    return S3D_OK;
}

int stage_init(s3d_ctx* c) {
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
    if (c->slab) c->halo = halo_for(c);
    for (int o = 0; o < c->noct; o++) {
        const int nzo = c->dims[o][2];
        c->full[o] = false;
        if (!c->slab) {
            c->p0[o] = c->za[o] = 0; c->p1[o] = c->zb[o] = nzo;
        } else {
            // plane k of octave o is owned iff own0 <= k * 2^o < own1 (== the owner of plane 2k one octave down)
            const int sh = 1 << o;
            c->p0[o] = std::min(nzo, (c->own0 + sh - 1) >> o);
            c->p1[o] = std::min(nzo, (c->own1 + sh - 1) >> o);
            if (o >= c->first_full) {  // replicated octave: every shard holds all planes
                c->full[o] = true;
                c->za[o] = 0; c->zb[o] = nzo;
            } else {
                c->za[o] = std::max(0, c->p0[o] - c->halo);
                c->zb[o] = std::min(nzo, c->p1[o] + c->halo);
                if (c->p1[o] <= c->p0[o]) { c->za[o] = c->zb[o] = c->p0[o]; }  // owns nothing here
            }
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
        for (int i = 0; i < 2; i++) S3D_CUDA(cudaMemsetAsync(c->d_tmp[i], 0xFF, tmp_elems * sizeof(float), st));
    }
    c->levels_alive = true;
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_thres, sizeof(float) * c->noct * L, st));
    {
        // the icosahedron table is the same for every handle: built and uploaded once per device, from a buffer that
        // outlives the copy (no stream synchronisation on the extraction path)
        static std::mutex mu;
        static MeshConst* dev_mesh[64] = {nullptr};
        static MeshConst host_copy;
        static bool host_ready = false;
        std::lock_guard<std::mutex> lk(mu);
        if (!host_ready) { host_mesh(&host_copy); host_ready = true; }
        MeshConst*& dm = dev_mesh[c->device & 63];
        if (!dm) {
            S3D_CUDA(cudaMalloc((void**)&dm, sizeof(MeshConst)));
            S3D_CUDA(cudaMemcpy(dm, &host_copy, sizeof(MeshConst), cudaMemcpyHostToDevice));
        }
        c->d_mesh = dm;  // shared, never freed
    }
    S3D_CUDA(cudaEventRecord(c->ev[1], st));
    c->stage = 1;
    return S3D_OK;
}

// Octave seed = even-index decimation of level L of the previous octave (:311) for planes [k0, k1) of octave o
// (a shard: its owned planes; halo planes of the seed come from their owners, s3d_slab.cu).
int stage_seed(s3d_ctx* c, int o, int k0, int k1) {
    cudaStream_t st = c->stream;
    const int L = c->L, G = c->G;
    const int nown = k1 - k0;
    if (o < 1 || nown <= 0) return S3D_OK;
    const int nx = c->dims[o][0], ny = c->dims[o][1];
    const float* src = c->gss[(o - 1) * G + L] + (size_t)(2 * k0 - c->za[o - 1]) * c->plane(o - 1);
    float* dst = c->gss[o * G] + (size_t)(k0 - c->za[o]) * c->plane(o);
    const size_t nout = (size_t)nown * c->plane(o);
    ProfScope ps(&c->prof, K_DOWNSAMPLE, 8.0 * nout);
    S3D_LAUNCH(downsample_kernel, s3d_blocks(nout, 256), 256, 0, st, src, c->dims[o - 1][0], c->dims[o - 1][1], dst, nx, ny, nown);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// Level i of octave o on the output planes [zlo, zhi): X -> Y -> Z (:609-617); the Z pass of level i >= 1 also emits
// DoG[i-1] = G[i-1] - G[i] on those planes and folds max|DoG| over them into its slot.  The X/Y passes are plane-local
// and run on the planes the Z pass reads, [zlo - hw, zhi + hw) clipped to the level; the source must be valid there.
int stage_level(s3d_ctx* c, int o, int i, int zlo, int zhi) {
    cudaStream_t st = c->stream;
    const int G = c->G, D = c->D;
    const int nx = c->dims[o][0], ny = c->dims[o][1], nzg = c->dims[o][2];
    const int za = c->za[o], nzl = c->zb[o] - c->za[o];
    if (zhi <= zlo || nzl <= 0) return S3D_OK;
    if (i == 0 && o > 0) return S3D_OK;  // the seed is produced by stage_seed
    const Taps& t = taps_of(c, i);
    float* dst = c->gss[o * G + i];
    const float* src = (o == 0 && i == 0) ? c->d_input : c->gss[o * G + i - 1];
    const int x0 = std::max(std::max(0, za), zlo - t.hw), x1 = std::min(std::min(nzg, za + nzl), zhi + t.hw);
    const size_t off = (size_t)(x0 - za) * c->plane(o);
    if (!blur_xy_pass(src + off, c->d_tmp[1] + off, nx, ny, x1 - x0, t, st, &c->prof)) {
        blur_pass(src + off, c->d_tmp[0] + off, nx, ny, x1 - x0, 0, t, 1, nullptr, nullptr, nullptr, st, &c->prof);
        blur_pass(c->d_tmp[0] + off, c->d_tmp[1] + off, nx, ny, x1 - x0, 1, t, 1, nullptr, nullptr, nullptr, st, &c->prof);
    }
    if (i >= 1)
        blur_z(c->d_tmp[1], dst, src, c->dog[o * D + i - 1], c->d_slots + 1 + o * D + i - 1, nx, ny, nzg, za, nzl, zlo, zhi, t, st, &c->prof);
    else
        blur_z(c->d_tmp[1], dst, nullptr, nullptr, nullptr, nx, ny, nzg, za, nzl, zlo, zhi, t, st, &c->prof);
    S3D_CUDA(cudaGetLastError());
    return S3D_OK;
}

// ---- capacities of the sparse stage ------------------------------------------------------------------------------
// Detection and survivor counts are produced on the device and consumed there (grids are fixed or persistent, kernels
// read the counts): a step is ONE uninterrupted stream, the host learns the counts in s3d_wait.  Buffers are therefore
// sized optimistically — from the densities this process has seen so far (times two), else from small defaults —
// and a run that exceeds them is repeated from the kept pyramid with exact sizes (s3d_wait), so results never depend
// on the guess.
static std::atomic<double> g_dens_extre{1.0 / 512.0}, g_dens_kps{1.0 / 4096.0};  // per owned voxel
static void learn_density(std::atomic<double>& a, double v) {
    double cur = a.load();
    while (v > cur && !a.compare_exchange_weak(cur, v)) {}
}

// Counter block of one run (ints, device): see s3d_wait.
enum { CT_NE = 0, CT_RECHECK = 1, CT_FLIPPED = 2, CT_NKPS = 3, CT_STAGED = 4, CT_WORK_A = 5, CT_WORK_B = 6, CT_REDO = 7, CT_N = 8 };

int stage_sparse(s3d_ctx* c) {
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
    c->own_total = own_total;
    {   // tests force tiny capacities to exercise the resize path
        const char* e1 = getenv("S3D_CAP_EXTRE");
        const char* e2 = getenv("S3D_CAP_KPS");
        if (c->cap_extre <= 0 && e1 && atoi(e1) > 0) c->cap_extre = atoi(e1);
        if (c->cap_kps <= 0 && e2 && atoi(e2) > 0) c->cap_kps = atoi(e2);
    }
    if (c->cap_extre <= 0) c->cap_extre = (int)std::min<double>(2.0e9 / 256, std::max(65536.0, 2.0 * g_dens_extre.load() * (double)own_total));
    if (c->cap_kps <= 0) c->cap_kps = (int)std::min<double>((double)c->cap_extre, std::max(8192.0, 2.0 * g_dens_kps.load() * (double)own_total));
    const int cap_e = c->cap_extre, cap_k = c->cap_kps;
    const int nblk = std::max<int>(1, (int)gb_base[(size_t)c->noct * L]);
    const unsigned stage_cap = (unsigned)cap_e;
    int *d_blk_cnt = nullptr, *d_blk_off = nullptr;
    StageEntry* d_stage = nullptr;
    Cand* d_cand = nullptr;
    S3D_CUDA(s3d::dev_alloc((void**)&d_blk_cnt, sizeof(int) * nblk, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_blk_off, sizeof(int) * nblk, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_counts, sizeof(int) * CT_N, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_stage, sizeof(StageEntry) * (size_t)stage_cap, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_cand, sizeof(Cand) * (size_t)cap_e, st));
    int* d_total = c->d_counts;
    unsigned* d_stage_count = (unsigned*)(d_total + CT_STAGED);
    S3D_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int) * CT_N, st));
    S3D_CUDA(cudaMemsetAsync(d_blk_cnt, 0, sizeof(int) * nblk, st));
    // One launch per (octave, level) unit.  (A single launch per octave over its three units was considered: the
    // kernel reads one level's threshold slot, and the launches of an octave already run back to back on the stream.)
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
        S3D_LAUNCH(scan_kernel, 1, 1024, 0, st, d_blk_cnt, d_blk_off, nblk, d_total + CT_NE);
    }
    {
        ProfScope ps(&c->prof, K_COMPACT, 0.0);
        S3D_LAUNCH(scatter_kernel, 148 * 4, 256, 0, st, d_stage, d_stage_count, stage_cap, d_blk_off, d_cand, cap_e);
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
        S3D_LAUNCH(wtab_kernel, dim3(2, c->noct * G), 256, 0, st, tab, c->noct, d_wtab);
    }
    int *d_recheck = nullptr;
    int* d_surv = nullptr;
    int* d_order = nullptr;
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_extre, sizeof(s3d_keypoint) * (size_t)cap_e, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_codes, sizeof(int) * (size_t)cap_e, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_xyz5, sizeof(int) * 5 * (size_t)cap_e, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_recheck, sizeof(int) * (size_t)cap_e, st));
    S3D_CUDA(s3d::dev_alloc((void**)&d_surv, sizeof(int) * (size_t)cap_e, st));
    {
        {
            ProfScope ps(&c->prof, K_ORIENT, 0.0);
            S3D_LAUNCH(orient_kernel, 148 * 32, 256, 0, st, d_cand, (const int*)(d_total + CT_NE), cap_e, tab, c->d_extre, c->d_codes, c->d_xyz5,
                       c->prm.max_eig_thres, c->prm.corner_thresh, 2e-3f, c->prm.exact_recheck ? d_recheck : (int*)nullptr,
                       d_total + CT_RECHECK);
        }
        ProfScope ps(&c->prof, K_ORIENT_EXACT, 0.0);
        if (c->prm.exact_recheck)
            S3D_LAUNCH(orient_exact_kernel, 148 * 6, kExactWarps * 32, 0, st, d_cand, tab, c->d_extre, c->d_codes, d_recheck,
                       d_total + CT_RECHECK, c->prm.max_eig_thres, c->prm.corner_thresh, d_total + CT_FLIPPED);
    }
    {
        ProfScope ps(&c->prof, K_SURVIVORS, 0.0);
        S3D_LAUNCH(survivors_kernel, 1, 1024, 0, st, c->d_codes, (const int*)(d_total + CT_NE), cap_e, d_surv, d_total + CT_NKPS);
        // heavy-first launch order of the descriptor CTAs (S3D_DESC_ORDER=0: list order)
        if (desc_order_enabled()) {
            S3D_CUDA(s3d::dev_alloc((void**)&d_order, sizeof(int) * (size_t)cap_e, st));
            S3D_LAUNCH(desc_order_kernel, 1, 1024, 0, st, c->d_extre, d_surv, d_total + CT_NKPS, G - 1, d_order);
        }
    }
    S3D_CUDA(cudaEventRecord(c->ev[4], st));

    // ---- Description -----------------------------------------------------------------------------
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_kps, sizeof(s3d_keypoint) * (size_t)cap_k, st));
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_desc, sizeof(float) * S3D_DESC_LEN * (size_t)cap_k, st));
    {
        static bool attr_set[64] = {false};
        if (c->device < 64 && !attr_set[c->device]) {
            S3D_CUDA(cudaFuncSetAttribute(describe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescSmem)));
            S3D_CUDA(cudaFuncSetAttribute(describe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescSmemQ)));
            attr_set[c->device] = true;
        }
        const int path = g_describe_path.load();
        // one CTA per keypoint SLOT (the count is on the device; surplus CTAs leave at once)
        const int grid_q = cap_k, grid_f = cap_k;
        if (path == 1) {  // FP32 staged/ordered accumulation for every keypoint
            ProfScope ps(&c->prof, K_DESCRIBE, 0.0);
            S3D_LAUNCH(describe_kernel<false>, grid_f, kDescWarps * 32, sizeof(DescSmem), st, c->d_extre, d_surv, cap_e, tab,
                       (const MeshConst*)c->d_mesh, c->d_kps, c->d_desc, (const int*)d_order, (const int*)(d_total + CT_NKPS), (int*)nullptr,
                       (int*)nullptr, 0.0f, cap_k);
        } else {
            // fixed-point atomics for all; the (normally empty) list of keypoints whose scale estimate was
            // too low is redone in FP32 — its grid reads the redo count on the device, so there is no host sync
            S3D_CUDA(s3d::dev_alloc((void**)&c->d_redo, sizeof(int) * ((size_t)cap_k + 1), st));
            {
                ProfScope ps(&c->prof, K_DESCRIBE, 0.0);
                S3D_LAUNCH(describe_kernel<true>, grid_q, kDescWarps * 32, sizeof(DescSmemQ), st, c->d_extre, d_surv, cap_e, tab,
                           (const MeshConst*)c->d_mesh, c->d_kps, c->d_desc, (const int*)d_order, (const int*)(d_total + CT_NKPS), c->d_redo,
                           d_total + CT_REDO, path == 2 ? 0.02f : kQMargin, cap_k);
            }
            ProfScope ps(&c->prof, K_DESCRIBE_REDO, 0.0);
            S3D_LAUNCH(describe_kernel<false>, grid_f, kDescWarps * 32, sizeof(DescSmem), st, c->d_extre, d_surv, cap_k, tab,
                       (const MeshConst*)c->d_mesh, c->d_kps, c->d_desc, (const int*)c->d_redo, (const int*)(d_total + CT_REDO), (int*)nullptr,
                       (int*)nullptr, 0.0f, cap_k);
        }
    }
    S3D_CUDA(cudaGetLastError());
    S3D_CUDA(cudaEventRecord(c->ev[5], st));

    // ---- Release_SIFT (:1659-1678): the pyramids are released in s3d_wait, once the counts have confirmed the capacities
    void* tmp[] = {d_blk_cnt, d_blk_off, d_stage, d_cand, d_recheck, d_surv, d_order, d_wtab};
    for (void* q : tmp) if (q) s3d::dev_free(q, st);
    S3D_CUDA(cudaEventRecord(c->ev[6], st));
    c->queued = true;
    c->stage = 100;
    return S3D_OK;
}

}  // namespace s3d

extern "C" {

static int run_impl(s3d_ctx* c) {
    if (c->ran || c->stage != 0) return fail(S3D_ERR_STATE, "s3d_run called twice on one handle (KpSiftAlgorithm is single-shot)");
    if (c->slab) return fail(S3D_ERR_STATE, "a z-slab shard is driven by s3d_slab_run / s3d_extract_multi, not by s3d_run");
    S3D_TRY(stage_init(c));
    for (int o = 0; o < c->noct; o++) {
        S3D_TRY(stage_seed(c, o, 0, c->dims[o][2]));
        for (int i = 0; i < c->G; i++) S3D_TRY(stage_level(c, o, i, 0, c->dims[o][2]));
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
    S3D_CUDA(cudaSetDevice(c->device));
    if (c->ran) {
        S3D_CUDA(cudaStreamSynchronize(c->stream));
        return S3D_OK;
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        // the ONLY host wait of a step: the counts the device produced and consumed on its own
        int h[CT_N] = {0};
        S3D_CUDA(cudaMemcpyAsync(h, c->d_counts, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        S3D_CUDA(cudaStreamSynchronize(c->stream));
        const int ne = h[CT_NE], nk = h[CT_NKPS];
        const bool fits = ne <= c->cap_extre && (unsigned)h[CT_STAGED] <= (unsigned)c->cap_extre && nk <= c->cap_kps;
        if (c->own_total > 0) {
            learn_density(g_dens_extre, (double)ne / (double)c->own_total);
            learn_density(g_dens_kps, (double)nk / (double)c->own_total);
        }
        if (fits) {
            c->n_extre = ne; c->n_rechecked = h[CT_RECHECK]; c->n_flipped = h[CT_FLIPPED]; c->n_kps = nk; c->n_desc_redo = h[CT_REDO];
            break;
        }
        if (attempt == 1 || !c->levels_alive)
            return fail(S3D_ERR_CAPACITY, "sparse stage: %d detections / %d keypoints exceed the buffers (%d / %d) after resizing", ne, nk,
                        c->cap_extre, c->cap_kps);
        // the optimistic buffers were too small for this volume: repeat the sparse stage from the kept pyramid with
        // exact sizes (the densities are learned, so the next volume of the job is sized right)
        void* old[] = {c->d_extre, c->d_codes, c->d_xyz5, c->d_kps, c->d_desc, c->d_redo, c->d_counts};
        for (void* q : old) if (q) s3d::dev_free(q, c->stream);
        c->d_extre = nullptr; c->d_codes = nullptr; c->d_xyz5 = nullptr; c->d_kps = nullptr; c->d_desc = nullptr; c->d_redo = nullptr;
        c->d_counts = nullptr;
        c->cap_extre = std::max(ne, h[CT_STAGED]) + 1024;
        c->cap_kps = std::max(nk, 1) + 1024;
        c->n_resized++;
        c->prof.resolve();
        for (int k = 0; k < K_NCLS; ++k) if (k >= K_DETECT) { c->prof.ms[k] = 0; c->prof.cnt[k] = 0; c->prof.bytes[k] = 0; }
        S3D_TRY(stage_sparse(c));
    }
    if (c->d_redo) { s3d::dev_free(c->d_redo, c->stream); c->d_redo = nullptr; }
    if (c->d_counts) { s3d::dev_free(c->d_counts, c->stream); c->d_counts = nullptr; }
    if (!c->prm.keep_levels) free_levels(c);
    {
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

// ---- z-slab shards: plane bookkeeping; the orchestration is in s3d_slab.cu ------------------

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

int s3d_slab_info(s3d_handle c, int* noct, int* halo, int* levels_per_octave, int* first_replicated_octave) {
    if (!c || !c->slab) return fail(S3D_ERR_ARG, "not a slab handle");
    if (c->stage < 1) return fail(S3D_ERR_STATE, "the shard has not been initialised");
    if (noct) *noct = c->noct;
    if (halo) *halo = c->halo;
    if (levels_per_octave) *levels_per_octave = c->G;
    if (first_replicated_octave) *first_replicated_octave = std::min(c->first_full, c->noct);
    return S3D_OK;
}

int s3d_slab_level_buffer(s3d_handle c, int which, int idx, float** d_ptr, int* ext4) {
    if (!c || !d_ptr || !ext4) return fail(S3D_ERR_ARG, "bad argument");
    if (c->stage < 1 || !c->levels_alive) return fail(S3D_ERR_STATE, "levels not allocated");
    const int per = which == 0 ? c->G : c->D;
    if (which < 0 || which > 1 || idx < 0 || idx >= c->noct * per) return fail(S3D_ERR_ARG, "level index out of range");
    const int o = idx / per;
    *d_ptr = which == 0 ? c->gss[idx] : c->dog[idx];
    ext4[0] = c->za[o]; ext4[1] = c->zb[o]; ext4[2] = c->p0[o]; ext4[3] = c->p1[o];
    return S3D_OK;
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
    out4[0] = c->n_rechecked; out4[1] = c->n_flipped; out4[2] = c->n_desc_redo; out4[3] = c->n_resized;
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
