// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" harness linked against the UNMODIFIED reference sources
// (/root/reference/3DSIFT/Src/*.cc compiled where they lie, see oracle/Makefile) so that
// tests/ and bench.py's cpu_baseline / --impl reference arm can drive the reference's own
// OpenMP implementation through ctypes.  Output: oracle/_ref/libsift3d_ref.so (git-ignored).
// Nothing in the product path (3dsift_b200/, include/) links, loads or calls this.
//
// Every entry point is a direct call into the reference's public surface:
//   CSIFT3DFactory::CreateCSIFT3D   cSIFT3D.cc:103-110
//   CSIFT3D::KpSiftAlgorithm        cSIFT3D.cc:165-235   (built with -DCHECK_ENABLE so the
//                                   pyramids survive, cSIFT3D.cc:223-225)
//   GET_GSS/GET_DOG/GET_LEVEL       cSIFT3D.h:169-177
//   GaussianSmooth_3D, DownSample_3D, Sub, Assign_Orientation_Imp, Extract_Descriptor_Imp
//                                   cSIFT3D.h:208-239
//   muBruteMatcher::*Match          cMatcher.cc:218-228
#include "cSIFT3D.h"
#include "cMatcher.h"
#include "cUtil.h"

#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <cstdio>
#include <cstring>
#include <vector>

using namespace CPUSIFT;

namespace {

// The reference keeps its state in protected members; a member-less subclass exposes them.
struct Peek : public CSIFT3D {
    using CSIFT3D::extre;
    using CSIFT3D::filter;
    using CSIFT3D::Host_Im;
    using CSIFT3D::octave_num;
    using CSIFT3D::num_kp_levels;
    using CSIFT3D::Gss_Pyramid;
    using CSIFT3D::DoG_Pyramid;
    using CSIFT3D::level_extrema;
    using CSIFT3D::mesh;
};

int g_quiet = 1;

// The reference prints from its hot path (cSIFT3D.cc:386,1010 ...); silence fd 1/2 around calls.
struct Quiet {
    int so = -1, se = -1;
    Quiet() {
        if (!g_quiet) return;
        fflush(stdout); fflush(stderr);
        std::cout.flush(); std::cerr.flush();
        int nul = open("/dev/null", O_WRONLY);
        so = dup(1); se = dup(2);
        dup2(nul, 1); dup2(nul, 2);
        close(nul);
    }
    ~Quiet() {
        if (so < 0) return;
        fflush(stdout); fflush(stderr);
        std::cout.flush(); std::cerr.flush();
        dup2(so, 1); dup2(se, 2);
        close(so); close(se);
    }
};

struct Ctx {
    CSIFT3D* s = nullptr;
    double t_create = 0, t_run = 0;
};

Peek* peek(Ctx* c) { return static_cast<Peek*>(c->s); }

void fill_tex(TexImage& t, const float* src, int nx, int ny, int nz, float unit) {
    t.SetImageSize(nx, ny, nz);
    t.SetImageUnit(unit, unit, unit);
    t.SetImageScale(1.0f);
    t.MallocArrayMemory();
    if (src) memcpy(t._Data, src, sizeof(float) * (size_t)nx * ny * nz);
}

}  // namespace

extern "C" {

void ref_set_quiet(int q) { g_quiet = q; }
int ref_sizeof_keypoint() { return (int)sizeof(Keypoint); }
int ref_sizeof_cvec() { return (int)sizeof(Cvec); }
int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) { sift_thread_num = n; omp_set_num_threads(n); } }

void* ref_create(const float* vol, int nx, int ny, int nz, int levels, float sigma, float sigma_n,
                 float peak, float eig, float corner) {
    Quiet q;
    Ctx* c = new Ctx;
    double t0 = omp_get_wtime();
    c->s = CSIFT3DFactory::CreateCSIFT3D(const_cast<float*>(vol), nx, ny, nz, levels, sigma, sigma_n,
                                         peak, eig, corner);
    c->t_create = omp_get_wtime() - t0;
    return c;
}

void* ref_create_default(const float* vol, int nx, int ny, int nz) {
    Quiet q;
    Ctx* c = new Ctx;
    double t0 = omp_get_wtime();
    c->s = CSIFT3DFactory::CreateCSIFT3D(const_cast<float*>(vol), nx, ny, nz);
    c->t_create = omp_get_wtime() - t0;
    return c;
}

void ref_run(void* h) {
    Quiet q;
    Ctx* c = (Ctx*)h;
    double t0 = omp_get_wtime();
    c->s->KpSiftAlgorithm();
    c->t_run = omp_get_wtime() - t0;
}

void ref_destroy(void* h) {
    Ctx* c = (Ctx*)h;
    delete c->s;
    delete c;
}

// which: 0 create, 1 run(total), 2 alloc, 3 gss, 4 dog, 5 detect, 6 orient, 7 desc
double ref_time(void* h, int which) {
    Ctx* c = (Ctx*)h;
    const SIFT_TimerPara& t = c->s->m_timer;
    switch (which) {
        case 0: return c->t_create;
        case 1: return c->t_run;
        case 2: return t.d_Allocation;
        case 3: return t.d_BuildGSS;
        case 4: return t.d_BuildDOG;
        case 5: return t.d_Detect;
        case 6: return t.d_AssignOrientation;
        case 7: return t.d_Extraction;
    }
    return -1;
}

int ref_num_octaves(void* h) { return peek((Ctx*)h)->octave_num; }

int ref_copy_input(void* h, float* out) {
    TexImage& t = peek((Ctx*)h)->Host_Im;
    memcpy(out, t._Data, sizeof(float) * (size_t)t._nx * t._ny * t._nz);
    return 0;
}

// which: 0 = GSS, 1 = DoG.  dims[3], meta[4] = {scale, ux, uy, uz}
int ref_level_info(void* h, int which, int idx, int* dims, float* meta) {
    std::vector<TexImage>* p = which == 0 ? ((Ctx*)h)->s->GET_GSS() : ((Ctx*)h)->s->GET_DOG();
    if (idx < 0 || idx >= (int)p->size()) return -1;
    TexImage& t = (*p)[idx];
    dims[0] = t._nx; dims[1] = t._ny; dims[2] = t._nz;
    meta[0] = t._s; meta[1] = t._ux; meta[2] = t._uy; meta[3] = t._uz;
    return 0;
}

int ref_copy_level(void* h, int which, int idx, float* out) {
    std::vector<TexImage>* p = which == 0 ? ((Ctx*)h)->s->GET_GSS() : ((Ctx*)h)->s->GET_DOG();
    if (idx < 0 || idx >= (int)p->size() || (*p)[idx]._Data == nullptr) return -1;
    TexImage& t = (*p)[idx];
    memcpy(out, t._Data, sizeof(float) * (size_t)t._nx * t._ny * t._nz);
    return 0;
}

// Raw detections after Assign_Orientation has run over them (rejected ones carry x=y=z=-1,
// cSIFT3D.cc:446-450) — includes str_tensor/win/eigvalue/eigvector debug fields.
int ref_num_extrema(void* h) { return (int)peek((Ctx*)h)->extre.size(); }
int ref_copy_extrema(void* h, void* out) {
    auto& v = peek((Ctx*)h)->extre;
    if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(Keypoint));
    return (int)v.size();
}

// Per-DoG-level raw detections in raster order (cSIFT3D.cc:412,419): x,y,z,octave,level as ints.
int ref_num_level_extrema(void* h) {
    size_t n = 0;
    for (auto& l : *((Ctx*)h)->s->GET_LEVEL()) n += l.size();
    return (int)n;
}
int ref_copy_level_extrema(void* h, int* out5) {
    size_t n = 0;
    for (auto& l : *((Ctx*)h)->s->GET_LEVEL())
        for (auto& k : l) {
            out5[5 * n + 0] = (int)k.x; out5[5 * n + 1] = (int)k.y; out5[5 * n + 2] = (int)k.z;
            out5[5 * n + 3] = k.octave; out5[5 * n + 4] = k.level;
            ++n;
        }
    return (int)n;
}

int ref_num_keypoints(void* h) { return (int)peek((Ctx*)h)->filter.size(); }
int ref_copy_keypoints(void* h, void* kp_out, float* desc_out) {
    std::vector<Keypoint> v = ((Ctx*)h)->s->GetKeypoints();  // by value, as a client would
    for (size_t i = 0; i < v.size(); ++i) {
        if (desc_out && v[i].desc) memcpy(desc_out + i * DESC_NUMEL, v[i].desc, sizeof(float) * DESC_NUMEL);
    }
    if (kp_out && !v.empty()) memcpy(kp_out, v.data(), v.size() * sizeof(Keypoint));
    return (int)v.size();
}

// ---- free kernels --------------------------------------------------------------------------

// GaussianSmooth_3D (cSIFT3D.cc:535-622) on a bare array.
void ref_gaussian_smooth(const float* src, int nx, int ny, int nz, float sigma, float* dst) {
    Quiet q;
    TexImage a, b;
    fill_tex(a, src, nx, ny, nz, 1.0f);
    fill_tex(b, nullptr, nx, ny, nz, 1.0f);
    GaussianSmooth_3D(&a, &b, sigma);
    memcpy(dst, b._Data, sizeof(float) * (size_t)nx * ny * nz);
}

// DownSample_3D (cSIFT3D.cc:506-533)
void ref_downsample(const float* src, int nx, int ny, int nz, float* dst) {
    TexImage a, b;
    fill_tex(a, src, nx, ny, nz, 1.0f);
    fill_tex(b, nullptr, nx / 2, ny / 2, nz / 2, 2.0f);
    DownSample_3D(&a, &b);
    memcpy(dst, b._Data, sizeof(float) * (size_t)(nx / 2) * (ny / 2) * (nz / 2));
}

// Initialize_geometry (cUtil.cc:113-175): v[20][3][3], idx[20][3]
void ref_mesh(float* v, int* idx) {
    Mesh m;
    Initialize_geometry(&m);
    for (int i = 0; i < ICOS_NFACES; ++i)
        for (int j = 0; j < 3; ++j) {
            v[(i * 3 + j) * 3 + 0] = m.tri[i].v[j].x;
            v[(i * 3 + j) * 3 + 1] = m.tri[i].v[j].y;
            v[(i * 3 + j) * 3 + 2] = m.tri[i].v[j].z;
            idx[i * 3 + j] = m.tri[i].idx[j];
        }
    free(m.tri);
}

// Assign_Orientation_Imp (cSIFT3D.cc:913-1138) + Extract_Descriptor_Imp (:1152-1381) for ONE
// keypoint on a bare level.  kp176 in/out; desc768 out (may be null to skip the descriptor).
int ref_orient_describe(const float* level, int nx, int ny, int nz, float unit, void* kp176,
                        float* desc768, float max_eig, float corner) {
    Quiet q;
    TexImage g;
    fill_tex(g, level, nx, ny, nz, unit);
    Keypoint* kp = (Keypoint*)kp176;
    Initialize_Keypoint(*kp);
    int res = Assign_Orientation_Imp(*kp, &g, 1.5f * kp->scale, max_eig, corner);
    if (res == 1 && desc768) {
        Mesh m;
        Initialize_geometry(&m);
        memset(desc768, 0, sizeof(float) * DESC_NUMEL);
        kp->desc = desc768;
        Extract_Descriptor_Imp(*kp, &g, &m);
        kp->desc = nullptr;
        free(m.tri);
    }
    return res;
}

// ---- matcher -------------------------------------------------------------------------------
// type: 1 inject, 2 biject, 3 enhanced (cMatcher.h:14-18).  Descriptors are row-major n x 768.
// Outputs (caller-allocated, length n_ref): gIdx (post-filter, cMatcher.cc:36), gDist, sIdx, sDist.
// pair_ref/pair_tar (length >= n_ref) receive the matched index pairs recovered from toCvec's
// coordinates (rx carries the keypoint index).  times[3] = matchTime, revMatchTime, totalTime.
int ref_match(int type, const float* ref_desc, int n_ref, const float* tar_desc, int n_tar, double thr,
              int* gIdx, float* gDist, int* sIdx, float* sDist, int* pair_ref, int* pair_tar,
              double* times) {
    Quiet q;
    std::vector<Keypoint> r(n_ref), t(n_tar);
    for (int i = 0; i < n_ref; ++i) {
        r[i].desc = const_cast<float*>(ref_desc) + (size_t)i * DESC_LENGTH;
        r[i].rx = (float)i; r[i].ry = 0; r[i].rz = 0;
    }
    for (int i = 0; i < n_tar; ++i) {
        t[i].desc = const_cast<float*>(tar_desc) + (size_t)i * DESC_LENGTH;
        t[i].rx = (float)i; t[i].ry = 0; t[i].rz = 0;
    }
    muBruteMatcher m;
    std::vector<Cvec> a, b;
    if (type == 1) m.injectMatch(a, b, r, t, thr);
    else if (type == 2) m.bijectMatch(a, b, r, t, thr);
    else m.enhancedMatch(a, b, r, t, thr);
    std::vector<int> gi = m.getGlodenIdx(), si = m.getSilverIdx();
    std::vector<float> gd = m.getGlodenDistSquare(), sd = m.getSilverDistSquare();
    for (int i = 0; i < n_ref; ++i) {
        if (gIdx) gIdx[i] = gi[i];
        if (sIdx) sIdx[i] = si[i];
        if (gDist) gDist[i] = gd[i];
        if (sDist) sDist[i] = sd[i];
    }
    for (size_t i = 0; i < a.size(); ++i) {
        if (pair_ref) pair_ref[i] = (int)a[i].x;
        if (pair_tar) pair_tar[i] = (int)b[i].x;
    }
    if (times) { times[0] = m.matchTime; times[1] = m.revMatchTime; times[2] = m.totalTime; }
    return (int)a.size();
}

}  // extern "C"
