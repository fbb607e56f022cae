// facade.cpp — the reference's C++ classes (include/3dsift/cSIFT3D.h, cMatcher.h) as a thin host
// layer over the C ABI (include/sift3d_b200.h).  No arithmetic of the hot path lives here.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>

#include <zlib.h>

#include <cstdint>
#include <vector>

#include "3dsift/Util/readNii.h"
#include "3dsift/cMatcher.h"
#include "3dsift/cSIFT3D.h"
#include "sift3d_b200.h"

namespace {

// Host descriptor blocks of live extractors -> their device copies.  Lets muBruteMatcher skip the
// host round trip when it is handed keypoints straight from GetKeypoints().
struct Resident {
    const float* d_desc;
    int n;
};
std::mutex g_reg_mu;
std::map<const float*, Resident> g_registry;

bool quiet() {
    static int q = -1;
    if (q < 0) {
        const char* e = getenv("SIFT3D_B200_VERBOSE");
        q = (e && e[0] && e[0] != '0') ? 0 : 1;
    }
    return q != 0;
}

}  // namespace

// ---- Util/common.h printers (reference: Src/Util/common.cpp:5-36) ---------------------------------
std::ostream& operator<<(std::ostream& os, const SIFT_TimerPara& st) {
    os << "\tTotal Time:" << st.d_TotalTime << "\n\tAllocation time:" << st.d_Allocation << "\n\tBuild-GSS time:" << st.d_BuildGSS
       << "\n\tBuild-DOG time:" << st.d_BuildDOG << "\n\tDetection time:" << st.d_Detect
       << "\n\tOrientation time:" << st.d_AssignOrientation << "\n\tExtract time:" << st.d_Extraction
       << "\n\tRelease time:" << st.d_release << "\n\tMemory ovhead time:" << st.d_memoryOverhead << std::endl;
    if (st.vD_octaveTime.size() == st.vD_octaveCompute.size())
        for (size_t i = 0; i < st.vD_octaveTime.size(); ++i)
            os << "\t----Octave " << i << "\ttime:" << st.vD_octaveTime[i] << "\tcompute-time:" << st.vD_octaveCompute[i] << std::endl;
    return os;
}

std::ostream& operator<<(std::ostream& os, const SIFT_PROCESS& sp) {
    os << "Reference:\n" << sp.REF << std::endl << "Target:\n" << sp.TAR << std::endl << "Match:\n" << sp.d_RegTime;
    return os;
}

// SIFT3D_B200_DEVICES="0,1,2,3": the devices one extraction / one match is spread over (the API
// stays the reference's, device selection comes from the environment).  One entry or unset: the single-device path.
static std::vector<int> multi_devices() {
    std::vector<int> out;
    const char* e = getenv("SIFT3D_B200_DEVICES");
    if (!e) return out;
    std::stringstream ss(e);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty()) out.push_back(atoi(tok.c_str()));
    if (out.size() < 2) out.clear();
    return out;
}

namespace CPUSIFT {

int sift_thread_num = 1;

// lexicographic (z, y, x) orders, Src/cSIFT3D.cc:40-75
bool cmp_kp(const Keypoint& a, const Keypoint& b) {
    if (a.rz != b.rz) return a.rz < b.rz;
    if (a.ry != b.ry) return a.ry < b.ry;
    return a.rx < b.rx;
}
bool cmp_kp_orig(const Keypoint& a, const Keypoint& b) {
    if (a.z != b.z) return a.z < b.z;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
}

struct CSIFT3D::Impl {
    s3d_handle h = nullptr;
    s3d_params prm;
    int status = S3D_OK;
    std::string err;
    bool ran = false, keep = false;
    int nx = 0, ny = 0, nz = 0;
    std::vector<float> volume;        // kept only until KpSiftAlgorithm when KeepLevels() may still change
    std::vector<Keypoint> filter;
    float* global_descriptor = nullptr;   // K x 768 host block the Keypoint::desc pointers borrow from
    std::vector<TexImage> gss, dog;
    std::vector<std::vector<Keypoint> > level_extrema;
    bool levels_fetched = false;

    void fail(int rc) {
        status = rc;
        err = s3d_last_error();
        std::cerr << "[sift3d_b200] " << err << std::endl;
    }
    void create() {
        if (h || volume.empty()) return;
        prm.keep_levels = keep ? 1 : 0;
        int rc = s3d_create(volume.data(), nx, ny, nz, &prm, &h);
        if (rc != S3D_OK) { h = nullptr; fail(rc); }
        std::vector<float>().swap(volume);
    }
};

CSIFT3D::CSIFT3D() : impl(new Impl) { s3d_default_params(&impl->prm); }

CSIFT3D::CSIFT3D(float* volume, int x_dim, int y_dim, int z_dim, int num_kp_levels_, float sigma_default_,
                 float sigma_n_default_, float peak_thresh_, float max_eig_thres_, float corner_thresh_)
    : impl(new Impl) {
    s3d_default_params(&impl->prm);
    impl->prm.num_kp_levels = num_kp_levels_;
    impl->prm.sigma_default = sigma_default_;
    impl->prm.sigma_n_default = sigma_n_default_;
    impl->prm.peak_thresh = peak_thresh_;
    impl->prm.max_eig_thres = max_eig_thres_;
    impl->prm.corner_thresh = corner_thresh_;
    const char* dev = getenv("SIFT3D_B200_DEVICE");
    if (dev && dev[0]) impl->prm.device = atoi(dev);
    const char* keep = getenv("SIFT3D_B200_KEEP_LEVELS");
    impl->keep = keep && keep[0] && keep[0] != '0';
    impl->nx = x_dim; impl->ny = y_dim; impl->nz = z_dim;
    // the reference's ctor copies the caller's buffer (Src/cSIFT3D.cc:161); the device copy is made
    // at the first use so that KeepLevels() can still be set between construction and the run
    if (volume && x_dim > 0 && y_dim > 0 && z_dim > 0) impl->volume.assign(volume, volume + (size_t)x_dim * y_dim * z_dim);
}

CSIFT3D::~CSIFT3D() {
    if (impl->global_descriptor) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_registry.erase(impl->global_descriptor);
    }
    if (impl->h) s3d_destroy(impl->h);
    free(impl->global_descriptor);
    delete impl;
}

void CSIFT3D::KeepLevels(bool keep) { impl->keep = keep; }
int CSIFT3D::LastStatus() const { return impl->status; }
const char* CSIFT3D::LastError() const { return impl->err.c_str(); }

void CSIFT3D::KpSiftAlgorithm() {
    if (impl->ran) return;  // single-shot
    impl->ran = true;
    int rc = S3D_OK;
    const std::vector<int> devs = impl->keep ? std::vector<int>() : multi_devices();
    if (!devs.empty() && !impl->h && !impl->volume.empty()) {
        // one volume over several GPUs: z-slabs with halo exchange between the devices (s3d_extract_multi); the merged
        // results live in the first shard's handle, which then answers like a single-device run.  (KeepLevels keeps the
        // single-device path: GET_GSS / GET_DOG return whole levels.)
        std::vector<s3d_handle> hs(devs.size(), nullptr);
        rc = s3d_extract_multi(impl->volume.data(), impl->nx, impl->ny, impl->nz, &impl->prm, devs.data(), (int)devs.size(), 1, hs.data());
        std::vector<float>().swap(impl->volume);
        if (rc != S3D_OK) return impl->fail(rc);
        impl->h = hs[0];
        for (size_t g = 1; g < hs.size(); ++g) s3d_destroy(hs[g]);
    } else {
        impl->create();
        if (!impl->h) return;
        rc = s3d_run(impl->h);
        if (rc != S3D_OK) return impl->fail(rc);
    }
    int n = 0;
    s3d_num_keypoints(impl->h, &n);
    impl->filter.resize(n);
    impl->global_descriptor = (float*)calloc((size_t)std::max(n, 1) * DESC_NUMEL, sizeof(float));
    static_assert(sizeof(Keypoint) == sizeof(s3d_keypoint), "record layouts must agree");
    rc = s3d_get_keypoints(impl->h, reinterpret_cast<s3d_keypoint*>(impl->filter.data()), impl->global_descriptor);
    if (rc != S3D_OK) return impl->fail(rc);
    for (int i = 0; i < n; ++i) impl->filter[i].desc = impl->global_descriptor + (size_t)i * DESC_NUMEL;  // Src/cSIFT3D.cc:495
    const float* dd = nullptr;
    int dn = 0;
    if (n > 0 && s3d_device_descriptors(impl->h, &dd, &dn) == S3D_OK && dd) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_registry[impl->global_descriptor] = Resident{dd, dn};
    }
    double t[10];
    s3d_get_timers(impl->h, t);
    m_timer.d_Allocation = t[0]; m_timer.d_BuildGSS = t[1]; m_timer.d_BuildDOG = t[2]; m_timer.d_Detect = t[3];
    m_timer.d_AssignOrientation = t[4]; m_timer.d_Extraction = t[5]; m_timer.d_release = t[6]; m_timer.d_TotalTime = t[7];
    m_timer.d_memoryOverhead = t[8] + t[9];
    if (!quiet()) {
        // SIFT3D_B200_VERBOSE=1: the lines the reference prints unconditionally while it runs (time_info
        // Src/cSIFT3D.cc:78-101, KpSiftAlgorithm :165-235), in its order and format, for scripts that scrape them.  The
        // stages are one fused submission here, so the lines appear together after the run, carrying the device times.
        int ne = 0;
        s3d_num_extrema(impl->h, &ne);
        auto stage_line = [](double seconds, const char* info) { std::cout << "\t\ttime:" << 1000.0 * seconds << "ms  ----" << info << std::endl; };
        stage_line(0.0, "----start");
        std::cout << "Initialization is OK" << std::endl;
        stage_line(t[0], "----Init done");
        stage_line(t[1], "----Build GSS");
        stage_line(t[2], "----Build DOG");
        stage_line(t[3], "----Detect keypoint");
        std::cout << "After detecting keypoints, kp size is : " << ne << std::endl;
        std::cout << "After Orientation, kp size is : " << n << std::endl;
        stage_line(t[4], "----Orientation");
        stage_line(t[5], "----Description");
        std::cout << "\ttotal time:" << t[7] << "s  ----" << "finish" << std::endl;
    }
}

void CSIFT3D::SetNumThreads(int t_num) { if (t_num > 0) sift_thread_num = t_num; }
std::vector<Keypoint> CSIFT3D::GetKeypoints() { return impl->filter; }
const float* CSIFT3D::DeviceDescriptors() const {
    const float* dd = nullptr;
    int n = 0;
    return (impl->h && s3d_device_descriptors(impl->h, &dd, &n) == S3D_OK) ? dd : nullptr;
}

void CSIFT3D::Initialize() { impl->create(); }
void CSIFT3D::Build_Gaussian_Scale_Space() { KpSiftAlgorithm(); }
void CSIFT3D::Build_DOG_Scale_Space() { KpSiftAlgorithm(); }
void CSIFT3D::Detect_KeyPoints() { KpSiftAlgorithm(); }
void CSIFT3D::Assign_Orientation() { KpSiftAlgorithm(); }
void CSIFT3D::Extract_Description() { KpSiftAlgorithm(); }
void CSIFT3D::Release_SIFT() {}

static void fetch_levels(s3d_handle h, int which, std::vector<TexImage>& out) {
    out.clear();
    for (int idx = 0;; ++idx) {
        int d[3];
        float meta[4];
        if (s3d_level_info(h, which, idx, d, meta) != S3D_OK) break;
        TexImage t(d[0], d[1], d[2]);
        t.SetImageScale(meta[0]);
        t.SetImageUnit(meta[1], meta[2], meta[3]);
        t.MallocArrayMemory();
        if (s3d_get_level(h, which, idx, t._Data) != S3D_OK) break;
        out.push_back(t);
    }
}

std::vector<TexImage>* CSIFT3D::GET_GSS() {
    if (impl->h && impl->ran && impl->gss.empty()) fetch_levels(impl->h, 0, impl->gss);
    return &impl->gss;
}
std::vector<TexImage>* CSIFT3D::GET_DOG() {
    if (impl->h && impl->ran && impl->dog.empty()) fetch_levels(impl->h, 1, impl->dog);
    return &impl->dog;
}
std::vector<std::vector<Keypoint> >* CSIFT3D::GET_LEVEL() {
    // per (octave, level) raw detections in raster order (Src/cSIFT3D.cc:412,419)
    if (impl->h && impl->ran && impl->level_extrema.empty()) {
        int ne = 0, noct = 0;
        s3d_num_extrema(impl->h, &ne);
        s3d_num_octaves(impl->h, &noct);
        std::vector<s3d_keypoint> kp(std::max(ne, 1));
        std::vector<int> codes(std::max(ne, 1)), xyz5(5 * (size_t)std::max(ne, 1));
        if (s3d_get_extrema(impl->h, kp.data(), codes.data(), xyz5.data()) == S3D_OK) {
            const int L = impl->prm.num_kp_levels;
            impl->level_extrema.assign((size_t)noct * L, std::vector<Keypoint>());
            for (int i = 0; i < ne; ++i) {
                Keypoint k;
                memcpy(&k, &kp[i], sizeof(k));
                k.x = (float)xyz5[5 * i]; k.y = (float)xyz5[5 * i + 1]; k.z = (float)xyz5[5 * i + 2];
                k.desc = nullptr;
                impl->level_extrema[(size_t)xyz5[5 * i + 3] * L + (xyz5[5 * i + 4] - 1)].push_back(k);
            }
        }
    }
    return &impl->level_extrema;
}

CSIFT3D* CSIFT3DFactory::CreateCSIFT3D(float* volume, int x_dim, int y_dim, int z_dim, int num_kp_levels, float sigma_default,
                                       float sigma_n_default, float peak_thresh, float max_eig_thres, float corner_thresh) {
    return new CSIFT3D(volume, x_dim, y_dim, z_dim, num_kp_levels, sigma_default, sigma_n_default, peak_thresh,
                       max_eig_thres, corner_thresh);
}

CSIFT3D* CSIFT3DFactory::CreateCSIFT3D(std::string path_, int num_kp_levels, float sigma_default, float sigma_n_default,
                                       float peak_thresh, float max_eig_thres, float corner_thresh) {
    // ReadMatrixFromDisk<float> (Include/Util/matrixIO3D.h:22-64): int m, n, p then m*n*p floats
    int dims[3] = {0, 0, 0};
    std::vector<float> vol;
    FILE* f = fopen(path_.c_str(), "rb");
    if (!f) {
        printf("Can't open input matrix file: %s.\n", path_.c_str());
    } else {
        if (fread(dims, sizeof(int), 3, f) == 3 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0) {
            vol.resize((size_t)dims[0] * dims[1] * dims[2]);
            if (fread(vol.data(), sizeof(float), vol.size(), f) != vol.size()) {
                printf("Error reading matrix from disk file: %s.\n", path_.c_str());
                vol.clear();
            }
        } else {
            printf("Error reading matrix header from disk file: %s.\n", path_.c_str());
        }
        fclose(f);
    }
    std::cout << dims[0] << " " << dims[1] << " " << dims[2] << std::endl;  // Src/cSIFT3D.cc:118
    return new CSIFT3D(vol.empty() ? nullptr : vol.data(), dims[0], dims[1], dims[2], num_kp_levels, sigma_default,
                       sigma_n_default, peak_thresh, max_eig_thres, corner_thresh);
}

void DownSample_3D(TexImage* src, TexImage* dst) {
    dst->SetImageSize(src->_nx / 2, src->_ny / 2, src->_nz / 2);
    dst->MallocArrayMemory();
    if (s3d_downsample(src->_Data, src->_nx, src->_ny, src->_nz, dst->_Data) != S3D_OK)
        std::cerr << "[sift3d_b200] " << s3d_last_error() << std::endl;
}

void GaussianSmooth_3D(TexImage* src, TexImage* dst, float sigma) {
    dst->SetImageSize(src->_nx, src->_ny, src->_nz);
    dst->SetImageUnit(src->_ux, src->_uy, src->_uz);
    dst->MallocArrayMemory();
    if (s3d_gaussian_smooth(src->_Data, src->_nx, src->_ny, src->_nz, sigma, dst->_Data) != S3D_OK)
        std::cerr << "[sift3d_b200] " << s3d_last_error() << std::endl;
}

// ---- matcher ------------------------------------------------------------------------------------

muBruteMatcher::muBruteMatcher() {}
float muBruteMatcher::getCalculationTime() { return totalTime; }
std::vector<float> muBruteMatcher::getGlodenDistSquare() { return gDist; }
std::vector<float> muBruteMatcher::getSilverDistSquare() { return sDist; }
std::vector<int> muBruteMatcher::getGlodenIdx() { return gIdx; }
std::vector<int> muBruteMatcher::getSilverIdx() { return sIdx; }

// Keypoint::desc pointers (Src/cMatcher.cc:20) -> one n x 768 block.  Returns the block and whether
// it lives on the device (descriptors of a live extractor, no copy) — else a host pointer, which
// is the caller's own memory when the pointers already form base + 768*i, or `scratch`.
// (*on_device < 0 on entry: the caller wants host memory whatever the registry knows — the multi-device matcher
// uploads a replica to every device itself)
static const float* gather(const std::vector<Keypoint>& kp, std::vector<float>& scratch, int* on_device) {
    const bool host_only = *on_device < 0;
    *on_device = host_only ? -1 : 0;
    const size_t n = kp.size();
    if (n == 0) return nullptr;
    bool contiguous = kp[0].desc != nullptr;
    for (size_t i = 1; i < n && contiguous; ++i) contiguous = kp[i].desc == kp[0].desc + i * DESC_LENGTH;
    if (contiguous) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        auto it = g_registry.find(kp[0].desc);
        if (*on_device >= 0 && it != g_registry.end() && (size_t)it->second.n == n) {
            *on_device = 1;
            return it->second.d_desc;
        }
        if (host_only) *on_device = 0;
        return kp[0].desc;
    }
    if (host_only) *on_device = 0;
    scratch.assign(n * DESC_LENGTH, 0.0f);
    for (size_t i = 0; i < n; ++i)
        if (kp[i].desc) memcpy(&scratch[i * DESC_LENGTH], kp[i].desc, sizeof(float) * DESC_LENGTH);
    return scratch.data();
}

void muBruteMatcher::run(int type, std::vector<Cvec>& refMatch, std::vector<Cvec>& tarMatch, const std::vector<Keypoint>& ref_kp,
                         const std::vector<Keypoint>& tar_kp, double thr) {
    const int n_ref = (int)ref_kp.size(), n_tar = (int)tar_kp.size();
    std::vector<float> s_ref, s_tar;
    const std::vector<int> devs = multi_devices();
    int ref_dev = devs.empty() ? 0 : -1, tar_dev = devs.empty() ? 0 : -1;
    const float* p_ref = gather(ref_kp, s_ref, &ref_dev);
    const float* p_tar = gather(tar_kp, s_tar, &tar_dev);
    // the reference (re)initialises its work vectors on every call (Src/cMatcher.cc:152-161)
    gDist.assign(n_ref, 0.0f); sDist.assign(n_ref, 0.0f); gIdx.assign(n_ref, -1); sIdx.assign(n_ref, -1);
    gDist2.assign(n_tar, 0.0f); sDist2.assign(n_tar, 0.0f); gIdx2.assign(n_tar, -1); sIdx2.assign(n_tar, -1);
    std::vector<int> pr(std::max(n_ref, 1)), pt(std::max(n_ref, 1));
    int np = 0;
    double times[3] = {0, 0, 0};
    if (!devs.empty())  // one match over several GPUs: database-sharded search, query-sharded exact re-rank (s3d_match_multi)
        status = s3d_match_multi(type, p_ref, n_ref, p_tar, n_tar, thr, devs.data(), (int)devs.size(), gIdx.data(), gDist.data(),
                                 sIdx.data(), sDist.data(), gIdx2.data(), gDist2.data(), sIdx2.data(), sDist2.data(), pr.data(),
                                 pt.data(), &np, times);
    else
        status = s3d_match_ex(type, p_ref, n_ref, ref_dev, p_tar, n_tar, tar_dev, thr, gIdx.data(), gDist.data(), sIdx.data(),
                              sDist.data(), gIdx2.data(), gDist2.data(), sIdx2.data(), sDist2.data(), pr.data(), pt.data(), &np, times);
    if (status != S3D_OK) {
        std::cerr << "[sift3d_b200] " << s3d_last_error() << std::endl;
        return;
    }
    for (int i = 0; i < np; ++i) {  // toCvec, Src/cMatcher.cc:99-112 (append)
        const Keypoint& a = ref_kp[pr[i]];
        const Keypoint& b = tar_kp[pt[i]];
        refMatch.push_back(Cvec(a.rx, a.ry, a.rz));
        tarMatch.push_back(Cvec(b.rx, b.ry, b.rz));
    }
    matchTime = (float)times[0]; revMatchTime = (float)times[1]; totalTime = (float)times[2];
}

void muBruteMatcher::injectMatch(std::vector<Cvec>& r, std::vector<Cvec>& t, const std::vector<Keypoint>& rk,
                                 const std::vector<Keypoint>& tk, const double thr) { run(1, r, t, rk, tk, thr); }
void muBruteMatcher::bijectMatch(std::vector<Cvec>& r, std::vector<Cvec>& t, const std::vector<Keypoint>& rk,
                                 const std::vector<Keypoint>& tk, const double thr) { run(2, r, t, rk, tk, thr); }
void muBruteMatcher::enhancedMatch(std::vector<Cvec>& r, std::vector<Cvec>& t, const std::vector<Keypoint>& rk,
                                   const std::vector<Keypoint>& tk, const double thr) { run(3, r, t, rk, tk, thr); }

// "%.5lf,%.5lf,%.5lf\n" per point (Src/cUtil.cc:938-954)
void write_sift_kp(std::vector<Cvec>& kp, const char* file_name) {
    FILE* f = fopen(file_name, "w");
    if (!f) return;
    for (const Cvec& c : kp) fprintf(f, "%.5lf,%.5lf,%.5lf\n", c.x, c.y, c.z);
    printf("sift keypoint size:%zd\n", kp.size());
    fclose(f);
}

void read_sift_kp(const char* file_name, std::vector<Cvec>& kp) {
    // one "x,y,z" per line, fields split at ',' and parsed with operator>>(float) as the reference's explode /
    // stringToNum<float> do (Src/cUtil.cc:956-1000); points are APPENDED (:1002-1016).  A line with fewer than three
    // fields reads past its vector in the reference; here the missing fields are 0.
    std::ifstream in(file_name);
    std::string line;
    int n = 0;
    while (std::getline(in, line)) {
        float v[3] = {0, 0, 0};
        std::stringstream ss(line);
        std::string tok;
        for (int k = 0; k < 3 && std::getline(ss, tok, ','); ++k) {
            std::istringstream iss(tok);
            float num = 0.0f;
            iss >> num;
            v[k] = num;
        }
        kp.push_back(Cvec(v[0], v[1], v[2]));
        ++n;
    }
    std::cout << "File:" << file_name << ", number of points:" << n << std::endl;
}

}  // namespace CPUSIFT

// ---- NIfTI ingest (reference: Src/Util/readNii.cpp:5-39 over layNii; own parser here) ---------------
namespace {

template <class T>
T nii_get(const unsigned char* p, bool swap) {
    unsigned char b[sizeof(T)];
    for (size_t i = 0; i < sizeof(T); ++i) b[i] = swap ? p[sizeof(T) - 1 - i] : p[i];
    T v;
    memcpy(&v, b, sizeof(T));
    return v;
}

template <class T>
void nii_convert(const unsigned char* src, size_t n, bool swap, float* dst) {
    for (size_t i = 0; i < n; ++i) dst[i] = static_cast<float>(nii_get<T>(src + i * sizeof(T), swap));
}

bool nii_fail(const char* fname, const char* why) {
    std::cerr << "[sift3d_b200] readNiiFile(" << fname << "): " << why << std::endl;
    return false;
}

}  // namespace

// Name of the voxel file of a NIfTI-1 .hdr/.img pair (nifti_findimgname, nifti2_io.cpp): same base name, .img or .img.gz.
static std::string nii_pair_image(const std::string& hdr_name) {
    std::string base = hdr_name;
    if (base.size() > 3 && base.compare(base.size() - 3, 3, ".gz") == 0) base.resize(base.size() - 3);
    if (base.size() > 4 && (base.compare(base.size() - 4, 4, ".hdr") == 0 || base.compare(base.size() - 4, 4, ".img") == 0)) base.resize(base.size() - 4);
    for (const char* ext : {".img", ".img.gz"}) {
        const std::string cand = base + ext;
        if (FILE* t = fopen(cand.c_str(), "rb")) { fclose(t); return cand; }
    }
    return std::string();
}

float* readNiiFile(const char* fname_in, int& nx, int& ny, int& nz) {
    nx = ny = nz = 0;
    // a pair may be named by its image file: the header is then <base>.hdr[.gz] (nifti_findhdrname)
    std::string fname_s = fname_in ? fname_in : "";
    {
        std::string base = fname_s;
        if (base.size() > 3 && base.compare(base.size() - 3, 3, ".gz") == 0) base.resize(base.size() - 3);
        if (base.size() > 4 && base.compare(base.size() - 4, 4, ".img") == 0) {
            base.resize(base.size() - 4);
            for (const char* ext : {".hdr", ".hdr.gz"}) {
                const std::string cand = base + ext;
                if (FILE* t = fopen(cand.c_str(), "rb")) { fclose(t); fname_s = cand; break; }
            }
        }
    }
    const char* fname = fname_s.c_str();
    // gzopen reads plain files transparently, so .nii and .nii.gz share one path
    gzFile f = gzopen(fname, "rb");
    if (!f) { nii_fail(fname, "cannot open"); return nullptr; }
    gzbuffer(f, 1 << 20);
    unsigned char hdr[540];
    if (gzread(f, hdr, 348) != 348) { gzclose(f); nii_fail(fname, "short header"); return nullptr; }
    int32_t sz = nii_get<int32_t>(hdr, false);
    bool swap = false;
    if (sz != 348 && sz != 540) {
        sz = nii_get<int32_t>(hdr, true);
        swap = true;
    }
    int64_t dim[8] = {0};
    int datatype = 0;
    int64_t vox_offset = 0;
    bool pair = false;  // NIfTI-1 two-file form: magic "ni1", voxels in <base>.img at vox_offset (usually 0)
    if (sz == 348) {  // NIfTI-1: dim[8] int16 @40, datatype int16 @70, vox_offset float @108, magic @344
        const bool single = hdr[344] == 'n' && hdr[345] == '+' && hdr[346] == '1';
        pair = hdr[344] == 'n' && hdr[345] == 'i' && hdr[346] == '1';
        if (!single && !pair) { gzclose(f); nii_fail(fname, "not a NIfTI-1 file (magic n+1 / ni1)"); return nullptr; }
        for (int i = 0; i < 8; ++i) dim[i] = nii_get<int16_t>(hdr + 40 + 2 * i, swap);
        datatype = nii_get<int16_t>(hdr + 70, swap);
        vox_offset = (int64_t)nii_get<float>(hdr + 108, swap);
    } else if (sz == 540) {  // NIfTI-2: magic @4, datatype int16 @12, dim[8] int64 @16, vox_offset int64 @168
        if (gzread(f, hdr + 348, 540 - 348) != 540 - 348) { gzclose(f); nii_fail(fname, "short NIfTI-2 header"); return nullptr; }
        if (!(hdr[4] == 'n' && hdr[5] == '+' && hdr[6] == '2')) { gzclose(f); nii_fail(fname, "not a single-file NIfTI-2 (magic n+2)"); return nullptr; }
        datatype = nii_get<int16_t>(hdr + 12, swap);
        for (int i = 0; i < 8; ++i) dim[i] = nii_get<int64_t>(hdr + 16 + 8 * i, swap);
        vox_offset = nii_get<int64_t>(hdr + 168, swap);
    } else {
        gzclose(f);
        nii_fail(fname, "sizeof_hdr is neither 348 nor 540");
        return nullptr;
    }
    if (dim[0] < 1 || dim[0] > 7 || dim[1] < 1) { gzclose(f); nii_fail(fname, "bad dim[]"); return nullptr; }
    const int64_t dx = dim[1], dy = dim[0] >= 2 ? dim[2] : 1, dz = dim[0] >= 3 ? dim[3] : 1;
    if (dx < 1 || dy < 1 || dz < 1 || dx > 0x7fffffff || dy > 0x7fffffff || dz > 0x7fffffff) { gzclose(f); nii_fail(fname, "bad volume dimensions"); return nullptr; }
    size_t bpv = 0;
    switch (datatype) {  // the types copy_nifti_as_float32 handles (laynii_lib.cpp:249-310) + float32 itself
        case 2: case 256: bpv = 1; break;             // uint8, int8
        case 4: case 512: bpv = 2; break;             // int16, uint16
        case 8: case 768: case 16: bpv = 4; break;    // int32, uint32, float32
        case 1024: case 1280: case 64: bpv = 8; break;  // int64, uint64, float64
        default: gzclose(f); nii_fail(fname, "unsupported datatype"); return nullptr;
    }
    int64_t have = sz;  // bytes of the voxel file consumed so far
    if (pair) {
        gzclose(f);
        const std::string img = nii_pair_image(fname_s);
        f = img.empty() ? nullptr : gzopen(img.c_str(), "rb");
        if (!f) { nii_fail(fname, "the .img file of the pair cannot be opened"); return nullptr; }
        gzbuffer(f, 1 << 20);
        have = 0;
        if (vox_offset < 0) vox_offset = 0;
    } else if (vox_offset < have) {
        vox_offset = have;  // the reference clamps a too-small offset to sizeof(header) (nifti_convert_n1hdr2nim / n2hdr2nim)
    }
    {   // skip the header extension (single file) / leading bytes (pair) up to vox_offset
        std::vector<unsigned char> skip((size_t)(vox_offset - have));
        if (!skip.empty() && gzread(f, skip.data(), (unsigned)skip.size()) != (int)skip.size()) { gzclose(f); nii_fail(fname, "short file (extension)"); return nullptr; }
    }
    const size_t n = (size_t)dx * dy * dz;
    std::vector<unsigned char> raw(n * bpv);
    size_t got = 0;
    while (got < raw.size()) {
        const unsigned want = (unsigned)std::min<size_t>(raw.size() - got, 1u << 30);
        const int r = gzread(f, raw.data() + got, want);
        if (r <= 0) break;
        got += (size_t)r;
    }
    gzclose(f);
    if (got != raw.size()) { nii_fail(fname, "short file (voxel data)"); return nullptr; }
    float* out = new float[n];
    switch (datatype) {
        case 2: nii_convert<uint8_t>(raw.data(), n, swap, out); break;
        case 256: nii_convert<int8_t>(raw.data(), n, swap, out); break;
        case 4: nii_convert<int16_t>(raw.data(), n, swap, out); break;
        case 512: nii_convert<uint16_t>(raw.data(), n, swap, out); break;
        case 8: nii_convert<int32_t>(raw.data(), n, swap, out); break;
        case 768: nii_convert<uint32_t>(raw.data(), n, swap, out); break;
        case 16: nii_convert<float>(raw.data(), n, swap, out); break;
        case 1024: nii_convert<int64_t>(raw.data(), n, swap, out); break;
        case 1280: nii_convert<uint64_t>(raw.data(), n, swap, out); break;
        case 64: nii_convert<double>(raw.data(), n, swap, out); break;
    }
    // copy_nifti_as_float32 ends with "Replace nans with zeros" (laynii_lib.cpp:303-308); readNiiFile only converts
    // inputs that are not float32 (Src/Util/readNii.cpp:16-20), so a float32 file keeps its NaNs
    if (datatype != 16)
        for (size_t i = 0; i < n; ++i)
            if (out[i] != out[i]) out[i] = 0.0f;
    nx = (int)dx; ny = (int)dy; nz = (int)dz;
    return out;
}

// C entry points for harnesses that cannot call the C++ signature (ctypes)
extern "C" {
S3D_API float* s3d_read_nii(const char* path, int* nx, int* ny, int* nz) {
    int a = 0, b = 0, c = 0;
    float* p = readNiiFile(path, a, b, c);
    if (nx) *nx = a;
    if (ny) *ny = b;
    if (nz) *nz = c;
    return p;
}
S3D_API void s3d_free_host(float* p) { delete[] p; }
}
