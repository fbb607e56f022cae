// 3dsift/cSIFT3D.h — drop-in C++ surface of the B200-native extractor.
//
// Client code written against /root/reference/3DSIFT/Include/cSIFT3D.h (e.g. 3DSIFT/Example.cpp:
// 21-44) compiles unchanged against this header and links libsift3d_b200.so:
//   namespace CPUSIFT, Cvec (cSIFT3D.h:38-49), Keypoint (:52-70, same 176-byte layout),
//   CSIFT3DFactory::CreateCSIFT3D (both overloads, same defaults :187-202), CSIFT3D::
//   KpSiftAlgorithm / SetNumThreads / GetKeypoints (:151-155), the public stage methods
//   (:158-165), GET_GSS / GET_DOG / GET_LEVEL (:169-177), m_timer (:143), cmp_kp / cmp_kp_orig
//   (:72-74), sift_thread_num (:35).
// The class body is a thin façade over the C ABI (include/sift3d_b200.h); all arithmetic runs in
// the sm_100a kernels.  There is no CPU fallback: without a usable B200 the factory prints the
// error and KpSiftAlgorithm() produces no keypoints (the reference's print-and-continue
// convention).
#ifndef S3D_FACADE_CSIFT3D_H
#define S3D_FACADE_CSIFT3D_H

#include <string>
#include <vector>

#include "Util/cTexImage.h"
#include "Util/common.h"

namespace CPUSIFT {

#define SIGMA_DEFAULT 1.6
#define SIGMA_N_DEFAULT 1.15
#define NUM_KP_LEVELS 3
#define PEAK_THRESH 0.1
#define EIG_THRES 0.9
#define CORNER_THRESH 0.4
#define IMG_BORDER 1
#define NHIST_PER_DIM 4
#define ICOS_NFACES 20
#define ICOS_NVERT 12
#define DESC_NUMEL (NHIST_PER_DIM * NHIST_PER_DIM * NHIST_PER_DIM * ICOS_NVERT)

// Kept for source compatibility (SetNumThreads writes it); the GPU path has no CPU worker threads.
SIFT_LIBRARY_API extern int sift_thread_num;

typedef struct _cCvec {
    float x, y, z;
    _cCvec(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {}
} Cvec;

typedef struct _cKeypoint {
    float x, y, z;          // voxel coordinates inside the octave (integers)
    float scale;
    int octave, level;
    float rx, ry, rz;       // coordinates in the input volume: (x,y,z) * 2^octave
    Cvec win;               // weighted mean gradient of the orientation window
    float eigvalue[3];      // ascending eigenvalues of the structure tensor
    float eigvector[9];     // their (pre-sign-fix) eigenvectors, one per row
    float Rotation[9];      // keypoint frame (transposed, as the reference leaves it)
    float str_tensor[9];
    float* desc = nullptr;  // 768 floats, owned by the extractor that produced the keypoint
} Keypoint;
static_assert(sizeof(Keypoint) == 176, "Keypoint layout must match the reference (Include/cSIFT3D.h:52-70)");

SIFT_LIBRARY_API bool cmp_kp(const Keypoint& a, const Keypoint& b);
SIFT_LIBRARY_API bool cmp_kp_orig(const Keypoint& a, const Keypoint& b);

// Plain types of the reference's header (Include/cSIFT3D.h:77-116), kept so that client sources naming them still
// compile.  The icosahedron (Tri / Mesh, 20 faces) lives in device constant data here, Image is the reference's own
// legacy container (dead code there too), EigenVal pairs an eigenvalue with its vector.
typedef struct _cTri {
    Cvec v[3];   // vertices
    int idx[3];  // index of each vertex in the solid
} Tri;

typedef struct _Mesh {
    Tri* tri;  // triangles
    int num;   // number of triangles
} Mesh;

typedef struct _cImage {
    float* data;
    int nx, ny, nz;
    size_t xs, ys, zs;  // strides: xs = 1, ys = nx, zs = nx * ny
    float ux, uy, uz;
    size_t size;        // total size in voxels
    float s;            // scale-space location
} Image;

typedef struct _cEigenVal {
    float val;
    float vec[3];
} EigenVal;

class SIFT_LIBRARY_API CSIFT3D {
public:
    SIFT_TimerPara m_timer;

    CSIFT3D();
    CSIFT3D(float* volume, int x_dim, int y_dim, int z_dim, int num_kp_levels_, float sigma_default_,
            float sigma_n_default_, float peak_thresh_, float max_eig_thres_, float corner_thresh_);
    ~CSIFT3D();
    CSIFT3D(const CSIFT3D&) = delete;
    CSIFT3D& operator=(const CSIFT3D&) = delete;

    void KpSiftAlgorithm();
    void SetNumThreads(int t_num);
    std::vector<Keypoint> GetKeypoints();

    // The reference exposes its stages publicly (cSIFT3D.h:158-165).  On the GPU the stages are
    // one fused submission, so each of these makes sure the whole pipeline has run (once).
    void Initialize();
    void Build_Gaussian_Scale_Space();
    void Build_DOG_Scale_Space();
    void Detect_KeyPoints();
    void Assign_Orientation();
    void Extract_Description();
    void Release_SIFT();
    void SetHostImNull() {}

    // Parity hooks (cSIFT3D.h:169-177).  Levels are only retained when the extractor was created
    // with KeepLevels(true) before KpSiftAlgorithm (== building the reference with CHECK_ENABLE).
    void KeepLevels(bool keep);
    std::vector<TexImage>* GET_GSS();
    std::vector<TexImage>* GET_DOG();
    std::vector<std::vector<Keypoint> >* GET_LEVEL();

    // Extras of this implementation
    const float* DeviceDescriptors() const;   // K x 768 floats in HBM (extract -> match handoff)
    int LastStatus() const;                    // 0 = ok, else a S3D_ERR_* code
    const char* LastError() const;

private:
    struct Impl;
    Impl* impl;
};

class SIFT_LIBRARY_API CSIFT3DFactory {
public:
    static CSIFT3D* CreateCSIFT3D(float* volume, int x_dim, int y_dim, int z_dim, int num_kp_levels = NUM_KP_LEVELS,
                                  float sigma_default = SIGMA_DEFAULT, float sigma_n_default = SIGMA_N_DEFAULT,
                                  float peak_thresh = PEAK_THRESH, float max_eigo_thres = EIG_THRES,
                                  float corner_thresh = CORNER_THRESH);
    // volume file: int m, n, p then m*n*p float32, x fastest (Include/Util/matrixIO3D.h:22-64)
    static CSIFT3D* CreateCSIFT3D(std::string path_, int num_kp_levels = NUM_KP_LEVELS, float sigma_default = SIGMA_DEFAULT,
                                  float sigma_n_default = SIGMA_N_DEFAULT, float peak_thresh = PEAK_THRESH,
                                  float max_eigo_thres = EIG_THRES, float corner_thresh = CORNER_THRESH);
};

// Free kernels of the reference that make sense on host buffers (cSIFT3D.h:208-214).  The reference declares the rest
// of its kernels (Sub, IsExtrema_neighbor, Assign_Orientation_Imp, Extract_Descriptor_Imp, ... :216-239) WITHOUT
// SIFT_LIBRARY_API: they are not exported from its DLL, so no client can link them and they are not part of the
// drop-in surface; their work is the device kernels' (parity hooks: s3d_blur_axis, s3d_get_level, s3d_get_extrema).
SIFT_LIBRARY_API void DownSample_3D(TexImage* src, TexImage* dst);
SIFT_LIBRARY_API void GaussianSmooth_3D(TexImage* src, TexImage* dst, float sigma);

}  // namespace CPUSIFT

#endif
