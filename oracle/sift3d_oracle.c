/* oracle/sift3d_oracle.c — TEST INFRASTRUCTURE ONLY (the "port" oracle).
 *
 * A plain-C CPU restatement of the 3DSIFT hot path, written from the reference's behaviour
 * (file:line cited per function, relative to /root/reference/3DSIFT/).  It exists so that
 * tests/ can check the CUDA path when the compiled reference (oracle/_ref) is not available and
 * so that every numerical rule the kernels must obey is written down once, readably.
 *
 * Pinned against: the reference itself.  The reference ships no tests or golden vectors
 * (SURVEY.md §4, §8c), so tests/test_oracle.py compares every function here with
 * oracle/_ref/libsift3d_ref.so (the unmodified reference compiled by oracle/Makefile) on seeded
 * inputs, and with the fixtures under tests/golden/ that were generated from it
 * (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * Build: make -C oracle port   (-O2 -ffp-contract=off: FP32 results must not be FMA-contracted,
 * the reference is built for baseline x86-64 where a*b+c is two roundings).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DESC_NUMEL 768
#define ICOS_NFACES 20
#define NHIST 4

/* Same layout as CPUSIFT::Keypoint (Include/cSIFT3D.h:52-70); sizeof == 176 on LP64. */
typedef struct {
    float x, y, z;
    float scale;
    int octave, level;
    float rx, ry, rz;
    float win[3];
    float eigvalue[3];
    float eigvector[9];
    float Rotation[9];
    float str_tensor[9];
    float* desc;
} orc_keypoint;

typedef struct {
    float v[3][3];
    int idx[3];
} orc_tri;

int orc_sizeof_keypoint(void) { return (int)sizeof(orc_keypoint); }

/* ---------------------------------------------------------------------------------------------
 * Scale-space constants.  Src/cSIFT3D.cc:270-287 (sigma table), :299 (base blur).
 * sig[0] = base blur of the input, sig[1..L+2] = incremental blur of level i.  Returns L+3.
 * ------------------------------------------------------------------------------------------- */
int orc_sigmas(int num_kp_levels, float sigma_default, float sigma_n_default, float* sig) {
    float k = (float)pow(2.0, 1.0 / num_kp_levels);
    float base = (float)(sigma_default * pow(2.0, -1.0 / 3.0));
    sig[0] = sqrtf(base * base - sigma_n_default * sigma_n_default);
    for (int i = 1; i < num_kp_levels + 3; i++) {
        float sig_prev = (float)(pow((double)k, (double)(i - 1)) * base);
        float sig_total = sig_prev * k;
        sig[i] = sqrtf(sig_total * sig_total - sig_prev * sig_prev);
    }
    return num_kp_levels + 3;
}

/* Level scale metadata.  Src/cUtil.cc:182,209-210. */
float orc_level_scale(int octave, int s, int num_kp_levels, float sigma_default) {
    double sigma0 = sigma_default * pow(2.0, -1.0 / 3.0);
    double scale_factor = pow(2.0, octave + (double)s / num_kp_levels);
    return (float)(scale_factor * sigma0);
}

/* Octave count.  Src/cSIFT3D.cc:254-255. */
int orc_num_octaves(int nx, int ny, int nz) {
    int m = nx < ny ? nx : ny;
    m = m < nz ? m : nz;
    return (int)log2f((float)m) - 3 + 1;
}

/* 1-D Gaussian taps.  Src/cSIFT3D.cc:541-572.  w must hold 2*hw+1 floats; returns hw. */
int orc_gauss_kernel(float sigma, float* w) {
    sigma = sigma > 0 ? sigma : 0;
    int t = (int)ceil(sigma * 3.0);
    const int hw = sigma > 0 ? (t > 1 ? t : 1) : 1;
    const int width = 2 * hw + 1;
    float acc = 0;
    for (int i = 0; i < width; i++) {
        float x = (float)(i - hw);
        x = (float)((double)x / ((double)sigma + DBL_EPSILON));
        w[i] = (float)exp(-0.5 * (double)x * (double)x);
        acc += w[i];
    }
    for (int i = 0; i < width; i++) w[i] /= acc;
    return hw;
}

/* Flat read exactly as TexImage::GetImageDataWithIdx (Include/Util/cTexImage.h:56-58) does it:
 * no bounds check, so an out-of-row index lands in the neighbouring row.  Indices outside the
 * whole buffer (true UB in the reference) read as the clamped end element. */
static inline float rd(const float* p, long i, long total) {
    if (i < 0) i = 0;
    if (i >= total) i = total - 1;
    return p[i];
}

/* One separable pass along `axis` (0=x,1=y,2=z) of an x-fastest volume.
 * Src/cSIFT3D.cc:624-790.  The reference runs Y and Z by transposing to X (:609-617); per output
 * the tap order and arithmetic are unchanged, so no transpose is needed (App. B Q6). */
void orc_blur_axis(const float* src, float* dst, int nx, int ny, int nz, int axis, const float* w, int hw) {
    const int dims[3] = {nx, ny, nz};
    const long strides[3] = {1, nx, (long)nx * ny};
    const int n = dims[axis];
    const long st = strides[axis];
    const long total = (long)nx * ny * nz;
    const int dim_end = n - 1;
    const int lo_int = hw, hi_int = n - 1 - (hw + 1); /* interior: [hw, n-hw-2]   :660-661 */
    const float conv_eps = 0.1f;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++) {
                const int c3[3] = {x, y, z};
                const int p = c3[axis];
                const long base = (long)x + (long)y * nx + (long)z * nx * ny - (long)p * st;
                float acc = 0.0f;
                if (p >= lo_int && p <= hi_int) {
                    /* interior, :682-719: frac == 0 so the sample is in[c] exactly */
                    for (int d = -hw; d <= hw; d++) {
                        float lo = src[base + (long)(p - d) * st];
                        float hi = src[base + (long)(p - d + 1) * st];
                        acc += w[d + hw] * ((1.0f - 0.0f) * lo + 0.0f * hi);
                    }
                } else {
                    /* boundary, :722-788 */
                    for (int d = -hw; d <= hw; d++) {
                        float c = (float)p - (float)d;
                        if (c < 0)
                            c = -1 * c;
                        else if (c >= dim_end)
                            c = (float)(2 * dim_end) - c - conv_eps;
                        int il = (int)c;
                        float frac = c - (float)il;
                        float lo = rd(src, base + (long)il * st, total);
                        float hi = rd(src, base + (long)(il + 1) * st, total);
                        acc += w[d + hw] * ((1.0f - frac) * lo + frac * hi);
                    }
                }
                dst[(long)x + (long)y * nx + (long)z * nx * ny] = acc;
            }
}

/* GaussianSmooth_3D.  Src/cSIFT3D.cc:535-622: X, then Y, then Z. */
void orc_gaussian_smooth(const float* src, float* dst, int nx, int ny, int nz, float sigma) {
    float w[64];
    int hw = orc_gauss_kernel(sigma, w);
    size_t n = (size_t)nx * ny * nz;
    float* tmp = (float*)malloc(n * sizeof(float));
    orc_blur_axis(src, dst, nx, ny, nz, 0, w, hw);
    orc_blur_axis(dst, tmp, nx, ny, nz, 1, w, hw);
    orc_blur_axis(tmp, dst, nx, ny, nz, 2, w, hw);
    free(tmp);
}

/* DownSample_3D.  Src/cSIFT3D.cc:506-533: dst(n,m,k) = src(2n,2m,2k), dst dims = src dims / 2. */
void orc_downsample(const float* src, int nx, int ny, int nz, float* dst) {
    int dx = nx / 2, dy = ny / 2, dz = nz / 2;
    for (int k = 0; k < dz; k++)
        for (int m = 0; m < dy; m++)
            for (int n = 0; n < dx; n++)
                dst[(size_t)n + (size_t)m * dx + (size_t)k * dx * dy] =
                    src[(size_t)(2 * n) + (size_t)(2 * m) * nx + (size_t)(2 * k) * nx * ny];
}

/* Sub.  Src/cSIFT3D.cc:849-882: dog = (cur - prev) * (-1). */
void orc_sub(const float* prev, const float* cur, float* dog, size_t n) {
    for (size_t i = 0; i < n; i++) dog[i] = (cur[i] - prev[i]) * (-1);
}

/* im_max_abs.  Src/cUtil.cc:587-605. */
float orc_max_abs(const float* p, size_t n) {
    float max = 0.0f;
    for (size_t i = 0; i < n; i++) {
        float tmp = p[i];
        max = (fabs(tmp) > max) ? (float)fabs(tmp) : max;
    }
    return max;
}

/* data_scale.  Src/cUtil.cc:536-564: in-place divide by global max|v|. */
void orc_data_scale(float* p, size_t n) {
    float max = orc_max_abs(p, n);
    for (size_t i = 0; i < n; i++) p[i] /= max;
}

/* Detection on one DoG level.  Src/cSIFT3D.cc:384-417 and IsExtrema_neighbor :884-911.
 * out_xyz receives up to cap (x,y,z) int triples in raster order; returns the total count. */
int orc_detect_level(const float* prev, const float* cur, const float* next, int nx, int ny, int nz,
                     float peak_thresh, int* out_xyz, int cap, float* thres_out) {
    size_t n = (size_t)nx * ny * nz;
    float thres = peak_thresh * orc_max_abs(cur, n);
    if (thres_out) *thres_out = thres;
    int cnt = 0;
    const long ys = nx, zs = (long)nx * ny;
    for (int z = 1; z < nz - 1; z++)
        for (int y = 1; y < ny - 1; y++)
            for (int x = 1; x < nx - 1; x++) {
                long i = x + y * ys + z * zs;
                float val = cur[i];
                if (!(val > thres || val < -thres)) continue;
                float t0 = prev[i], t1 = cur[i - 1], t2 = cur[i + 1], t3 = cur[i + ys], t4 = cur[i - ys],
                      t5 = cur[i + zs], t6 = cur[i - zs], t7 = next[i];
                int mn = val < t0 && val < t1 && val < t2 && val < t3 && val < t4 && val < t5 && val < t6 && val < t7;
                int mx = val > t0 && val > t1 && val > t2 && val > t3 && val > t4 && val > t5 && val > t6 && val > t7;
                if (mn || mx) {
                    if (cnt < cap) {
                        out_xyz[3 * cnt] = x; out_xyz[3 * cnt + 1] = y; out_xyz[3 * cnt + 2] = z;
                    }
                    cnt++;
                }
            }
    return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * Icosahedron mesh.  Src/cUtil.cc:19-55 (tables), :113-175 (Initialize_geometry).
 * ------------------------------------------------------------------------------------------- */
static const double orc_gr = 1.6180339887;
static const int orc_faces[60] = {0, 1, 8, 0, 8, 4, 0, 4, 5, 0, 5, 9, 0, 9, 1, 1, 6, 8, 8, 6, 10, 8, 10, 4, 4, 10, 2,
                                  4, 2, 5, 5, 2, 11, 5, 11, 9, 9, 11, 7, 9, 7, 1, 1, 7, 6, 3, 6, 7, 3, 7, 11, 3, 11, 2,
                                  3, 2, 10, 3, 10, 6};

void orc_mesh(orc_tri* tri) {
    const double vert[36] = {0, 1, orc_gr, 0, -1, orc_gr, 0, 1, -orc_gr, 0, -1, -orc_gr, 1, orc_gr, 0, -1, orc_gr, 0,
                             1, -orc_gr, 0, -1, -orc_gr, 0, orc_gr, 0, 1, -orc_gr, 0, 1, orc_gr, 0, -1, -orc_gr, 0, -1};
    for (int i = 0; i < ICOS_NFACES; i++) {
        orc_tri* t = tri + i;
        for (int j = 0; j < 3; j++) {
            t->idx[j] = orc_faces[i * 3 + j];
            float* v = t->v[j];
            v[0] = (float)vert[t->idx[j] * 3 + 0];
            v[1] = (float)vert[t->idx[j] * 3 + 1];
            v[2] = (float)vert[t->idx[j] * 3 + 2];
            double mag = (double)sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); /* float norm, :76 */
            double s = 1.0 / mag;
            v[0] = (float)(v[0] * s); v[1] = (float)(v[1] * s); v[2] = (float)(v[2] * s);
        }
        float a[3], b[3], n[3];
        for (int c = 0; c < 3; c++) {
            a[c] = t->v[2][c] - t->v[1][c];
            b[c] = t->v[1][c] - t->v[0][c];
        }
        n[0] = a[1] * b[2] - a[2] * b[1];
        n[1] = a[2] * b[0] - a[0] * b[2];
        n[2] = a[0] * b[1] - a[1] * b[0];
        if (n[0] * t->v[0][0] + n[1] * t->v[0][1] + n[2] * t->v[0][2] < 0) {
            /* swaps the vertices but NOT idx (App. B Q13) */
            for (int c = 0; c < 3; c++) {
                float tmp = t->v[0][c];
                t->v[0][c] = t->v[1][c];
                t->v[1][c] = tmp;
            }
        }
    }
}

void orc_mesh_flat(float* v, int* idx) {
    orc_tri tri[ICOS_NFACES];
    orc_mesh(tri);
    for (int i = 0; i < ICOS_NFACES; i++)
        for (int j = 0; j < 3; j++) {
            for (int c = 0; c < 3; c++) v[(i * 3 + j) * 3 + c] = tri[i].v[j][c];
            idx[i * 3 + j] = tri[i].idx[j];
        }
}

/* ---------------------------------------------------------------------------------------------
 * Symmetric 3x3 eigen-decomposition in double (cyclic Jacobi).  The reference calls
 * Eigen::EigenSolver<Matrix3d> (Src/cSIFT3D.cc:1016-1029; vendored Eigen 3.3.7,
 * 3party/Eigen/Eigen/src/Eigenvalues/EigenSolver.h:379) — a general real solver applied to a
 * symmetric matrix; eigenvalues agree to ~1e-15 relative and eigenvectors up to sign, which the
 * reference then fixes itself (:1089-1108).  Columns of V are unit eigenvectors.
 * ------------------------------------------------------------------------------------------- */
static void orc_eig3(const double A[9], double val[3], double V[9]) {
    double a[3][3] = {{A[0], A[1], A[2]}, {A[3], A[4], A[5]}, {A[6], A[7], A[8]}};
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) {
        val[i] = a[i][i];
        double nn = sqrt(v[0][i] * v[0][i] + v[1][i] * v[1][i] + v[2][i] * v[2][i]);
        for (int k = 0; k < 3; k++) V[k * 3 + i] = v[k][i] / nn;
    }
}

/* Window bounds shared by orientation and descriptor.  Src/cSIFT3D.cc:939-955, :1182-1198. */
static void orc_window(float c, float r_over_u, int n, int* s, int* e) {
    int a = (int)floorf(c - r_over_u);
    a = a > 1 ? a : 1;
    int b = (int)ceilf(c + r_over_u);
    b = b < (n - 2) ? b : n - 1 - 1;
    *s = a;
    *e = b;
}

/* Initialize_Keypoint.  Src/cUtil.cc:449-463. */
static void orc_init_keypoint(orc_keypoint* kp) {
    kp->rx = kp->ry = kp->rz = -1.0f;
    for (int i = 0; i < 9; i++) {
        kp->Rotation[i] = 0;
        kp->str_tensor[i] = 0;
    }
}

/* Assign_Orientation_Imp.  Src/cSIFT3D.cc:913-1138.  Returns 1 / -1 / -2 / -3.
 * `unit` = ux = uy = uz of the level (2^octave). */
int orc_orient(orc_keypoint* kp, const float* g, int nx, int ny, int nz, float unit, float max_eig_ratio,
               float corner_thresh) {
    const float ori_sig_fctr = 1.5f, ori_rad_fctr = 3.0f, ori_grad_thresh = 1E-10f;
    const float sigma = ori_sig_fctr * kp->scale; /* :442 */
    const float win_radius = sigma * ori_rad_fctr;
    const float cx = kp->x, cy = kp->y, cz = kp->z;
    const float u = unit;
    const long ys = nx, zs = (long)nx * ny;
    float* T = kp->str_tensor;
    float wx = 0, wy = 0, wz = 0;
    int xs, xe, y0, y1, z0, z1;
    orc_window(cx, win_radius / u, nx, &xs, &xe);
    orc_window(cy, win_radius / u, ny, &y0, &y1);
    orc_window(cz, win_radius / u, nz, &z0, &z1);
    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++)
            for (int x = xs; x <= xe; x++) {
                float dx = ((float)x - cx) * u, dy = ((float)y - cy) * u, dz = ((float)z - cz) * u;
                float sq = dx * dx + dy * dy + dz * dz;
                if (sq > win_radius * win_radius) continue;
                float weight = expf((float)(-0.5 * sq / (sigma * sigma)));
                long i = x + y * ys + z * zs;
                float vx = (float)(0.5 * (g[i + 1] - g[i - 1]));
                float vy = (float)(0.5 * (g[i + ys] - g[i - ys]));
                float vz = (float)(0.5 * (g[i + zs] - g[i - zs]));
                vx *= 1.0f / u; vy *= 1.0f / u; vz *= 1.0f / u;
                T[0] += vx * vx * weight;
                T[1] += vx * vy * weight;
                T[2] += vx * vz * weight;
                T[4] += vy * vy * weight;
                T[5] += vy * vz * weight;
                T[8] += vz * vz * weight;
                wx += vx * weight; wy += vy * weight; wz += vz * weight;
            }
    T[3] = T[1]; T[6] = T[2]; T[7] = T[5];
    kp->win[0] = wx; kp->win[1] = wy; kp->win[2] = wz;
    if (wx * wx + wy * wy + wz * wz < ori_grad_thresh) return -1;

    double A[9], val[3], V[9];
    for (int i = 0; i < 9; i++) A[i] = T[i];
    orc_eig3(A, val, V);
    struct { float val; float vec[3]; } EV[3], tmp;
    for (int i = 0; i < 3; i++) {
        EV[i].val = (float)val[i];
        for (int k = 0; k < 3; k++) EV[i].vec[k] = (float)V[k * 3 + i];
    }
    for (int i = 1; i < 3; i++) /* ascending by value, :1050 */
        for (int j = i; j > 0 && EV[j].val < EV[j - 1].val; j--) {
            tmp = EV[j]; EV[j] = EV[j - 1]; EV[j - 1] = tmp;
        }
    for (int i = 0; i < 3; i++) {
        kp->eigvalue[i] = EV[i].val;
        for (int k = 0; k < 3; k++) kp->eigvector[i * 3 + k] = EV[i].vec[k];
    }
    if (fabs(EV[0].val / EV[1].val) > max_eig_ratio || fabs(EV[1].val / EV[2].val) > max_eig_ratio) return -2;
    /* DistinctEig, :1140-1150 */
    if (fabs(EV[0].val - EV[1].val) < DBL_EPSILON || fabs(EV[0].val - EV[2].val) < DBL_EPSILON ||
        fabs(EV[2].val - EV[1].val) < DBL_EPSILON)
        return -2;

    float d_NORM = sqrtf(wx * wx + wy * wy + wz * wz);
    float corner_score = FLT_MAX;
    for (int i = 2; i > 0; i--) {
        float ex = EV[i].vec[0], ey = EV[i].vec[1], ez = EV[i].vec[2];
        float d = ex * wx + ey * wy + ez * wz;
        float q_NORM = sqrtf(ex * ex + ey * ey + ez * ez);
        float cos_ang = d / (d_NORM * q_NORM);
        float abs_cos_ang = (float)fabs(cos_ang);
        corner_score = corner_score < abs_cos_ang ? corner_score : abs_cos_ang;
        float sgn = d > 0.0 ? 1.0f : -1.0f;
        EV[i].vec[0] *= sgn; EV[i].vec[1] *= sgn; EV[i].vec[2] *= sgn;
    }
    if (corner_score < corner_thresh) return -3;
    const float* v1 = EV[2].vec;
    const float* v2 = EV[1].vec;
    float vr[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    float* R = kp->Rotation;
    R[0] = v1[0]; R[1] = v2[0]; R[2] = vr[0];
    R[3] = v1[1]; R[4] = v2[1]; R[5] = vr[1];
    R[6] = v1[2]; R[7] = v2[2]; R[8] = vr[2];
    return 1;
}

/* cart2bary.  Src/cSIFT3D.cc:1592-1637 (Moller-Trumbore from the origin along `cart`). */
static int orc_cart2bary(const float* cart, const orc_tri* tri, float* bary, float* k) {
    const float(*v)[3] = tri->v;
    float e1[3], e2[3], t[3], p[3], q[3];
    for (int c = 0; c < 3; c++) {
        e1[c] = v[1][c] - v[0][c];
        e2[c] = v[2][c] - v[0][c];
        t[c] = (float)(v[0][c] * (-1.0));
    }
    p[0] = cart[1] * e2[2] - cart[2] * e2[1];
    p[1] = cart[2] * e2[0] - cart[0] * e2[2];
    p[2] = cart[0] * e2[1] - cart[1] * e2[0];
    q[0] = t[1] * e1[2] - t[2] * e1[1];
    q[1] = t[2] * e1[0] - t[0] * e1[2];
    q[2] = t[0] * e1[1] - t[1] * e1[0];
    float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    const float bary_eps = (float)(FLT_EPSILON * 1E1);
    if (fabsf(det) < bary_eps) return -1;
    float det_inv = (float)(1.0 / det);
    bary[1] = det_inv * (p[0] * t[0] + p[1] * t[1] + p[2] * t[2]);
    bary[2] = det_inv * (cart[0] * q[0] + cart[1] * q[1] + cart[2] * q[2]);
    bary[0] = 1 - bary[1] - bary[2];
    *k = det_inv * (q[0] * e2[0] + q[1] * e2[1] + q[2] * e2[2]);
    return 0;
}

/* Check_intersect_faces.  Src/cSIFT3D.cc:1542-1573: first passing face wins (App. B Q14). */
static int orc_intersect(const orc_tri* mesh, const float* grad, float* bary) {
    const float bary_eps = (float)(FLT_EPSILON * 1E1);
    if (grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2] < bary_eps) return -1;
    for (int i = 0; i < ICOS_NFACES; i++) {
        float k;
        if (orc_cart2bary(grad, mesh + i, bary, &k) < 0) continue;
        if (bary[0] < -bary_eps || bary[1] < -bary_eps || bary[2] < -bary_eps || k < 0) continue;
        return i;
    }
    return -1;
}

/* normailize_desc.  Src/cSIFT3D.cc:1639-1656. */
static void orc_normalize_desc(float* desc) {
    float norm = 0.0f;
    for (int i = 0; i < DESC_NUMEL; i++) norm += desc[i] * desc[i];
    norm = (float)(sqrtf(norm) + DBL_EPSILON);
    for (int i = 0; i < DESC_NUMEL; i++) {
        float norm_inv = (float)(1.0 / norm);
        desc[i] *= norm_inv;
    }
}

/* Extract_Descriptor_Imp + Trilinear_interpolation_over_desc_debug.
 * Src/cSIFT3D.cc:1152-1381, :1450-1540.  kp->Rotation is transposed in place (:1214, Q11);
 * desc (768 floats) must be zeroed by the caller (calloc at :486). */
void orc_describe(orc_keypoint* kp, const float* g, int nx, int ny, int nz, float unit, float* desc) {
    static orc_tri mesh[ICOS_NFACES];
    static int mesh_ready = 0;
#pragma omp critical(orc_mesh_init)
    if (!mesh_ready) { orc_mesh(mesh); mesh_ready = 1; }

    const float desc_sig_fctr = 7.071067812f, desc_rad_fctr = 2.0f;
    const float trunc_thresh = (float)(0.2 * 128 / DESC_NUMEL);
    const float sigma = kp->scale * desc_sig_fctr;
    const float win_radius = desc_rad_fctr * sigma;
    const float desc_hw = (float)(win_radius / sqrt(2.0));
    const float desc_width = 2.0f * desc_hw;
    const float desc_bin_fctr = (float)NHIST / desc_width;
    const float coord_factor = (float)pow(2.0, kp->octave);
    const float cx = kp->x, cy = kp->y, cz = kp->z, u = unit;
    const long ys = nx, zs = (long)nx * ny;
    int xs, xe, y0, y1, z0, z1;
    orc_window(cx, win_radius / u, nx, &xs, &xe);
    orc_window(cy, win_radius / u, ny, &y0, &y1);
    orc_window(cz, win_radius / u, nz, &z0, &z1);
    float* R = kp->Rotation;
    { float t; t = R[1]; R[1] = R[3]; R[3] = t; t = R[2]; R[2] = R[6]; R[6] = t; t = R[5]; R[5] = R[7]; R[7] = t; }

    for (int z = z0; z <= z1; z++)
        for (int y = y0; y <= y1; y++)
            for (int x = xs; x <= xe; x++) {
                float dx = ((float)x - cx) * u, dy = ((float)y - cy) * u, dz = ((float)z - cz) * u;
                float sq = dx * dx + dy * dy + dz * dz;
                if (sq > win_radius * win_radius) continue;
                float vb[3];
                vb[0] = (R[0] * dx + R[1] * dy + R[2] * dz + desc_hw) * desc_bin_fctr;
                vb[1] = (R[3] * dx + R[4] * dy + R[5] * dz + desc_hw) * desc_bin_fctr;
                vb[2] = (R[6] * dx + R[7] * dy + R[8] * dz + desc_hw) * desc_bin_fctr;
                vb[0] -= 0.5f; vb[1] -= 0.5f; vb[2] -= 0.5f;
                if (vb[0] <= -0.5f || vb[1] <= -0.5f || vb[2] <= -0.5f || vb[0] >= 3.5f || vb[1] >= 3.5f ||
                    vb[2] >= 3.5f)
                    continue;
                float weight = expf(-0.5f * sq / (sigma * sigma));
                long i = x + y * ys + z * zs;
                float gx = (float)(0.5 * (g[i + 1] - g[i - 1]));
                float gy = (float)(0.5 * (g[i + ys] - g[i - ys]));
                float gz = (float)(0.5 * (g[i + zs] - g[i - zs]));
                gx *= 1.0f / u; gy *= 1.0f / u; gz *= 1.0f / u;
                gx = (float)(gx * (double)weight); gy = (float)(gy * (double)weight); gz = (float)(gz * (double)weight);
                float gr[3];
                gr[0] = R[0] * gx + R[1] * gy + R[2] * gz;
                gr[1] = R[3] * gx + R[4] * gy + R[5] * gz;
                gr[2] = R[6] * gx + R[7] * gy + R[8] * gz;

                /* trilinear + icosahedron binning, :1450-1540 */
                float dv[3] = {vb[0] - floorf(vb[0]), vb[1] - floorf(vb[1]), vb[2] - floorf(vb[2])};
                float bary[3];
                int face = orc_intersect(mesh, gr, bary);
                if (face < 0) continue;
                float mag = sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]);
                for (int ddx = 0; ddx < 2; ddx++)
                    for (int ddy = 0; ddy < 2; ddy++)
                        for (int ddz = 0; ddz < 2; ddz++) {
                            int bx = (int)vb[0] + ddx, by = (int)vb[1] + ddy, bz = (int)vb[2] + ddz; /* Q12 */
                            if (bx < 0 || by < 0 || bz < 0 || bx >= NHIST || by >= NHIST || bz >= NHIST) continue;
                            int hist = bx + by * NHIST + bz * NHIST * NHIST;
                            float wt = (float)(((ddx == 0) ? (1.0 - dv[0]) : dv[0]) * ((ddy == 0) ? (1.0 - dv[1]) : dv[1]) *
                                               ((ddz == 0) ? (1.0 - dv[2]) : dv[2]));
                            desc[hist * 12 + mesh[face].idx[0]] += mag * wt * bary[0];
                            desc[hist * 12 + mesh[face].idx[1]] += mag * wt * bary[1];
                            desc[hist * 12 + mesh[face].idx[2]] += mag * wt * bary[2];
                        }
            }
    orc_normalize_desc(desc);
    for (int i = 0; i < DESC_NUMEL; i++) desc[i] = desc[i] < trunc_thresh ? desc[i] : trunc_thresh;
    orc_normalize_desc(desc);
    kp->rx = kp->x * coord_factor;
    kp->ry = kp->y * coord_factor;
    kp->rz = kp->z * coord_factor;
}

/* ---------------------------------------------------------------------------------------------
 * Whole extraction, CSIFT3D ctor + KpSiftAlgorithm.  Src/cSIFT3D.cc:146-163, :165-235.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int nx, ny, nz, L, noct;
    float sigma_default, sigma_n, peak, eig, corner;
    float* input;     /* normalised copy */
    float** gss;      /* noct*(L+3) */
    float** dog;      /* noct*(L+2) */
    int* dims;        /* noct*3 */
    orc_keypoint* extre; int n_extre; int* ret;   /* raw detections after orientation */
    orc_keypoint* kps; int n_kps; float* desc;     /* survivors */
} orc_ctx;

void* orc_create(const float* vol, int nx, int ny, int nz, int L, float sigma, float sigma_n, float peak, float eig,
                 float corner) {
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->nx = nx; c->ny = ny; c->nz = nz; c->L = L;
    c->sigma_default = sigma; c->sigma_n = sigma_n; c->peak = peak; c->eig = eig; c->corner = corner;
    size_t n = (size_t)nx * ny * nz;
    c->input = (float*)malloc(n * sizeof(float));
    memcpy(c->input, vol, n * sizeof(float));
    orc_data_scale(c->input, n);
    return c;
}

void orc_run(void* h) {
    orc_ctx* c = (orc_ctx*)h;
    const int L = c->L, G = L + 3, D = L + 2;
    c->noct = orc_num_octaves(c->nx, c->ny, c->nz);
    c->gss = (float**)calloc((size_t)c->noct * G, sizeof(float*));
    c->dog = (float**)calloc((size_t)c->noct * D, sizeof(float*));
    c->dims = (int*)calloc((size_t)c->noct * 3, sizeof(int));
    float sig[16];
    orc_sigmas(L, c->sigma_default, c->sigma_n, sig);
    int nx = c->nx, ny = c->ny, nz = c->nz;
    for (int o = 0; o < c->noct; o++) {
        c->dims[o * 3] = nx; c->dims[o * 3 + 1] = ny; c->dims[o * 3 + 2] = nz;
        size_t n = (size_t)nx * ny * nz;
        for (int i = 0; i < G; i++) c->gss[o * G + i] = (float*)calloc(n ? n : 1, sizeof(float));
        for (int i = 0; i < D; i++) c->dog[o * D + i] = (float*)calloc(n ? n : 1, sizeof(float));
        nx /= 2; ny /= 2; nz /= 2;
    }
    /* Build_Gaussian_Scale_Space :268-319 */
    for (int o = 0; o < c->noct; o++) {
        const int* d = c->dims + o * 3;
        for (int i = 0; i < G; i++) {
            if (o == 0 && i == 0)
                orc_gaussian_smooth(c->input, c->gss[0], d[0], d[1], d[2], sig[0]);
            else if (i == 0)
                orc_downsample(c->gss[(o - 1) * G + L], d[-3], d[-2], d[-1], c->gss[o * G]);
            else
                orc_gaussian_smooth(c->gss[o * G + i - 1], c->gss[o * G + i], d[0], d[1], d[2], sig[i]);
        }
    }
    /* Build_DOG_Scale_Space :346-360 */
    for (int o = 0; o < c->noct; o++) {
        const int* d = c->dims + o * 3;
        size_t n = (size_t)d[0] * d[1] * d[2];
        for (int i = 1; i < G; i++) orc_sub(c->gss[o * G + i - 1], c->gss[o * G + i], c->dog[o * D + i - 1], n);
    }
    /* Detect_KeyPoints :362-425 */
    int cap = 1 << 16, cnt = 0;
    c->extre = (orc_keypoint*)malloc(sizeof(orc_keypoint) * cap);
    for (int o = 0; o < c->noct; o++) {
        const int* d = c->dims + o * 3;
        for (int i = 1; i < D - 1; i++) {
            size_t n = (size_t)d[0] * d[1] * d[2];
            int* xyz = (int*)malloc(sizeof(int) * 3 * (n ? n : 1));
            int m = orc_detect_level(c->dog[o * D + i - 1], c->dog[o * D + i], c->dog[o * D + i + 1], d[0], d[1], d[2],
                                     c->peak, xyz, (int)n, NULL);
            while (cnt + m > cap) {
                cap *= 2;
                c->extre = (orc_keypoint*)realloc(c->extre, sizeof(orc_keypoint) * cap);
            }
            for (int k = 0; k < m; k++) {
                orc_keypoint* kp = c->extre + cnt++;
                memset(kp, 0, sizeof(*kp));
                kp->x = (float)xyz[3 * k]; kp->y = (float)xyz[3 * k + 1]; kp->z = (float)xyz[3 * k + 2];
                kp->octave = o; kp->level = i;
                kp->scale = orc_level_scale(o, i, L, c->sigma_default);
                orc_init_keypoint(kp);
            }
            free(xyz);
        }
    }
    c->n_extre = cnt;
    /* Assign_Orientation :427-482 */
    c->ret = (int*)malloc(sizeof(int) * (cnt ? cnt : 1));
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < cnt; i++) {
        orc_keypoint* kp = c->extre + i;
        const int* d = c->dims + kp->octave * 3;
        float unit = (float)(1 << kp->octave);
        int res = orc_orient(kp, c->gss[kp->octave * G + kp->level], d[0], d[1], d[2], unit, c->eig, c->corner);
        c->ret[i] = res;
        if (res < 1) kp->x = kp->y = kp->z = -1.0f;
    }
    int nk = 0;
    for (int i = 0; i < cnt; i++) nk += !(c->extre[i].x < 0);
    c->n_kps = nk;
    c->kps = (orc_keypoint*)malloc(sizeof(orc_keypoint) * (nk ? nk : 1));
    c->desc = (float*)calloc((size_t)DESC_NUMEL * (nk ? nk : 1), sizeof(float));
    nk = 0;
    for (int i = 0; i < cnt; i++)
        if (!(c->extre[i].x < 0)) c->kps[nk++] = c->extre[i];
    /* Extract_Description :484-502 */
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < c->n_kps; i++) {
        orc_keypoint* kp = c->kps + i;
        const int* d = c->dims + kp->octave * 3;
        kp->desc = c->desc + (size_t)i * DESC_NUMEL;
        orc_describe(kp, c->gss[kp->octave * G + kp->level], d[0], d[1], d[2], (float)(1 << kp->octave), kp->desc);
    }
}

int orc_noct(void* h) { return ((orc_ctx*)h)->noct; }
int orc_level_dims(void* h, int o, int* d) {
    orc_ctx* c = (orc_ctx*)h;
    if (o < 0 || o >= c->noct) return -1;
    d[0] = c->dims[o * 3]; d[1] = c->dims[o * 3 + 1]; d[2] = c->dims[o * 3 + 2];
    return 0;
}
int orc_copy_level(void* h, int which, int idx, float* out) {
    orc_ctx* c = (orc_ctx*)h;
    int per = which == 0 ? c->L + 3 : c->L + 2;
    if (idx < 0 || idx >= c->noct * per) return -1;
    const int* d = c->dims + (idx / per) * 3;
    memcpy(out, which == 0 ? c->gss[idx] : c->dog[idx], sizeof(float) * (size_t)d[0] * d[1] * d[2]);
    return 0;
}
int orc_copy_input(void* h, float* out) {
    orc_ctx* c = (orc_ctx*)h;
    memcpy(out, c->input, sizeof(float) * (size_t)c->nx * c->ny * c->nz);
    return 0;
}
int orc_num_extrema(void* h) { return ((orc_ctx*)h)->n_extre; }
int orc_copy_extrema(void* h, void* out, int* ret) {
    orc_ctx* c = (orc_ctx*)h;
    memcpy(out, c->extre, sizeof(orc_keypoint) * c->n_extre);
    if (ret) memcpy(ret, c->ret, sizeof(int) * c->n_extre);
    return c->n_extre;
}
int orc_num_keypoints(void* h) { return ((orc_ctx*)h)->n_kps; }
int orc_copy_keypoints(void* h, void* kp, float* desc) {
    orc_ctx* c = (orc_ctx*)h;
    if (kp) memcpy(kp, c->kps, sizeof(orc_keypoint) * c->n_kps);
    if (desc) memcpy(desc, c->desc, sizeof(float) * DESC_NUMEL * (size_t)c->n_kps);
    return c->n_kps;
}
void orc_destroy(void* h) {
    orc_ctx* c = (orc_ctx*)h;
    if (c->gss) for (int i = 0; i < c->noct * (c->L + 3); i++) free(c->gss[i]);
    if (c->dog) for (int i = 0; i < c->noct * (c->L + 2); i++) free(c->dog[i]);
    free(c->gss); free(c->dog); free(c->dims); free(c->input);
    free(c->extre); free(c->ret); free(c->kps); free(c->desc);
    free(c);
}

/* ---------------------------------------------------------------------------------------------
 * Matcher.  Src/cMatcher.cc.
 * ------------------------------------------------------------------------------------------- */

/* calMatches :40-79 with KP_squareSum :17-23 (float product, double running sum). */
void orc_cal_matches(const float* q, int nq, const float* db, int nd, const int* mask, float* gDist, float* sDist,
                     int* gIdx, int* sIdx) {
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < nq; i++) {
        if (mask && mask[i] == 0) {
            gIdx[i] = -1;
            continue;
        }
        const float* a = q + (size_t)i * DESC_NUMEL;
        double d1 = FLT_MIN, d2 = FLT_MIN;
        int i1 = -1, i2 = -1;
        for (int j = 0; j < nd; j++) {
            const float* b = db + (size_t)j * DESC_NUMEL;
            double s = 0;
            for (int k = 0; k < DESC_NUMEL; k++) s += a[k] * b[k];
            if (s > d1) { d2 = d1; i2 = i1; d1 = s; i1 = j; }
            else if (s > d2) { d2 = s; i2 = j; }
        }
        gDist[i] = (float)(2 - 2 * d1);
        sDist[i] = (float)(2 - 2 * d2);
        gIdx[i] = i1;
        sIdx[i] = i2;
    }
}

/* filter :81-97 */
void orc_filter(int* gIdx, const float* gDist, const float* sDist, int n, double thr) {
    const double thresSquare = thr * thr;
    for (int i = 0; i < n; i++) {
        if (gIdx[i] < 0) continue;
        float d1 = gDist[i], d2 = sDist[i];
        if (d1 / d2 >= thresSquare) gIdx[i] *= -1;
    }
}

/* bijectMatchBase :146-215.  type 1 inject / 2 biject / 3 enhanced.  All outputs caller-allocated:
 * forward arrays length n_ref, reverse arrays length n_tar (may be NULL), pairs length n_ref.
 * Returns the number of matched pairs. */
int orc_match(int type, const float* ref, int n_ref, const float* tar, int n_tar, double thr, int* gIdx, float* gDist,
              int* sIdx, float* sDist, int* gIdx2, float* gDist2, int* sIdx2, float* sDist2, int* pair_ref,
              int* pair_tar) {
    int* own[8] = {0};
    #define ORC_OWN(p, T, n, slot) if (!(p)) { p = (T*)malloc(sizeof(T) * ((n) ? (n) : 1)); own[slot] = (int*)(p); }
    ORC_OWN(gIdx, int, n_ref, 0) ORC_OWN(gDist, float, n_ref, 1) ORC_OWN(sIdx, int, n_ref, 2) ORC_OWN(sDist, float, n_ref, 3)
    ORC_OWN(gIdx2, int, n_tar, 4) ORC_OWN(gDist2, float, n_tar, 5) ORC_OWN(sIdx2, int, n_tar, 6) ORC_OWN(sDist2, float, n_tar, 7)
    for (int i = 0; i < n_ref; i++) { gIdx[i] = -1; sIdx[i] = -1; gDist[i] = 0; sDist[i] = 0; }
    for (int i = 0; i < n_tar; i++) { gIdx2[i] = -1; sIdx2[i] = -1; gDist2[i] = 0; sDist2[i] = 0; }
    orc_cal_matches(ref, n_ref, tar, n_tar, NULL, gDist, sDist, gIdx, sIdx);
    orc_filter(gIdx, gDist, sDist, n_ref, thr);
    if (type != 1) {
        const int maskThres = (type == 2 ? 0 : 1);
        int* counts = (int*)calloc(n_tar ? n_tar : 1, sizeof(int));
        for (int i = 0; i < n_ref; i++) /* countMatched :114-120 */
            if (gIdx[i] >= 0) counts[gIdx[i]] += 1;
        for (int j = 0; j < n_tar; j++) counts[j] = counts[j] > maskThres ? 1 : 0; /* toMask :122-131 */
        orc_cal_matches(tar, n_tar, ref, n_ref, counts, gDist2, sDist2, gIdx2, sIdx2);
        orc_filter(gIdx2, gDist2, sDist2, n_tar, thr);
        for (int i = 0; i < n_ref; i++) { /* bijectFilter :133-144 */
            int m = gIdx[i];
            if (m < 0 || counts[m] == 0) continue;
            if (gIdx2[m] != i) gIdx[i] *= -1;
        }
        free(counts);
    }
    int np = 0;
    for (int i = 0; i < n_ref; i++) { /* toCvec :99-112 */
        int j = gIdx[i];
        if (j < 0) continue;
        if (pair_ref) pair_ref[np] = i;
        if (pair_tar) pair_tar[np] = j;
        np++;
    }
    for (int s = 0; s < 8; s++) free(own[s]);
    return np;
}
