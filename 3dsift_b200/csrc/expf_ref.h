// expf_ref.h — float exp() that reproduces glibc's expf bit-for-bit on the argument range the
// 3DSIFT hot path uses ([-88, 0]); shared by the CUDA kernels and a host-side unit test.
//
// Why: the reference weights every window voxel with expf(...) from the host libm
// (Src/cSIFT3D.cc:971 orientation, :1312 descriptor).  CUDA's expf is a different approximation
// (up to 2 ulp), which perturbs every weight.  glibc >= 2.28 computes expf in double with a
// 32-entry 2^(i/32) table and a cubic (sysdeps/ieee754/flt-32/e_expf.c, from ARM
// optimized-routines); evaluating the same recipe in double on the device gives the same float
// except where the pre-rounding double differs in its last bits *and* sits on a float rounding
// boundary (~1e-8 of arguments; FMA vs non-FMA host builds differ from each other at that rate).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define S3D_HD __host__ __device__ __forceinline__
#else
#define S3D_HD static inline
#endif

// T[i] = bits(2^(i/32)) - (i << 47), correctly rounded (generated with 60-digit decimals)
#define S3D_EXP2F_TAB_INIT \
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL

static const uint64_t s3d_exp2f_tab_host[32] = {S3D_EXP2F_TAB_INIT};
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t s3d_exp2f_tab_dev[32] = {S3D_EXP2F_TAB_INIT};
#endif
#if defined(__CUDA_ARCH__)
#define S3D_EXP2F_TAB s3d_exp2f_tab_dev
#else
#define S3D_EXP2F_TAB s3d_exp2f_tab_host
#endif

S3D_HD float s3d_expf_ref(float x) {
    // Arguments on the hot path are in [-4.5, 0]; clamp so the table arithmetic stays in range.
    if (!(x > -87.0f)) return 0.0f;
    if (x > 87.0f) x = 87.0f;
    const double InvLn2N = 0x1.71547652b82fep+0 * 32;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32;
    double xd = (double)x;
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki;
#if defined(__CUDA_ARCH__)
    ki = (uint64_t)__double_as_longlong(kd);
#else
    union { double d; uint64_t u; } cv;
    cv.d = kd;
    ki = cv.u;
#endif
    kd -= SHIFT;
    double r = z - kd;
    uint64_t t = S3D_EXP2F_TAB[ki % 32];
    t += ki << (52 - 5);
    double s;
#if defined(__CUDA_ARCH__)
    s = __longlong_as_double((long long)t);
    // explicit non-fused arithmetic: this header may be compiled with FMA contraction on
    double zz = __dadd_rn(__dmul_rn(C0, r), C1);
    double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(zz, r2), y);
    y = __dmul_rn(y, s);
#else
    cv.u = t;
    s = cv.d;
    double zz = C0 * r + C1;
    double r2 = r * r;
    double y = C2 * r + 1;
    y = zz * r2 + y;
    y = y * s;
#endif
    return (float)y;
}
