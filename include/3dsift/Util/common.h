// 3dsift/Util/common.h — timers and the export macro of the drop-in C++ surface.
// Source-compatible with /root/reference/3DSIFT/Include/Util/common.h:4-63 (same type and field
// names), with a portable visibility macro in place of the MSVC-only __declspec pair (:4-8).
#ifndef S3D_FACADE_COMMON_H
#define S3D_FACADE_COMMON_H

#include <chrono>
#include <iostream>
#include <vector>

#if defined(_WIN32)
#  ifdef SIFT_LIBRARY_EXPORTS
#    define SIFT_LIBRARY_API __declspec(dllexport)
#  else
#    define SIFT_LIBRARY_API __declspec(dllimport)
#  endif
#else
#  define SIFT_LIBRARY_API __attribute__((visibility("default")))
#endif

template <class T>
T getDoubleMill(std::chrono::high_resolution_clock::time_point start, std::chrono::high_resolution_clock::time_point end) {
    return std::chrono::duration<T, std::milli>(end - start).count();
}

// Per-stage wall times of one extraction, in seconds (reference: filled from omp_get_wtime deltas,
// Src/cSIFT3D.cc:169-233; here: from CUDA events on the extractor's stream).
struct SIFT_LIBRARY_API SIFT_TimerPara {
    double d_TotalTime = 0;
    double d_Allocation = 0;
    double d_BuildGSS = 0;
    double d_BuildDOG = 0;   // 0 on this implementation: the DoG is fused into the last blur pass
    double d_Detect = 0;
    double d_AssignOrientation = 0;
    double d_Extraction = 0;
    double d_release = 0;
    double d_memoryOverhead = 0;  // host<->device copies
    std::vector<double> vD_octaveTime;
    std::vector<double> vD_octaveCompute;
    double getAllComputeTime() { return d_BuildGSS + d_BuildDOG + d_Detect + d_AssignOrientation + d_Extraction; }
};

struct SIFT_LIBRARY_API SIFT_Dev_TimerPara {
    int devId = 0;
    double d_Allocation = 0, d_hostCopy = 0, d_BuildGSS = 0, d_Detect = 0, d_KeypointCompute = 0, d_downsample = 0;
};

struct SIFT_LIBRARY_API SIFT_PROCESS {
    SIFT_TimerPara REF;
    SIFT_TimerPara TAR;
    double d_RegTime = 0;
};

SIFT_LIBRARY_API std::ostream& operator<<(std::ostream& os, const SIFT_TimerPara& st);
SIFT_LIBRARY_API std::ostream& operator<<(std::ostream& os, const SIFT_PROCESS& sp);

#endif
