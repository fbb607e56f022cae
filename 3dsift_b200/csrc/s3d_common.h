// s3d_common.h — shared host-side plumbing for libsift3d_b200.so (error slot, launch counter).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "sift3d_b200.h"

namespace s3d {

extern std::atomic<uint64_t> g_launches;
int fail(int code, const char* fmt, ...);
void clear_error();
// Make `device` current (or keep the current one for -1) and check it is an sm_100 part.
int use_device(int device, int* resolved);

}  // namespace s3d

#define S3D_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return s3d::fail(S3D_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,             \
                             cudaGetErrorString(e_));                                               \
    } while (0)

#define S3D_TRY(expr)              \
    do {                           \
        int r_ = (expr);           \
        if (r_ != S3D_OK) return r_; \
    } while (0)

// Every kernel launch of the library goes through this macro so bench.py can report gpu_launches.
#define S3D_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        s3d::g_launches.fetch_add(1, std::memory_order_relaxed);           \
    } while (0)

static inline unsigned s3d_blocks(size_t n, unsigned per_block) {
    return (unsigned)((n + per_block - 1) / per_block);
}
