/* Force-included portability shim so the UNMODIFIED reference sources build with g++ on Linux.
 * Test infrastructure only (oracle/_ref); nothing here is part of the product.
 * Neutralises MSVC-only spellings used by /root/reference/3DSIFT/Src/cUtil.cc:608-1283
 * (sprintf_s / fopen_s / errno_t in the dead debug writers) and supplies headers the
 * reference forgets to include (cTexImage.h:15 size_t, cSIFT3D.cc:23 FLT_EPSILON, :161 memcpy). */
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <cmath>
#include <cerrno>
#include <cstdarg>
typedef int errno_t;
static inline errno_t fopen_s(FILE** f, const char* name, const char* mode) {
    *f = fopen(name, mode);
    return *f ? 0 : errno;
}
template <size_t N>
static inline int sprintf_s(char (&buf)[N], const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    int r = vsnprintf(buf, N, fmt, ap);
    va_end(ap);
    return r;
}
