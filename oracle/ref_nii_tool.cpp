// ref_nii_tool — TEST INFRASTRUCTURE (never linked into the product).  Two verbs over the UNMODIFIED reference
// sources, compiled where they lie (oracle/Makefile target `nii`):
//   write <file> <datatype> <nx> <ny> <nz> <seed> [nifti2]   a volume written by layNii's own writer
//                                                             (nifti_image_write, 3party/layNii/dep/nifti2_io.cpp)
//   read  <file> <out.bin>                                    the reference's readNiiFile (Src/Util/readNii.cpp:5-39)
//                                                             -> "int nx ny nz" + float32 voxels
// tests/golden/make_nii_golden.py drives it to produce the committed fixtures under tests/golden/nii/.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "Include/Util/readNii.h"                // (first: nifti2_io.h defines min / max as macros)
#include "3party/layNii/dep/laynii_lib.h"

static uint64_t lcg(uint64_t& s) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 11; }

template <class T>
static void fill(void* data, size_t n, uint64_t seed, double lo, double hi) {
    T* p = static_cast<T*>(data);
    for (size_t i = 0; i < n; ++i) {
        const double u = (double)(lcg(seed) % 1000003) / 1000003.0;
        p[i] = (T)(lo + (hi - lo) * u);
    }
}

int main(int argc, char** argv) {
    if (argc >= 4 && !strcmp(argv[1], "read")) {
        int nx = 0, ny = 0, nz = 0;
        float* d = readNiiFile(argv[2], nx, ny, nz);
        FILE* f = fopen(argv[3], "wb");
        if (!d || !f) return 3;
        const int dims[3] = {nx, ny, nz};
        fwrite(dims, sizeof(int), 3, f);
        fwrite(d, sizeof(float), (size_t)nx * ny * nz, f);
        fclose(f);
        delete[] d;
        return 0;
    }
    if (argc >= 8 && !strcmp(argv[1], "write")) {
        const int dt = atoi(argv[3]);
        const int64_t dims[8] = {3, atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), 1, 1, 1, 1};
        uint64_t seed = strtoull(argv[7], nullptr, 10);
        nifti_image* nim = nifti_make_new_nim(dims, dt, 1);
        if (!nim) return 4;
        const size_t n = (size_t)nim->nvox;
        switch (dt) {
            case 2: fill<uint8_t>(nim->data, n, seed, 0, 255); break;
            case 256: fill<int8_t>(nim->data, n, seed, -128, 127); break;
            case 4: fill<int16_t>(nim->data, n, seed, -3000, 3000); break;
            case 512: fill<uint16_t>(nim->data, n, seed, 0, 65535); break;
            case 8: fill<int32_t>(nim->data, n, seed, -2.0e9, 2.0e9); break;
            case 768: fill<uint32_t>(nim->data, n, seed, 0, 4.0e9); break;
            case 1024: fill<int64_t>(nim->data, n, seed, -9.0e15, 9.0e15); break;
            case 1280: fill<uint64_t>(nim->data, n, seed, 0, 1.8e16); break;
            case 16: fill<float>(nim->data, n, seed, -1000, 1000); break;
            case 64: {
                fill<double>(nim->data, n, seed, -1.0e6, 1.0e6);
                double* p = static_cast<double*>(nim->data);  // NaNs: the reference zeroes them for non-float32 inputs
                for (size_t i = 3; i < n; i += 17) p[i] = std::numeric_limits<double>::quiet_NaN();
                break;
            }
            default: return 5;
        }
        nim->scl_slope = 2.0f; nim->scl_inter = 7.0f;  // must be ignored on the way in
        const bool v2 = argc > 8 && !strcmp(argv[8], "nifti2");
        const bool pair = strstr(argv[2], ".hdr") != nullptr;
        nim->nifti_type = v2 ? NIFTI_FTYPE_NIFTI2_1 : (pair ? NIFTI_FTYPE_NIFTI1_2 : NIFTI_FTYPE_NIFTI1_1);
        if (nifti_set_filenames(nim, argv[2], 0, 1)) return 6;
        nifti_set_iname_offset(nim, v2 ? 2 : 1);
        nifti_image_write(nim);
        nifti_image_free(nim);
        return 0;
    }
    fprintf(stderr, "usage: %s write <file> <datatype> <nx> <ny> <nz> <seed> [nifti2] | read <file> <out.bin>\n", argv[0]);
    return 2;
}
