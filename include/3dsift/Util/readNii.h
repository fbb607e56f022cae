// readNii.h — volume ingest from NIfTI files, same entry point as the reference's
// Include/Util/readNii.h:6 (readNiiFile, implemented there on top of layNii/nifti2 + zlib,
// Src/Util/readNii.cpp:5-39).  This implementation parses the format itself (facade.cpp):
// single-file NIfTI-1 ("n+1") and NIfTI-2 ("n+2"), either byte order, plain or gzip-compressed
// (.nii / .nii.gz), every scalar datatype the reference converts (copy_nifti_as_float32,
// 3party/layNii/dep/laynii_lib.cpp:226-310): values are cast to float32 as they are stored —
// scl_slope / scl_inter are NOT applied, as in the reference.  Only the first volume (nx*ny*nz
// voxels, x fastest) is returned.
#ifndef __READ_NII_H__
#define __READ_NII_H__

#include "common.h"

// Returns a new float[nx*ny*nz] the caller releases with delete[] (reference: readNii.cpp:25), or
// nullptr after printing a message (the reference dereferences a null image in that case).
SIFT_LIBRARY_API float* readNiiFile(const char* fname, int& nx, int& ny, int& nz);

#endif
