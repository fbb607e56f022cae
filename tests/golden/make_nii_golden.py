"""NIfTI golden fixtures, produced by the REFERENCE's own code where /root/reference exists:

    make -C oracle nii && python tests/golden/make_nii_golden.py

* the FILES are written by layNii's writer (nifti_image_write, /root/reference/3DSIFT/3party/layNii/dep/nifti2_io.cpp)
  through oracle/_ref/ref_nii_tool — every scalar type the reference converts (laynii_lib.cpp:249-310), plain and
  gzip, NIfTI-1 and NIfTI-2 single files and a NIfTI-1 .hdr/.img pair, float64 with NaNs;
* the EXPECTED float32 volumes come from the reference's readNiiFile (Src/Util/readNii.cpp:5-39) reading those files.

Both are committed under tests/golden/nii/ (a few hundred bytes each); tests/test_nii.py compares the product's
readNiiFile with them on any machine."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "nii")
TOOL = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "ref_nii_tool")

# (file name, datatype code, nx, ny, nz, seed, extra)
CASES = [
    ("u8.nii", 2, 7, 6, 5, 1, ""), ("i8.nii", 256, 7, 6, 5, 2, ""), ("i16.nii", 4, 7, 6, 5, 3, ""), ("u16.nii", 512, 7, 6, 5, 4, ""),
    ("i32.nii", 8, 7, 6, 5, 5, ""), ("u32.nii", 768, 7, 6, 5, 6, ""), ("i64.nii", 1024, 7, 6, 5, 7, ""), ("u64.nii", 1280, 7, 6, 5, 8, ""),
    ("f32.nii", 16, 7, 6, 5, 9, ""), ("f64_nan.nii", 64, 7, 6, 5, 10, ""),
    ("gz_i16.nii.gz", 4, 9, 4, 11, 11, ""), ("gz_f64_nan.nii.gz", 64, 9, 4, 11, 12, ""), ("gz_f32.nii.gz", 16, 12, 3, 4, 13, ""),
    ("f32_v2.nii", 16, 5, 8, 3, 14, "nifti2"), ("gz_u8_v2.nii.gz", 2, 5, 8, 3, 15, "nifti2"),
    ("pair_i16.hdr", 4, 6, 5, 4, 16, ""), ("pair_f64.hdr", 64, 6, 5, 4, 17, ""),
]


def main():
    if not os.path.exists(TOOL):
        sys.exit(f"{TOOL} missing: run `make -C oracle nii` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    expected = {}
    for name, dt, nx, ny, nz, seed, extra in CASES:
        path = os.path.join(OUT, name)
        subprocess.check_call([TOOL, "write", path, str(dt), str(nx), str(ny), str(nz), str(seed)] + ([extra] if extra else []))
        tmp = path + ".ref.bin"
        subprocess.check_call([TOOL, "read", path, tmp], stdout=subprocess.DEVNULL)
        with open(tmp, "rb") as f:
            dims = np.fromfile(f, np.int32, 3)
            vol = np.fromfile(f, np.float32)
        os.remove(tmp)
        assert tuple(dims) == (nx, ny, nz) and vol.size == nx * ny * nz
        expected[name] = vol.reshape(nz, ny, nx)
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    print(f"wrote {len(CASES)} NIfTI files + expected.npz under {OUT}")


if __name__ == "__main__":
    main()
