"""Volume ingest from NIfTI: the façade's own parser behind the reference's
readNiiFile entry (Include/Util/readNii.h:6; the reference reads through layNii + zlib and casts
every scalar type to float32 as stored, Src/Util/readNii.cpp:16-20, laynii_lib.cpp:226-310).
No GPU needed."""
import gzip
import struct

import numpy as np
import pytest

CODES = {np.dtype("uint8"): 2, np.dtype("int16"): 4, np.dtype("int32"): 8, np.dtype("float32"): 16, np.dtype("float64"): 64,
         np.dtype("int8"): 256, np.dtype("uint16"): 512, np.dtype("uint32"): 768, np.dtype("int64"): 1024, np.dtype("uint64"): 1280}


def nifti1_bytes(vol, big_endian=False, ext_bytes=0, slope=2.0):
    e = ">" if big_endian else "<"
    nz, ny, nx = vol.shape
    h = bytearray(348)
    struct.pack_into(e + "i", h, 0, 348)
    struct.pack_into(e + "8h", h, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into(e + "h", h, 70, CODES[vol.dtype])
    struct.pack_into(e + "h", h, 72, vol.dtype.itemsize * 8)
    struct.pack_into(e + "f", h, 108, 352.0 + ext_bytes)
    struct.pack_into(e + "f", h, 112, slope)          # scl_slope: must be ignored, as in the reference
    h[344:348] = b"n+1\0"
    data = vol.astype(vol.dtype.newbyteorder(e)).tobytes()
    return bytes(h) + b"\0\0\0\0" + b"\x07" * ext_bytes + data


def nifti2_bytes(vol):
    nz, ny, nx = vol.shape
    h = bytearray(540)
    struct.pack_into("<i", h, 0, 540)
    h[4:12] = b"n+2\0\r\n\x1a\n"
    struct.pack_into("<h", h, 12, CODES[vol.dtype])
    struct.pack_into("<h", h, 14, vol.dtype.itemsize * 8)
    struct.pack_into("<8q", h, 16, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<q", h, 168, 544)
    return bytes(h) + b"\0\0\0\0" + vol.tobytes()


@pytest.mark.parametrize("dtype", ["float32", "int16", "uint8", "float64", "uint16", "int32", "int8", "uint32", "int64", "uint64"])
def test_nifti1_every_scalar_type(s3d, tmp_path, dtype):
    rng = np.random.default_rng(3)
    vol = (rng.standard_normal((5, 6, 7)) * 50 + 60).astype(dtype)
    p = tmp_path / "v.nii"
    p.write_bytes(nifti1_bytes(vol))
    got = s3d.readNiiFile(str(p))
    assert got.dtype == np.float32 and got.shape == (5, 6, 7)
    assert np.array_equal(got, vol.astype(np.float32))      # static_cast<float>, scl_slope not applied


def test_gzip_big_endian_extension_and_nifti2(s3d, tmp_path):
    rng = np.random.default_rng(4)
    vol = rng.integers(-3000, 3000, size=(9, 4, 11)).astype(np.int16)
    p = tmp_path / "v.nii.gz"
    with gzip.open(p, "wb") as f:
        f.write(nifti1_bytes(vol, big_endian=True, ext_bytes=32))
    assert np.array_equal(s3d.readNiiFile(str(p)), vol.astype(np.float32))
    v2 = rng.standard_normal((3, 8, 5)).astype(np.float32)
    p2 = tmp_path / "v2.nii"
    p2.write_bytes(nifti2_bytes(v2))
    assert np.array_equal(s3d.readNiiFile(str(p2)), v2)


def test_bad_files_fail_loudly(s3d, tmp_path):
    p = tmp_path / "bad.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(s3d.S3DError):
        s3d.readNiiFile(str(p))
    with pytest.raises(s3d.S3DError):
        s3d.readNiiFile(str(tmp_path / "missing.nii"))
    vol = np.zeros((4, 4, 4), np.float32)
    p3 = tmp_path / "short.nii"
    p3.write_bytes(nifti1_bytes(vol)[:-10])
    with pytest.raises(s3d.S3DError):
        s3d.readNiiFile(str(p3))


@pytest.mark.gpu
def test_nifti_volume_through_the_extractor(s3d, synth, tmp_path):
    vol = synth.v_blobs(48, seed=2)
    p = tmp_path / "vol.nii.gz"
    with gzip.open(p, "wb") as f:
        f.write(nifti1_bytes(vol))
    a = s3d.CSIFT3DFactory.CreateCSIFT3D(s3d.readNiiFile(str(p)))
    a.KpSiftAlgorithm()
    b = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    b.KpSiftAlgorithm()
    assert len(a.GetKeypoints()) == len(b.GetKeypoints()) > 0 and np.array_equal(a.descriptors, b.descriptors)


# ---- fixtures produced by the reference's own code (layNii writer + readNiiFile): tests/golden/make_nii_golden.py --------

import os

NII_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nii")


def _golden_cases():
    exp = np.load(os.path.join(NII_DIR, "expected.npz"))
    return sorted(exp.files)


@pytest.mark.parametrize("name", _golden_cases())
def test_files_written_by_laynii_read_like_the_reference(s3d, name):
    """Every scalar type, gzip, NIfTI-2, the .hdr/.img pair and float64 NaNs (zeroed by the reference's conversion,
    laynii_lib.cpp:303-308): the product's parser must return the reference's readNiiFile output bit for bit."""
    want = np.load(os.path.join(NII_DIR, "expected.npz"))[name]
    got = s3d.readNiiFile(os.path.join(NII_DIR, name))
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got, want), name
    if "nan" in name:
        assert not np.isnan(got).any() and (got == 0).sum() >= got.size // 17


def test_pair_named_by_its_image_file(s3d):
    want = np.load(os.path.join(NII_DIR, "expected.npz"))["pair_i16.hdr"]
    assert np.array_equal(s3d.readNiiFile(os.path.join(NII_DIR, "pair_i16.img")), want)


def test_float32_keeps_nans_and_small_vox_offset_is_clamped_to_the_header(s3d, tmp_path):
    """readNiiFile converts (and de-NaNs) only inputs that are not float32 (Src/Util/readNii.cpp:16-20); a vox_offset
    below the header size is clamped to sizeof(header) = 348, not to 352 (nifti_convert_n1hdr2nim)."""
    vol = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    vol[1, 2, 3] = np.nan
    raw = bytearray(nifti1_bytes(vol))
    struct.pack_into("<f", raw, 108, 0.0)                       # vox_offset 0 in a single-file NIfTI-1
    p = tmp_path / "v.nii"
    p.write_bytes(bytes(raw[:348]) + bytes(raw[352:]))          # voxels directly after the 348-byte header
    got = s3d.readNiiFile(str(p))
    assert np.array_equal(got, vol, equal_nan=True) and np.isnan(got[1, 2, 3])
