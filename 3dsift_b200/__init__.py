"""3dsift_b200 — B200-native (sm_100a) implementation of the 3DSIFT hot path.

The directory name starts with a digit, so import it with
``importlib.import_module("3dsift_b200")`` (tests/conftest.py does).  The product is
``lib/libsift3d_b200.so`` (hand-written CUDA + C ABI, include/sift3d_b200.h) plus the C++ façade
headers under include/3dsift/; this package is the Python host mirror used by tests and bench.
"""
from .api import (  # noqa: F401
    CSIFT3D, CSIFT3DFactory, DESC_LENGTH, KP_DTYPE, S3DError, DownSample_3D, GaussianSmooth_3D, blur_axis, check,
    device_count, launch_count, lib, match_stats, muBruteMatcher, readNiiFile, read_matrix_from_disk, selftest, set_describe_path, set_match_path, trim_cache,
    write_matrix_to_disk,
)
from . import synth  # noqa: F401
