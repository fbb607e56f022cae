"""Host-side mirror of the reference's public interface over the C ABI (include/sift3d_b200.h).

Names, argument meaning and defaults follow /root/reference/3DSIFT/Include/cSIFT3D.h:118-204 and
Include/cMatcher.h:12-88, so tests read like client code of the reference
(3DSIFT/Example.cpp:21-44):

    sift = CSIFT3DFactory.CreateCSIFT3D(volume)      # volume: float32 [nz, ny, nx], x fastest
    sift.KpSiftAlgorithm()
    kps = sift.GetKeypoints()                         # structured array, KP_DTYPE (176-byte records)
    m = muBruteMatcher()
    ref_xyz, tar_xyz = m.enhancedMatch(kps_ref, kps_tar, 0.85)

Everything here is plumbing: the arithmetic runs in the hand-written sm_100a kernels of
libsift3d_b200.so.  There is no CPU fallback — a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libsift3d_b200.so")

DESC_LENGTH = 768  # Include/cMatcher.h:8

# CPUSIFT::Keypoint, Include/cSIFT3D.h:52-70 (sizeof == 176 on LP64)
KP_DTYPE = np.dtype(
    [("x", "f4"), ("y", "f4"), ("z", "f4"), ("scale", "f4"), ("octave", "i4"), ("level", "i4"),
     ("rx", "f4"), ("ry", "f4"), ("rz", "f4"), ("win", "f4", 3), ("eigvalue", "f4", 3),
     ("eigvector", "f4", 9), ("Rotation", "f4", 9), ("str_tensor", "f4", 9), ("desc", "u8")],
    align=True)
assert KP_DTYPE.itemsize == 176


class S3DError(RuntimeError):
    pass


def _rebuild_keypoints(raw, desc):
    kp = np.frombuffer(bytearray(raw), dtype=KP_DTYPE).view(KeypointArray)
    return _bind_descriptors(kp, desc, rows=None)


def _bind_descriptors(kp, desc, rows):
    """Point kp['desc'] at rows of `desc` (rows None: keep each record's row index relative to the old block)
    and make `desc` live as long as kp (and every view / copy / slice of it)."""
    kp = kp.view(KeypointArray)
    if rows is None:
        p = kp["desc"].astype(np.int64)
        rows = (p - (int(p.min()) if len(p) else 0)) // (DESC_LENGTH * 4)
    kp["desc"] = desc.ctypes.data + np.asarray(rows, dtype=np.uint64) * np.uint64(DESC_LENGTH * 4)
    kp._desc_owner = desc
    return kp


class KeypointArray(np.ndarray):
    """KP_DTYPE records whose `desc` pointers BORROW from a descriptor block (Keypoint::desc borrows from
    CSIFT3D::global_descriptor, Src/cSIFT3D.cc:486,495).  The reference frees that block in ~CSIFT3D; here
    the array (and every view, slice or copy of it) holds a strong reference to the block, so a helper that
    returns only GetKeypoints() cannot leave dangling pointers, and pickling (all_gather_object) ships the
    block along and re-points the records in the receiving process."""

    _desc_owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._desc_owner = getattr(obj, "_desc_owner", None)

    def __reduce__(self):
        own = self._desc_owner
        if own is None or self.dtype != KP_DTYPE:
            return super().__reduce__()
        flat = np.ascontiguousarray(self).reshape(-1)
        base, n = own.ctypes.data, len(flat)
        rows = (flat["desc"].astype(np.int64) - base) // (DESC_LENGTH * 4) if n else np.zeros(0, np.int64)
        block = np.ascontiguousarray(own.reshape(-1, DESC_LENGTH)[rows]) if n else np.zeros((0, DESC_LENGTH), np.float32)
        plain = flat.view(np.ndarray).copy()
        plain["desc"] = np.arange(n, dtype=np.uint64) * np.uint64(DESC_LENGTH * 4)
        return (_rebuild_keypoints, (plain.tobytes(), block))


class s3d_params(C.Structure):
    _fields_ = [("num_kp_levels", C.c_int), ("sigma_default", C.c_float), ("sigma_n_default", C.c_float),
                ("peak_thresh", C.c_float), ("max_eig_thres", C.c_float), ("corner_thresh", C.c_float),
                ("device", C.c_int), ("keep_levels", C.c_int), ("exact_recheck", C.c_int), ("profile", C.c_int),
                ("stream", C.c_void_p)]


_lib = None


def lib():
    """Load libsift3d_b200.so (built in-tree by 3dsift_b200/build.py).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise S3DError(f"{LIB_PATH} is missing: run `python 3dsift_b200/build.py` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_void_p, C.c_void_p
    L.s3d_last_error.restype = C.c_char_p
    L.s3d_launch_count.restype = C.c_uint64
    L.s3d_default_params.argtypes = [C.POINTER(s3d_params)]
    L.s3d_default_params.restype = None
    L.s3d_selftest.argtypes = [C.c_int]
    L.s3d_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.POINTER(vp)]
    L.s3d_create_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.POINTER(vp)]
    L.s3d_create_async.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.POINTER(vp)]
    for n in ("s3d_run", "s3d_run_async", "s3d_wait"):
        getattr(L, n).argtypes = [vp]
    L.s3d_destroy.argtypes = [vp]
    L.s3d_destroy.restype = None
    L.s3d_num_octaves.argtypes = [vp, C.POINTER(C.c_int)]
    L.s3d_level_dims.argtypes = [vp, C.c_int, C.POINTER(C.c_int * 3)]
    L.s3d_num_keypoints.argtypes = [vp, C.POINTER(C.c_int)]
    L.s3d_get_keypoints.argtypes = [vp, vp, fp]
    L.s3d_get_keypoints_async.argtypes = [vp, vp, fp]
    L.s3d_sync.argtypes = [vp]
    L.s3d_trim_cache.argtypes = [C.c_int, C.POINTER(C.c_ulonglong)]
    L.s3d_num_extrema.argtypes = [vp, C.POINTER(C.c_int)]
    L.s3d_get_extrema.argtypes = [vp, vp, ip, ip]
    L.s3d_get_level.argtypes = [vp, C.c_int, C.c_int, fp]
    L.s3d_get_input.argtypes = [vp, fp]
    L.s3d_get_thresholds.argtypes = [vp, fp, C.c_int]
    L.s3d_get_timers.argtypes = [vp, C.POINTER(C.c_double * 10)]
    L.s3d_get_kernel_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_int), vp, vp, vp]
    L.s3d_kernel_class_name.argtypes = [C.c_int]
    L.s3d_kernel_class_name.restype = C.c_char_p
    L.s3d_gaussian_smooth.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_float, fp]
    L.s3d_blur_axis.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp]
    L.s3d_downsample.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp]
    L.s3d_match.argtypes = [C.c_int, fp, C.c_int, fp, C.c_int, C.c_double] + [vp] * 12
    L.s3d_match_device.argtypes = [C.c_int, fp, C.c_int, fp, C.c_int, C.c_double] + [vp] * 12
    L.s3d_top2_device.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int] + [vp] * 6
    L.s3d_top2_merge_device.argtypes = [C.c_int, C.c_int] + [vp] * 10
    L.s3d_ratio_filter_device.argtypes = [vp, vp, vp, C.c_int, C.c_double, vp]
    L.s3d_count_mask_device.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp]
    L.s3d_biject_filter_device.argtypes = [vp, C.c_int, vp, vp, vp]
    L.s3d_pairs_device.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.s3d_match_ex.argtypes = [C.c_int, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double] + [vp] * 12
    L.s3d_set_match_path.argtypes = [C.c_int]
    L.s3d_match_sharded.argtypes = [vp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_double] + [vp] * 12
    L.s3d_match_multi.argtypes = [C.c_int, fp, C.c_int, fp, C.c_int, C.c_double, C.POINTER(C.c_int), C.c_int] + [vp] * 12
    L.s3d_read_nii.restype = C.POINTER(C.c_float)
    L.s3d_read_nii.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.s3d_free_host.restype = None
    L.s3d_free_host.argtypes = [C.POINTER(C.c_float)]
    L.s3d_set_describe_path.argtypes = [C.c_int]
    L.s3d_get_counters.argtypes = [vp, C.POINTER(C.c_int)]
    L.s3d_match_stats.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.c_int]
    L.s3d_match_stats.restype = None
    L.s3d_level_info.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.s3d_device_descriptors.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int)]
    L.s3d_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    ip = C.POINTER(C.c_int)
    L.s3d_slab_extent.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.c_int, ip]
    L.s3d_slab_bounds.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip]
    L.s3d_slab_first_replicated.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), ip]
    L.s3d_slab_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.c_int, ip]
    L.s3d_comm_unique_id.argtypes = [vp]
    L.s3d_comm_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.s3d_comm_destroy.argtypes = [vp]
    L.s3d_comm_destroy.restype = None
    L.s3d_comm_info.argtypes = [vp, ip, ip, ip]
    L.s3d_comm_traffic.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.s3d_slab_run.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.POINTER(vp)]
    L.s3d_slab_create.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), C.POINTER(vp)]
    L.s3d_slab_execute.argtypes = [vp, vp]
    L.s3d_slab_execute_async.argtypes = [vp, vp]
    L.s3d_slab_gather.argtypes = [vp, vp, C.c_int, C.c_int]
    L.s3d_slab_phases.argtypes = [vp, C.POINTER(C.c_double * 8)]
    L.s3d_extract_multi.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(s3d_params), ip, C.c_int, C.c_int, C.POINTER(vp)]
    L.s3d_slab_info.argtypes = [vp, ip, ip, ip, ip]
    L.s3d_slab_level_buffer.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), ip]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise S3DError(f"libsift3d_b200 error {rc}: {lib().s3d_last_error().decode(errors='replace')}")


def _ptr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def device_count():
    return int(lib().s3d_device_count())


def launch_count():
    return int(lib().s3d_launch_count())


def selftest(device=-1):
    check(lib().s3d_selftest(device))


def trim_cache(device=-1):
    """Return the library's cached device blocks to the driver; returns the number of bytes that were cached."""
    n = C.c_ulonglong()
    check(lib().s3d_trim_cache(device, C.byref(n)))
    return n.value


MATCH_AUTO, MATCH_EXACT, MATCH_TENSOR, MATCH_TENSOR_SINGLE, MATCH_TENSOR_PAIR = 0, 1, 2, 3, 4


def set_match_path(path):
    """0 auto / 1 exact CUDA-core kernel / 2 tensor-core candidate pass (variant by size) / 3 tensor cores,
    one CTA per tile / 4 tensor cores, CTA pairs with resident query tile (results are identical)."""
    check(lib().s3d_set_match_path(int(path)))


def readNiiFile(path):
    """readNiiFile (Include/Util/readNii.h:6): NIfTI-1/-2 (.nii / .nii.gz) -> float32 volume [nz, ny, nx].
    Needs no GPU.  Raises S3DError when the file cannot be read."""
    nx, ny, nz = C.c_int(), C.c_int(), C.c_int()
    p = lib().s3d_read_nii(os.fsencode(path), C.byref(nx), C.byref(ny), C.byref(nz))
    if not p:
        raise S3DError(f"readNiiFile({path}) failed")
    try:
        n = nx.value * ny.value * nz.value
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(nz.value, ny.value, nx.value).copy()
    finally:
        lib().s3d_free_host(p)


DESC_FIXED, DESC_FP32, DESC_FORCE_REDO = 0, 1, 2


def set_describe_path(path):
    """0 fixed-point atomics + FP32 redo (default) / 1 FP32 ordered accumulation / 2 forced redo (tests)."""
    check(lib().s3d_set_describe_path(int(path)))


def match_stats(reset=False):
    """(rows searched on the tensor-core path, rows that needed the exact fallback)."""
    a, b = C.c_ulonglong(), C.c_ulonglong()
    lib().s3d_match_stats(C.byref(a), C.byref(b), int(reset))
    return int(a.value), int(b.value)


# Defaults of CSIFT3DFactory::CreateCSIFT3D, Include/cSIFT3D.h:187-202 (values :13-21)
NUM_KP_LEVELS, SIGMA_DEFAULT, SIGMA_N_DEFAULT = 3, 1.6, 1.15
PEAK_THRESH, EIG_THRES, CORNER_THRESH = 0.1, 0.9, 0.4


class CSIFT3D:
    """CPUSIFT::CSIFT3D (Include/cSIFT3D.h:118-183) over an s3d_handle."""

    def __init__(self, volume, x_dim=None, y_dim=None, z_dim=None, num_kp_levels=NUM_KP_LEVELS,
                 sigma_default=SIGMA_DEFAULT, sigma_n_default=SIGMA_N_DEFAULT, peak_thresh=PEAK_THRESH,
                 max_eig_thres=EIG_THRES, corner_thresh=CORNER_THRESH, *, device=-1, keep_levels=False,
                 exact_recheck=True, profile=False, stream=None, async_upload=False):
        L = lib()
        self._h = C.c_void_p()
        p = s3d_params()
        L.s3d_default_params(C.byref(p))
        p.num_kp_levels, p.sigma_default, p.sigma_n_default = num_kp_levels, sigma_default, sigma_n_default
        p.peak_thresh, p.max_eig_thres, p.corner_thresh = peak_thresh, max_eig_thres, corner_thresh
        p.device, p.keep_levels, p.exact_recheck = device, int(keep_levels), int(exact_recheck)
        p.profile = int(profile)
        p.stream = C.c_void_p(int(stream)) if stream else None
        self.num_kp_levels = num_kp_levels
        on_device = hasattr(volume, "is_cuda") and volume.is_cuda
        if isinstance(volume, np.ndarray):
            volume = np.ascontiguousarray(volume, dtype=np.float32)
        if x_dim is None:
            z_dim, y_dim, x_dim = (int(v) for v in volume.shape)
        self.dims = (int(x_dim), int(y_dim), int(z_dim))
        if on_device:
            # s3d_create_device reads raw float32 memory: refuse anything else instead of misreading it
            import torch
            if volume.dtype != torch.float32 or not volume.is_contiguous():
                raise S3DError(f"device volume must be a contiguous float32 tensor (got {volume.dtype}, contiguous={volume.is_contiguous()})")
            if p.device < 0:
                p.device = volume.device.index
            if not stream:
                # without an explicit stream the handle works on a private non-blocking stream, which is not ordered
                # against the stream that produced `volume`: finish that work first
                torch.cuda.current_stream(volume.device).synchronize()
            check(L.s3d_create_device(_ptr(volume), x_dim, y_dim, z_dim, C.byref(p), C.byref(self._h)))
        elif async_upload:
            # the upload is only enqueued: `volume` (pinned) must stay alive until KpSiftAlgorithm returns
            self._pinned = volume
            check(L.s3d_create_async(_ptr(volume), x_dim, y_dim, z_dim, C.byref(p), C.byref(self._h)))
        else:
            check(L.s3d_create(_ptr(volume), x_dim, y_dim, z_dim, C.byref(p), C.byref(self._h)))
        self._kp = None
        self._desc = None

    # -- reference surface -------------------------------------------------------------------
    def KpSiftAlgorithm(self):
        check(lib().s3d_run(self._h))

    def SetNumThreads(self, t_num):  # Include/cSIFT3D.h:154 — no CPU threads on this path
        pass

    def GetKeypoints(self, kp_out=None, desc_out=None):
        """Structured KP_DTYPE array; `desc` fields point into self.descriptors (kept alive by
        this object, as the reference keeps global_descriptor, Src/cSIFT3D.cc:486,495)."""
        n = self.num_keypoints()
        kp = kp_out if kp_out is not None else np.zeros(max(n, 1), KP_DTYPE)
        desc = desc_out if desc_out is not None else np.zeros((max(n, 1), DESC_LENGTH), np.float32)
        check(lib().s3d_get_keypoints(self._h, _ptr(kp), _ptr(desc)))
        if isinstance(kp, np.ndarray):
            kp = kp[:n]
            desc = desc[:n]
            # the returned array keeps `desc` alive (the pointers must not outlive the block)
            kp = _bind_descriptors(kp, desc, rows=np.arange(n))
        self._kp, self._desc = kp, desc
        return kp

    @property
    def descriptors(self):
        if self._desc is None:
            self.GetKeypoints()
        return self._desc

    # -- asynchronous halves ---------------------------------------------------------------------
    def run_async(self):
        check(lib().s3d_run_async(self._h))

    def wait(self):
        check(lib().s3d_wait(self._h))

    def get_keypoints_async(self, kp_ptr, desc_ptr):
        """Enqueue the D2H copies of the records / descriptors into (pinned) host memory; sync() completes them."""
        check(lib().s3d_get_keypoints_async(self._h, kp_ptr, desc_ptr))

    def sync(self):
        check(lib().s3d_sync(self._h))

    # -- parity hooks (GET_GSS / GET_DOG / GET_LEVEL, Include/cSIFT3D.h:169-177) -----------------
    def num_keypoints(self):
        n = C.c_int()
        check(lib().s3d_num_keypoints(self._h, C.byref(n)))
        return n.value

    def num_octaves(self):
        n = C.c_int()
        check(lib().s3d_num_octaves(self._h, C.byref(n)))
        return n.value

    def level_dims(self, octave):
        d = (C.c_int * 3)()
        check(lib().s3d_level_dims(self._h, octave, C.byref(d)))
        return tuple(d)

    def _level(self, which, idx):
        per = self.num_kp_levels + (3 if which == 0 else 2)
        nx, ny, nz = self.level_dims(idx // per)
        out = np.empty((nz, ny, nx), np.float32)
        check(lib().s3d_get_level(self._h, which, idx, _ptr(out)))
        return out

    def GET_GSS(self, idx):
        return self._level(0, idx)

    def GET_DOG(self, idx):
        return self._level(1, idx)

    def input(self):
        nx, ny, nz = self.dims
        out = np.empty((nz, ny, nx), np.float32)
        check(lib().s3d_get_input(self._h, _ptr(out)))
        return out

    def extrema(self):
        """(records, codes, xyz5): raw detections after orientation, `extre`/RET/level_extrema."""
        n = C.c_int()
        check(lib().s3d_num_extrema(self._h, C.byref(n)))
        n = n.value
        kp = np.zeros(max(n, 1), KP_DTYPE)
        codes = np.zeros(max(n, 1), np.int32)
        xyz5 = np.zeros((max(n, 1), 5), np.int32)
        check(lib().s3d_get_extrema(self._h, _ptr(kp), _ptr(codes), _ptr(xyz5)))
        return kp[:n], codes[:n], xyz5[:n]

    def thresholds(self):
        n = self.num_octaves() * self.num_kp_levels
        out = np.zeros(n, np.float32)
        check(lib().s3d_get_thresholds(self._h, _ptr(out), n))
        return out

    @property
    def m_timer(self):
        """SIFT_TimerPara (Include/Util/common.h:22-41) as a dict of seconds."""
        t = (C.c_double * 10)()
        check(lib().s3d_get_timers(self._h, C.byref(t)))
        names = ["d_Allocation", "d_BuildGSS", "d_BuildDOG", "d_Detect", "d_AssignOrientation", "d_Extraction",
                 "d_release", "d_TotalTime", "d_h2d", "d_d2h"]
        return dict(zip(names, list(t)))

    def kernel_stats(self):
        """{class name: dict(ms, launches, alg_bytes)} of the last run (needs profile=True)."""
        cap = 32
        n = C.c_int()
        ms = np.zeros(cap, np.float64); cnt = np.zeros(cap, np.int64); by = np.zeros(cap, np.float64)
        check(lib().s3d_get_kernel_stats(self._h, cap, C.byref(n), _ptr(ms), _ptr(cnt), _ptr(by)))
        return {lib().s3d_kernel_class_name(i).decode(): dict(ms=float(ms[i]), launches=int(cnt[i]), alg_bytes=float(by[i]))
                for i in range(n.value) if cnt[i]}

    def counters(self):
        """dict(orient_rechecked, orient_flipped, desc_redo): safeguard bookkeeping of the last run."""
        out = (C.c_int * 4)()
        check(lib().s3d_get_counters(self._h, out))
        return dict(orient_rechecked=int(out[0]), orient_flipped=int(out[1]), desc_redo=int(out[2]), sparse_resized=int(out[3]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().s3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CSIFT3DFactory:
    """CPUSIFT::CSIFT3DFactory (Include/cSIFT3D.h:184-204)."""

    @staticmethod
    def CreateCSIFT3D(volume, x_dim=None, y_dim=None, z_dim=None, num_kp_levels=NUM_KP_LEVELS,
                      sigma_default=SIGMA_DEFAULT, sigma_n_default=SIGMA_N_DEFAULT, peak_thresh=PEAK_THRESH,
                      max_eigo_thres=EIG_THRES, corner_thresh=CORNER_THRESH, **kw):
        if isinstance(volume, (str, os.PathLike)):
            volume = read_matrix_from_disk(volume)
        return CSIFT3D(volume, x_dim, y_dim, z_dim, num_kp_levels, sigma_default, sigma_n_default, peak_thresh,
                       max_eigo_thres, corner_thresh, **kw)


def read_matrix_from_disk(path):
    """ReadMatrixFromDisk<float> (Include/Util/matrixIO3D.h:22-64): int m, n, p header then m*n*p
    float32, as consumed by CreateCSIFT3D(std::string) (Src/cSIFT3D.cc:112-125: m->x, n->y, p->z)."""
    with open(path, "rb") as f:
        m, n, p = np.fromfile(f, dtype=np.int32, count=3)
        data = np.fromfile(f, dtype=np.float32, count=int(m) * int(n) * int(p))
    if data.size != int(m) * int(n) * int(p):
        raise S3DError(f"{path}: truncated matrix file")
    return data.reshape(int(p), int(n), int(m))


def write_matrix_to_disk(path, vol):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, ny, nx = vol.shape
    with open(path, "wb") as f:
        np.array([nx, ny, nz], np.int32).tofile(f)
        vol.tofile(f)


def _desc_matrix(kps):
    """Gather Keypoint::desc pointers (Src/cMatcher.cc:20) into one n x 768 array."""
    if isinstance(kps, np.ndarray) and kps.dtype == KP_DTYPE:
        n = len(kps)
        out = np.empty((n, DESC_LENGTH), np.float32)
        if n == 0:
            return out
        p = kps["desc"].astype(np.uint64)
        own = getattr(kps, "_desc_owner", None)
        if own is None:
            raise S3DError("keypoint records carry raw `desc` addresses but no owner of the descriptor block: pass the "
                           "array returned by GetKeypoints() (or a slice / copy of it), or an n x 768 float32 matrix")
        lo, hi = own.ctypes.data, own.ctypes.data + own.nbytes
        if int(p.min()) < lo or int(p.max()) + DESC_LENGTH * 4 > hi:
            raise S3DError("keypoint `desc` addresses lie outside the descriptor block they were bound to")
        if np.all(np.diff(p.astype(np.int64)) == DESC_LENGTH * 4):  # base + 768*i: one block copy
            C.memmove(out.ctypes.data, int(p[0]), n * DESC_LENGTH * 4)
        else:
            for i in range(n):
                C.memmove(out.ctypes.data + i * DESC_LENGTH * 4, int(p[i]), DESC_LENGTH * 4)
        return out
    return np.ascontiguousarray(kps, dtype=np.float32).reshape(-1, DESC_LENGTH)


class muBruteMatcher:
    """CPUSIFT::muBruteMatcher (Include/cMatcher.h:12-88)."""

    inject, biject, enhanced = 1, 2, 3

    def __init__(self):
        self.matchTime = self.revMatchTime = self.totalTime = 0.0
        self._r = None

    def _run(self, mtype, ref_kp, tar_kp, thresHold):
        ref = _desc_matrix(ref_kp)
        tar = _desc_matrix(tar_kp)
        n_ref, n_tar = len(ref), len(tar)
        A = lambda n, t: np.zeros(max(n, 1), t)
        r = dict(gIdx=A(n_ref, np.int32), gDist=A(n_ref, np.float32), sIdx=A(n_ref, np.int32), sDist=A(n_ref, np.float32),
                 gIdx2=A(n_tar, np.int32), gDist2=A(n_tar, np.float32), sIdx2=A(n_tar, np.int32), sDist2=A(n_tar, np.float32),
                 pr=A(n_ref, np.int32), pt=A(n_ref, np.int32))
        npairs = C.c_int()
        times = (C.c_double * 3)()
        check(lib().s3d_match(mtype, _ptr(ref), n_ref, _ptr(tar), n_tar, float(thresHold), _ptr(r["gIdx"]), _ptr(r["gDist"]),
                              _ptr(r["sIdx"]), _ptr(r["sDist"]), _ptr(r["gIdx2"]), _ptr(r["gDist2"]), _ptr(r["sIdx2"]),
                              _ptr(r["sDist2"]), _ptr(r["pr"]), _ptr(r["pt"]), C.cast(C.byref(npairs), C.c_void_p),
                              C.cast(C.byref(times), C.c_void_p)))
        n = npairs.value
        self.matchTime, self.revMatchTime, self.totalTime = times[0], times[1], times[2]
        self._r = {k: v[: (n_ref if not k.endswith("2") else n_tar)] for k, v in r.items() if k not in ("pr", "pt")}
        self.pairs = np.stack([r["pr"][:n], r["pt"][:n]], 1)
        # toCvec (Src/cMatcher.cc:99-112): coordinates (rx, ry, rz) of the matched pairs
        if isinstance(ref_kp, np.ndarray) and ref_kp.dtype == KP_DTYPE and isinstance(tar_kp, np.ndarray) and tar_kp.dtype == KP_DTYPE:
            rm = np.stack([ref_kp[c][self.pairs[:, 0]] for c in ("rx", "ry", "rz")], 1) if n else np.zeros((0, 3), np.float32)
            tm = np.stack([tar_kp[c][self.pairs[:, 1]] for c in ("rx", "ry", "rz")], 1) if n else np.zeros((0, 3), np.float32)
            return rm, tm
        return self.pairs[:, 0].copy(), self.pairs[:, 1].copy()

    def injectMatch(self, ref_kp, tar_kp, thresHold=0.85):
        return self._run(self.inject, ref_kp, tar_kp, thresHold)

    def bijectMatch(self, ref_kp, tar_kp, thresHold=0.85):
        return self._run(self.biject, ref_kp, tar_kp, thresHold)

    def enhancedMatch(self, ref_kp, tar_kp, thresHold=0.85):
        return self._run(self.enhanced, ref_kp, tar_kp, thresHold)

    def getCalculationTime(self):
        return self.totalTime

    def getGlodenDistSquare(self):
        return self._r["gDist"]

    def getSilverDistSquare(self):
        return self._r["sDist"]

    def getGlodenIdx(self):
        return self._r["gIdx"]

    def getSilverIdx(self):
        return self._r["sIdx"]

    def reverse(self):
        return {k: self._r[k] for k in ("gIdx2", "gDist2", "sIdx2", "sDist2")}


# ---- free kernels (Include/cSIFT3D.h:208-239) ----------------------------------------------------

def GaussianSmooth_3D(vol, sigma):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, ny, nx = vol.shape
    out = np.empty_like(vol)
    check(lib().s3d_gaussian_smooth(_ptr(vol), nx, ny, nz, float(sigma), _ptr(out)))
    return out


def blur_axis(vol, axis, w, hw, variant=1):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    nz, ny, nx = vol.shape
    out = np.empty_like(vol)
    check(lib().s3d_blur_axis(_ptr(vol), nx, ny, nz, axis, _ptr(w), hw, variant, _ptr(out)))
    return out


def DownSample_3D(vol):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, ny, nx = vol.shape
    out = np.empty((nz // 2, ny // 2, nx // 2), np.float32)
    check(lib().s3d_downsample(_ptr(vol), nx, ny, nz, _ptr(out)))
    return out
