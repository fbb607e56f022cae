"""ctypes bindings for the two CHECKERS — test infrastructure only.

* ``Ref``  -> oracle/_ref/libsift3d_ref.so : the unmodified reference compiled by oracle/Makefile
             (kind "reference").
* ``Port`` -> oracle/libsift3d_oracle.so   : the C restatement sift3d_oracle.c (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product (3dsift_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsift3d_ref.so")
PORT_SO = os.path.join(HERE, "libsift3d_oracle.so")

KP_DTYPE = np.dtype(
    [("x", "f4"), ("y", "f4"), ("z", "f4"), ("scale", "f4"), ("octave", "i4"), ("level", "i4"),
     ("rx", "f4"), ("ry", "f4"), ("rz", "f4"), ("win", "f4", 3), ("eigvalue", "f4", 3),
     ("eigvector", "f4", 9), ("Rotation", "f4", 9), ("str_tensor", "f4", 9), ("desc", "u8")],
    align=True)
assert KP_DTYPE.itemsize == 176

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
DEFAULTS = dict(levels=3, sigma=1.6, sigma_n=1.15, peak=0.1, eig=0.9, corner=0.4)  # cSIFT3D.h:13-21


def build(which="all"):
    """(Re)build the checkers.  `ref` is only rebuilt where /root/reference exists."""
    subprocess.check_call(["make", "-s", "-C", HERE, which])


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


class _Extraction:
    """Result bundle common to both checkers."""

    def __init__(self):
        self.noct = 0
        self.dims = []          # per octave (nx, ny, nz)
        self.keypoints = None   # KP_DTYPE array (survivors, raster order)
        self.desc = None        # K x 768
        self.extrema = None     # KP_DTYPE array of raw detections after orientation
        self.times = {}
        self._level = None

    def gss(self, idx):
        return self._level(0, idx)

    def dog(self, idx):
        return self._level(1, idx)


class Ref:
    """The compiled reference (oracle/_ref)."""

    kind = "reference"

    def __init__(self):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libsift3d_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(REF_SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_run.argtypes = [C.c_void_p]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_time.restype = C.c_double
        L.ref_time.argtypes = [C.c_void_p, C.c_int]
        for name in ("ref_num_octaves", "ref_num_extrema", "ref_num_keypoints", "ref_num_level_extrema"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_level_info.argtypes = [C.c_void_p, C.c_int, C.c_int, _i, _f]
        L.ref_copy_level.argtypes = [C.c_void_p, C.c_int, C.c_int, _f]
        L.ref_copy_input.argtypes = [C.c_void_p, _f]
        L.ref_copy_extrema.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_copy_level_extrema.argtypes = [C.c_void_p, _i]
        L.ref_copy_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_gaussian_smooth.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_float, _f]
        L.ref_downsample.argtypes = [_f, C.c_int, C.c_int, C.c_int, _f]
        L.ref_mesh.argtypes = [_f, _i]
        L.ref_orient_describe.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.ref_match.argtypes = [C.c_int, _f, C.c_int, _f, C.c_int, C.c_double] + [C.c_void_p] * 7
        assert L.ref_sizeof_keypoint() == 176 and L.ref_sizeof_cvec() == 12

    def threads(self):
        return int(self.lib.ref_max_threads())

    def set_threads(self, n):
        """sift_thread_num + omp_set_num_threads (the reference's SetNumThreads, Src/cSIFT3D.cc:1680-1684)."""
        self.lib.ref_set_threads.argtypes = [C.c_int]
        self.lib.ref_set_threads.restype = None
        self.lib.ref_set_threads(int(n))

    def extract(self, vol, keep_levels=True, **params):
        p = dict(DEFAULTS, **params)
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        L = self.lib
        h = L.ref_create(vol, nx, ny, nz, p["levels"], p["sigma"], p["sigma_n"], p["peak"], p["eig"], p["corner"])
        L.ref_run(h)
        r = _Extraction()
        r.noct = L.ref_num_octaves(h)
        G = p["levels"] + 3
        meta = np.zeros(4, np.float32)
        d = np.zeros(3, np.int32)
        for o in range(r.noct):
            L.ref_level_info(h, 0, o * G, d, meta)
            r.dims.append(tuple(int(v) for v in d))
        r.input = np.empty_like(vol)
        L.ref_copy_input(h, r.input)
        n = L.ref_num_extrema(h)
        r.extrema = np.zeros(n, KP_DTYPE)
        if n:
            L.ref_copy_extrema(h, _fp(r.extrema))
        n = L.ref_num_level_extrema(h)
        r.level_extrema = np.zeros((n, 5), np.int32)
        if n:
            L.ref_copy_level_extrema(h, r.level_extrema)
        k = L.ref_num_keypoints(h)
        r.keypoints = np.zeros(k, KP_DTYPE)
        r.desc = np.zeros((k, 768), np.float32)
        if k:
            L.ref_copy_keypoints(h, _fp(r.keypoints), _fp(r.desc))
        r.times = {n_: L.ref_time(h, i) for i, n_ in enumerate(
            ["create", "run", "alloc", "gss", "dog", "detect", "orient", "desc"])}
        levels = {}
        if keep_levels:
            for which, per in ((0, G), (1, G - 1)):
                for idx in range(r.noct * per):
                    L.ref_level_info(h, which, idx, d, meta)
                    a = np.empty((int(d[2]), int(d[1]), int(d[0])), np.float32)
                    L.ref_copy_level(h, which, idx, a)
                    levels[(which, idx)] = a
        r._level = lambda which, idx: levels[(which, idx)]
        L.ref_destroy(h)
        return r

    def time_extract(self, vol, **params):
        """Wall time of CreateCSIFT3D + KpSiftAlgorithm (seconds) and the keypoint count."""
        r = self.extract(vol, keep_levels=False, **params)
        return r.times["create"] + r.times["run"], len(r.keypoints), r.times

    def gaussian_smooth(self, vol, sigma):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        out = np.empty_like(vol)
        self.lib.ref_gaussian_smooth(vol, nx, ny, nz, float(sigma), out)
        return out

    def downsample(self, vol):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        out = np.empty((nz // 2, ny // 2, nx // 2), np.float32)
        self.lib.ref_downsample(vol, nx, ny, nz, out)
        return out

    def mesh(self):
        v = np.zeros((20, 3, 3), np.float32)
        idx = np.zeros((20, 3), np.int32)
        self.lib.ref_mesh(v, idx)
        return v, idx

    def orient_describe(self, level, unit, x, y, z, scale, octave=0, eig=0.9, corner=0.4, want_desc=True):
        level = np.ascontiguousarray(level, dtype=np.float32)
        nz, ny, nx = level.shape
        kp = np.zeros(1, KP_DTYPE)
        kp["x"], kp["y"], kp["z"], kp["scale"], kp["octave"] = x, y, z, scale, octave
        desc = np.zeros(768, np.float32)
        res = self.lib.ref_orient_describe(level, nx, ny, nz, float(unit), _fp(kp), _fp(desc) if want_desc else None,
                                           eig, corner)
        return res, kp[0], desc

    def match(self, mtype, ref, tar, thr=0.85):
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        tar = np.ascontiguousarray(tar, dtype=np.float32)
        n_ref, n_tar = len(ref), len(tar)
        g = np.zeros(max(n_ref, 1), np.int32); s = np.zeros(max(n_ref, 1), np.int32)
        gd = np.zeros(max(n_ref, 1), np.float32); sd = np.zeros(max(n_ref, 1), np.float32)
        pr = np.zeros(max(n_ref, 1), np.int32); pt = np.zeros(max(n_ref, 1), np.int32)
        times = np.zeros(3, np.float64)
        n = self.lib.ref_match(mtype, ref, n_ref, tar, n_tar, float(thr), _fp(g), _fp(gd), _fp(s), _fp(sd),
                               _fp(pr), _fp(pt), _fp(times))
        return dict(gIdx=g[:n_ref], gDist=gd[:n_ref], sIdx=s[:n_ref], sDist=sd[:n_ref],
                    pairs=np.stack([pr[:n], pt[:n]], 1), times=times)


class Port:
    """The C restatement (oracle/sift3d_oracle.c)."""

    kind = "port"

    def __init__(self):
        if not have_port():
            build("port")
        L = self.lib = C.CDLL(PORT_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [_f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
        for name in ("orc_run", "orc_destroy", "orc_noct", "orc_num_extrema", "orc_num_keypoints"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.orc_level_dims.argtypes = [C.c_void_p, C.c_int, _i]
        L.orc_copy_level.argtypes = [C.c_void_p, C.c_int, C.c_int, _f]
        L.orc_copy_input.argtypes = [C.c_void_p, _f]
        L.orc_copy_extrema.argtypes = [C.c_void_p, C.c_void_p, _i]
        L.orc_copy_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gaussian_smooth.argtypes = [_f, _f, C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_blur_axis.argtypes = [_f, _f, C.c_int, C.c_int, C.c_int, C.c_int, _f, C.c_int]
        L.orc_gauss_kernel.argtypes = [C.c_float, _f]
        L.orc_sigmas.argtypes = [C.c_int, C.c_float, C.c_float, _f]
        L.orc_level_scale.restype = C.c_float
        L.orc_level_scale.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        L.orc_downsample.argtypes = [_f, C.c_int, C.c_int, C.c_int, _f]
        L.orc_mesh_flat.argtypes = [_f, _i]
        L.orc_max_abs.restype = C.c_float
        L.orc_max_abs.argtypes = [_f, C.c_size_t]
        L.orc_detect_level.argtypes = [_f, _f, _f, C.c_int, C.c_int, C.c_int, C.c_float, _i, C.c_int, C.c_void_p]
        L.orc_orient.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_describe.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, C.c_int, C.c_float, _f]
        L.orc_match.argtypes = [C.c_int, _f, C.c_int, _f, C.c_int, C.c_double] + [C.c_void_p] * 10
        L.orc_cal_matches.argtypes = [_f, C.c_int, _f, C.c_int, C.c_void_p, _f, _f, _i, _i]
        assert L.orc_sizeof_keypoint() == 176

    def threads(self):
        return os.cpu_count() or 1

    def sigmas(self, levels=3, sigma=1.6, sigma_n=1.15):
        s = np.zeros(levels + 3, np.float32)
        self.lib.orc_sigmas(levels, sigma, sigma_n, s)
        return s

    def gauss_kernel(self, sigma):
        w = np.zeros(64, np.float32)
        hw = self.lib.orc_gauss_kernel(float(sigma), w)
        return w[: 2 * hw + 1].copy(), hw

    def level_scale(self, o, s, levels=3, sigma=1.6):
        return float(self.lib.orc_level_scale(o, s, levels, sigma))

    def blur_axis(self, vol, axis, w, hw):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        out = np.empty_like(vol)
        self.lib.orc_blur_axis(vol, out, nx, ny, nz, axis, np.ascontiguousarray(w, np.float32), hw)
        return out

    def gaussian_smooth(self, vol, sigma):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        out = np.empty_like(vol)
        self.lib.orc_gaussian_smooth(vol, out, nx, ny, nz, float(sigma))
        return out

    def downsample(self, vol):
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        out = np.empty((nz // 2, ny // 2, nx // 2), np.float32)
        self.lib.orc_downsample(vol, nx, ny, nz, out)
        return out

    def mesh(self):
        v = np.zeros((20, 3, 3), np.float32)
        idx = np.zeros((20, 3), np.int32)
        self.lib.orc_mesh_flat(v, idx)
        return v, idx

    def detect_level(self, prev, cur, nxt, peak=0.1):
        nz, ny, nx = cur.shape
        cap = cur.size
        out = np.zeros((cap, 3), np.int32)
        n = self.lib.orc_detect_level(np.ascontiguousarray(prev), np.ascontiguousarray(cur), np.ascontiguousarray(nxt),
                                      nx, ny, nz, peak, out, cap, None)
        return out[:n].copy()

    def orient_describe(self, level, unit, x, y, z, scale, octave=0, eig=0.9, corner=0.4, want_desc=True):
        level = np.ascontiguousarray(level, dtype=np.float32)
        nz, ny, nx = level.shape
        kp = np.zeros(1, KP_DTYPE)
        kp["x"], kp["y"], kp["z"], kp["scale"], kp["octave"] = x, y, z, scale, octave
        kp["rx"] = kp["ry"] = kp["rz"] = -1
        desc = np.zeros(768, np.float32)
        res = self.lib.orc_orient(_fp(kp), level, nx, ny, nz, float(unit), eig, corner)
        if res == 1 and want_desc:
            self.lib.orc_describe(_fp(kp), level, nx, ny, nz, float(unit), desc)
        return res, kp[0], desc

    def extract(self, vol, keep_levels=True, **params):
        p = dict(DEFAULTS, **params)
        vol = np.ascontiguousarray(vol, dtype=np.float32)
        nz, ny, nx = vol.shape
        L = self.lib
        h = L.orc_create(vol, nx, ny, nz, p["levels"], p["sigma"], p["sigma_n"], p["peak"], p["eig"], p["corner"])
        L.orc_run(h)
        r = _Extraction()
        r.noct = L.orc_noct(h)
        G = p["levels"] + 3
        d = np.zeros(3, np.int32)
        for o in range(r.noct):
            L.orc_level_dims(h, o, d)
            r.dims.append(tuple(int(v) for v in d))
        r.input = np.empty_like(vol)
        L.orc_copy_input(h, r.input)
        n = L.orc_num_extrema(h)
        r.extrema = np.zeros(n, KP_DTYPE)
        r.ret = np.zeros(max(n, 1), np.int32)
        if n:
            L.orc_copy_extrema(h, _fp(r.extrema), r.ret)
        r.ret = r.ret[:n]
        k = L.orc_num_keypoints(h)
        r.keypoints = np.zeros(k, KP_DTYPE)
        r.desc = np.zeros((k, 768), np.float32)
        if k:
            L.orc_copy_keypoints(h, _fp(r.keypoints), _fp(r.desc))
        levels = {}
        if keep_levels:
            for which, per in ((0, G), (1, G - 1)):
                for idx in range(r.noct * per):
                    nxo, nyo, nzo = r.dims[idx // per]
                    a = np.empty((nzo, nyo, nxo), np.float32)
                    L.orc_copy_level(h, which, idx, a)
                    levels[(which, idx)] = a
        r._level = lambda which, idx: levels[(which, idx)]
        L.orc_destroy(h)
        return r

    def match(self, mtype, ref, tar, thr=0.85):
        ref = np.ascontiguousarray(ref, dtype=np.float32)
        tar = np.ascontiguousarray(tar, dtype=np.float32)
        n_ref, n_tar = len(ref), len(tar)
        A = lambda n, t: np.zeros(max(n, 1), t)
        g, s, gd, sd = A(n_ref, np.int32), A(n_ref, np.int32), A(n_ref, np.float32), A(n_ref, np.float32)
        g2, s2, gd2, sd2 = A(n_tar, np.int32), A(n_tar, np.int32), A(n_tar, np.float32), A(n_tar, np.float32)
        pr, pt = A(n_ref, np.int32), A(n_ref, np.int32)
        n = self.lib.orc_match(mtype, ref, n_ref, tar, n_tar, float(thr), _fp(g), _fp(gd), _fp(s), _fp(sd),
                               _fp(g2), _fp(gd2), _fp(s2), _fp(sd2), _fp(pr), _fp(pt))
        return dict(gIdx=g[:n_ref], gDist=gd[:n_ref], sIdx=s[:n_ref], sDist=sd[:n_ref],
                    gIdx2=g2[:n_tar], gDist2=gd2[:n_tar], sIdx2=s2[:n_tar], sDist2=sd2[:n_tar],
                    pairs=np.stack([pr[:n], pt[:n]], 1))


def best():
    """The strongest checker available: the compiled reference if present, else the port."""
    return Ref() if have_ref() else Port()
