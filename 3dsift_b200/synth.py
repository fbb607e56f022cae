"""Seeded synthetic inputs shared by tests/, bench.py and the oracle drivers.

The reference ships no data (README.md:66 points at an external download), so every parity and
benchmark input is generated here, byte-identically for the CPU oracle and the GPU path.

* ``v_blobs(n, seed)``       — float32 n^3 volume: isotropic Gaussian blobs + N(0, 0.01^2) noise.
* ``v_blobs_pair(n, seed)``  — (ref, tar): tar is the same blob list under a small rigid motion
                               (<=10 deg rotation about a random axis, <=3 voxel translation) with
                               an independent noise seed.  Isotropic blobs stay isotropic under a
                               rigid motion, so the pair is rendered analytically (no resampling).
* ``v_ct(n, seed)``          — CT-like: piecewise-constant ellipsoids + blob texture + noise, >= 0.
* ``d_synth(k, seed)``       — k x 768 descriptor set shaped like the reference's output
                               (cSIFT3D.cc:1350-1358): non-negative, ~25 % non-zeros, L2-normalised,
                               clamped at 0.2*128/768, renormalised.
* ``d_synth_pair(k, seed)``  — (ref, tar): tar = permuted noisy copy for 70 % of rows + 30 % fresh.
"""
from __future__ import annotations

import numpy as np

TRUNC = np.float32(0.2 * 128 / 768)


def _render_blobs(shape, centers, sigmas, amps, out=None):
    nz, ny, nx = shape
    vol = np.zeros(shape, dtype=np.float32) if out is None else out
    for (cx, cy, cz), s, a in zip(centers, sigmas, amps):
        r = int(np.ceil(3.5 * s))
        x0, x1 = max(0, int(np.floor(cx)) - r), min(nx, int(np.floor(cx)) + r + 2)
        y0, y1 = max(0, int(np.floor(cy)) - r), min(ny, int(np.floor(cy)) + r + 2)
        z0, z1 = max(0, int(np.floor(cz)) - r), min(nz, int(np.floor(cz)) + r + 2)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            continue
        inv = np.float32(-0.5 / (s * s))
        gx = np.exp(inv * (np.arange(x0, x1, dtype=np.float32) - np.float32(cx)) ** 2)
        gy = np.exp(inv * (np.arange(y0, y1, dtype=np.float32) - np.float32(cy)) ** 2)
        gz = np.exp(inv * (np.arange(z0, z1, dtype=np.float32) - np.float32(cz)) ** 2)
        vol[z0:z1, y0:y1, x0:x1] += (np.float32(a) * gz)[:, None, None] * gy[None, :, None] * gx[None, None, :]
    return vol


def _blob_list(shape, seed):
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    k = int(np.ceil(nx * ny * nz / 4096.0))
    centers = rng.uniform(0, 1, size=(k, 3)) * np.array([nx, ny, nz])
    sigmas = rng.uniform(1.5, 4.5, size=k)
    amps = rng.uniform(0.3, 1.3, size=k)
    return centers, sigmas, amps


def _as_shape(n):
    if isinstance(n, (tuple, list)):
        nx, ny, nz = n
        return (int(nz), int(ny), int(nx))
    return (int(n), int(n), int(n))


def _noise(shape, seed, sd=0.01):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(size=shape, dtype=np.float32) * np.float32(sd)


def v_blobs(n, seed=0):
    """float32 volume indexed [z, y, x] (x fastest — the reference's layout, cTexImage.cc:28-30).
    ``n`` is an int (cube) or (nx, ny, nz)."""
    shape = _as_shape(n)
    c, s, a = _blob_list(shape, seed)
    vol = _render_blobs(shape, c, s, a)
    vol += _noise(shape, seed + 1_000_003)
    return np.ascontiguousarray(vol)


def _rigid(seed, shape, max_deg=10.0, max_shift=3.0):
    rng = np.random.default_rng(seed + 77)
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    t = rng.uniform(-max_shift, max_shift, size=3)
    nz, ny, nx = shape
    ctr = np.array([nx, ny, nz]) / 2.0
    return R, t, ctr


def v_blobs_pair(n, seed=0):
    shape = _as_shape(n)
    c, s, a = _blob_list(shape, seed)
    ref = _render_blobs(shape, c, s, a)
    ref += _noise(shape, seed + 1_000_003)
    R, t, ctr = _rigid(seed, shape)
    c2 = (c - ctr) @ R.T + ctr + t
    tar = _render_blobs(shape, c2, s, a)
    tar += _noise(shape, seed + 2_000_003)
    return np.ascontiguousarray(ref), np.ascontiguousarray(tar)


def v_ct(n, seed=0):
    shape = _as_shape(n)
    nz, ny, nx = shape
    rng = np.random.default_rng(seed + 5)
    vol = np.zeros(shape, dtype=np.float32)
    zz = np.arange(nz, dtype=np.float32)[:, None, None]
    yy = np.arange(ny, dtype=np.float32)[None, :, None]
    xx = np.arange(nx, dtype=np.float32)[None, None, :]
    dens = [0.2, 0.5, 1.0]
    for i in range(12):
        cx, cy, cz = rng.uniform(0.2, 0.8, 3) * np.array([nx, ny, nz])
        rx, ry, rz = rng.uniform(0.08, 0.3, 3) * np.array([nx, ny, nz])
        m = ((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 + ((zz - cz) / rz) ** 2 <= 1.0
        vol[m] = np.float32(dens[i % 3])
    c, s, a = _blob_list(shape, seed)
    tex = _render_blobs(shape, c, s, a * 0.15)
    vol += tex
    vol += _noise(shape, seed + 3_000_003)
    np.maximum(vol, 0, out=vol)
    return np.ascontiguousarray(vol)


def _finish_desc(d):
    d = d.astype(np.float32)
    d /= (np.linalg.norm(d, axis=1, keepdims=True) + np.float32(1e-12))
    np.minimum(d, TRUNC, out=d)
    d /= (np.linalg.norm(d, axis=1, keepdims=True) + np.float32(1e-12))
    return np.ascontiguousarray(d.astype(np.float32))


def d_synth(k, seed=0, chunk=65536):
    rng = np.random.default_rng(seed)
    out = np.empty((k, 768), dtype=np.float32)
    for s in range(0, k, chunk):
        e = min(k, s + chunk)
        v = rng.standard_exponential(size=(e - s, 768), dtype=np.float32)
        v *= rng.random(size=(e - s, 768), dtype=np.float32) < 0.25
        out[s:e] = _finish_desc(v)
    return out


def d_synth_pair(k, seed=0, k_tar=None):
    """Returns (ref, tar, truth) where truth[i] = index in tar of ref row i's noisy copy or -1."""
    k_tar = k if k_tar is None else k_tar
    ref = d_synth(k, seed)
    rng = np.random.default_rng(seed + 99)
    n_copy = min(int(0.7 * k_tar), k)
    src = rng.permutation(k)[:n_copy]
    tar = d_synth(k_tar, seed + 12345)
    pos = rng.permutation(k_tar)[:n_copy]
    noisy = ref[src] * (1.0 + 0.1 * rng.standard_normal(size=(n_copy, 768), dtype=np.float32))
    np.maximum(noisy, 0, out=noisy)
    tar[pos] = _finish_desc(noisy)
    truth = np.full(k, -1, dtype=np.int64)
    truth[src] = pos
    return ref, np.ascontiguousarray(tar), truth


def d_synth_pair_device(k, seed=0, k_tar=None, device="cuda", chunk=65536):
    """D-synth(K) generated on the GPU with torch (same recipe as d_synth_pair, different random
    stream): for the matching sizes where building 2 x K x 768 floats with numpy would dominate a
    benchmark's wall time.  Returns (ref, tar, truth) as device tensors."""
    import torch
    k_tar = k if k_tar is None else k_tar
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def finish(d):
        d = d / (d.norm(dim=1, keepdim=True) + 1e-12)
        d = torch.clamp(d, max=float(TRUNC))
        return d / (d.norm(dim=1, keepdim=True) + 1e-12)

    def synth(n):
        out = torch.empty((n, 768), dtype=torch.float32, device=device)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            v = torch.empty((e - s, 768), dtype=torch.float32, device=device).exponential_(1.0, generator=g)
            v *= torch.rand((e - s, 768), device=device, generator=g) < 0.25
            out[s:e] = finish(v)
        return out

    ref, tar = synth(k), synth(k_tar)
    n_copy = min(int(0.7 * k_tar), k)
    src = torch.randperm(k, device=device, generator=g)[:n_copy]
    pos = torch.randperm(k_tar, device=device, generator=g)[:n_copy]
    for s in range(0, n_copy, chunk):
        e = min(n_copy, s + chunk)
        noisy = ref[src[s:e]] * (1.0 + 0.1 * torch.randn((e - s, 768), device=device, generator=g))
        tar[pos[s:e]] = finish(torch.clamp(noisy, min=0))
    truth = torch.full((k,), -1, dtype=torch.int64, device=device)
    truth[src] = pos
    return ref, tar, truth
