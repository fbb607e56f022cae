"""GPU parity of the sparse stages (orientation, descriptors, GetKeypoints) through the C ABI.
Bars (BASELINE.json north_star): accept/reject mask and order bit-exact with any flips
enumerated; coordinates within 1e-3 voxel (they are integers: exact); descriptor cosine >= 0.9999."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cos(a, b):
    """Row-wise cosine; two all-zero descriptors (a window whose gradients all fall below the
    reference's |grad|^2 >= 1.19e-6 floor, Src/cSIFT3D.cc:1544) count as equal."""
    na, nb = np.linalg.norm(a, axis=1), np.linalg.norm(b, axis=1)
    both_zero = (na == 0) & (nb == 0)
    den = np.where(both_zero, 1.0, na * nb)
    return np.where(both_zero, 1.0, (a * b).sum(1) / den)


def _compare_sparse(s3d, r, sift, desc_cos=0.9999):
    kp, codes, xyz5 = sift.extrema()
    assert len(kp) == len(r.extrema)
    ref_rej = r.extrema["x"] < 0
    flips = np.flatnonzero((codes != 1) != ref_rej)
    # enumerate (the north star asks for flips to be listed, not hidden)
    for i in flips:
        print(f"orientation flip at detection {i}: xyz5={xyz5[i]} gpu code={codes[i]} eig={kp['eigvalue'][i]} "
              f"ref eig={r.extrema['eigvalue'][i]}")
    assert len(flips) == 0, f"{len(flips)} accept/reject flips of {len(kp)}"
    # debug fields carried on every keypoint (Include/cSIFT3D.h:60-67)
    # FP32 sums over ~1e4 voxels in a different order: compare relative to each tensor's scale
    tscale = np.abs(r.extrema["str_tensor"]).max(1, keepdims=True) + 1e-30
    assert (np.abs(kp["str_tensor"] - r.extrema["str_tensor"]) / tscale).max() < 2e-4
    wscale = np.abs(r.extrema["win"]).max(1, keepdims=True) + 1e-7
    assert (np.abs(kp["win"] - r.extrema["win"]) / wscale).max() < 2e-3
    ok = codes != -1            # eigen fields are only written past the weak-gradient test
    # eigenvalues of an FP32-summed tensor are conditioned by the tensor's scale (Weyl), not by their own size: the
    # smallest one of a nearly singular tensor has no relative accuracy
    escale = np.abs(r.extrema["eigvalue"][ok]).max(1, keepdims=True) + 1e-30
    assert (np.abs(kp["eigvalue"][ok] - r.extrema["eigvalue"][ok]) / escale).max() < 5e-4
    ev_g = kp["eigvector"][ok].reshape(-1, 3, 3)
    ev_r = r.extrema["eigvector"][ok].reshape(-1, 3, 3)
    sep = np.abs(np.diff(r.extrema["eigvalue"][ok], axis=1)).min(1) / np.abs(r.extrema["eigvalue"][ok]).max(1)
    good = sep > 1e-2           # eigenvectors are only well conditioned for separated eigenvalues
    dots = np.abs((ev_g * ev_r).sum(2))[good]   # pre-sign-fix vectors: equal up to sign (Q11)
    assert dots.min() > 1 - 1e-3

    kps = sift.GetKeypoints()
    desc = sift.descriptors
    assert len(kps) == len(r.keypoints)
    for f in ("x", "y", "z", "rx", "ry", "rz", "scale", "octave", "level"):
        assert np.array_equal(kps[f], r.keypoints[f]), f
    np.testing.assert_allclose(kps["Rotation"], r.keypoints["Rotation"], atol=2e-3)   # transposed (Q11)
    if len(kps):
        cos = _cos(desc, r.desc)
        print(f"descriptors: n={len(kps)} cos min={cos.min():.7f} mean={cos.mean():.7f} max|diff|={np.abs(desc - r.desc).max():.3g}")
        assert cos.min() >= desc_cos
        nrm = np.linalg.norm(desc, axis=1)
        assert np.allclose(nrm[nrm > 0], 1.0, atol=1e-4) and np.array_equal(nrm == 0, np.linalg.norm(r.desc, axis=1) == 0)
        assert (desc >= 0).all()
        # desc pointers borrow from the extractor-owned contiguous block (Q17)
        assert np.array_equal(np.diff(kps["desc"].astype(np.int64)), np.full(len(kps) - 1, 768 * 4))
    return len(kps)


@pytest.mark.parametrize("shape,seed", [((64, 64, 64), 1), ((48, 40, 32), 3), ((72, 64, 80), 4)])
def test_keypoints_and_descriptors(s3d, synth, checker, shape, seed):
    vol = synth.v_blobs(shape, seed=seed)
    r = checker.extract(vol, keep_levels=False)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    sift.KpSiftAlgorithm()
    _compare_sparse(s3d, r, sift)


def test_config0_128_cube_pair(s3d, synth, checker):
    """BASELINE.json configs[0]: 128^3 blobs+noise volume and a rotated/translated copy,
    KpSiftAlgorithm on both + enhancedMatch(0.85), against the reference's own path."""
    ref_v, tar_v = synth.v_blobs_pair(128, seed=0)
    out = []
    for v in (ref_v, tar_v):
        r = checker.extract(v, keep_levels=False)
        sift = s3d.CSIFT3DFactory.CreateCSIFT3D(v)
        sift.KpSiftAlgorithm()
        n = _compare_sparse(s3d, r, sift)
        assert n > 100
        out.append((r, sift, sift.GetKeypoints()))
    (r0, s0, k0), (r1, s1, k1) = out
    m = s3d.muBruteMatcher()
    rm, tm = m.enhancedMatch(k0, k1, 0.85)
    # the oracle matcher on the ORACLE's descriptors: match index lists must be identical
    want = checker.match(3, r0.desc, r1.desc, 0.85)
    assert np.array_equal(m.pairs, want["pairs"]), (len(m.pairs), len(want["pairs"]))
    assert len(m.pairs) > 20
    assert np.array_equal(rm, np.stack([k0[c][m.pairs[:, 0]] for c in ("rx", "ry", "rz")], 1))
    print("config0: keypoints", len(k0), len(k1), "matches", len(m.pairs))


def test_extract_matches_golden_sparse(s3d, oracle_mod):
    g = np.load(os.path.join(GOLD, "extract.npz"))
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(g["vol"])
    sift.KpSiftAlgorithm()
    gk = np.ascontiguousarray(g["keypoints"]).view(oracle_mod.KP_DTYPE).reshape(-1)
    ge = np.ascontiguousarray(g["extrema"]).view(oracle_mod.KP_DTYPE).reshape(-1)
    kp, codes, _ = sift.extrema()
    assert np.array_equal(codes != 1, ge["x"] < 0)
    kps = sift.GetKeypoints()
    assert len(kps) == len(gk)
    for f in ("x", "y", "z", "rx", "ry", "rz", "octave", "level"):
        assert np.array_equal(kps[f], gk[f])
    assert _cos(sift.descriptors, g["desc"]).min() >= 0.9999


def test_exact_recheck_off_still_close(s3d, synth, checker):
    vol = synth.v_blobs(64, seed=6)
    r = checker.extract(vol, keep_levels=False)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol, exact_recheck=False)
    sift.KpSiftAlgorithm()
    _, codes, _ = sift.extrema()
    assert ((codes != 1) != (r.extrema["x"] < 0)).sum() <= 1


def test_device_resident_input(s3d, synth):
    torch = pytest.importorskip("torch")
    vol = synth.v_blobs(64, seed=8)
    a = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    a.KpSiftAlgorithm()
    b = s3d.CSIFT3DFactory.CreateCSIFT3D(torch.from_numpy(vol).cuda())
    b.KpSiftAlgorithm()
    ka, kb = a.GetKeypoints(), b.GetKeypoints()
    assert len(ka) == len(kb) and np.array_equal(a.descriptors, b.descriptors)   # run-to-run deterministic


def test_run_is_deterministic(s3d, synth):
    vol = synth.v_ct(64, seed=1)
    outs = []
    for _ in range(2):
        s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
        s.KpSiftAlgorithm()
        k = s.GetKeypoints()
        outs.append((k["x"].copy(), s.descriptors.copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.fixture()
def desc_path_reset(s3d):
    yield
    s3d.set_describe_path(s3d.api.DESC_FIXED)


@pytest.mark.parametrize("path", ["fixed", "fp32", "forced_redo"])
def test_descriptor_paths_meet_the_tolerance(s3d, synth, checker, desc_path_reset, path):
    """The fixed-point (integer shared-memory atomics) descriptor kernel, the FP32 ordered kernel and
    the redo hand-over between them all meet cosine >= 0.9999 against the reference."""
    s3d.set_describe_path({"fixed": s3d.api.DESC_FIXED, "fp32": s3d.api.DESC_FP32, "forced_redo": s3d.api.DESC_FORCE_REDO}[path])
    vol = synth.v_blobs((72, 64, 80), seed=11)
    r = checker.extract(vol, keep_levels=False)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    sift.KpSiftAlgorithm()
    n = _compare_sparse(s3d, r, sift)
    redo = sift.counters()["desc_redo"]
    print(f"describe path {path}: {n} keypoints, {redo} redone in FP32")
    if path == "fixed":
        assert redo <= max(1, n // 20)          # the scale estimate normally holds
    elif path == "fp32":
        assert redo == 0
    else:
        assert redo >= n // 2                   # the tiny margin really exercises the hand-over


def test_fixed_point_descriptor_on_low_contrast_volume(s3d, synth, checker, desc_path_reset):
    """One bright outlier voxel sets the normalisation, so every gradient in the rest of the volume
    is a few percent of full scale — just above the reference's own |grad|^2 >= 1.19e-6 floor
    (Src/cSIFT3D.cc:1544), below which nothing contributes at all: the per-keypoint fixed-point scale
    must keep such descriptors accurate."""
    vol = synth.v_blobs(64, seed=21) * np.float32(0.04)
    vol[5, 6, 7] = 1.0
    r = checker.extract(vol, keep_levels=False)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    sift.KpSiftAlgorithm()
    _compare_sparse(s3d, r, sift)


def test_undersized_sparse_buffers_are_resized_and_results_unchanged(s3d, synth, monkeypatch):
    """A step has no host round trip: detection and keypoint buffers are sized optimistically and the counts stay on the
    device until s3d_wait.  When a volume exceeds the buffers the sparse stage is repeated from the kept pyramid with
    exact sizes: forced here with absurdly small capacities, the results must be those of the default run."""
    vol = synth.v_blobs(64, seed=3)
    a = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    a.KpSiftAlgorithm()
    ka, da = a.GetKeypoints(), a.descriptors
    ea = a.extrema()
    assert a.counters()["sparse_resized"] == 0 and len(ka) > 8
    monkeypatch.setenv("S3D_CAP_EXTRE", "16")
    monkeypatch.setenv("S3D_CAP_KPS", "4")
    b = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    b.KpSiftAlgorithm()
    assert b.counters()["sparse_resized"] == 1
    kb, db = b.GetKeypoints(), b.descriptors
    eb = b.extrema()
    assert len(ka) == len(kb) and np.array_equal(da, db)
    for f in ka.dtype.names:
        if f != "desc":
            assert np.array_equal(ka[f], kb[f]), f
    assert np.array_equal(ea[1], eb[1]) and np.array_equal(ea[2], eb[2])
    monkeypatch.setenv("S3D_CAP_EXTRE", "100000")     # detections fit, keypoints do not
    monkeypatch.setenv("S3D_CAP_KPS", "4")
    c = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    c.KpSiftAlgorithm()
    assert c.counters()["sparse_resized"] == 1 and np.array_equal(c.descriptors, da)
