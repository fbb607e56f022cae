"""The bench volume itself — V-blobs 512^3 seed 0, BASELINE.json configs[2] / the configuration the headline metric is
quoted on — through the whole extraction on one B200, against the compiled reference on the same bytes (one reference
run: a few minutes of host time, the longest test of the suite, so the file sorts last).  Bars as everywhere: the
detection list and its order bit-exact, accept/reject codes equal with every flip enumerated, keypoint records exact,
descriptor cosine >= 0.9999.  Seven octaves, every fast kernel at full width, ~48 k detections, ~9.7 k keypoints."""
import time

import numpy as np
import pytest

from test_gpu_sparse import _cos

pytestmark = pytest.mark.gpu


def test_config2_512_bench_volume(s3d, synth, refimpl):
    vol = synth.v_blobs(512, seed=0)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    sift.KpSiftAlgorithm()
    assert sift.num_octaves() == 7
    kp, codes, xyz5 = sift.extrema()
    kps = sift.GetKeypoints()
    desc = sift.descriptors.copy()
    thr = sift.thresholds()
    sift.close()
    t0 = time.time()
    r = refimpl.extract(vol, keep_levels=False)
    print(f"reference 512^3: {time.time() - t0:.1f} s on {refimpl.threads()} threads, stage times {r.times}")
    assert np.array_equal(xyz5, r.level_extrema), (len(xyz5), len(r.level_extrema))
    ref_rej = r.extrema["x"] < 0
    flips = np.flatnonzero((codes != 1) != ref_rej)
    for i in flips[:20]:
        print(f"orientation flip at detection {i}: xyz5={xyz5[i]} gpu code={codes[i]}")
    assert len(flips) == 0, f"{len(flips)} accept/reject flips of {len(kp)}"
    assert len(kps) == len(r.keypoints) and len(kps) > 5000
    for f in ("x", "y", "z", "rx", "ry", "rz", "scale", "octave", "level"):
        assert np.array_equal(kps[f], r.keypoints[f]), f
    assert np.abs(kps["Rotation"] - r.keypoints["Rotation"]).max() <= 2e-3
    cos = _cos(desc, r.desc)
    print(f"512^3 V-blobs: {len(xyz5)} detections, {len(kps)} keypoints, descriptor cosine min {cos.min():.7f}, "
          f"thresholds {thr[:3]}")
    assert cos.min() >= 0.9999
