"""BASELINE.json configs[1]: a synthetic 256^3 CT-like volume through the whole extraction on one B200, against the
compiled reference on the same bytes (about 15 s of host time).  The north-star bars at full width: detection list and
order bit-exact (88 987 detections), accept/reject codes equal with flips enumerated, keypoint records exact, descriptor
cosine >= 0.9999.  (The debug fields of the rejected detections - eigenvalues of nearly singular structure tensors - are
compared in test_gpu_sparse.py on smaller volumes; with 9e4 detections their relative tolerances are not meaningful.)
At this size every fast kernel of the pyramid runs at full width and six octaves are built."""
import numpy as np
import pytest

from test_gpu_sparse import _cos

pytestmark = pytest.mark.gpu


def test_config1_256_ct_volume(s3d, synth, refimpl):
    vol = synth.v_ct(256, seed=0)
    r = refimpl.extract(vol, keep_levels=False)
    sift = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    sift.KpSiftAlgorithm()
    assert sift.num_octaves() == 6
    kp, codes, xyz5 = sift.extrema()
    assert np.array_equal(xyz5, r.level_extrema), (len(xyz5), len(r.level_extrema))
    ref_rej = r.extrema["x"] < 0
    flips = np.flatnonzero((codes != 1) != ref_rej)
    for i in flips[:20]:
        print(f"orientation flip at detection {i}: xyz5={xyz5[i]} gpu code={codes[i]}")
    assert len(flips) == 0, f"{len(flips)} accept/reject flips of {len(kp)}"
    kps = sift.GetKeypoints()
    assert len(kps) == len(r.keypoints) and len(kps) > 500
    for f in ("x", "y", "z", "rx", "ry", "rz", "scale", "octave", "level"):
        assert np.array_equal(kps[f], r.keypoints[f]), f
    cos = _cos(sift.descriptors, r.desc)
    print(f"256^3 V-CT: {len(xyz5)} detections, {len(kps)} keypoints, descriptor cosine min {cos.min():.7f}")
    assert cos.min() >= 0.9999
    sift.close()
