"""Every run-time kernel variant must give the SAME bits.  The switches are read once per process
(S3D_BLUR_XY, S3D_ZVAR, S3D_DESC_ORDER), so each variant runs in its own interpreter on one seeded volume
and reports digests of: two Gaussian levels, two DoG levels, the detection list, the keypoint records and
the descriptors.  The default configuration is also compared with the oracle in the other GPU test files;
here the point is variant == default, bit for bit."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import hashlib, importlib, json, sys
import numpy as np
sys.path.insert(0, %r)
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
out = {}
for n in (96, 64):
    v = synth.v_blobs(n, seed=7)
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(v, keep_levels=True)
    s.KpSiftAlgorithm()
    h = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    kps = s.GetKeypoints()
    rec = np.stack([kps[f] for f in ("x", "y", "z", "scale", "octave", "level")], 1) if len(kps) else np.zeros((0, 6))
    out[str(n)] = dict(g1=h(s.GET_GSS(1)), g5=h(s.GET_GSS(5)), g8=h(s.GET_GSS(8)), d0=h(s.GET_DOG(0)), d4=h(s.GET_DOG(4)),
                       nk=len(kps), rec=h(rec), rot=h(kps["Rotation"]) if len(kps) else "", desc=h(s.descriptors))
    s.close()
print("RESULT " + json.dumps(out))
""" % ROOT


def run_variant(env):
    e = dict(os.environ)
    for k in ("S3D_BLUR_XY", "S3D_ZVAR", "S3D_DESC_ORDER", "S3D_DESC_PATH"):
        e.pop(k, None)
    e.update(env)
    p = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.fixture(scope="module")
def default_result():
    r = run_variant({})
    assert r["96"]["nk"] > 0, "the variant volume must yield keypoints"
    return r


@pytest.mark.parametrize("env", [
    {"S3D_BLUR_XY": "0"},   # separate X and Y passes
    {"S3D_BLUR_XY": "1"},   # unrolled fused X+Y kernel up to hw 6, separate passes at hw 8
    {"S3D_BLUR_XY": "3"},   # compact-code fused X+Y kernel for every hw
    {"S3D_ZVAR": "0"},      # register-ring march instead of the cp.async ring
    {"S3D_DESC_ORDER": "0"},  # descriptor CTAs in list order
    {"S3D_BLUR_XY": "0", "S3D_ZVAR": "0", "S3D_DESC_ORDER": "0"},
], ids=lambda e: ",".join(f"{k[4:]}={v}" for k, v in e.items()))
def test_variant_equals_default(default_result, env):
    got = run_variant(env)
    assert got == default_result
