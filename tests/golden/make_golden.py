"""Generate tests/golden/*.npz from the compiled reference (oracle/_ref/libsift3d_ref.so).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these fixtures — outputs of
the UNMODIFIED reference on small seeded inputs, inputs included — are what pins the oracle port
and the CUDA path on machines where the reference sources are absent.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref as O  # noqa: E402

synth = importlib.import_module("3dsift_b200.synth")


def main():
    O.build("ref")
    R = O.Ref()
    # (a) GaussianSmooth_3D on an odd-sized volume, the three sigma classes of SURVEY.md §8c
    rng = np.random.default_rng(7)
    vol = rng.standard_normal((19, 23, 37)).astype(np.float32)
    sig = np.array([0.538701117, 0.97329402, 2.45254731], np.float32)
    np.savez_compressed(os.path.join(HERE, "blur.npz"), vol=vol, sigmas=sig,
                        out=np.stack([R.gaussian_smooth(vol, float(s)) for s in sig]))
    # (b) full extraction on a small non-cubic volume
    v = synth.v_blobs((48, 40, 32), seed=3)
    r = R.extract(v)
    G = 6
    np.savez_compressed(
        os.path.join(HERE, "extract.npz"), vol=v, noct=r.noct, dims=np.array(r.dims, np.int32), input=r.input,
        gss1=r.gss(1), gss3=r.gss(3), gss5=r.gss(5), gss_o1_2=r.gss(G + 2), dog2=r.dog(2), dog_o1_1=r.dog(5 + 1),
        level_extrema=r.level_extrema, extrema=r.extrema.view(np.uint8).reshape(len(r.extrema), 176),
        keypoints=r.keypoints.view(np.uint8).reshape(len(r.keypoints), 176), desc=r.desc)
    print("extract:", r.noct, r.dims, len(r.extrema), "extrema", len(r.keypoints), "keypoints")
    # (c) matcher with the index-0 quirk, duplicates (ties) and an all-zero row
    ref, tar, _ = synth.d_synth_pair(60, seed=5, k_tar=50)
    tar[7] = tar[3]          # duplicate database rows: ties resolve to the lowest index
    ref[11] = 0.0            # dot == 0 <= FLT_MIN: never matches (idx -1, dist 2)
    out = {}
    for t, name in ((1, "inject"), (2, "biject"), (3, "enhanced")):
        m = R.match(t, ref, tar, 0.85)
        for k in ("gIdx", "gDist", "sIdx", "sDist", "pairs"):
            out[f"{name}_{k}"] = m[k]
    np.savez_compressed(os.path.join(HERE, "match.npz"), ref=ref, tar=tar, **out)
    # (d) icosahedron mesh after Initialize_geometry
    mv, mi = R.mesh()
    np.savez_compressed(os.path.join(HERE, "mesh.npz"), v=mv, idx=mi)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
