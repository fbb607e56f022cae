"""One (or a few) HBM-resident extraction steps for ncu captures:  python scripts/profile_step.py 512 2"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
vol = torch.from_numpy(synth.v_blobs(n, seed=0)).cuda()
for _ in range(reps):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    s.KpSiftAlgorithm()
    print("keypoints", s.num_keypoints(), {k: round(v * 1e3, 3) for k, v in s.m_timer.items()})
    s.close()
