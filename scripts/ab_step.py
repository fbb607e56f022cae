"""A/B helper: a few HBM-resident 512^3 extraction steps with per-kernel event timing, one line per
kernel class.  Run-time switches come from the environment (S3D_BLUR_XY, S3D_ZVAR, S3D_DESC_ORDER ...):
    S3D_ZVAR=0 python scripts/ab_step.py 512 4
"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cache = f"/tmp/vblobs_{n}.npy"
if os.path.exists(cache):
    v = np.load(cache)
else:
    v = synth.v_blobs(n, seed=0)
    np.save(cache, v)
vol = torch.from_numpy(v).cuda()
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("S3D_"))
for _ in range(3):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol); s.KpSiftAlgorithm(); s.close()
acc = {}
tot = 0.0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
for _ in range(reps):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol, profile=True, stream=torch.cuda.current_stream().cuda_stream)
    s.KpSiftAlgorithm()
    for k, st in s.kernel_stats().items():
        acc[k] = acc.get(k, 0.0) + st["ms"]
    nk = s.num_keypoints()
    s.close()
ev1.record()
torch.cuda.synchronize()
print(f"[{tag}] step {ev0.elapsed_time(ev1) / reps:.3f} ms, {nk} keypoints | " +
      " ".join(f"{k}={v / reps:.3f}" for k, v in acc.items() if v / reps >= 0.02))
