"""Experiment: throughput of T host threads, each extracting HBM-resident 512^3 volumes on its own handle and stream
(the C calls release the GIL), against one thread.    python scripts/ab_threads.py 512 8"""
import importlib, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cache = f"/tmp/vblobs_{n}.npy"
if os.path.exists(cache):
    v = np.load(cache)
else:
    v = synth.v_blobs(n, seed=0); np.save(cache, v)
vol = torch.from_numpy(v).cuda()
for _ in range(3):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol); s.KpSiftAlgorithm(); s.close()

def worker(k, out):
    nk = 0
    for _ in range(k):
        s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)   # private stream per handle
        s.KpSiftAlgorithm()
        nk = s.num_keypoints()
        s.close()
    out.append(nk)

for T in (1, 2, 3, 4):
    torch.cuda.synchronize()
    out = []
    th = [threading.Thread(target=worker, args=(reps, out)) for _ in range(T)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"threads={T}: {T * reps} volumes in {dt * 1e3:.1f} ms -> {dt * 1e3 / (T * reps):.2f} ms/volume, keypoints {out}")
