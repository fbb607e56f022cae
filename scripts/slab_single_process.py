"""Single-process multi-device z-slab extraction (s3d_extract_multi: one host thread per shard, peer copies):
    python scripts/slab_single_process.py N [size]
parity against the unsharded run on device 0 and wall-clock / per-phase timing."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
D = importlib.import_module("3dsift_b200.dist")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
vol = torch.from_numpy(synth.v_blobs(size, seed=0)).pin_memory().numpy()
ref = s3d.CSIFT3DFactory.CreateCSIFT3D(vol); ref.KpSiftAlgorithm(); kr = ref.GetKeypoints(); dr = ref.descriptors
out = {"n": n, "size": size}
devs = list(range(n))
for rep in range(4):
    t = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = D.extract_slabs(vol, shards=n, devices=devs, timing=t, with_extrema=False)
    dt = time.perf_counter() - t0
    out[f"rep{rep}"] = {"wall_ms": round(dt * 1e3, 2), "phases_ms": [{k: round(v, 2) for k, v in p.items()} for p in t["per_shard"]]}
same = all(np.array_equal(r["kp"][f], kr[f]) for f in kr.dtype.names if f != "desc") and np.array_equal(r["desc"], dr)
out["equal_to_unsharded"] = bool(same)
out["keypoints"] = int(len(r["kp"]))
print(json.dumps(out))
