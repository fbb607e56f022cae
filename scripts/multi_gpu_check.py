"""Multi-GPU parity + timing of the sharded paths, one process per GPU over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py [--size 256] [--big 512] [--match 50000] [--match-big 200000]

1. z-slab extraction of ONE volume (halo planes by NCCL send/recv, maxima by all-reduce) must equal the
   unsharded extraction on rank 0: same detections, codes, keypoint records, descriptors, same order.
2. Sharded matcher (searched set split over the ranks, top-2 lists all-gathered and merged under
   (dot desc, index asc)) must equal the single-GPU enhancedMatch bit for bit.
3. Timings (device-synchronised wall clock, max over ranks) of both at a larger size.
Rank 0 prints one JSON line."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--big", type=int, default=512)
ap.add_argument("--match", type=int, default=50000)
ap.add_argument("--match-big", type=int, default=200000)
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
D = importlib.import_module("3dsift_b200.dist")
s3d.selftest(local)
out = {"world": world}


def same_records(x, y):
    return all(np.array_equal(x[f], y[f]) for f in x.dtype.names if f != "desc")


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
    return r, best


# ---- 1. z-slab extraction == unsharded -------------------------------------------------------------
vol = synth.v_blobs(a.size, seed=5)
sl = D.extract_slabs(vol)
if rank == 0:
    ref = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    ref.KpSiftAlgorithm()
    kp_r, codes_r, xyz_r = ref.extrema()
    ok = (np.array_equal(sl["xyz5"], xyz_r) and np.array_equal(sl["codes"], codes_r) and same_records(sl["extrema"], kp_r)
          and same_records(sl["kp"], ref.GetKeypoints()) and np.array_equal(sl["desc"], ref.descriptors))
    out["slab_parity"] = {"size": a.size, "keypoints": int(len(sl["kp"])), "detections": int(len(xyz_r)), "equal_to_unsharded": bool(ok)}
    ref.close()
del sl

# ---- 2. sharded matcher == single GPU --------------------------------------------------------------
r_, t_, _ = synth.d_synth_pair(a.match, seed=3)
sm = D.match_sharded(3, r_, t_, 0.85)
if rank == 0:
    m = s3d.muBruteMatcher()
    m.enhancedMatch(r_, t_, 0.85)
    ok = (np.array_equal(sm["pairs"], m.pairs) and np.array_equal(sm["gIdx"], m.getGlodenIdx())
          and np.array_equal(sm["gDist"], m.getGlodenDistSquare()) and np.array_equal(sm["sIdx"], m.getSilverIdx()))
    out["match_parity"] = {"n": a.match, "pairs": int(len(m.pairs)), "equal_to_single_gpu": bool(ok)}

# ---- 3. timings ------------------------------------------------------------------------------------
big = torch.from_numpy(synth.v_blobs(a.big, seed=0)).pin_memory().numpy()   # numpy view of pinned host memory
res, sec = timed(lambda: D.extract_slabs(big, with_extrema=False), reps=3)
phases = {}
D.extract_slabs(big, timing=phases, with_extrema=False)
out["slab_timing"] = {"size": a.big, "seconds": sec, "phases_ms": {k: round(v, 3) for k, v in phases.items() if not isinstance(v, list)}, "mvoxels_per_s": big.size / sec / 1e6, "keypoints": int(len(res["kp"])),
                      "note": "pinned host volume in, merged keypoints + descriptors out on the host of rank 0 (includes the per-shard upload and the NCCL result gather)"}
del res
dr, dt_, _ = synth.d_synth_pair_device(a.match_big, seed=9)   # same seed on every rank: replicated, HBM-resident sets
res, sec = timed(lambda: D.match_sharded(3, dr, dt_, 0.85))
out["match_timing"] = {"n": a.match_big, "seconds": sec, "pairs_per_s": a.match_big ** 2 / sec, "pairs": int(len(res["pairs"])),
                       "note": "descriptor sets resident in HBM on every rank, host results out"}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
