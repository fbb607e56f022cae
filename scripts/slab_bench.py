"""The z-slab leg alone (fast A/B of transport settings), one process per GPU:
    python -m torch.distributed.run --nproc-per-node N ... scripts/slab_bench.py [--size 512] [--steps 10] [--tag name]
Prints one line on rank 0: throughput ms/volume (two volumes in flight), single-volume latency, per-phase ms (max over ranks)."""
import argparse, ctypes as C, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512); ap.add_argument("--steps", type=int, default=10); ap.add_argument("--tag", default="")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s3d = importlib.import_module("3dsift_b200"); synth = importlib.import_module("3dsift_b200.synth"); D = importlib.import_module("3dsift_b200.dist")
api, L = s3d.api, s3d.lib()
n = a.size
comm = D.nccl_comm()
o0, o1 = C.c_int(), C.c_int()
api.check(L.s3d_slab_bounds(n, world, rank, C.byref(o0), C.byref(o1)))
vol = synth.v_blobs(n, seed=0)
d_own = torch.from_numpy(np.ascontiguousarray(vol[o0.value:o1.value])).cuda()
p = D._params(dict(device=local))
def create():
    h = C.c_void_p(); api.check(L.s3d_slab_create(comm._c, d_own.data_ptr(), 1, n, n, n, C.byref(p), C.byref(h))); return h
def steps(k, keep=None):
    infl = []
    def fin(h):
        api.check(L.s3d_wait(h)); api.check(L.s3d_slab_gather(comm._c, h, 0, 0))
        if keep is not None and not infl: keep.append(h)
        else: L.s3d_destroy(h)
    cur = create()
    for i in range(k):
        nxt = create() if i + 1 < k else None
        api.check(L.s3d_slab_execute_async(comm._c, cur)); infl.append(cur)
        if len(infl) >= 2: fin(infl.pop(0))
        cur = nxt
    while infl: fin(infl.pop(0))
def bar(): dist.barrier(); torch.cuda.synchronize()
def mx(x):
    t = torch.tensor(x, dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.tolist()
steps(4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bar(); e0.record(); steps(a.steps); torch.cuda.synchronize(); e1.record(); bar()
thr = mx([e0.elapsed_time(e1) / a.steps])[0]
lat, ph = [], None
for _ in range(5):
    keep = []
    bar(); t0 = time.perf_counter(); steps(1, keep); torch.cuda.synchronize(); lat.append((time.perf_counter() - t0) * 1e3)
    sh = D.SlabShard(keep[0], (n, n, n), rank); ph = sh.phases(); nk = sh.num_keypoints(); sh.close()
latm = mx([float(np.median(lat))])[0]
phm = mx([ph[k] for k in ("normalize", "pyramid", "halo", "sparse", "gather")])
if rank == 0:
    print(json.dumps({"tag": a.tag, "world": world, "ms_per_volume": round(thr, 3), "latency_ms": round(latm, 3),
                      "phases_max_ms": dict(zip(("normalize", "pyramid", "halo", "sparse", "gather"), [round(v, 3) for v in phm])), "keypoints": nk}), flush=True)
dist.destroy_process_group()
