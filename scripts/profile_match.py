"""One HBM-resident enhancedMatch for ncu captures:  python scripts/profile_match.py 20000 [path]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
if len(sys.argv) > 2:
    s3d.set_match_path(int(sys.argv[2]))
L = s3d.lib()
ref, tar, _ = synth.d_synth_pair(n, seed=100)
d_ref, d_tar = torch.from_numpy(ref).cuda(), torch.from_numpy(tar).cuda()
nr, nt = len(ref), len(tar)
I = lambda m: torch.empty(max(m, 1), dtype=torch.int32, device="cuda")
F = lambda m: torch.empty(max(m, 1), dtype=torch.float32, device="cuda")
bufs = [I(nr), F(nr), I(nr), F(nr), I(nt), F(nt), I(nt), F(nt), I(nr), I(nr), I(1)]
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s3d.check(L.s3d_match_device(3, d_ref.data_ptr(), nr, d_tar.data_ptr(), nt, 0.85, *[b.data_ptr() for b in bufs], st))
    e1.record(); torch.cuda.synchronize()
    print("enhancedMatch", nr, "x", nt, "ms", e0.elapsed_time(e1), "matches", int(bufs[10].item()), "stats", s3d.match_stats())
