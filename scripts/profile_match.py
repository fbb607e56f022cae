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
if n >= 200000:
    d_ref, d_tar, truth = synth.d_synth_pair_device(n, seed=100)
else:
    ref, tar, truth = synth.d_synth_pair(n, seed=100)
    d_ref, d_tar, truth = torch.from_numpy(ref).cuda(), torch.from_numpy(tar).cuda(), torch.from_numpy(truth).cuda()
nr, nt = len(d_ref), len(d_tar)
I = lambda m: torch.empty(max(m, 1), dtype=torch.int32, device="cuda")
F = lambda m: torch.empty(max(m, 1), dtype=torch.float32, device="cuda")
bufs = [I(nr), F(nr), I(nr), F(nr), I(nt), F(nt), I(nt), F(nt), I(nr), I(nr), I(1)]
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s3d.check(L.s3d_match_device(3, d_ref.data_ptr(), nr, d_tar.data_ptr(), nt, 0.85, *[b.data_ptr() for b in bufs], st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    npairs = int(bufs[10].item())
    pr, pt = bufs[8][:npairs].long(), bufs[9][:npairs].long()
    hit = float((truth[pr] == pt).float().mean().item()) if npairs else 0.0
    print("enhancedMatch", nr, "x", nt, "ms", round(ms, 3), "pairs/s %.3e" % (nr * nt / (ms * 1e-3)), "alg TFLOP/s %.1f" % (2 * 768 * nr * nt / (ms * 1e-3) / 1e12),
          "matches", npairs, "true-pair fraction %.4f" % hit, "stats", s3d.match_stats())
