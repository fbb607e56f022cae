"""Host-side split of the e2e leg of bench.py (pinned volume -> create_async -> run -> get -> close)."""
import importlib, os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
s3d = importlib.import_module("3dsift_b200"); synth = importlib.import_module("3dsift_b200.synth")
L = s3d.lib()
n = 512
vol = synth.v_blobs(n, seed=0)
h_vol = torch.from_numpy(vol).pin_memory()
h_kp = torch.empty((20000, 176), dtype=torch.uint8).pin_memory(); h_desc = torch.empty((20000, 768), dtype=torch.float32).pin_memory()
def upload():
    return s3d.CSIFT3DFactory.CreateCSIFT3D(h_vol, x_dim=n, y_dim=n, z_dim=n, device=0, async_upload=True)
def run(k_steps, tag):
    T = {"upload": 0.0, "run": 0.0, "get": 0.0, "close": 0.0}
    torch.cuda.synchronize(); t00 = time.perf_counter()
    t0 = time.perf_counter(); cur = upload(); T["upload"] += time.perf_counter() - t0
    for i in range(k_steps):
        t0 = time.perf_counter(); nxt = upload() if i + 1 < k_steps else None; t1 = time.perf_counter()
        cur.KpSiftAlgorithm(); t2 = time.perf_counter()
        s3d.check(L.s3d_get_keypoints(cur._h, h_kp.data_ptr(), h_desc.data_ptr())); t3 = time.perf_counter()
        tm = cur.m_timer
        cur.close(); t4 = time.perf_counter()
        T["upload"] += t1 - t0; T["run"] += t2 - t1; T["get"] += t3 - t2; T["close"] += t4 - t3
        cur = nxt
    torch.cuda.synchronize(); tot = time.perf_counter() - t00
    print(tag, "ms/step %.2f" % (tot / k_steps * 1e3), {k: round(v / k_steps * 1e3, 2) for k, v in T.items()}, "device", round(tm["d_TotalTime"] * 1e3, 2), "h2d", round(tm["d_h2d"] * 1e3, 2))
run(3, "warm")
run(6, "no sampler")
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-lms", "50", "-i", "0"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
run(6, "with nvidia-smi -lms 50")
p.terminate()
time.sleep(0.5)
run(6, "no sampler again")
