"""Where does the host time of an e2e step go?  Wall-clock split of upload / run / fetch / close over 12 steps."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
s3d = importlib.import_module("3dsift_b200")
synth = importlib.import_module("3dsift_b200.synth")
L = s3d.lib()
n = 512
cache = f"/tmp/vblobs_{n}.npy"
v = np.load(cache) if os.path.exists(cache) else synth.v_blobs(n, seed=0)
h_vol = torch.from_numpy(v).pin_memory()
h_kp = torch.empty((20000, 176), dtype=torch.uint8).pin_memory()
h_desc = torch.empty((20000, 768), dtype=torch.float32).pin_memory()
up = lambda: s3d.CSIFT3DFactory.CreateCSIFT3D(h_vol, x_dim=n, y_dim=n, z_dim=n, device=0, async_upload=True)
T = {"upload": [], "run": [], "fetch": [], "close": [], "step": [], "device": []}
cur = up()
for i in range(14):
    t0 = time.perf_counter()
    nxt = up()
    t1 = time.perf_counter()
    cur.KpSiftAlgorithm()
    t2 = time.perf_counter()
    k = cur.num_keypoints()
    s3d.check(L.s3d_get_keypoints(cur._h, h_kp.data_ptr(), h_desc.data_ptr()))
    t3 = time.perf_counter()
    dev = cur.m_timer["d_TotalTime"] * 1e3
    cur.close()
    t4 = time.perf_counter()
    cur = nxt
    if i >= 4:
        for key, a, b in (("upload", t0, t1), ("run", t1, t2), ("fetch", t2, t3), ("close", t3, t4), ("step", t0, t4)):
            T[key].append((b - a) * 1e3)
        T["device"].append(dev)
cur.close()
print({k: round(float(np.mean(x)), 3) for k, x in T.items()})
