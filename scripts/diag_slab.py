import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
s3d = importlib.import_module("3dsift_b200"); synth = importlib.import_module("3dsift_b200.synth"); D = importlib.import_module("3dsift_b200.dist")
n, G = int(sys.argv[1]), int(sys.argv[2])
vol = synth.v_blobs(n, seed=5)
ref = s3d.CSIFT3DFactory.CreateCSIFT3D(vol, keep_levels=True); ref.KpSiftAlgorithm()
kp_r, codes_r, xyz_r = ref.extrema()
sl = D.extract_slabs(vol, shards=G, params=dict(keep_levels=1), keep=True)
print("detections", len(xyz_r), len(sl["xyz5"]), "equal xyz5", np.array_equal(sl["xyz5"], xyz_r))
if len(xyz_r) == len(sl["xyz5"]):
    bad = np.flatnonzero((sl["xyz5"] != xyz_r).any(1)); print("xyz5 mismatches", len(bad), xyz_r[bad[:5]], sl["xyz5"][bad[:5]])
    badc = np.flatnonzero(sl["codes"] != codes_r); print("code mismatches", len(badc), xyz_r[badc[:10]], codes_r[badc[:10]], sl["codes"][badc[:10]])
else:
    a = set(map(tuple, xyz_r)); b = set(map(tuple, sl["xyz5"])); print("only ref", sorted(a - b)[:10], "only slab", sorted(b - a)[:10])
kps = ref.GetKeypoints()
print("kps", len(kps), len(sl["kp"]))
if len(kps) == len(sl["kp"]):
    d = np.abs(sl["desc"] - ref.descriptors).max(1); badd = np.flatnonzero(d > 0); print("desc mismatches", len(badd), [(int(kps["octave"][i]), int(kps["level"][i]), float(kps["z"][i]), float(d[i])) for i in badd[:10]])
    for f in kps.dtype.names:
        if f != "desc" and not np.array_equal(kps[f], sl["kp"][f]): print("field differs", f)
# levels
Gl, Dl = 6, 5
for g, sh in sl["shards"].items():
    for o in range(sh.noct):
        for which, per in ((0, Gl), (1, Dl)):
            for i in range(per):
                got, (za, zb, p0, p1) = sh.get_level_host(which, o * per + i)
                if got is None or p1 <= p0: continue
                want = (ref.GET_GSS if which == 0 else ref.GET_DOG)(o * per + i)
                got = got.reshape(zb - za, *want.shape[1:])
                if not np.array_equal(got[p0 - za:p1 - za], want[p0:p1]):
                    bz = [int(z) for z in range(p0, p1) if not np.array_equal(got[z - za], want[z])]
                    print("level mismatch shard", g, "oct", o, "gss" if which == 0 else "dog", i, "ext", (za, zb, p0, p1), "planes", bz[:8])
print("---- full local extents, shard", int(sys.argv[3]) if len(sys.argv) > 3 else 4)
g = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sh = sl["shards"][g]
for o in range(sh.noct):
    for which, per in ((0, Gl), (1, Dl)):
        for i in range(per):
            got, (za, zb, p0, p1) = sh.get_level_host(which, o * per + i)
            if got is None: continue
            want = (ref.GET_GSS if which == 0 else ref.GET_DOG)(o * per + i)
            got = got.reshape(zb - za, *want.shape[1:])
            bz = [int(z) for z in range(za, zb) if not np.array_equal(got[z - za], want[z], equal_nan=False)]
            print("oct", o, "gss" if which == 0 else "dog", i, "ext", (za, zb, p0, p1), "bad planes", (bz[0], bz[-1], len(bz)) if bz else None)
for (gg, o, lvl, zlist) in ((4, 2, 0, (58, 59, 63)), (7, 3, 0, (18, 19, 20))):
    sh = sl["shards"][gg]
    got, (za, zb, p0, p1) = sh.get_level_host(0, o * Gl + lvl)
    want = ref.GET_GSS(o * Gl + lvl)
    got = got.reshape(zb - za, *want.shape[1:])
    for z in zlist:
        a, b = got[z - za], want[z]
        print("shard", gg, "oct", o, "gss", lvl, "plane", z, "nan", int(np.isnan(a).sum()), "of", a.size, "maxdiff", float(np.nanmax(np.abs(a - b))), "got", a.ravel()[:3], "want", b.ravel()[:3])
    # does the plane equal some OTHER plane of the reference?
    for z in zlist[1:2]:
        eq = [int(k) for k in range(want.shape[0]) if np.array_equal(got[z - za], want[k])]
        print("  plane", z, "equals reference planes", eq)
