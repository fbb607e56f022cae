"""World-size-2 gloo tests (CPU) of the multi-GPU orchestration in 3dsift_b200/dist.py: shard
bounds, global index offsets, all-gather layout, the (dot desc, index asc) merge contract and the
replicated filters — driven with a checker-backed ops object (the CUDA primitives need a GPU and are
covered by tests/test_gpu_match.py::test_sharded_database_merge_equals_unsharded)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLT_MIN = float(np.finfo(np.float32).tiny)


class NumpyOps:
    """Reference semantics of the primitives (Src/cMatcher.cc), numpy only — test side."""

    def to_device(self, a):
        return np.ascontiguousarray(a, dtype=np.float32)

    def top2(self, q, db, offset, mask):
        nq = len(q)
        d1 = np.full(nq, FLT_MIN, np.float64); d2 = np.full(nq, FLT_MIN, np.float64)
        i1 = np.full(nq, -1, np.int32); i2 = np.full(nq, -1, np.int32)
        for i in range(nq):
            if mask is not None and mask[i] == 0:
                continue
            for j in range(len(db)):
                s = float(np.add.accumulate((q[i] * db[j]).astype(np.float64))[-1])   # float product, sequential double sum
                if s > d1[i]:
                    d2[i], i2[i], d1[i], i1[i] = d1[i], i1[i], s, j + offset
                elif s > d2[i]:
                    d2[i], i2[i] = s, j + offset
        return d1, i1, d2, i2

    def merge(self, D1, I1, D2, I2, mask):
        parts, nq = D1.shape
        gD = np.zeros(nq, np.float32); sD = np.zeros(nq, np.float32)
        gI = np.full(nq, -1, np.int32); sI = np.full(nq, -1, np.int32)
        for i in range(nq):
            if mask is not None and mask[i] == 0:
                continue
            c = [(D1[p, i], I1[p, i]) for p in range(parts) if I1[p, i] >= 0] + [(D2[p, i], I2[p, i]) for p in range(parts) if I2[p, i] >= 0]
            c.sort(key=lambda t: (-t[0], t[1]))
            b1 = c[0] if len(c) > 0 else (FLT_MIN, -1)
            b2 = c[1] if len(c) > 1 else (FLT_MIN, -1)
            gD[i], gI[i], sD[i], sI[i] = np.float32(2 - 2 * b1[0]), b1[1], np.float32(2 - 2 * b2[0]), b2[1]
        return gD, gI, sD, sI

    def ratio_filter(self, gI, gD, sD, thr):
        with np.errstate(divide="ignore", invalid="ignore"):
            rej = (gI >= 0) & ((gD / sD).astype(np.float64) >= thr * thr)
        gI[rej] *= -1

    def count_mask(self, gI, n_tar, count_thres):
        cnt = np.bincount(gI[gI >= 0], minlength=n_tar)
        return (cnt > count_thres).astype(np.int32)

    def biject_filter(self, gI, mask, gI2):
        for i in range(len(gI)):
            m = gI[i]
            if m < 0 or mask[m] == 0:
                continue
            if gI2[m] != i:
                gI[i] *= -1

    def all_gather(self, x, group=None):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(x))
        outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(outs, t, group=group)
        return np.stack([o.numpy() for o in outs])

    def to_numpy(self, x):
        return np.asarray(x)


def _worker(rank, world, port, ref, tar, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("3dsift_b200.dist")
    res = {t: d.match_sharded(t, ref, tar, 0.85, ops=NumpyOps()) for t in (1, 2, 3)}
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    d = importlib.import_module("3dsift_b200.dist")
    assert d.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert d.shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    assert d.shard_bounds(0, 2) == [0, 0, 0]
    for n in (1, 7, 64, 1000):
        for w in (1, 2, 4, 8):
            b = d.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and all(0 <= b[i + 1] - b[i] <= -(-n // w) for i in range(w))


@pytest.mark.timeout(300)
def test_sharded_match_world2_gloo_equals_oracle(port, synth):
    import torch.multiprocessing as mp
    ref, tar, _ = synth.d_synth_pair(36, seed=8, k_tar=41)
    tar[30] = tar[5]     # a tie whose two members land in different shards (bounds [0,21,41])
    ref[4] = 0.0         # never matches
    want = {t: port.match(t, ref, tar, 0.85) for t in (1, 2, 3)}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, ref, tar, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for t in (1, 2, 3):
            for k in ("gIdx", "sIdx", "gDist", "sDist", "pairs"):
                assert np.array_equal(got[rank][t][k], want[t][k]), (rank, t, k)
            if t != 1:
                searched = want[t]["gIdx2"] != -1
                assert np.array_equal(got[rank][t]["gIdx2"], want[t]["gIdx2"])
                assert np.array_equal(got[rank][t]["gDist2"][searched], want[t]["gDist2"][searched])
    assert want[1]["gIdx"][4] == -1 and abs(want[1]["gIdx"]).max() > 21   # both shards contribute winners


def test_single_process_path_equals_oracle(port, synth):
    d = importlib.import_module("3dsift_b200.dist")
    ref, tar, _ = synth.d_synth_pair(20, seed=9, k_tar=17)
    for t in (1, 3):
        got = d.match_sharded(t, ref, tar, 0.85, ops=NumpyOps())
        want = port.match(t, ref, tar, 0.85)
        assert np.array_equal(got["gIdx"], want["gIdx"]) and np.array_equal(got["pairs"], want["pairs"])


# ---- z-slab sharding: plane bookkeeping and the halo exchange (CPU, gloo) ---------------------------------

def test_slab_bounds_and_extents(s3d):
    import ctypes as C
    d = importlib.import_module("3dsift_b200.dist")
    L = s3d.lib()
    p = s3d.api.s3d_params()
    L.s3d_default_params(C.byref(p))
    for nz, G in ((512, 8), (96, 3), (70, 2), (64, 4), (20, 8)):
        b = d.slab_bounds(nz, G)
        assert b[0] == 0 and b[-1] == nz and all(b[i] <= b[i + 1] for i in range(G)) and all(v % 2 == 0 for v in b[:-1])
        nzo = nz
        for o in range(4):
            owned = []
            for g in range(G):
                if b[g + 1] <= b[g]:
                    continue
                e = (C.c_int * 4)()
                assert L.s3d_slab_extent(nz, b[g], b[g + 1], C.byref(p), o, e) == 0
                za, zb, p0, p1 = (int(v) for v in e)
                assert 0 <= za <= p0 <= p1 <= zb <= nzo
                if p1 > p0:
                    assert za == max(0, p0 - 38) and zb == min(nzo, p1 + 38)     # default parameters: halo 38
                    assert (p0 << o) >= b[g] and ((p1 - 1) << o) < b[g + 1]
                owned.append((p0, p1))
            assert owned[0][0] == 0 and owned[-1][1] == nzo                       # the owned ranges tile [0, nz_o)
            assert all(owned[i][1] == owned[i + 1][0] for i in range(len(owned) - 1))
            nzo //= 2


def test_exchange_plan_covers_every_halo_plane_once():
    d = importlib.import_module("3dsift_b200.dist")
    exts = [(0, 20, 0, 8), (0, 28, 8, 16), (4, 32, 16, 24), (12, 32, 24, 32)]    # halo 12 > slab 8: multi-hop
    for me in range(4):
        plan = d.exchange_plan(exts, [0, 1, 2, 3], me)
        got = sorted((k0, k1) for s, dd, k0, k1, kind in plan if dd == me and kind == "recv")
        za, zb, p0, p1 = exts[me]
        need = set(range(za, p0)) | set(range(p1, zb))
        have = [k for k0, k1 in got for k in range(k0, k1)]
        assert sorted(have) == sorted(need)
        sends = [(s, dd, k0, k1) for s, dd, k0, k1, kind in plan if kind == "send"]
        assert all(s == me and exts[me][2] <= k0 < k1 <= exts[me][3] for s, dd, k0, k1 in sends)
    allcopy = d.exchange_plan(exts, [0, 0, 0, 0], 0)
    assert all(kind == "copy" for *_, kind in allcopy) and len(allcopy) > 0


def _halo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module("3dsift_b200.dist")
    exts = [(0, 14, 0, 10), (6, 20, 10, 20)]                   # 20 planes, halo 4
    plane = 6
    truth = torch.arange(20 * plane, dtype=torch.float32).view(20, plane)
    za, zb, p0, p1 = exts[rank]
    buf = torch.full((zb - za, plane), -1.0)
    buf[p0 - za:p1 - za] = truth[p0:p1]                        # only the owned planes are valid
    d.exchange_halos({rank: buf}, exts, [0, 1], rank)
    q.put((rank, bool(torch.equal(buf, truth[za:zb]))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_halo_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: True, 1: True}


def test_merge_shard_results_restores_reference_order(s3d):
    d = importlib.import_module("3dsift_b200.dist")
    KP = s3d.api.KP_DTYPE

    def part(rows):                                            # rows: (octave, level, z)
        kp = np.zeros(len(rows), KP)
        xyz5 = np.zeros((len(rows), 5), np.int32)
        for i, (o, l, z) in enumerate(rows):
            kp[i]["octave"], kp[i]["level"], kp[i]["z"] = o, l, z
            xyz5[i] = (0, 0, z, o, l)
        return dict(kp=kp, desc=np.zeros((len(rows), 768), np.float32), extrema=kp.copy(), codes=np.ones(len(rows), np.int32), xyz5=xyz5)

    a = part([(0, 1, 3), (0, 1, 5), (0, 2, 1), (1, 1, 2)])
    b = part([(0, 1, 40), (0, 3, 33), (1, 1, 20), (1, 1, 21)])
    m = d.merge_shard_results([a, b])
    got = [(int(r["octave"]), int(r["level"]), int(r["z"])) for r in m["kp"]]
    assert got == [(0, 1, 3), (0, 1, 5), (0, 1, 40), (0, 2, 1), (0, 3, 33), (1, 1, 2), (1, 1, 20), (1, 1, 21)]
    assert [tuple(r[2:]) for r in m["xyz5"]] == [(z, o, l) for o, l, z in got]


# ---- the library's own plane bookkeeping (s3d_slab.cu; host logic, no GPU) ---------------------------------------

def test_library_bounds_equal_python_bounds(s3d):
    import ctypes as C
    d = importlib.import_module("3dsift_b200.dist")
    L = s3d.lib()
    for nz, G in ((512, 8), (512, 2), (96, 3), (70, 2), (64, 4), (20, 8), (9, 1)):
        b = d.slab_bounds(nz, G)
        for r in range(G):
            o0, o1 = C.c_int(), C.c_int()
            assert L.s3d_slab_bounds(nz, G, r, C.byref(o0), C.byref(o1)) == 0
            assert (o0.value, o1.value) == (b[r], b[r + 1])


def test_library_halo_plan_pairs_up_and_covers_the_halo(s3d):
    """Every receive of rank b from rank a has the matching send of a to b at the same position of the (a, b) sequence
    (the pairing contract of Comm::p2p), receives cover exactly the halo planes other shards own, and sends only name
    owned planes."""
    d = importlib.import_module("3dsift_b200.dist")
    for nz, G, o, depth in ((512, 8, 0, 9), (512, 8, 1, 38), (512, 8, 2, 38), (512, 4, 2, -1), (96, 3, 0, 5), (200, 8, 0, 30),
                            (72, 3, 1, 38), (64, 4, 0, -1)):
        nzo = nz >> o
        b = d.slab_bounds(nz, G)
        own = [(min(nzo, (b[r] + (1 << o) - 1) >> o), min(nzo, (b[r + 1] + (1 << o) - 1) >> o)) for r in range(G)]
        plans = [d.plan_from_library(nz, G, r, o, depth) for r in range(G)]
        for r in range(G):
            p0, p1 = own[r]
            if depth < 0:
                need = set(range(0, p0)) | set(range(p1, nzo))
            elif p1 > p0:
                need = set(range(max(0, p0 - depth), p0)) | set(range(p1, min(nzo, p1 + depth)))
            else:
                need = set()
            got = [k for peer, kind, k0, k1 in plans[r] if kind == "recv" for k in range(k0, k1)]
            assert sorted(got) == sorted(need), (nz, G, o, depth, r)
            for peer, kind, k0, k1 in plans[r]:
                if kind == "send":
                    assert own[r][0] <= k0 < k1 <= own[r][1]
                else:
                    assert own[peer][0] <= k0 < k1 <= own[peer][1]
        for a in range(G):
            for bb in range(G):
                if a == bb:
                    continue
                sends = [(k0, k1) for peer, kind, k0, k1 in plans[a] if kind == "send" and peer == bb]
                recvs = [(k0, k1) for peer, kind, k0, k1 in plans[bb] if kind == "recv" and peer == a]
                assert sends == recvs, (nz, G, o, depth, a, bb)


def test_library_first_replicated_octave(s3d, monkeypatch):
    import ctypes as C
    L = s3d.lib()

    def first(n, world):
        o = C.c_int()
        assert L.s3d_slab_first_replicated(n, n, n, world, None, C.byref(o)) == 0
        return o.value
    monkeypatch.delenv("S3D_SLAB_MIN_NZ", raising=False)
    assert first(512, 1) == 7                       # one shard: nothing is replicated
    assert first(512, 2) == 3                       # 64^3 and below are computed by every shard
    assert first(512, 8) == 3
    assert first(256, 8) == 2
    assert first(64, 2) == 0                        # below the sharding threshold: the input itself is gathered
    monkeypatch.setenv("S3D_SLAB_MIN_NZ", "0")
    assert first(512, 8) == 3                       # 64 planes / 8 shards = 8 < 10 planes per shard
    assert first(512, 2) == 5                       # 16 planes / 2 = 8
    assert first(64, 2) == 2
