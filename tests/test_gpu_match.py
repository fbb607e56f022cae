"""GPU parity of the matcher (muBruteMatcher) through the C ABI: gIdx / sIdx / distances / match
lists bit-exact against the oracle, including the reference's quirks (App. B Q18-Q22)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = {1: "injectMatch", 2: "bijectMatch", 3: "enhancedMatch"}


def _run(s3d, t, ref, tar, thr=0.85):
    m = s3d.muBruteMatcher()
    getattr(m, NAMES[t])(ref, tar, thr)
    return m


@pytest.mark.parametrize("t", [1, 2, 3])
def test_match_matches_golden(s3d, t):
    g = np.load(os.path.join(GOLD, "match.npz"))
    name = {1: "inject", 2: "biject", 3: "enhanced"}[t]
    m = _run(s3d, t, g["ref"], g["tar"])
    assert np.array_equal(m.getGlodenIdx(), g[f"{name}_gIdx"])
    assert np.array_equal(m.getSilverIdx(), g[f"{name}_sIdx"])
    assert np.array_equal(m.getGlodenDistSquare(), g[f"{name}_gDist"])
    assert np.array_equal(m.getSilverDistSquare(), g[f"{name}_sDist"])
    assert np.array_equal(m.pairs, g[f"{name}_pairs"])


@pytest.mark.parametrize("t", [1, 2, 3])
@pytest.mark.parametrize("n_ref,n_tar", [(300, 257), (1000, 1100), (65, 700)])
def test_match_vs_oracle(s3d, synth, checker, t, n_ref, n_tar):
    ref, tar, _ = synth.d_synth_pair(n_ref, seed=n_ref + t, k_tar=n_tar)
    want = checker.match(t, ref, tar, 0.85)
    m = _run(s3d, t, ref, tar)
    assert np.array_equal(m.getGlodenIdx(), want["gIdx"])
    assert np.array_equal(m.getSilverIdx(), want["sIdx"])
    assert np.array_equal(m.getGlodenDistSquare(), want["gDist"])
    assert np.array_equal(m.getSilverDistSquare(), want["sDist"])
    assert np.array_equal(m.pairs, want["pairs"])
    assert len(m.pairs) > 0


def test_match_reverse_arrays_vs_port(s3d, synth, port):
    ref, tar, _ = synth.d_synth_pair(400, seed=21, k_tar=380)
    for t in (2, 3):
        want = port.match(t, ref, tar, 0.85)
        m = _run(s3d, t, ref, tar)
        rev = m.reverse()
        searched = want["gIdx2"] != -1
        assert np.array_equal(rev["gIdx2"], want["gIdx2"])
        assert np.array_equal(rev["sIdx2"][searched], want["sIdx2"][searched])
        assert np.array_equal(rev["gDist2"][searched], want["gDist2"][searched])


def test_ties_zero_rows_and_index0(s3d, port):
    rng = np.random.default_rng(1)
    tar = np.abs(rng.standard_normal((40, 768))).astype(np.float32)
    tar /= np.linalg.norm(tar, axis=1, keepdims=True)
    tar[9] = tar[4]                      # exact duplicates: lowest index wins, duplicate is second (Q18)
    tar[30] = tar[4]
    ref = tar[[4, 0, 12, 0]].copy()
    ref[2] = 0                           # dot == 0 <= FLT_MIN: idx -1, dist 2
    ref[3] = (0.51 * tar[0] + 0.49 * tar[1]) / np.linalg.norm(0.51 * tar[0] + 0.49 * tar[1])   # ambiguous -> index 0 (Q19)
    for t in (1, 2, 3):
        want = port.match(t, ref, tar, 0.85)
        m = _run(s3d, t, ref, tar)
        assert np.array_equal(m.getGlodenIdx(), want["gIdx"]) and np.array_equal(m.getSilverIdx(), want["sIdx"])
        assert np.array_equal(m.getGlodenDistSquare(), want["gDist"])
        assert np.array_equal(m.pairs, want["pairs"])
    m = _run(s3d, 1, ref, tar)
    assert m.getGlodenIdx()[0] in (4, -4) and m.getSilverIdx()[0] == 9
    assert m.getGlodenIdx()[2] == -1 and m.getGlodenDistSquare()[2] == 2.0
    assert m.getGlodenIdx()[3] == 0 and [3, 0] in m.pairs.tolist()


def test_empty_and_ragged_inputs(s3d, port):
    rng = np.random.default_rng(2)
    a = np.abs(rng.standard_normal((5, 768))).astype(np.float32)
    e = np.zeros((0, 768), np.float32)
    for t in (1, 2, 3):
        m = _run(s3d, t, e, a)
        assert len(m.pairs) == 0 and len(m.getGlodenIdx()) == 0
        m = _run(s3d, t, a, e)
        assert len(m.pairs) == 0 and np.array_equal(m.getGlodenIdx(), np.full(5, -1))
        assert np.array_equal(m.getGlodenDistSquare(), np.full(5, 2.0, np.float32))
        m = _run(s3d, t, a, a[:1])           # single database row
        want = port.match(t, a, a[:1], 0.85)
        assert np.array_equal(m.getGlodenIdx(), want["gIdx"]) and np.array_equal(m.pairs, want["pairs"])


def test_threshold_sweep(s3d, synth, port):
    ref, tar, _ = synth.d_synth_pair(200, seed=33, k_tar=210)
    for thr in (0.5, 0.7, 0.95, 1.0):
        want = port.match(3, ref, tar, thr)
        m = _run(s3d, 3, ref, tar, thr)
        assert np.array_equal(m.getGlodenIdx(), want["gIdx"]) and np.array_equal(m.pairs, want["pairs"])


def test_sharded_database_merge_equals_unsharded(s3d, synth, port):
    """shard the searched set, per-shard top-2 with global indices, merge under
    (dot desc, index asc).  Single-process logical shards; the NCCL path uses the same entry points."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    ref, tar, _ = synth.d_synth_pair(500, seed=44, k_tar=777)
    tar[500] = tar[3]                                 # a tie straddling two shards
    want = port.match(1, ref, tar, 0.85)
    L = s3d.lib()
    dq = torch.from_numpy(ref).cuda()
    nq = len(ref)
    for shards in (1, 2, 3, 8):
        bounds = np.linspace(0, len(tar), shards + 1).astype(int)
        d1 = torch.empty((shards, nq), dtype=torch.float64, device="cuda"); d2 = torch.empty_like(d1)
        i1 = torch.empty((shards, nq), dtype=torch.int32, device="cuda"); i2 = torch.empty_like(i1)
        for s in range(shards):
            db = torch.from_numpy(tar[bounds[s]:bounds[s + 1]]).cuda()
            s3d.check(L.s3d_top2_device(dq.data_ptr(), nq, db.data_ptr(), len(db), int(bounds[s]), None,
                                        d1[s].data_ptr(), i1[s].data_ptr(), d2[s].data_ptr(), i2[s].data_ptr(), None))
        torch.cuda.synchronize()
        gD = torch.empty(nq, dtype=torch.float32, device="cuda"); sD = torch.empty_like(gD)
        gI = torch.empty(nq, dtype=torch.int32, device="cuda"); sI = torch.empty_like(gI)
        s3d.check(L.s3d_top2_merge_device(shards, nq, d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(), None,
                                          gD.data_ptr(), gI.data_ptr(), sD.data_ptr(), sI.data_ptr(), None))
        s3d.check(L.s3d_ratio_filter_device(gI.data_ptr(), gD.data_ptr(), sD.data_ptr(), nq, 0.85, None))
        torch.cuda.synchronize()
        assert np.array_equal(gI.cpu().numpy(), want["gIdx"]), shards
        assert np.array_equal(sI.cpu().numpy(), want["sIdx"]), shards
        assert np.array_equal(gD.cpu().numpy(), want["gDist"]), shards


def test_match_sharded_cudaops_single_rank(s3d, synth, port):
    dist_mod = __import__("importlib").import_module("3dsift_b200.dist")
    ref, tar, _ = synth.d_synth_pair(300, seed=55, k_tar=280)
    for t in (1, 2, 3):
        got = dist_mod.match_sharded(t, ref, tar, 0.85)
        want = port.match(t, ref, tar, 0.85)
        for k in ("gIdx", "sIdx", "gDist", "sDist", "pairs"):
            assert np.array_equal(got[k], want[k]), (t, k)


def _nccl_worker(rank, world, port_no, ref, tar, q):
    import os, sys, importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = importlib.import_module("3dsift_b200.dist")
    res = {t: d.match_sharded(t, ref, tar, 0.85) for t in (1, 3)}
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_match_sharded_nccl_two_gpus(s3d, synth, port):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ref, tar, _ = synth.d_synth_pair(600, seed=66, k_tar=650)
    want = {t: port.match(t, ref, tar, 0.85) for t in (1, 3)}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, 29650, ref, tar, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        for t in (1, 3):
            for k in ("gIdx", "sIdx", "gDist", "sDist", "pairs"):
                assert np.array_equal(got[rank][t][k], want[t][k]), (rank, t, k)


# ---- the library's own sharded matcher (s3d_match_multi: database-sharded candidate pass, query-sharded exact re-rank) ----

@pytest.mark.parametrize("shards", [2, 3, 8])
@pytest.mark.parametrize("path", ["exact", "tensor"])
def test_match_multi_logical_shards_equal_single_gpu(s3d, synth, port, shards, path):
    """G logical shards on ONE device (host threads, device copies; the NCCL path runs the same code over another
    transport): all three match types, bit-equal to the oracle and to the unsharded call, incl. a tie across shards,
    uneven block sizes and (tensor path) the candidate exchange + merged guard."""
    d = __import__("importlib").import_module("3dsift_b200.dist")
    n_ref, n_tar = (2300, 2111) if path == "tensor" else (300, 277)
    ref, tar, _ = synth.d_synth_pair(n_ref, seed=71, k_tar=n_tar)
    tar[n_tar - 1] = tar[3]                              # a tie straddling the first and the last shard
    s3d.set_match_path(s3d.api.MATCH_TENSOR if path == "tensor" else s3d.api.MATCH_EXACT)
    try:
        for t in (1, 2, 3):
            want = port.match(t, ref, tar, 0.85)
            s3d.match_stats(reset=True)
            got = d.match_multi(t, ref, tar, 0.85, shards=shards)
            for k in ("gIdx", "sIdx", "gDist", "sDist", "pairs"):
                assert np.array_equal(got[k], want[k]), (shards, path, t, k)
            if t != 1:
                for k in ("gIdx2", "sIdx2", "gDist2", "sDist2"):
                    assert np.array_equal(got[k], want[k]), (shards, path, t, k)
            rows, fb = s3d.match_stats()
            if path == "tensor":
                assert rows >= n_ref, "the tensor-core pass did not run"   # every shard re-ranks its block: the blocks add up
    finally:
        s3d.set_match_path(s3d.api.MATCH_AUTO)


def test_match_multi_signed_sets_take_the_exact_kernel_on_every_shard(s3d):
    d = __import__("importlib").import_module("3dsift_b200.dist")
    rng = np.random.default_rng(9)
    a = rng.standard_normal((2100, 768)).astype(np.float32); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = (a[rng.permutation(2100)] + 0.05 * rng.standard_normal((2100, 768)).astype(np.float32)); b /= np.linalg.norm(b, axis=1, keepdims=True)
    b = b.astype(np.float32)
    b[:, 0] = np.abs(b[:, 0])
    s3d.set_match_path(s3d.api.MATCH_EXACT)
    m = s3d.muBruteMatcher(); m.enhancedMatch(a, b, 0.85)
    s3d.set_match_path(s3d.api.MATCH_AUTO)
    s3d.match_stats(reset=True)
    got = d.match_multi(3, a, b, 0.85, shards=4)
    assert np.array_equal(got["gIdx"], m.getGlodenIdx()) and np.array_equal(got["pairs"], m.pairs)
    assert s3d.match_stats()[0] == 0


def _nccl_c_worker(rank, world, port_no, ref, tar, q):
    import os, sys, importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = importlib.import_module("3dsift_b200.dist")
    dr, dt = torch.from_numpy(ref).cuda(), torch.from_numpy(tar).cuda()
    res = {}
    for t in (1, 3):
        r = d.match_sharded_c(t, dr, dt, 0.85)
        res[t] = {k: v.cpu().numpy() for k, v in r.items()}
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_match_sharded_library_nccl_two_gpus(s3d, synth, port):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ref, tar, _ = synth.d_synth_pair(2300, seed=67, k_tar=2250)
    want = {t: port.match(t, ref, tar, 0.85) for t in (1, 3)}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_c_worker, args=(r, 2, 29660, ref, tar, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank in (0, 1):
        for t in (1, 3):
            for k in ("gIdx", "sIdx", "gDist", "sDist", "pairs"):
                assert np.array_equal(got[rank][t][k], want[t][k]), (rank, t, k)
