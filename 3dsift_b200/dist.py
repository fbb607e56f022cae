"""Multi-GPU partitioning of the hot path — one process per GPU, torch.distributed.

* ``extract_batch``   independent volumes: volume b goes to rank b mod world; no collective on the
                      data path (results are gathered as host objects at the end).
* ``match_sharded``   the SEARCHED set (tar in the forward pass, ref in the masked reverse pass,
                      Src/cMatcher.cc:58) is split into contiguous index ranges, one per rank; the
                      query set is replicated.  Every rank computes per-query top-2 (dot, global
                      index) over its range, the partial lists are all-gathered (16 B/query/rank)
                      and merged under the total order (dot desc, index asc) — identical to the
                      reference's sequential strict-'>' scan — then the cheap filters run
                      replicated on every rank.

The compute primitives are injected (``ops``): ``CudaOps`` drives the C ABI on CUDA tensors over
NCCL; the CPU tests drive the same orchestration over gloo with a checker-backed ops object.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def shard_bounds(n, world):
    """Contiguous, balanced index ranges: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


class CudaOps:
    """The C-ABI primitives on torch CUDA tensors (current device / current stream)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.L = api.lib()
        self.dev = torch.device("cuda", torch.cuda.current_device())

    def to_device(self, a):
        if self.torch.is_tensor(a):  # already a (device-resident) descriptor block: no host round trip
            return a.to(self.dev, dtype=self.torch.float32).contiguous()
        return self.torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(self.dev)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.dev)

    def top2(self, q, db, offset, mask):
        t = self.torch
        nq = q.shape[0]
        d1, d2 = self.empty(nq, t.float64), self.empty(nq, t.float64)
        i1, i2 = self.empty(nq, t.int32), self.empty(nq, t.int32)
        api.check(self.L.s3d_top2_device(q.data_ptr(), nq, db.data_ptr() if db.shape[0] else None, db.shape[0], int(offset),
                                         mask.data_ptr() if mask is not None else None, d1.data_ptr(), i1.data_ptr(),
                                         d2.data_ptr(), i2.data_ptr(), self._stream()))
        return d1, i1, d2, i2

    def merge(self, d1, i1, d2, i2, mask):
        t = self.torch
        parts, nq = d1.shape
        gD, sD = self.empty(nq, t.float32), self.empty(nq, t.float32)
        gI, sI = t.full((nq,), -1, dtype=t.int32, device=self.dev), t.full((nq,), -1, dtype=t.int32, device=self.dev)
        gD.zero_(); sD.zero_()
        api.check(self.L.s3d_top2_merge_device(parts, nq, d1.data_ptr(), i1.data_ptr(), d2.data_ptr(), i2.data_ptr(),
                                               mask.data_ptr() if mask is not None else None, gD.data_ptr(), gI.data_ptr(),
                                               sD.data_ptr(), sI.data_ptr(), self._stream()))
        return gD, gI, sD, sI

    def ratio_filter(self, gI, gD, sD, thr):
        api.check(self.L.s3d_ratio_filter_device(gI.data_ptr(), gD.data_ptr(), sD.data_ptr(), gI.shape[0], float(thr), self._stream()))

    def count_mask(self, gI, n_tar, count_thres):
        mask = self.empty(max(n_tar, 1), self.torch.int32)
        api.check(self.L.s3d_count_mask_device(gI.data_ptr(), gI.shape[0], mask.data_ptr(), n_tar, count_thres, self._stream()))
        return mask[:n_tar]

    def biject_filter(self, gI, mask, gI2):
        api.check(self.L.s3d_biject_filter_device(gI.data_ptr(), gI.shape[0], mask.data_ptr(), gI2.data_ptr(), self._stream()))

    def all_gather(self, x, group=None):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        out = self.empty((world,) + tuple(x.shape), x.dtype)
        dist.all_gather_into_tensor(out.view(-1), x.contiguous().view(-1), group=group)
        return out

    def to_numpy(self, x):
        return x.cpu().numpy()


def _search(ops, q, db_full_len, db_shard, lo, mask, group, world):
    """One search direction: local top-2 over this rank's shard, all-gather, merge."""
    d1, i1, d2, i2 = ops.top2(q, db_shard, lo, mask)
    if world == 1:
        D1, I1, D2, I2 = d1[None], i1[None], d2[None], i2[None]
    else:
        D1, I1, D2, I2 = (ops.all_gather(x, group) for x in (d1, i1, d2, i2))
    return ops.merge(D1, I1, D2, I2, mask)


def match_sharded(mtype, ref, tar, thr=0.85, ops=None, group=None):
    """bijectMatchBase (Src/cMatcher.cc:146-215) with the searched set sharded over the ranks of
    ``group``.  ``ref`` / ``tar`` are the FULL n x 768 arrays (replicated on every rank).
    Returns a dict of numpy arrays identical on every rank: gIdx (post-filter), gDist, sIdx, sDist,
    pairs (and gIdx2 ... for biject/enhanced)."""
    import torch.distributed as dist
    ops = ops or CudaOps()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_ref, n_tar = len(ref), len(tar)
    bt, br = shard_bounds(n_tar, world), shard_bounds(n_ref, world)
    d_ref, d_tar = ops.to_device(ref), ops.to_device(tar)
    # forward: queries = ref (replicated), database = this rank's slice of tar
    gD, gI, sD, sI = _search(ops, d_ref, n_tar, d_tar[bt[rank]:bt[rank + 1]], bt[rank], None, group, world)
    ops.ratio_filter(gI, gD, sD, thr)
    out = {}
    if mtype != 1:
        mask = ops.count_mask(gI, n_tar, 0 if mtype == 2 else 1)
        # reverse: queries = tar (replicated, masked), database = this rank's slice of ref
        gD2, gI2, sD2, sI2 = _search(ops, d_tar, n_ref, d_ref[br[rank]:br[rank + 1]], br[rank], mask, group, world)
        ops.ratio_filter(gI2, gD2, sD2, thr)
        ops.biject_filter(gI, mask, gI2)
        out.update(gIdx2=ops.to_numpy(gI2), gDist2=ops.to_numpy(gD2), sIdx2=ops.to_numpy(sI2), sDist2=ops.to_numpy(sD2))
    g = ops.to_numpy(gI)
    keep = np.flatnonzero(g >= 0)
    out.update(gIdx=g, gDist=ops.to_numpy(gD), sIdx=ops.to_numpy(sI), sDist=ops.to_numpy(sD),
               pairs=np.stack([keep, g[keep]], 1).astype(np.int32))
    return out


def match_sharded_c(mtype, d_ref, d_tar, thr=0.85, group=None, comm=None, stream=None):
    """s3d_match_sharded: bijectMatchBase (Src/cMatcher.cc:146-215) over the library's NCCL communicator of ``group`` —
    database-sharded candidate pass, query-sharded exact re-rank, all inside the library.  ``d_ref`` / ``d_tar``: the
    FULL n x 768 float32 CUDA tensors (replicated on every rank).  Returns CUDA tensors, complete on every rank:
    dict(gIdx, gDist, sIdx, sDist, gIdx2, gDist2, sIdx2, sDist2, pairs [n_pairs, 2])."""
    import torch
    comm = comm or nccl_comm(group)
    n_ref, n_tar = int(d_ref.shape[0]), int(d_tar.shape[0])
    dev = d_ref.device
    I = lambda m: torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    F = lambda m: torch.empty(max(m, 1), dtype=torch.float32, device=dev)
    b = dict(gIdx=I(n_ref), gDist=F(n_ref), sIdx=I(n_ref), sDist=F(n_ref), gIdx2=I(n_tar), gDist2=F(n_tar), sIdx2=I(n_tar),
             sDist2=F(n_tar), pr=I(n_ref), pt=I(n_ref), np=I(1))
    st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    api.check(api.lib().s3d_match_sharded(comm._c, mtype, d_ref.data_ptr(), n_ref, d_tar.data_ptr(), n_tar, float(thr),
                                          *[b[k].data_ptr() for k in ("gIdx", "gDist", "sIdx", "sDist", "gIdx2", "gDist2", "sIdx2",
                                                                      "sDist2", "pr", "pt", "np")], C.c_void_p(st or 1)))
    n = int(b["np"].item())
    out = {k: b[k][: (n_ref if not k.endswith("2") else n_tar)] for k in ("gIdx", "gDist", "sIdx", "sDist", "gIdx2", "gDist2", "sIdx2", "sDist2")}
    out["pairs"] = torch.stack([b["pr"][:n], b["pt"][:n]], 1)
    return out


def match_multi(mtype, ref, tar, thr=0.85, devices=None, shards=None):
    """s3d_match_multi: the same in ONE process over several devices (host threads + peer copies), host arrays in and
    out.  ``devices`` = list of device ordinals, or ``shards`` = G logical shards on the current device (CI check)."""
    ref = np.ascontiguousarray(ref, dtype=np.float32).reshape(-1, api.DESC_LENGTH)
    tar = np.ascontiguousarray(tar, dtype=np.float32).reshape(-1, api.DESC_LENGTH)
    n_ref, n_tar = len(ref), len(tar)
    G = len(devices) if devices is not None else int(shards or 1)
    devs = (C.c_int * G)(*devices) if devices is not None else None
    A = lambda n, t: np.zeros(max(n, 1), t)
    r = dict(gIdx=A(n_ref, np.int32), gDist=A(n_ref, np.float32), sIdx=A(n_ref, np.int32), sDist=A(n_ref, np.float32),
             gIdx2=A(n_tar, np.int32), gDist2=A(n_tar, np.float32), sIdx2=A(n_tar, np.int32), sDist2=A(n_tar, np.float32),
             pr=A(n_ref, np.int32), pt=A(n_ref, np.int32))
    npairs = C.c_int()
    times = (C.c_double * 3)()
    api.check(api.lib().s3d_match_multi(mtype, api._ptr(ref), n_ref, api._ptr(tar), n_tar, float(thr), devs, G,
                                        *[api._ptr(r[k]) for k in ("gIdx", "gDist", "sIdx", "sDist", "gIdx2", "gDist2", "sIdx2", "sDist2", "pr", "pt")],
                                        C.cast(C.byref(npairs), C.c_void_p), C.cast(C.byref(times), C.c_void_p)))
    n = npairs.value
    out = {k: v[: (n_ref if not k.endswith("2") else n_tar)] for k, v in r.items() if k not in ("pr", "pt")}
    out["pairs"] = np.stack([r["pr"][:n], r["pt"][:n]], 1)
    out["seconds"] = float(times[2])
    return out


def extract_batch(volumes, group=None, **params):
    """Independent volumes over the ranks (volume b -> rank b mod world).  Returns on every rank the
    list [(keypoints KP_DTYPE array, descriptors K x 768)] in input order."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = []
    for b in range(rank, len(volumes), world):
        s = api.CSIFT3DFactory.CreateCSIFT3D(volumes[b], **params)
        s.KpSiftAlgorithm()
        kp = s.GetKeypoints()
        mine.append((b, kp.copy(), s.descriptors.copy()))
        s.close()
    if world == 1:
        gathered = [mine]
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
    out = [None] * len(volumes)
    for part in gathered:
        for b, kp, desc in part:
            out[b] = (kp, desc)
    return out


# ---- one large volume in z-slabs (BASELINE.json configs[2]) ------------------------

def slab_bounds(nz, shards):
    """Owned octave-0 plane ranges: contiguous, balanced, even starts (so decimation pairs stay together)."""
    b = [min(nz, (nz * r // shards) & ~1) for r in range(shards)] + [nz]
    b[0] = 0
    return b


def needed_from(ext_src, ext_dst):
    """Plane intervals [k0, k1) that shard `dst` must receive from shard `src` for one level: the part
    of dst's halo ([za, p0) and [p1, zb)) that src owns.  ext = (za, zb, p0, p1) in global planes."""
    za, zb, p0, p1 = ext_dst
    s0, s1 = ext_src[2], ext_src[3]
    out = []
    for lo, hi in ((za, p0), (p1, zb)):
        k0, k1 = max(lo, s0), min(hi, s1)
        if k1 > k0:
            out.append((k0, k1))
    return out


class SlabShard:
    """One shard's handle of a z-slab extraction (s3d_slab_run / s3d_extract_multi): the shard's own results (the merged
    results on the gather root), its local level buffers when the run kept them, its plane bookkeeping."""

    def __init__(self, handle, dims, gid=0):
        self._h = handle if isinstance(handle, C.c_void_p) else C.c_void_p(handle)
        self.gid = gid
        self.nx, self.ny, self.nz = dims
        n, h, g, f = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        api.check(api.lib().s3d_slab_info(self._h, C.byref(n), C.byref(h), C.byref(g), C.byref(f)))
        self.noct, self.halo, self.G, self.first_replicated = n.value, h.value, g.value, f.value

    def extent(self, which, idx):
        ptr = C.c_void_p()
        ext = (C.c_int * 4)()
        api.check(api.lib().s3d_slab_level_buffer(self._h, which, idx, C.byref(ptr), ext))
        return tuple(int(v) for v in ext)

    def get_level_host(self, which, idx):
        """(local planes [za, zb) of the level as a [zb-za, ny_o*nx_o] array or None, (za, zb, p0, p1))."""
        za, zb, p0, p1 = self.extent(which, idx)
        per = self.G if which == 0 else self.G - 1
        o = idx // per
        plane = (self.nx >> o) * (self.ny >> o)
        if zb <= za:
            return None, (za, zb, p0, p1)
        out = np.empty(((zb - za), plane), np.float32)
        api.check(api.lib().s3d_get_level(self._h, which, idx, api._ptr(out)))
        return out, (za, zb, p0, p1)

    def results(self, with_extrema=True):
        L = api.lib()
        n = C.c_int()
        api.check(L.s3d_num_keypoints(self._h, C.byref(n)))
        k = n.value
        kp = np.zeros(max(k, 1), api.KP_DTYPE)
        desc = np.zeros((max(k, 1), api.DESC_LENGTH), np.float32)
        api.check(L.s3d_get_keypoints(self._h, api._ptr(kp), api._ptr(desc)))
        out = dict(kp=kp[:k], desc=desc[:k])
        if with_extrema:
            api.check(L.s3d_num_extrema(self._h, C.byref(n)))
            e = n.value
            ex = np.zeros(max(e, 1), api.KP_DTYPE)
            codes = np.zeros(max(e, 1), np.int32)
            xyz5 = np.zeros((max(e, 1), 5), np.int32)
            api.check(L.s3d_get_extrema(self._h, api._ptr(ex), api._ptr(codes), api._ptr(xyz5)))
            out.update(extrema=ex[:e], codes=codes[:e], xyz5=xyz5[:e])
        return out

    def num_keypoints(self):
        n = C.c_int()
        api.check(api.lib().s3d_num_keypoints(self._h, C.byref(n)))
        return n.value

    def phases(self):
        """Device time per phase of this shard's run, ms: all-reduce(max) + normalise, pyramid (incl. halo exchanges),
        window halos, all-reduce + sparse stages, gather, and the upload (allocation + copy of the owned planes + max|v|)."""
        ms = (C.c_double * 8)()
        api.check(api.lib().s3d_slab_phases(self._h, ms))
        return dict(zip(("normalize", "pyramid", "halo", "sparse", "gather", "upload"), (float(v) for v in ms[:6])))

    def timers(self):
        t = (C.c_double * 10)()
        api.check(api.lib().s3d_get_timers(self._h, C.byref(t)))
        return list(t)

    def close(self):
        if self._h is not None and self._h.value:
            api.lib().s3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NcclComm:
    """s3d_comm over NCCL for one process per GPU: rank 0 draws the unique id inside the library, torch.distributed
    carries its 128 bytes to the other ranks (any group backend), ncclCommInitRank happens inside the library."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        L = api.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        uid = (C.c_ubyte * 128)()
        box = [None]
        if self.rank == 0:
            api.check(L.s3d_comm_unique_id(uid))
            box[0] = bytes(uid)
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._c = C.c_void_p()
        dev = torch.cuda.current_device() if device is None else device
        api.check(L.s3d_comm_create(uid, self.world, self.rank, dev, C.byref(self._c)))

    def traffic(self):
        a, b = C.c_ulonglong(), C.c_ulonglong()
        api.check(api.lib().s3d_comm_traffic(self._c, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if self._c is not None and self._c.value:
            api.lib().s3d_comm_destroy(self._c)
            self._c = C.c_void_p()


_COMMS = {}


def nccl_comm(group=None):
    """The process's s3d communicator for `group` (created on first use, kept for the life of the process)."""
    key = id(group)
    if key not in _COMMS:
        _COMMS[key] = NcclComm(group)
    return _COMMS[key]


def plan_from_library(nz, world, rank, octave, depth):
    """s3d_slab_plan: [(peer, 'recv' | 'send', k0, k1)] — the transfers of one halo fill as `rank` sees them."""
    cap = 4 * world + 8
    q = (C.c_int * (4 * cap))()
    n = C.c_int()
    api.check(api.lib().s3d_slab_plan(nz, world, rank, octave, depth, q, cap, C.byref(n)))
    return [(int(q[4 * i]), "send" if q[4 * i + 1] else "recv", int(q[4 * i + 2]), int(q[4 * i + 3])) for i in range(n.value)]


def exchange_plan(exts, owner_rank, my_rank):
    """Halo fill of one level as a list of transfers, identical (and identically ordered) on every rank.
    exts[g] = (za, zb, p0, p1) of shard g; owner_rank[g] = process that holds shard g.
    Returns [(src_gid, dst_gid, k0, k1, kind)] with kind in {"copy", "send", "recv"} for this rank."""
    plan = []
    for d in range(len(exts)):
        for s in range(len(exts)):
            if s == d:
                continue
            for k0, k1 in needed_from(exts[s], exts[d]):
                src_here, dst_here = owner_rank[s] == my_rank, owner_rank[d] == my_rank
                if src_here and dst_here:
                    plan.append((s, d, k0, k1, "copy"))
                elif src_here:
                    plan.append((s, d, k0, k1, "send"))
                elif dst_here:
                    plan.append((s, d, k0, k1, "recv"))
    return plan


def exchange_halos(levels, exts, owner_rank, my_rank, group=None):
    """levels: {gid: 2-D tensor [zb-za, plane]} for the shards held by this rank (all the same level).
    Fills every held shard's halo planes from their owners: device copies between shards of this
    process, NCCL/gloo send/recv (one batch) between processes."""
    import torch.distributed as dist
    ops = []
    for s, d, k0, k1, kind in exchange_plan(exts, owner_rank, my_rank):
        if kind == "copy":
            levels[d][k0 - exts[d][0]:k1 - exts[d][0]].copy_(levels[s][k0 - exts[s][0]:k1 - exts[s][0]])
        elif kind == "send":
            ops.append(dist.P2POp(dist.isend, levels[s][k0 - exts[s][0]:k1 - exts[s][0]], owner_rank[d], group))
        else:
            ops.append(dist.P2POp(dist.irecv, levels[d][k0 - exts[d][0]:k1 - exts[d][0]], owner_rank[s], group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def merge_shard_results(parts, num_kp_levels=3):
    """Concatenate shard outputs in the reference's order (octave, level, z, y, x): shards partition
    z, each shard's lists are already in raster order, so per (octave, level) unit the global list is
    the shards' segments in shard order (App. B Q16)."""
    def unit_of(o, l):
        return o.astype(np.int64) * (num_kp_levels + 1) + l
    out = {}
    kp_units = [unit_of(p["kp"]["octave"], p["kp"]["level"]) for p in parts]
    ex_units = [unit_of(p["xyz5"][:, 3], p["xyz5"][:, 4]) for p in parts]
    def order(units):
        key = np.concatenate(units) if units else np.zeros(0, np.int64)
        shard = np.concatenate([np.full(len(u), i, np.int64) for i, u in enumerate(units)]) if units else key
        pos = np.concatenate([np.arange(len(u), dtype=np.int64) for u in units]) if units else key
        return np.lexsort((pos, shard, key))
    ko, eo = order(kp_units), order(ex_units)
    out["kp"] = np.concatenate([p["kp"] for p in parts])[ko]
    out["desc"] = np.concatenate([p["desc"] for p in parts])[ko]
    out["extrema"] = np.concatenate([p["extrema"] for p in parts])[eo]
    out["codes"] = np.concatenate([p["codes"] for p in parts])[eo]
    out["xyz5"] = np.concatenate([p["xyz5"] for p in parts])[eo]
    return out


def _params(params):
    p = api.s3d_params()
    api.lib().s3d_default_params(C.byref(p))
    for k, v in (params or {}).items():
        setattr(p, k, v)
    return p


def extract_slabs(volume, shards=None, group=None, params=None, keep=False, timing=None, with_extrema=True, devices=None,
                  on_device=False):
    """Full extraction of ONE volume split into z-slabs, orchestrated inside the library.

    * distributed (torch.distributed initialised, ``shards`` None): one shard per rank of ``group`` over the library's
      NCCL communicator; every rank passes the same ``volume`` ([nz, ny, nx] float32, host — pinned for an asynchronous
      upload — or, with on_device, a CUDA tensor); only the rank's own planes are read.  Rank 0 of the group
      receives the merged result, the other ranks their own part (``out["merged"]`` tells which).
    * single process (``shards`` = G): G shards driven by G host threads of this process, shard g on
      ``devices[g]`` (default: all on the current device — logical shards, the CI check that sharded == unsharded).

    ``timing`` (a dict) receives the device milliseconds per phase of this rank's shard (see SlabShard.phases).
    Returns dict(kp, desc[, extrema, codes, xyz5], merged[, shards, bounds]) in the reference's order."""
    import torch.distributed as dist
    L = api.lib()
    distributed = shards is None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    nz, ny, nx = (int(v) for v in volume.shape)
    p = _params(params)
    if distributed:
        comm = nccl_comm(group)
        o0, o1 = C.c_int(), C.c_int()
        api.check(L.s3d_slab_bounds(nz, comm.world, comm.rank, C.byref(o0), C.byref(o1)))
        own = volume[o0.value:o1.value]
        if isinstance(own, np.ndarray):
            own = np.ascontiguousarray(own, dtype=np.float32)
        h = C.c_void_p()
        api.check(L.s3d_slab_run(comm._c, api._ptr(own), int(bool(on_device)), nx, ny, nz, C.byref(p), C.byref(h)))
        sh = SlabShard(h, (nx, ny, nz), comm.rank)
        api.check(L.s3d_slab_gather(comm._c, h, 0, int(with_extrema)))
        out = sh.results(with_extrema)
        out["merged"] = comm.rank == 0
        if timing is not None:
            timing.update(sh.phases())
        if keep:
            out["shards"] = {comm.rank: sh}
        else:
            sh.close()
        return out
    G = int(shards or 1)
    vol = np.ascontiguousarray(volume, dtype=np.float32) if isinstance(volume, np.ndarray) else volume
    hs = (C.c_void_p * G)()
    devs = (C.c_int * G)(*devices) if devices is not None else None
    api.check(L.s3d_extract_multi(api._ptr(vol), nx, ny, nz, C.byref(p), devs, G, int(with_extrema), hs))
    sh = {g: SlabShard(C.c_void_p(hs[g]), (nx, ny, nz), g) for g in range(G)}
    out = sh[0].results(with_extrema)
    out["merged"] = True
    if timing is not None:
        timing.update(sh[0].phases())
        timing["per_shard"] = [sh[g].phases() for g in range(G)]
    if keep:
        out["shards"] = sh
        out["bounds"] = slab_bounds(nz, G)
    else:
        for g in range(G):
            sh[g].close()
    return out
