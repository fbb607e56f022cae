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


class _DevMem:
    """Zero-copy view of raw device memory for torch (``torch.as_tensor(_DevMem(...), device='cuda')``)."""

    def __init__(self, ptr, nelem, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(nelem),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class SlabShard:
    """One shard of a z-slab extraction: an s3d slab handle plus its plane bookkeeping."""

    def __init__(self, gid, vol_ext, dims, own, params=None, device=-1, stream=None):
        L = api.lib()
        self.gid = gid
        self.nx, self.ny, self.nz = dims
        self.own = own
        p = api.s3d_params()
        L.s3d_default_params(C.byref(p))
        for k, v in (params or {}).items():
            setattr(p, k, v)
        p.device = device
        p.stream = C.c_void_p(int(stream)) if stream else None
        self._p = p
        self._h = C.c_void_p()
        on_dev = hasattr(vol_ext, "is_cuda") and vol_ext.is_cuda
        api.check(L.s3d_slab_create(api._ptr(vol_ext), int(on_dev), self.nx, self.ny, self.nz, own[0], own[1], C.byref(p),
                                    C.byref(self._h)))

    def local_max(self):
        v = C.c_float()
        api.check(api.lib().s3d_slab_local_max(self._h, C.byref(v)))
        return float(v.value)

    def begin(self, gmax):
        api.check(api.lib().s3d_slab_begin(self._h, C.c_float(gmax)))
        n, h, g = C.c_int(), C.c_int(), C.c_int()
        api.check(api.lib().s3d_slab_info(self._h, C.byref(n), C.byref(h), C.byref(g)))
        self.noct, self.halo, self.G = n.value, h.value, g.value

    def seed(self, o):
        api.check(api.lib().s3d_slab_seed(self._h, o))

    def octave(self, o):
        api.check(api.lib().s3d_slab_octave(self._h, o))

    def level(self, which, idx):
        """(torch tensor [local planes, ny_o*nx_o] aliasing the level buffer, (za, zb, p0, p1))."""
        import torch
        ptr = C.c_void_p()
        ext = (C.c_int * 4)()
        api.check(api.lib().s3d_slab_level_buffer(self._h, which, idx, C.byref(ptr), ext))
        za, zb, p0, p1 = (int(v) for v in ext)
        per = self.G if which == 0 else self.G - 1
        o = idx // per
        plane = (self.nx >> o) * (self.ny >> o)
        n = (zb - za) * plane
        if n == 0:
            return None, (za, zb, p0, p1)
        t = torch.as_tensor(_DevMem(ptr.value, n), device="cuda").view(zb - za, plane)
        return t, (za, zb, p0, p1)

    def maxima(self):
        n = self.noct * (self.G - 1)
        out = np.zeros(n, np.float32)
        api.check(api.lib().s3d_slab_get_maxima(self._h, api._ptr(out), n))
        return out

    def set_maxima(self, m):
        m = np.ascontiguousarray(m, np.float32)
        api.check(api.lib().s3d_slab_set_maxima(self._h, api._ptr(m), len(m)))

    def finish(self):
        api.check(api.lib().s3d_slab_finish(self._h))

    def results(self):
        L = api.lib()
        n = C.c_int()
        api.check(L.s3d_num_keypoints(self._h, C.byref(n)))
        k = n.value
        kp = np.zeros(max(k, 1), api.KP_DTYPE)
        desc = np.zeros((max(k, 1), api.DESC_LENGTH), np.float32)
        api.check(L.s3d_get_keypoints(self._h, api._ptr(kp), api._ptr(desc)))
        api.check(L.s3d_num_extrema(self._h, C.byref(n)))
        e = n.value
        ex = np.zeros(max(e, 1), api.KP_DTYPE)
        codes = np.zeros(max(e, 1), np.int32)
        xyz5 = np.zeros((max(e, 1), 5), np.int32)
        api.check(L.s3d_get_extrema(self._h, api._ptr(ex), api._ptr(codes), api._ptr(xyz5)))
        return dict(kp=kp[:k], desc=desc[:k], extrema=ex[:e], codes=codes[:e], xyz5=xyz5[:e])

    def device_results(self):
        """The shard's result arrays as torch tensors aliasing device memory (valid until close()):
        dict(kp uint8 [k,176], desc float32 [k,768], extrema uint8 [e,176], codes int32 [e,1], xyz5 int32 [e,5])."""
        import torch
        ptrs = (C.c_void_p * 5)()
        k, e = C.c_int(), C.c_int()
        api.check(api.lib().s3d_device_results(self._h, ptrs, C.byref(k), C.byref(e)))
        k, e = k.value, e.value
        spec = (("kp", k, 176, "|u1", torch.uint8), ("desc", k, api.DESC_LENGTH, "<f4", torch.float32),
                ("extrema", e, 176, "|u1", torch.uint8), ("codes", e, 1, "<i4", torch.int32), ("xyz5", e, 5, "<i4", torch.int32))
        out = {}
        for (name, n, w, ts, dt), p in zip(spec, ptrs):
            if n > 0 and p:
                out[name] = torch.as_tensor(_DevMem(p, n * w, ts), device="cuda").view(n, w)
            else:
                out[name] = torch.empty((0, w), dtype=dt, device="cuda")
        return out

    def get_level_host(self, which, idx):
        t, ext = self.level(which, idx)
        return (None if t is None else t.cpu().numpy()), ext

    def close(self):
        if self._h is not None and self._h.value:
            api.lib().s3d_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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


_PINNED = {}


def _to_host(t):
    """Device tensor -> fresh numpy array through a cached pinned staging buffer (a pageable .cpu() of the
    30 MB descriptor block runs at a few GB/s; pinned D2H + one host memcpy is several times faster)."""
    import torch
    n = t.numel() * t.element_size()
    if n == 0:
        return t.cpu().numpy()
    buf = _PINNED.get("buf")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 22), dtype=torch.uint8).pin_memory()
        _PINNED["buf"] = buf
    view = buf[:n].view(t.dtype).view(t.shape)
    view.copy_(t.contiguous(), non_blocking=False)
    return view.numpy().copy()


def _gather_merge_device(shard_results, world, group, num_kp_levels, with_extrema=True):
    """All ranks' shard results merged in the reference's order (octave, level, z, y, x), on the GPU:
    padded all-gathers of the device result arrays over NCCL, a stable sort of the per-row unit key
    (rows arrive shard-major and in raster order inside a shard, so a STABLE sort by unit is the whole
    merge, App. B Q16), one gather, one device->host copy per array.
    shard_results: list of device_results() dicts of the shards held by this process, in shard order."""
    import torch
    import torch.distributed as dist
    cat = {n: torch.cat([r[n] for r in shard_results]) if shard_results else None for n in ("kp", "desc", "extrema", "codes", "xyz5")}
    k = int(cat["kp"].shape[0]) if shard_results else 0
    e = int(cat["extrema"].shape[0]) if shard_results else 0
    if world > 1:
        cnt = torch.tensor([k, e], dtype=torch.int64, device="cuda")
        cnts = torch.empty((world, 2), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(cnts.view(-1), cnt, group=group)
        cnts = cnts.cpu()
        kmax, emax = int(cnts[:, 0].max()), int(cnts[:, 1].max())
    widths = dict(kp=(176, torch.uint8), desc=(api.DESC_LENGTH, torch.float32), extrema=(176, torch.uint8), codes=(1, torch.int32),
                  xyz5=(5, torch.int32))
    out = {}
    names = ("kp", "desc", "extrema", "codes", "xyz5") if with_extrema else ("kp", "desc")
    flat = {}
    for name in names:
        w, dt = widths[name]
        mine = cat[name] if shard_results else torch.empty((0, w), dtype=dt, device="cuda")
        if world == 1:
            flat[name] = mine
            continue
        is_k = name in ("kp", "desc")
        rows, n = (kmax, k) if is_k else (emax, e)
        buf = torch.zeros((max(rows, 1), w), dtype=dt, device="cuda")
        buf[:n] = mine
        allb = torch.empty((world,) + tuple(buf.shape), dtype=dt, device="cuda")
        dist.all_gather_into_tensor(allb.view(-1), buf.view(-1), group=group)
        valid = (torch.arange(max(rows, 1), device="cuda")[None, :] < cnts[:, 0 if is_k else 1].cuda()[:, None]).view(-1)
        flat[name] = allb.view(-1, w)[valid]
    # unit keys: octave and level sit at int32 words 4 and 5 of a keypoint record; xyz5 carries them as columns 3, 4
    kw = flat["kp"].contiguous().view(torch.int32).view(-1, 44)
    ko = torch.sort(kw[:, 4].long() * (num_kp_levels + 1) + kw[:, 5].long(), stable=True).indices
    out["kp"] = _to_host(flat["kp"][ko]).view(api.KP_DTYPE).reshape(-1)
    out["desc"] = _to_host(flat["desc"][ko])
    if with_extrema:
        x5 = flat["xyz5"]
        eo = torch.sort(x5[:, 3].long() * (num_kp_levels + 1) + x5[:, 4].long(), stable=True).indices
        out["extrema"] = _to_host(flat["extrema"][eo]).view(api.KP_DTYPE).reshape(-1)
        out["codes"] = _to_host(flat["codes"][eo]).reshape(-1)
        out["xyz5"] = _to_host(x5[eo])
    return out


def extract_slabs(volume, shards=None, group=None, params=None, keep=False, timing=None, with_extrema=True):
    """Full extraction of ONE volume split into z-slabs.

    * distributed (torch.distributed initialised, ``shards`` None): one shard per rank of ``group``;
      every rank passes the same host ``volume`` ([nz, ny, nx] float32; only its own planes + halo
      are uploaded) and receives the merged result.
    * single process (``shards`` = G): G logical shards on the current device, same code path with
      device copies instead of send/recv — the CI check that sharded == unsharded.

    ``with_extrema`` = False skips gathering the per-detection debug records (extrema, codes, xyz5).
    ``timing`` (a dict) receives device-synchronised wall seconds per phase: upload, pyramid (blur
    chain + seed halo exchanges + scalar all-reduces), halo (descriptor-window planes), sparse, gather.

    Returns dict(kp, desc, extrema, codes, xyz5[, shards]) in the reference's order."""
    import time

    import torch
    import torch.distributed as dist

    def tick(name, _t=[None]):
        if timing is None:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None and _t[0] is not None:
            timing[name] = timing.get(name, 0.0) + now - _t[0]
        _t[0] = now

    tick(None)
    distributed = shards is None and dist.is_initialized() and dist.get_world_size(group) > 1
    if distributed:
        world, me = dist.get_world_size(group), dist.get_rank(group)
        G = world
        owner = list(range(G))
    else:
        G = int(shards or 1)
        world, me = 1, 0
        owner = [0] * G
    nz, ny, nx = (int(v) for v in volume.shape)
    bounds = slab_bounds(nz, G)
    held = [g for g in range(G) if owner[g] == me and bounds[g + 1] > bounds[g]]
    L = api.lib()
    p = api.s3d_params()
    L.s3d_default_params(C.byref(p))
    for k, v in (params or {}).items():
        setattr(p, k, v)
    nlev = p.num_kp_levels

    def ext_of(g, o):
        e = (C.c_int * 4)()
        if bounds[g + 1] <= bounds[g]:
            return (0, 0, 0, 0)
        api.check(L.s3d_slab_extent(nz, bounds[g], bounds[g + 1], C.byref(p), o, e))
        return tuple(int(v) for v in e)

    # the shards run on torch's current stream so that the plane copies / NCCL calls below are
    # ordered with the kernels (0 = the legacy default stream: pass its explicit handle, 0x1)
    stream = torch.cuda.current_stream().cuda_stream or 1
    sh = {}
    for g in held:
        za, zb, _, _ = ext_of(g, 0)
        sh[g] = SlabShard(g, np.ascontiguousarray(volume[za:zb], dtype=np.float32), (nx, ny, nz), (bounds[g], bounds[g + 1]),
                          params, device=torch.cuda.current_device(), stream=stream)
    tick("upload")
    # global max|v| (data_scale, Src/cUtil.cc:538-550)
    m = max([sh[g].local_max() for g in held] or [0.0])
    if distributed:
        t = torch.tensor([m], dtype=torch.float32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        m = float(t.item())
    for g in held:
        sh[g].begin(m)
    any_s = sh[held[0]] if held else None
    noct = any_s.noct if any_s else 0
    Gl = any_s.G if any_s else nlev + 3
    if distributed:  # ranks that hold nothing still take part in the collectives
        t = torch.tensor([noct, Gl], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        noct, Gl = int(t[0]), int(t[1])
    for o in range(1, noct):
        for g in held:
            sh[g].seed(o)
        exts = [ext_of(g, o) for g in range(G)]
        lv = {g: sh[g].level(0, o * Gl)[0] for g in held}
        exchange_halos(lv, exts, owner, me, group)
        for g in held:
            sh[g].octave(o)
    # global max|DoG| per level (Detect_KeyPoints threshold, Src/cSIFT3D.cc:384-385)
    mx = np.zeros(noct * (Gl - 1), np.float32)
    for g in held:
        mx = np.maximum(mx, sh[g].maxima())
    if distributed:
        t = torch.from_numpy(mx).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        mx = t.cpu().numpy()
    for g in held:
        sh[g].set_maxima(mx)
    tick("pyramid")
    # descriptor windows reach into the neighbours' planes: fill the halos of levels 1..L of every octave
    for o in range(noct):
        exts = [ext_of(g, o) for g in range(G)]
        for i in range(1, nlev + 1):
            lv = {g: sh[g].level(0, o * Gl + i)[0] for g in held}
            exchange_halos(lv, exts, owner, me, group)
    tick("halo")
    for g in held:
        sh[g].finish()
    tick("sparse")
    # results stay on the device until they are merged: all-gather + stable sort + one D2H per array
    out = _gather_merge_device([sh[g].device_results() for g in held], world if distributed else 1, group, nlev, with_extrema)
    tick("gather")
    if keep:
        out["shards"] = sh
        out["bounds"] = bounds
    else:
        for g in held:
            sh[g].close()
    return out
