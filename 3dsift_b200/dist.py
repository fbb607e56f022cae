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
