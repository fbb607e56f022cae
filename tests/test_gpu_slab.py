"""z-slab sharded extraction must equal the unsharded extraction: every level
bit-exact on the owned planes, the same detections, codes, keypoint records and descriptors in the
same order.  G logical shards run on ONE device here (device copies instead of NCCL send/recv; the
plan, the stages and the kernels are the ones the multi-process path uses)."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
os.environ["S3D_SLAB_POISON"] = "1"     # level buffers start as NaN: an unfilled halo plane cannot pass by luck


def _unsharded(s3d, vol):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol, keep_levels=True)
    s.KpSiftAlgorithm()
    return s


def _same_records(a, b):
    fields = [f for f in a.dtype.names if f != "desc"]
    return all(np.array_equal(a[f], b[f]) for f in fields)


# (160, ...) with 4 or 3 shards: the top shard's local plane count (78 / 92) lies in another FP32 binade than
# the global 160, so the right-edge blend fraction of the z pass (c' = 2(n-1) - c - 0.1f, "whatever FP32
# gives") only matches the reference when it is evaluated in GLOBAL coordinates; (.., 46, 50) takes the
# generic kernel (nx % 4 != 0).
# min_nz = the S3D_SLAB_MIN_NZ switch: 0 shards every octave whose slabs are at least 10 planes thick (halo exchanges
# per level, window halos, then the gather into the replicated octaves), 1000 replicates from octave 0 on (the input
# itself is gathered), 96 is the default rule (these volumes: octave 0 sharded when nz >= 96).
@pytest.mark.parametrize("shape,shards,min_nz", [((96, 64, 72), 2, 0), ((96, 64, 72), 3, 0), ((128, 48, 64), 4, 0), ((70, 66, 68), 2, 0),
                                                 ((160, 48, 64), 4, 0), ((160, 46, 50), 3, 0), ((64, 96, 160), 5, 0),
                                                 ((96, 64, 72), 3, 1000), ((64, 64, 128), 4, 96), ((48, 56, 200), 8, 0)])
@pytest.mark.parametrize("group", [1, 3, 6])
def test_slabs_equal_unsharded(s3d, synth, shape, shards, min_nz, group, monkeypatch):
    d = importlib.import_module("3dsift_b200.dist")
    monkeypatch.setenv("S3D_SLAB_MIN_NZ", str(min_nz))
    monkeypatch.setenv("S3D_SLAB_GROUP", str(group))      # levels per halo exchange (1 = an exchange before every level)
    vol = synth.v_blobs(shape, seed=11)
    ref = _unsharded(s3d, vol)
    out = d.extract_slabs(vol, shards=shards, params=dict(keep_levels=1), keep=True)
    assert out["merged"]
    G, D = 6, 5
    nchecked = 0
    for g, sh in out["shards"].items():
        for o in range(sh.noct):
            for which, per in ((0, G), (1, D)):
                for i in range(per):
                    got, (za, zb, p0, p1) = sh.get_level_host(which, o * per + i)
                    if got is None or p1 <= p0:
                        continue
                    if o >= sh.first_replicated:
                        assert (za, zb) == (0, ref.level_dims(o)[2])          # replicated octave: every plane, on every shard
                        p0, p1 = za, zb
                    want = (ref.GET_GSS if which == 0 else ref.GET_DOG)(o * per + i)
                    got = got.reshape(zb - za, *want.shape[1:])
                    assert np.array_equal(got[p0 - za:p1 - za], want[p0:p1]), (g, o, which, i)
                    nchecked += 1
    assert nchecked > 20
    kp_r, codes_r, xyz_r = ref.extrema()
    assert np.array_equal(out["xyz5"], xyz_r)
    assert np.array_equal(out["codes"], codes_r)
    assert _same_records(out["extrema"], kp_r)
    kps = ref.GetKeypoints()
    assert len(out["kp"]) == len(kps) and len(kps) > 20
    assert _same_records(out["kp"], kps)
    assert np.array_equal(out["desc"], ref.descriptors)
    for sh in out["shards"].values():
        sh.close()


def test_shards_on_explicit_devices_and_phase_times(s3d, synth):
    """The devices list of s3d_extract_multi (here: the same device three times) and the per-phase device times."""
    d = importlib.import_module("3dsift_b200.dist")
    vol = synth.v_blobs((64, 64, 128), seed=4)
    ref = _unsharded(s3d, vol)
    t = {}
    out = d.extract_slabs(vol, shards=3, devices=[0, 0, 0], timing=t, with_extrema=False)
    assert _same_records(out["kp"], ref.GetKeypoints()) and np.array_equal(out["desc"], ref.descriptors)
    assert "extrema" not in out
    assert len(t["per_shard"]) == 3 and all(p["pyramid"] > 0 and p["sparse"] > 0 for p in t["per_shard"])


def test_single_shard_is_the_plain_path(s3d, synth):
    d = importlib.import_module("3dsift_b200.dist")
    vol = synth.v_blobs((64, 64, 64), seed=2)
    ref = _unsharded(s3d, vol)
    out = d.extract_slabs(vol, shards=1)
    assert _same_records(out["kp"], ref.GetKeypoints()) and np.array_equal(out["desc"], ref.descriptors)
