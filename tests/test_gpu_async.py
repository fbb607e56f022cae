"""The asynchronous halves of the C ABI and the device block cache: results fetched with
s3d_get_keypoints_async + s3d_sync equal the blocking s3d_get_keypoints; handles whose lifetimes overlap on
different streams (the e2e pipeline of bench.py) give the same bits as one handle at a time; s3d_trim_cache
returns the cached blocks and extraction still works afterwards."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _extract_blocking(s3d, vol):
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    s.KpSiftAlgorithm()
    kp = s.GetKeypoints().copy()
    desc = s.descriptors.copy()
    s.close()
    return kp, desc


def _fields_equal(a, b):
    return all(np.array_equal(a[f], b[f]) for f in a.dtype.names if f != "desc")


def test_async_fetch_equals_blocking(s3d, synth):
    import torch
    vol = synth.v_blobs(64, seed=3)
    want_kp, want_desc = _extract_blocking(s3d, vol)
    assert len(want_kp) > 0
    s = s3d.CSIFT3DFactory.CreateCSIFT3D(vol)
    s.run_async()
    s.wait()
    n = s.num_keypoints()
    assert n == len(want_kp)
    h_kp = torch.zeros((n, 176), dtype=torch.uint8).pin_memory()
    h_desc = torch.zeros((n, 768), dtype=torch.float32).pin_memory()
    s.get_keypoints_async(h_kp.data_ptr(), h_desc.data_ptr())
    s.sync()
    got_kp = np.frombuffer(h_kp.numpy().tobytes(), dtype=s3d.api.KP_DTYPE)
    assert _fields_equal(got_kp, want_kp)
    assert np.array_equal(h_desc.numpy(), want_desc)
    s.close()


def test_overlapping_handles_equal_serial(s3d, synth):
    """Three volumes in the bench's e2e pattern: next volume uploading while the current one runs and the previous
    one's results are still in flight."""
    import torch
    vols = [synth.v_blobs(64, seed=10 + i) for i in range(4)]
    want = [_extract_blocking(s3d, v) for v in vols]
    pinned = [torch.from_numpy(v).pin_memory() for v in vols]
    cap = max(len(k) for k, _ in want) + 8
    h_kp = [torch.zeros((cap, 176), dtype=torch.uint8).pin_memory() for _ in range(2)]
    h_desc = [torch.zeros((cap, 768), dtype=torch.float32).pin_memory() for _ in range(2)]
    up = lambda i: s3d.CSIFT3DFactory.CreateCSIFT3D(pinned[i], x_dim=64, y_dim=64, z_dim=64, async_upload=True)
    got = [None] * len(vols)

    def finish(h, i, n):
        h.sync()
        kp = np.frombuffer(h_kp[i & 1].numpy()[:n].tobytes(), dtype=s3d.api.KP_DTYPE).copy()
        got[i] = (kp, h_desc[i & 1].numpy()[:n].copy())
        h.close()

    for rep in range(2):   # second round runs entirely on recycled blocks
        cur, pend = up(0), None
        for i in range(len(vols)):
            nxt = up(i + 1) if i + 1 < len(vols) else None
            cur.KpSiftAlgorithm()
            if pend is not None:
                finish(*pend)
            n = cur.num_keypoints()
            cur.get_keypoints_async(h_kp[i & 1].data_ptr(), h_desc[i & 1].data_ptr())
            pend = (cur, i, n)
            cur = nxt
        finish(*pend)
        for i, (kp, desc) in enumerate(got):
            assert len(kp) == len(want[i][0]), (rep, i)
            assert _fields_equal(kp, want[i][0]), (rep, i)
            assert np.array_equal(desc, want[i][1]), (rep, i)


@pytest.mark.skipif(os.environ.get("S3D_ALLOC") == "pool", reason="block cache switched off")
def test_trim_cache(s3d, synth):
    vol = synth.v_blobs(48, seed=1)
    want_kp, want_desc = _extract_blocking(s3d, vol)
    cached = s3d.trim_cache()
    assert cached > 0, "a finished extraction leaves its blocks in the cache"
    assert s3d.trim_cache() == 0
    kp, desc = _extract_blocking(s3d, vol)
    assert _fields_equal(kp, want_kp) and np.array_equal(desc, want_desc)
