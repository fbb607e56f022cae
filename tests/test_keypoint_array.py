"""GetKeypoints() returns records whose `desc` pointers borrow from a descriptor block (Src/cSIFT3D.cc:486,495).
The Python mirror ties the block's lifetime to the returned array: slices, copies and pickled copies
stay valid after the extractor — and the original block — are gone; records without an owner are refused."""
import gc
import importlib
import pickle

import numpy as np
import pytest

api = importlib.import_module("3dsift_b200.api")


def _make(n=6, seed=0):
    rng = np.random.default_rng(seed)
    kp = np.zeros(n, api.KP_DTYPE)
    kp["x"] = np.arange(n)
    desc = rng.random((n, api.DESC_LENGTH), dtype=np.float32)
    return api._bind_descriptors(kp, desc, rows=np.arange(n)), desc.copy()


def test_views_and_copies_keep_the_block_alive():
    k, want = _make()
    s, c, f = k[1:4], k.copy(), k[[5, 0, 2]]
    del k
    gc.collect()
    assert np.array_equal(api._desc_matrix(s), want[1:4])
    assert np.array_equal(api._desc_matrix(c), want)
    assert np.array_equal(api._desc_matrix(f), want[[5, 0, 2]])


def test_pickle_ships_the_block_and_repoints():
    k, want = _make()
    blob = pickle.dumps(k[[4, 1]])
    del k
    gc.collect()
    r = pickle.loads(blob)
    assert isinstance(r, api.KeypointArray) and np.array_equal(r["x"], [4, 1])
    assert np.array_equal(api._desc_matrix(r), want[[4, 1]])
    assert np.array_equal(np.diff(r["desc"].astype(np.int64)), [api.DESC_LENGTH * 4])


def test_records_without_owner_are_refused():
    k, _ = _make()
    plain = np.array(k.view(np.ndarray))        # raw addresses, no owner
    with pytest.raises(api.S3DError):
        api._desc_matrix(plain)
    bad = k.copy()
    bad["desc"][0] += 1 << 40
    with pytest.raises(api.S3DError):
        api._desc_matrix(bad)


def test_plain_matrices_still_accepted():
    a = np.random.rand(3, 768).astype(np.float32)
    assert np.array_equal(api._desc_matrix(a), a)
