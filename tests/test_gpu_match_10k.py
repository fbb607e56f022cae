"""Matcher parity at the size SURVEY.md §7.1 step 5 names: 10 k x 10 k D-synth sets, all three match types, the exact
CUDA-core kernel and both tensor-core candidate kernels, against the compiled reference's muBruteMatcher
(Src/cMatcher.cc:146-215) — indices, float distances and pair lists bit for bit.  One reference run per type
(the reference recomputes the forward search each time: ~10-20 s each on the box's cores)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAMES = {1: "injectMatch", 2: "bijectMatch", 3: "enhancedMatch"}


@pytest.fixture(scope="module")
def sets10k(synth):
    return synth.d_synth_pair(10000, seed=4242)


@pytest.fixture(scope="module")
def want10k(sets10k, checker):
    ref, tar, _ = sets10k
    return {t: checker.match(t, ref, tar, 0.85) for t in (1, 2, 3)}


@pytest.mark.parametrize("path", ["exact", "single", "pair"])
def test_match_10k_vs_reference(s3d, sets10k, want10k, path):
    ref, tar, truth = sets10k
    s3d.set_match_path({"exact": s3d.api.MATCH_EXACT, "single": s3d.api.MATCH_TENSOR_SINGLE, "pair": s3d.api.MATCH_TENSOR_PAIR}[path])
    try:
        for t in (1, 2, 3):
            s3d.match_stats(reset=True)
            m = s3d.muBruteMatcher()
            getattr(m, NAMES[t])(ref, tar, 0.85)
            want = want10k[t]
            assert np.array_equal(m.getGlodenIdx(), want["gIdx"]), (path, t)
            assert np.array_equal(m.getSilverIdx(), want["sIdx"]), (path, t)
            assert np.array_equal(m.getGlodenDistSquare(), want["gDist"]), (path, t)
            assert np.array_equal(m.getSilverDistSquare(), want["sDist"]), (path, t)
            assert np.array_equal(m.pairs, want["pairs"]), (path, t)
            rows, fb = s3d.match_stats()
            if path != "exact":
                assert rows >= len(ref)
            print(f"{path} {NAMES[t]}: {len(m.pairs)} pairs, tc rows {rows}, fallback rows {fb}, {m.totalTime * 1e3:.1f} ms")
        hit = truth[m.pairs[:, 0]] == m.pairs[:, 1]
        assert hit.mean() > 0.99
    finally:
        s3d.set_match_path(s3d.api.MATCH_AUTO)


def _signed_sets(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, 768)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    perm = rng.permutation(n)
    b = (a[perm] + 0.05 * rng.standard_normal((n, 768)).astype(np.float32))
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    return a, b.astype(np.float32)


@pytest.mark.parametrize("kind", ["signed", "large", "nonfinite"])
def test_tensor_path_refuses_inputs_outside_its_guard(s3d, kind):
    """The tensor-core guard (a8*(1+1.2e-3)+1e-6) is only a bound for non-negative descriptors with entries <= 1
    (the reference's own output, Src/cSIFT3D.cc:1350-1358).  s3d_match takes arbitrary floats: sets with negative,
    large or non-finite entries must be routed to the exact kernel, at sizes where `auto` would pick tensor cores
    (>= 4e6 pairs), and give the exact kernel's bits."""
    a, b = _signed_sets(2200, seed=3)           # 4.84e6 pairs
    if kind == "large":
        a, b = np.abs(a) * 1500.0, np.abs(b) * 1500.0
    elif kind == "nonfinite":
        a, b = np.abs(a), np.abs(b)
        b[17, 5] = np.inf
    out = {}
    for path in (s3d.api.MATCH_EXACT, s3d.api.MATCH_AUTO, s3d.api.MATCH_TENSOR):
        s3d.set_match_path(path)
        s3d.match_stats(reset=True)
        m = s3d.muBruteMatcher()
        m.enhancedMatch(a, b, 0.85)
        out[path] = (m.getGlodenIdx().copy(), m.getSilverIdx().copy(), m.getGlodenDistSquare().copy(), m.pairs.copy(), s3d.match_stats())
    s3d.set_match_path(s3d.api.MATCH_AUTO)
    ex = out[s3d.api.MATCH_EXACT]
    for path in (s3d.api.MATCH_AUTO, s3d.api.MATCH_TENSOR):
        got = out[path]
        assert np.array_equal(got[0], ex[0]) and np.array_equal(got[1], ex[1])
        assert np.array_equal(got[2], ex[2], equal_nan=True) and np.array_equal(got[3], ex[3])
        assert got[4][0] == 0, f"{kind}: {got[4][0]} rows went through the tensor-core pass"
    if kind == "signed":
        assert len(ex[3]) > 1500


def test_tensor_path_still_used_for_reference_shaped_sets(s3d, synth):
    ref, tar, _ = synth.d_synth_pair(2200, seed=5)
    s3d.match_stats(reset=True)
    m = s3d.muBruteMatcher()
    m.enhancedMatch(ref, tar, 0.85)
    rows, fb = s3d.match_stats()
    assert rows >= 2200
