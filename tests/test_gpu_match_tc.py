"""The tensor-core matcher path (tcgen05 candidate pass + exact re-rank + guard/fallback) must give
the same integers and the same float distances as the exact path and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NAMES = {1: "injectMatch", 2: "bijectMatch", 3: "enhancedMatch"}


@pytest.fixture(params=["single", "pair"])
def tc_path(s3d, request):
    """Both candidate kernels: one CTA per tile (tc_topk_kernel) and CTA pairs with the query tile
    resident in shared memory (tc_pair_topk_kernel)."""
    s3d.set_match_path(s3d.api.MATCH_TENSOR_SINGLE if request.param == "single" else s3d.api.MATCH_TENSOR_PAIR)
    s3d.match_stats(reset=True)
    yield
    s3d.set_match_path(s3d.api.MATCH_AUTO)


def _run(s3d, t, ref, tar, thr=0.85):
    m = s3d.muBruteMatcher()
    getattr(m, NAMES[t])(ref, tar, thr)
    return m


@pytest.mark.parametrize("t", [1, 2, 3])
@pytest.mark.parametrize("n_ref,n_tar", [(300, 257), (1000, 1100), (129, 513), (2500, 3000)])
def test_tc_path_vs_oracle(s3d, synth, checker, tc_path, t, n_ref, n_tar):
    ref, tar, _ = synth.d_synth_pair(n_ref, seed=n_ref + t, k_tar=n_tar)
    want = checker.match(t, ref, tar, 0.85)
    m = _run(s3d, t, ref, tar)
    assert np.array_equal(m.getGlodenIdx(), want["gIdx"])
    assert np.array_equal(m.getSilverIdx(), want["sIdx"])
    assert np.array_equal(m.getGlodenDistSquare(), want["gDist"])
    assert np.array_equal(m.getSilverDistSquare(), want["sDist"])
    assert np.array_equal(m.pairs, want["pairs"])
    rows, fb = s3d.match_stats()
    assert rows >= n_ref                      # the tensor-core path really ran
    print(f"tc rows {rows}, exact-fallback rows {fb}")


def test_tc_path_quirks(s3d, port, tc_path):
    rng = np.random.default_rng(1)
    tar = np.abs(rng.standard_normal((600, 768))).astype(np.float32)
    tar /= np.linalg.norm(tar, axis=1, keepdims=True)
    tar[9] = tar[4]; tar[300] = tar[4]; tar[599] = tar[4]      # exact duplicates incl. another N-tile
    ref = tar[[4, 0, 12, 0, 599]].copy()
    ref[2] = 0                                                    # all-zero row: idx -1, dist 2
    ref[3] = (0.51 * tar[0] + 0.49 * tar[1]) / np.linalg.norm(0.51 * tar[0] + 0.49 * tar[1])
    ref = np.concatenate([ref, tar[20:200]])
    for t in (1, 2, 3):
        want = port.match(t, ref, tar, 0.85)
        m = _run(s3d, t, ref, tar)
        assert np.array_equal(m.getGlodenIdx(), want["gIdx"]) and np.array_equal(m.getSilverIdx(), want["sIdx"])
        assert np.array_equal(m.getGlodenDistSquare(), want["gDist"])
        assert np.array_equal(m.pairs, want["pairs"])


def test_tc_equals_exact_at_20k(s3d, synth, tc_path):
    """Beyond what the CPU oracle finishes quickly: the two GPU paths must agree bit for bit."""
    ref, tar, truth = synth.d_synth_pair(20000, seed=7)
    m_tc = _run(s3d, 3, ref, tar)
    rows, fb = s3d.match_stats()
    s3d.set_match_path(s3d.api.MATCH_EXACT)
    m_ex = _run(s3d, 3, ref, tar)
    assert np.array_equal(m_tc.getGlodenIdx(), m_ex.getGlodenIdx())
    assert np.array_equal(m_tc.getSilverIdx(), m_ex.getSilverIdx())
    assert np.array_equal(m_tc.getGlodenDistSquare(), m_ex.getGlodenDistSquare())
    assert np.array_equal(m_tc.getSilverDistSquare(), m_ex.getSilverDistSquare())
    assert np.array_equal(m_tc.pairs, m_ex.pairs)
    hit = truth[m_tc.pairs[:, 0]] == m_tc.pairs[:, 1]
    assert hit.mean() > 0.99 and len(m_tc.pairs) > 10000
    assert fb <= 0.01 * rows                  # the candidate pass itself is right (fallback is the exception)
    print(f"20k x 20k: tc rows {rows}, exact-fallback rows {fb} ({100.0 * fb / max(rows, 1):.2f} %), time {m_tc.totalTime * 1e3:.1f} ms "
          f"vs exact {m_ex.totalTime * 1e3:.1f} ms")
