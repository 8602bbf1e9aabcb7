"""-m gpu: the drop-in entry points end to end against (i) the golden fixtures produced by the real
reference (tests/golden/, oracle/make_golden.py), (ii) the scikit-learn oracle on seeded inputs, and
(iii) size-independent properties at larger N."""
import os
import os.path as osp

import numpy as np
import pytest
import torch

from oracle import reference_calls as RC
from oracle import restatement as RS
from videotofaces_b200.synth import blobs, eval_case, plant_dupes
from videotofaces_b200 import dupes as vdupes, grouping as vgrouping
from videotofaces_b200.main import video_to_faces

pytestmark = pytest.mark.gpu


def fake_faces(out_dir, n):
    fdir = osp.join(out_dir, 'faces')
    os.makedirs(fdir, exist_ok=True)
    paths = [osp.join(fdir, 'f%06d.jpg' % i) for i in range(n)]
    for p in paths:
        with open(p, 'wb') as f:
            f.write(b'x')
    return paths


def gold(golden_dir, name):
    return np.load(osp.join(golden_dir, name + '.npz'))


def read(golden_dir, name):
    with open(osp.join(golden_dir, name)) as f:
        return f.read()


@pytest.mark.parametrize('name,gen,thr', [
    ('dedup_400x512', lambda: plant_dupes(blobs(400, 512, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0], 0.2),
    ('dedup_300x768', lambda: plant_dupes(blobs(300, 768, kc=3, noise=1.0, seed=3, normalize=False), frac=10, seed=2,
                                          normalize=False)[0], 0.25),
])
def test_remove_dupes_overall_golden(tmp_path, golden_dir, capsys, name, gen, thr):
    g = gold(golden_dir, name)
    X = gen()
    paths = fake_faces(str(tmp_path), X.shape[0])
    os.makedirs(tmp_path / 'intermediate', exist_ok=True)
    Xk, goods = vdupes.remove_dupes_overall(X, paths, ('enc', thr, True, str(tmp_path)))
    kept = np.asarray([int(osp.basename(p)[1:7]) for p in goods])
    assert np.array_equal(kept, g['kept'])                               # exact duplicate set
    assert Xk.dtype == np.float32 and np.array_equal(Xk, X[g['kept']])
    assert 'Removed %u near-duplicates' % (X.shape[0] - kept.size) in capsys.readouterr().out
    assert sorted(os.listdir(tmp_path / 'intermediate' / 'dupes3')) == sorted(
        'f%06d.jpg' % i for i in range(X.shape[0]) if i not in set(kept.tolist()))
    got = open(tmp_path / 'intermediate' / 'log_dupes3.csv').read().replace(str(tmp_path), '<OUT>').strip().split('\n')
    want = read(golden_dir, name + '.log_dupes3.csv').strip().split('\n')
    assert got[0] == want[0] and len(got) == len(want)
    same_partner = 0
    for a, b in zip(got[1:], want[1:]):
        a, b = a.split(','), b.split(',')
        assert a[0] == b[0] and a[3] == b[3]
        assert abs(float(a[2]) - float(b[2])) <= 1e-4 * float(b[2]) + 3e-7   # logged distance
        same_partner += a[1] == b[1]
    assert same_partner >= 0.99 * (len(want) - 1)


def test_remove_dupes_overall_hash_golden(tmp_path, golden_dir, capsys):
    """'hash' branch (detection.py:57-59 -> dupes.py:55-58): bit-exact against the real reference, log included."""
    from oracle.make_hashes import make_hashes
    g = gold(golden_dir, 'hash_500')
    H = make_hashes(500, 7)
    names = [osp.basename(p) for p in fake_faces(str(tmp_path), 500)]
    os.makedirs(tmp_path / 'intermediate', exist_ok=True)
    Hk, goods = vdupes.remove_dupes_overall(H, names, ('hash', 8, True, str(tmp_path)))
    kept = np.asarray([int(n[1:7]) for n in goods])
    assert np.array_equal(kept, g['kept']) and np.array_equal(Hk, H[g['kept']])
    assert open(tmp_path / 'intermediate' / 'log_dupes2.csv').read() == read(golden_dir, 'hash_500.log_dupes2.csv')
    assert sorted(os.listdir(tmp_path / 'intermediate' / 'dupes2')) == sorted(
        'f%06d.jpg' % i for i in range(500) if i not in set(kept.tolist()))
    assert 'Removed' not in capsys.readouterr().out                     # dupes.py:90 prints for 'enc' only
    # a size the Python-callable reference would need hours for: compare with the numpy restatement on a slice
    Hb = make_hashes(20000, 11)
    mins, inds, idx = vdupes.get_engine().dedup_hash(Hb, 8)
    rm, ri, _ = RS.hash_dedup_core(Hb[:3000], 8)
    assert np.array_equal(mins[:3000], rm) and np.array_equal(inds[:3000], ri)


@pytest.mark.parametrize('name,thr', [('classify_300x512_r32', 0.55), ('classify_300x512_r32_nothr', -1)])
def test_classify_golden(tmp_path, golden_dir, name, thr):
    g = gold(golden_dir, name)
    B = blobs(332, 512, kc=8, noise=1.0, seed=0)
    X, R = np.ascontiguousarray(B[:300]), np.ascontiguousarray(B[300:])
    paths = fake_faces(str(tmp_path), 300)
    classes = ['c%02d' % i for i in range(32)]
    inds, classes_out = vgrouping.classify(X, R, classes, thr, True, paths, str(tmp_path))
    assert inds.dtype == np.int64 and np.array_equal(inds, g['inds'])     # identical labels
    assert classes_out is classes and (classes[-1] == 'other') == (thr != -1)
    got = open(tmp_path / 'faces' / 'log_classification.csv').read().strip().split('\n')
    want = read(golden_dir, name + '.log_classification.csv').strip().split('\n')
    assert got[0] == want[0] and len(got) == len(want)
    for a, b in zip(got[1:], want[1:]):
        a, b = a.split(','), b.split(',')
        assert a[0] == b[0] and a[-1] == b[-1]
        assert np.abs(np.asarray(a[1:-1], float) - np.asarray(b[1:-1], float)).max() <= 1.0001e-4   # '%.4f' rounding


CLUSTER_CASES = {
    'cluster_60x512': (lambda: blobs(60, 512, kc=3, noise=0.5, seed=0), range(2, 7)),
    'cluster_300x768': (lambda: blobs(300, 768, kc=5, noise=1.0, seed=2, normalize=False), range(2, 6)),
    'cluster_c1_8x512': (lambda: blobs(8, 512, kc=2, noise=0.5, seed=0), range(2, 7)),
    'cluster_2000x512': (lambda: blobs(2000, 512, kc=8, noise=1.0, seed=0), range(2, 7)),
}


@pytest.mark.parametrize('name', sorted(CLUSTER_CASES))
def test_cluster_faces_golden(tmp_path, golden_dir, capsys, name):
    """cluster_faces end to end (k-means++ on the device included): folders, log and chosen k."""
    gen, ks = CLUSTER_CASES[name]
    g = gold(golden_dir, name)
    X = gen()
    paths = fake_faces(str(tmp_path), X.shape[0])
    vgrouping.cluster_faces(list(paths), X, (list(ks), True, 0, True, str(tmp_path)))
    out = capsys.readouterr().out
    want_log = read(golden_dir, name + '.log_clustering.csv').strip().split('\n')
    got_log = open(tmp_path / 'faces' / 'log_clustering.csv').read().strip().split('\n')
    assert got_log[0] == want_log[0] and len(got_log) == len(want_log)
    for a, b, k in zip(got_log[1:], want_log[1:], ks):
        a, b = a.split(','), b.split(',')
        assert a[0] == b[0]
        np.testing.assert_allclose([float(v) for v in a[1:]], [float(v) for v in b[1:]], rtol=1e-4)   # north_star tolerance
        lab = np.full(X.shape[0], -1)
        for j in range(k):
            for fn in os.listdir(tmp_path / 'faces' / ('G%u' % k) / str(j)):
                lab[int(fn[1:7])] = j
        assert (lab == g['labels_k%d' % k]).mean() >= 0.95 and (lab == g['labels_k%d' % k]).mean() == 1.0
    assert not [p for p in paths if osp.exists(p)]                        # originals removed (grouping.py:136-137)
    assert 'Grouped %u images into %s folders:' % (X.shape[0], '/'.join(str(k) for k in ks)) in out


@pytest.mark.parametrize('name,thr', [('classify_faces_300x512', 0.9), ('classify_faces_300x512_nothr', -1)])
def test_classify_faces_golden(tmp_path, golden_dir, capsys, name, thr):
    """classify_faces (grouping.py:69-89) on the CUDA engine against the real reference: console text, folder of every
    file, classification log."""
    import dropin_cases as DC
    DC.check_classify_faces(tmp_path, golden_dir, capsys, name, thr)


def test_classification_mode_golden(tmp_path, golden_dir, capsys):
    """video_to_faces(mode='grouping', group_mode='classification') (main.py:13-82): dedup, classify_faces, 'Done'."""
    import dropin_cases as DC
    DC.check_classification_mode(tmp_path, golden_dir, capsys)


@pytest.mark.parametrize('thr', [-0.5, 0.0, None, -1, 2.5])
def test_classify_threshold_switch(engine, thr):
    """grouping.py:53 `if thr and thr != -1`: a negative threshold other than -1 is ON (every row -> 'other')."""
    from videotofaces_b200.synth import classify_faces_case
    X, R, names = classify_faces_case()
    classes = list(names)
    inds, _ = vgrouping.classify(X, R, classes, thr, False, None, None)
    assert np.array_equal(inds, RC.classify_core(X, R, len(names), thr)[1])
    assert (classes[-1] == 'other') == bool(thr and thr != -1)


def test_zero_row_inputs(engine):
    """Empty shards / empty inputs: every entry returns without touching memory."""
    K = engine.K
    E = torch.zeros((0, 512), dtype=torch.float32, device='cuda')
    R = torch.from_numpy(blobs(5, 512)).cuda()
    assert tuple(K.normalize_rows(E).shape) == (0, 512)
    dist, lab = K.cosine_classify(E, K.normalize_rows(R), 0.5)
    assert tuple(dist.shape) == (0, 5) and lab.numel() == 0
    d2, l2 = engine.classify(E, R, 0.5)
    assert d2.shape == (0, 5) and l2.shape == (0,)
    assert engine.dedup(E, 0.2)[2].size == 0


@pytest.mark.parametrize('N,d', [(500, 100), (5000, 72), (300, 20)])
def test_unaligned_dimension_and_odd_sizes(engine, N, d):
    """Feature dimensions that are not multiples of 16/64 (padded with zero columns) through every entry."""
    X = blobs(N, d, kc=4, noise=1.0, seed=2, normalize=False)
    ctx, res = engine.kmeans_sweep(X, [2, 3], 0)
    scores = engine.cluster_scores(ctx, res)
    ref = RC.cluster_core(X, [2, 3], 0, inits=[X[r.seeds] for r in res])
    for r, lab, cen, sc, rsc in zip(res, ref['labels'], ref['centers'], scores, ref['scores']):
        assert (r.labels.cpu().numpy() == lab).mean() >= 0.999 and tuple(r.centers.shape) == cen.shape
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-4)
    Xd = plant_dupes(X, frac=10, seed=1, normalize=False)[0]
    assert np.array_equal(engine.dedup(Xd, 0.2)[2], RC.dedup_core(Xd, 0.2)[2])
    assert np.array_equal(engine.dedup(Xd, 0.2, force='tc')[2], RC.dedup_core(Xd, 0.2)[2])
    tc = engine.silhouette_scores(ctx, [r.labels for r in res], [2, 3], force='tc')
    np.testing.assert_allclose(tc, [s[1] for s in ref['scores']], rtol=2e-4)
    assert tuple(engine.delete_rows(Xd, np.array([1, 5])).shape) == (N - 2, d)
    assert np.array_equal(engine.classify(X[:100], X[200:205], 0.6)[1], RC.classify_core(X[:100], X[200:205], 5, 0.6)[1])


def test_lockstep_sweep_equals_per_k_loops(engine):
    """The lock-step Lloyd (all k share each pass) against one loop per k: identical labels, n_iter, centres."""
    X = blobs(20000, 512, kc=8, noise=1.0, seed=0)
    ks = [2, 3, 4, 5, 6, 9]
    engine.lockstep = True
    ctx, a = engine.kmeans_sweep(X, ks, 0)
    engine.lockstep = False
    _, b = engine.kmeans_sweep(X, ks, 0, ctx=ctx)
    engine.lockstep = True
    for ra, rb in zip(a, b):
        assert torch.equal(ra.labels, rb.labels) and ra.n_iter == rb.n_iter and ra.strict == rb.strict
        assert torch.equal(ra.centers, rb.centers) and abs(ra.inertia - rb.inertia) <= 1e-9 * rb.inertia
        assert torch.equal(ra.counts, rb.counts)
    sa = engine.ch_db_scores_multi(ctx, [r.labels for r in a], ks, [r.counts for r in a])
    for r, (ch, db) in zip(a, sa):
        rch, rdb = RC.cluster_scores(X[:0], None)[1:] if False else (None, None)
    ref = [(RC.sklearn.metrics.calinski_harabasz_score(X, r.labels.cpu().numpy()),
            RC.sklearn.metrics.davies_bouldin_score(X, r.labels.cpu().numpy())) for r in a]
    np.testing.assert_allclose(np.asarray(sa), np.asarray(ref), rtol=1e-5)


def test_four_rows_error_like_reference(tmp_path):
    X = blobs(4, 512, kc=2, noise=0.5, seed=0)
    paths = fake_faces(str(tmp_path), 4)
    with pytest.raises(ValueError, match='Number of labels is 4'):
        vgrouping.cluster_faces(paths, X, (list(range(2, 7)), False, 0, False, str(tmp_path)))


def test_video_to_faces_grouping_mode(tmp_path, capsys):
    """main.py:64-80 with precomputed embeddings: dedup then clustering picks the reference's k."""
    X = plant_dupes(blobs(200, 512, kc=3, noise=0.5, seed=4), frac=20, seed=1)[0]
    fake_faces(str(tmp_path), 200)
    video_to_faces(out_dir=str(tmp_path), mode='grouping', style='live', clusters='2-5', enc_dup_thr=0.2,
                   group_log=True, features=X)
    out = capsys.readouterr().out
    _, _, ridx = RC.dedup_core(X, 0.2)
    Xk = np.delete(X, ridx, axis=0)
    ref = RC.cluster_core(Xk, [2, 3, 4, 5], 0)
    assert 'Removed %u near-duplicates' % ridx.size in out
    assert 'The number of groups chosen: %u' % ref['best_k'] in out
    sizes = sorted(len(os.listdir(tmp_path / 'faces' / str(j))) for j in range(ref['best_k']))
    assert sizes == sorted(np.bincount(ref['labels'][ref['clusters'].index(ref['best_k'])]).tolist())


@pytest.mark.parametrize('N,d,ks,normalize', [(8192, 512, [2, 3, 4], True), (6000, 768, [2, 5], False), (5000, 1024, [3], False)])
def test_sweep_vs_oracle_medium(engine, N, d, ks, normalize):
    """Tensor-core silhouette + CH + DB + Lloyd on sizes the oracle finishes in seconds."""
    X = blobs(N, d, kc=6, noise=1.0, seed=1, normalize=normalize)
    ctx, res = engine.kmeans_sweep(X, ks, 0)
    scores = engine.cluster_scores(ctx, res)
    inits = [X[r.seeds] for r in res]
    ref = RC.cluster_core(X, ks, 0, inits=inits)                           # same initial centroids
    for r, lab, inertia, n_iter, sc, rsc in zip(res, ref['labels'], ref['inertia'], ref['n_iter'], scores, ref['scores']):
        agree = (r.labels.cpu().numpy() == lab).mean()
        assert agree >= 0.95 and agree >= 0.9999, agree
        assert r.n_iter == n_iter
        assert abs(r.inertia - inertia) <= 1e-4 * inertia
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-4)
    for r, k in zip(res, ks):                                              # the seeding itself: same rows as sklearn's k-means++
        assert np.array_equal(r.seeds, RC.kmeanspp_indices(X, k, 0)), (k, r.seeds)


def test_full_size_properties_200k(engine):
    """Size-independent checks at a size the oracle cannot sweep: (a) for a few 64-row tiles the tensor
    core cell sums reproduce the exact-kernel row sums against all columns; (b) Lloyd inertia equals a
    direct recomputation and every row sits at its nearest centre; (c) planted duplicates are found."""
    N, d, ks = 200_000, 512, [2, 3, 4, 5, 6]
    X = blobs(N, d, kc=8, noise=1.0, seed=0)
    ctx, res = engine.kmeans_sweep(X, ks, 0)
    sil = engine.silhouette_scores(ctx, [r.labels for r in res], ks, [r.counts for r in res])
    K = engine.K
    sqn = K.row_sqnorms(ctx.Xc)
    for r, s in zip(res, sil):
        rowsum = K.zeros((N, r.k), torch.float64)
        rows = [(0, 64), (64 * 1000, 64 * 1001), (N - 64, N)]
        vals = []
        for b, e in rows:
            K.sil_rowsums_exact(ctx.Xc, sqn, r.labels, r.k, b, e, rowsum)
            out = K.zeros((1,), torch.float64)
            K.sil_finish_rows(rowsum, r.labels, r.counts, b, e, out)
            vals.append(out.item() / (e - b))
        assert -1 <= s <= 1
        # (a) the mean silhouette of 192 exactly computed rows is a sample of the full mean
        assert abs(np.mean(vals) - s) < 0.05
        # (b)
        C = r.centers - ctx.mean32
        lab2 = torch.full((N,), -1, dtype=torch.int32, device='cuda')
        K.lloyd_assign(ctx.Xc, C.contiguous(), lab2, K.zeros((r.k,), torch.int64), K.zeros((1,), torch.int64))
        assert (lab2 == r.labels).float().mean().item() >= 0.9999
        direct = ((ctx.Xc.double() - C.double()[r.labels.long()]) ** 2).sum().item()
        assert abs(direct - r.inertia) <= 1e-6 * direct
    inertias = [r.inertia for r in res]
    assert all(a > b for a, b in zip(inertias, inertias[1:]))              # more clusters, less inertia
    Xd, src, dst = plant_dupes(X[:50_000], frac=50, seed=1)
    mins, inds, idx = engine.dedup(Xd, 0.2)
    close = np.einsum('ij,ij->i', Xd[src], Xd[dst]) >= 0.8 + 1e-5         # planted pairs clearly under the threshold
    assert np.isin(dst[close], idx).all()
    sub = np.sort(np.random.default_rng(0).choice(50_000, 200, replace=False))
    D = RS.cosine_distances(Xd[sub], Xd)
    for r_i, i in enumerate(sub):
        if i == 0:
            continue
        assert abs(D[r_i, :i].min() - mins[i]) <= 1e-6                   # exact rows against the oracle


@pytest.mark.parametrize('name,excl', [('testmode_500x512', False), ('testmode_500x512_excl', True)])
def test_test_grouping_golden(tmp_path, golden_dir, capsys, name, excl):
    """test_grouping (grouping.py:140-155) through video_to_faces(_test_enc=True): the printed accuracy / rand score /
    silhouette line and the classification log equal the real reference's."""
    import json
    X, R, gt = eval_case(excl)
    paths = fake_faces(str(tmp_path), X.shape[0])
    with open(tmp_path / 'labels.txt', 'w') as f:
        f.write('\n'.join(str(int(g)) for g in gt) + '\n')
    ref_dir = tmp_path / 'ref'
    for i in range(R.shape[0]):
        os.makedirs(ref_dir / ('c%02d' % i))
        with open(ref_dir / ('c%02d' % i) / 'a.jpg', 'wb') as f:
            f.write(b'x')
    video_to_faces(mode='grouping', style='live', out_dir=str(tmp_path), input_path=str(tmp_path), clusters='2-6',
                   ref_dir=str(ref_dir), enc_oth_thr=0.97, random_state=0, _test_enc=True, _test_exclude_other=excl,
                   features=X, ref_features=R)
    out = capsys.readouterr().out
    with open(osp.join(golden_dir, 'MANIFEST.json')) as f:
        want = json.load(f)['cases'][name]['stdout']
    assert out.endswith(want), out
    assert read(str(tmp_path / 'faces'), 'log_classification.csv') == read(golden_dir, name + '.log_classification.csv')


@pytest.mark.parametrize('N,ka,kb', [(1_000_000, 6, 6), (300_001, 200, 150), (5, 3, 2), (70_000, 1, 7)])
def test_rand_score_equals_sklearn(engine, N, ka, kb):
    """Contingency histogram on the device (shared-memory and global-atomic forms) + exact integer pair counts."""
    import sklearn.metrics
    rng = np.random.default_rng(N)
    a = rng.integers(0, ka, N) * 3 - 7                       # arbitrary integers, like labels.txt
    b = rng.integers(0, kb, N).astype(np.int32)
    b[: N // 3] = (a[: N // 3] // 3 + 2) % kb                # correlated part so the score is not just chance level
    got = engine.rand_score(a, torch.from_numpy(b).cuda())
    assert got == sklearn.metrics.rand_score(a, b)
    assert got == RS.rand_score(a, b)


def test_contingency_counts_out_of_range_labels(engine):
    a = torch.tensor([0, 1, 2, 5, -1, 1], dtype=torch.int32).cuda()
    b = torch.tensor([0, 1, 0, 0, 0, 9], dtype=torch.int32).cuda()
    table, bad = engine.K.contingency(a, b, 3, 2)
    assert int(bad.item()) == 3
    assert table.cpu().tolist() == [[1, 0], [0, 1], [1, 0]]
