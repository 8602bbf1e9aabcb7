"""Host orchestration (engine.py) driven by the CPU stand-in kernels and checked against the
scikit-learn oracle: k-means++ protocol and seed prefix sharing, the Lloyd loop and its stopping
rules, empty-cluster relocation, the cell layout of the one-sweep silhouette, CH/DB, dedup key
decoding, argument parsing.  CPU only -- the CUDA kernels themselves are checked by the -m gpu tests."""
import numpy as np
import pytest
import torch

from fake_kernels import FakeKernels
from oracle import reference_calls as RC
from videotofaces_b200 import prep
from videotofaces_b200.engine import GroupingEngine, n_local_trials, _ceil_to_f32
from videotofaces_b200.dist import shard_bounds, cyclic_tiles
from videotofaces_b200.synth import blobs, plant_dupes


@pytest.fixture()
def eng():
    return GroupingEngine(FakeKernels())


def test_kmeanspp_seeds_match_sklearn(eng):
    X = blobs(3000, 512, kc=8, noise=1.0, seed=0)
    ctx = eng.prepare(X)
    for k in (2, 6):
        assert np.array_equal(eng.kmeans_plusplus(ctx, k, 0), RC.kmeanspp_indices(X, k, 0))


def test_seed_prefix_sharing(eng):
    """Seeds of k are a prefix of k+1's while n_local_trials is equal."""
    assert [n_local_trials(k) for k in (2, 3, 7, 8, 20, 21)] == [2, 3, 3, 4, 4, 5]
    X = blobs(1500, 512, kc=8, noise=1.0, seed=3)
    ctx = eng.prepare(X)
    s7 = eng.kmeans_plusplus(ctx, 7, 5)
    for k in (3, 5):
        assert np.array_equal(RC.kmeanspp_indices(X, k, 5), s7[:k])


@pytest.mark.parametrize('gen,ks', [
    (lambda: blobs(2000, 512, kc=8, noise=1.0, seed=0), [2, 3, 4, 5, 6]),
    (lambda: blobs(300, 768, kc=5, noise=1.0, seed=2, normalize=False), [2, 3, 4, 5]),
])
def test_sweep_matches_reference_calls(eng, gen, ks):
    X = gen()
    ctx, res = eng.kmeans_sweep(X, ks, 0)
    ref = RC.cluster_core(X, ks, 0)
    scores = eng.cluster_scores(ctx, res)
    for r, lab, inertia, n_iter, sc, rsc in zip(res, ref['labels'], ref['inertia'], ref['n_iter'], scores, ref['scores']):
        assert np.array_equal(r.labels.numpy(), lab)
        assert r.n_iter == n_iter
        np.testing.assert_allclose(r.inertia, inertia, rtol=1e-4)
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-4)
    assert max(scores, key=lambda s: s[1])[0] == ref['best_k']


def test_cells_layout_and_tc_path_agree_with_exact(eng):
    """build_cells: every 32-column chunk holds one cell, every row appears once; the cell-sum
    silhouette (fp16 operands) equals the exact one within 1e-4."""
    X = blobs(700, 128, kc=5, noise=1.0, seed=1)
    ks = [2, 3, 5]
    ctx, res = eng.kmeans_sweep(X, ks, 0)
    labels = [r.labels for r in res]
    perm, chunk_cell, c2c, G, P = eng.build_cells(labels, ks)
    pm = perm.numpy()
    assert P % 256 == 0 and sorted(pm[pm >= 0].tolist()) == list(range(700))
    cc = np.repeat(chunk_cell.numpy(), 32)
    for j, lab in enumerate(labels):
        live = pm >= 0
        assert np.array_equal(c2c[j].numpy()[cc[live]], lab.numpy()[pm[live]])
    exact = eng.silhouette_scores(ctx, labels, ks, force='exact')
    tc = eng.silhouette_scores(ctx, labels, ks, force='tc')
    eng.sym = True
    tc_sym = eng.silhouette_scores(ctx, labels, ks, force='tc')
    eng.sym = None
    np.testing.assert_allclose(tc_sym, tc, rtol=1e-6)
    ref = [RC.cluster_scores(X, l.numpy())[0] for l in labels]
    np.testing.assert_allclose(exact, ref, rtol=1e-5)
    np.testing.assert_allclose(tc, ref, rtol=1e-4)


def test_empty_cluster_relocation(eng):
    """SK/cluster/tests/test_k_means.py:85-113: a centre that attracts nothing gets the farthest point."""
    X = np.zeros((10, 4), dtype=np.float32)
    X[:, 0] = [-10, -9.5, -9, -8.5, -8, -1, 1, 9, 9.5, 10]
    init = np.zeros((3, 4), dtype=np.float32)
    init[:, 0] = [-10, -10, -10]
    ctx, res = eng.kmeans_sweep(X, [3], 0, inits=[init])
    ref = RC.kmeans_fit(X, 3, 0, init=init)
    assert np.unique(res[0].labels.numpy()).size == 3
    np.testing.assert_allclose(res[0].inertia, ref.inertia_, rtol=1e-5)
    assert np.array_equal(res[0].labels.numpy(), ref.labels_)


def test_empty_cluster_relocation_wide_path(eng):
    """The same through engine._lloyd_wide: the relocation works on the fixed-point sums converted to fp64 and leaves them
    (and the labels they belong to) untouched, as scikit-learn leaves labels_old untouched."""
    X = np.zeros((10, 64), dtype=np.float32)
    X[:, 0] = [-10, -9.5, -9, -8.5, -8, -1, 1, 9, 9.5, 10]
    init = np.zeros((3, 64), dtype=np.float32)
    init[:, 0] = [-10, -10, -10]
    eng.wide = True
    ctx, res = eng.kmeans_sweep(X, [3], 0, inits=[init])
    eng.wide = None
    ref = RC.kmeans_fit(X, 3, 0, init=init)
    assert np.unique(res[0].labels.numpy()).size == 3
    np.testing.assert_allclose(res[0].inertia, ref.inertia_, rtol=1e-5)
    assert np.array_equal(res[0].labels.numpy(), ref.labels_) and res[0].n_iter == ref.n_iter_


def test_dedup_and_classify_decode(eng):
    X = plant_dupes(blobs(400, 512, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0]
    mins, inds, idx = eng.dedup(X, 0.2)
    rmins, rinds, ridx = RC.dedup_core(X, 0.2)
    assert np.array_equal(idx, ridx) and mins[0] == np.float32(10000.0) and inds[0] == 0
    np.testing.assert_allclose(mins, rmins, rtol=1e-5, atol=2e-7)
    assert np.array_equal(eng.delete_rows(X, idx).numpy(), np.delete(X, idx, axis=0))
    B = blobs(332, 512, kc=8, noise=1.0, seed=0)
    for thr in (0.55, -1, None, 0):
        dist, lab = eng.classify(B[:300], B[300:], thr)
        _, rlab = RC.classify_core(B[:300], B[300:], 32, thr)
        assert np.array_equal(lab, rlab)


def test_hash_dedup_decode(eng):
    from oracle.make_hashes import make_hashes
    H = make_hashes(300, 3)
    mins, inds, idx = eng.dedup_hash(H, 8)
    rmins, rinds, ridx = RC.hash_dedup_core(H, 8)
    assert mins.dtype == np.uint16 and np.array_equal(mins, rmins) and np.array_equal(inds, rinds) and np.array_equal(idx, ridx)


def test_unaligned_feature_dimension(eng):
    """d = 100: the engine pads to 112 zero columns; labels, centres (sliced back) and scores are the oracle's."""
    X = blobs(400, 100, kc=4, noise=1.0, seed=2, normalize=False)
    ctx, res = eng.kmeans_sweep(X, [2, 3], 0)
    ref = RC.cluster_core(X, [2, 3], 0)
    scores = eng.cluster_scores(ctx, res)
    for r, lab, cen, sc, rsc in zip(res, ref['labels'], ref['centers'], scores, ref['scores']):
        assert np.array_equal(r.labels.numpy(), lab) and r.centers.shape == cen.shape
        np.testing.assert_allclose(r.centers.numpy(), cen, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-4)
    Xd = plant_dupes(X, frac=10, seed=1, normalize=False)[0]
    assert np.array_equal(eng.dedup(Xd, 0.2)[2], RC.dedup_core(Xd, 0.2)[2])
    assert eng.delete_rows(Xd, np.array([1, 5])).shape == (398, 100)
    assert np.array_equal(eng.classify(X[:100], X[200:205], 0.6)[1], RC.classify_core(X[:100], X[200:205], 5, 0.6)[1])


def test_four_rows_fail_like_the_reference(eng):
    X = blobs(4, 512, kc=2, noise=0.5, seed=0)
    ctx, res = eng.kmeans_sweep(X, [2, 3, 4], 0)
    with pytest.raises(ValueError, match='Number of labels is 4'):
        eng.cluster_scores(ctx, res)


def test_ceil_to_f32_and_shards():
    v = np.array([0.1, 1.0, 1e-9 + 1.0, 123456.789])
    t = _ceil_to_f32(v)
    assert (t.astype(np.float64) >= v).all() and (np.nextafter(t, np.float32(-np.inf)).astype(np.float64) < v).all()
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(1000, 4, 64) == [0, 256, 512, 768, 1000]
    assert shard_bounds(100, 4, 64) == [0, 64, 100, 100, 100]
    assert cyclic_tiles(7, 1, 3) == [1, 4]


def test_get_clusters_parsing(capsys):
    """prep.py:48-66."""
    assert prep.get_clusters(None) == list(range(2, 9))
    assert prep.get_clusters('2-6') == [2, 3, 4, 5, 6]
    assert prep.get_clusters('7,3,3,5') == [3, 5, 7]
    assert prep.get_clusters(4) == [4]
    assert prep.get_clusters('6-2') is None and prep.get_clusters('abc') is None
    assert 'incorrent value for clusters' in capsys.readouterr().out


@pytest.mark.parametrize('name,excl', [('testmode_500x512', False), ('testmode_500x512_excl', True)])
def test_test_grouping_host_logic(tmp_path, golden_dir, capsys, eng, name, excl):
    """test_grouping through the drop-in module on the stand-in back end: console text and classification log equal
    the real reference's (tests/golden, oracle/make_golden.py:golden_testmode)."""
    import json
    import os
    import os.path as osp
    from videotofaces_b200.synth import eval_case
    from videotofaces_b200 import grouping as vgrouping, runtime
    X, R, gt = eval_case(excl)
    fdir = tmp_path / 'faces'
    os.makedirs(fdir)
    paths = [str(fdir / ('f%06d.jpg' % i)) for i in range(X.shape[0])]
    with open(tmp_path / 'labels.txt', 'w') as f:
        f.write('\n'.join(str(int(g)) for g in gt) + '\n')
    refs = [('c%02d' % i, ['unused.jpg']) for i in range(R.shape[0])]
    runtime.reset_engine(eng)
    try:
        vgrouping.test_grouping(paths, refs, ('live', 'facenet_vgg', 'cpu', str(tmp_path), excl, 16, None, 0.97, 0), X=X, R=R)
    finally:
        runtime.reset_engine(None)
    with open(osp.join(golden_dir, 'MANIFEST.json')) as f:
        assert capsys.readouterr().out == json.load(f)['cases'][name]['stdout']
    with open(fdir / 'log_classification.csv') as f, open(osp.join(golden_dir, name + '.log_classification.csv')) as g:
        assert f.read() == g.read()


def test_rand_score_host_arithmetic(eng):
    import sklearn.metrics
    rng = np.random.default_rng(3)
    for n, ka, kb in [(1000, 5, 5), (17, 3, 9), (4000, 40, 2)]:
        a, b = rng.integers(0, ka, n) * 2 + 1, rng.integers(0, kb, n).astype(np.int32)
        assert eng.rand_score(a, torch.from_numpy(b)) == sklearn.metrics.rand_score(a, b)
    assert eng.rand_score(np.arange(6), torch.arange(6, dtype=torch.int32)) == 1.0      # every row its own cluster
    assert eng.rand_score(np.zeros(6, int), torch.zeros(6, dtype=torch.int32)) == 1.0   # nothing split


@pytest.mark.parametrize('N,d,kc,ks', [(5000, 64, 4, [2, 3, 4]), (700, 128, 5, [2, 5, 9])])
def test_wide_lloyd_host_loop(eng, N, d, kc, ks):
    """engine._lloyd_wide (tensor-core E-step, fixed-point M-step from change lists) against the lock-step loop and the
    scikit-learn oracle; 5000 rows put the first iteration over the change-list cap (sums from scratch) and the later
    ones under it (delta), 700 rows run on change lists only."""
    X = blobs(N, d, kc=kc, noise=1.0, seed=N % 7, normalize=False)
    ctx, base = eng.kmeans_sweep(X, ks, 0)
    eng.wide, eng.profile = True, True
    _, wide = eng.kmeans_sweep(X, ks, 0, ctx=ctx)
    stats = eng.wide_stats
    eng.wide, eng.profile = None, False
    ref = RC.cluster_core(X, ks, 0, inits=[X[r.seeds] for r in base], with_scores=False)
    for r, w, lab, it in zip(base, wide, ref['labels'], ref['n_iter']):
        assert torch.equal(r.labels, w.labels) and r.n_iter == w.n_iter == it and r.strict == w.strict
        assert np.array_equal(w.labels.numpy(), lab)
        assert torch.equal(r.counts, w.counts)
        np.testing.assert_allclose(w.centers.numpy(), r.centers.numpy(), rtol=0, atol=1e-6)
        want = torch.zeros_like(w.csums).index_add_(0, w.labels.to(torch.int64), ctx.Xc.to(torch.float64))
        np.testing.assert_allclose(w.csums.numpy(), want.numpy(), rtol=0, atol=1e-7 * float(want.abs().max()))
    assert stats['iters'] == max(w.n_iter for w in wide) and stats['cap'] == max(1024, N // 4)
    assert stats['listed'][0] == (0 if N > 1024 else N * len(ks))      # first iteration: every row "changes" from -1
    if max(w.n_iter for w in wide) > 2:
        assert any(v > 0 for v in stats['listed'][1:])                 # later iterations run from the change lists


def test_wide_column_table_layout():
    """engine._wide_columns: every centre of every labelling of the batch appears exactly once, in order, with its local
    index and labelling in the info word, the last centre of each labelling flagged, padding up to whole 256-column tiles."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        nL = int(rng.integers(1, 32))
        ks = [int(k) for k in rng.integers(2, 101, nL)]
        off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
        batch = sorted(rng.choice(nL, size=int(rng.integers(1, nL + 1)), replace=False).tolist())
        src, info, col0 = GroupingEngine._wide_columns(off, ks, batch)
        n_cols = sum(ks[L] for L in batch)
        assert src.size % 256 == 0 and src.size - n_cols < 256 and (src[n_cols:] == -1).all()
        c = 0
        for L in range(nL):
            if L not in batch:
                assert col0[L] == -1
                continue
            assert col0[L] == c
            assert np.array_equal(src[c:c + ks[L]], np.arange(off[L], off[L + 1]))
            assert np.array_equal((info[c:c + ks[L]] >> 4) & 0xffff, np.arange(ks[L]))
            assert ((info[c:c + ks[L]] >> 20) == L).all()
            assert np.array_equal((info[c:c + ks[L]] >> 1) & 1, np.arange(ks[L]) == ks[L] - 1)
            c += ks[L]


def test_pack_groups_and_fixed_point_unit(eng):
    for ks, d in [([2, 3, 4, 5, 6], 512), (list(range(2, 33)), 1024), ([100, 100, 3], 0), ([7] * 70, 64)]:
        groups = eng._pack_groups(ks, d)
        assert [j for g in groups for j in g] == list(range(len(ks)))          # contiguous, complete, in order
        for g in groups:
            tot = sum(ks[j] for j in g)
            assert len(g) <= 32 and (len(g) == 1 or (tot <= 96 and tot * d * 4 <= 160 * 1024))
    # the fixed-point unit leaves head room: N * max|x| * 2^S < 2^62, and resolves far below fp32
    X = blobs(3000, 64, kc=3, noise=1.0, seed=1, normalize=False) * np.float32(37.5)
    ctx = eng.prepare(X)
    wx = eng._wide_operand(ctx)
    amax = float(np.abs(ctx.Xc.numpy()).max())
    assert ctx.N * amax * wx['two_s'] < 2.0 ** 62 and wx['two_s'] * amax > 2.0 ** 36
    assert wx['scale'] * amax <= 2048.0 and wx['scale'] * amax > 1024.0


def test_wide_lloyd_more_than_32_labellings(eng):
    """A k-range with more than 32 fits runs as several wide groups (the kernels take 32 labellings per call); every fit
    must come back complete and self-consistent: labels in range, counts = bincount, sums = the labels' sums, centres =
    means (strict convergence), and a final E-step with the returned centres must reproduce the labels."""
    from oracle import restatement as RS
    ks = list(range(2, 38))                                   # 36 labellings -> groups of 32 + 4
    X = blobs(600, 64, kc=6, noise=0.6, seed=11, normalize=False)
    eng.wide = True
    ctx, res = eng.kmeans_sweep(X, ks, 0)
    eng.wide = None
    assert [r.k for r in res] == ks
    Xc = ctx.Xc.numpy()
    for r in res:
        lab = r.labels.numpy()
        assert lab.min() >= 0 and lab.max() < r.k and r.n_iter >= 1
        assert np.array_equal(r.counts.numpy(), np.bincount(lab, minlength=r.k))
        want = np.zeros((r.k, 64))
        np.add.at(want, lab, Xc.astype(np.float64))
        np.testing.assert_allclose(r.csums.numpy(), want, rtol=0, atol=1e-7 * np.abs(want).max())
        if r.strict:
            C = (r.centers.numpy() - ctx.mean32.numpy()[None, :])
            np.testing.assert_allclose(C[r.counts.numpy() > 0], (want / np.maximum(r.counts.numpy(), 1)[:, None])[r.counts.numpy() > 0],
                                       rtol=0, atol=2e-6)
            again = RS.lloyd_assign(Xc, C.astype(np.float32))
            assert (again == lab).mean() >= 0.995                  # (rows within an ulp of a bisector may flip)
    inert = [r.inertia for r in res]
    assert inert[0] > inert[5] > inert[-1]


def test_strict_fits_leave_no_pass_over_X_for_ch_and_db(eng):
    """A fit that stops on 'labels unchanged' has its per-cluster sums (last M-step) and its distances to the centres
    (inertia pass) already: CH / DB of such fits take no further pass over X and equal the two-pass form."""
    X = blobs(3000, 512, kc=6, noise=0.7, seed=4)
    ks = [2, 3, 4, 5, 6]
    ctx, res = eng.kmeans_sweep(X, ks, 0)
    assert all(r.strict for r in res)
    assert all(r.csums is not None and r.disp is not None for r in res)
    calls = []
    K = eng.K
    orig_disp, orig_sums = K.cluster_dispersion_multi, K.cluster_sums_multi
    K.cluster_dispersion_multi = lambda *a, **k: (calls.append('disp'), orig_disp(*a, **k))[1]
    K.cluster_sums_multi = lambda *a, **k: (calls.append('sums'), orig_sums(*a, **k))[1]
    try:
        labels, counts = [r.labels for r in res], [r.counts for r in res]
        fast = eng.ch_db_scores_multi(ctx, labels, ks, counts, [r.csums for r in res], [r.disp for r in res])
        assert calls == []
        slow = eng.ch_db_scores_multi(ctx, labels, ks, counts)
        assert calls == ['sums', 'disp']
        one_pass = eng.ch_db_scores_multi(ctx, labels, ks, counts, [r.csums for r in res])      # sums reused, distances recomputed
        assert calls == ['sums', 'disp', 'disp']
    finally:
        K.cluster_dispersion_multi, K.cluster_sums_multi = orig_disp, orig_sums
    np.testing.assert_allclose(np.asarray(fast), np.asarray(one_pass), rtol=1e-11)   # fp32(mean) vs mean: second order
    np.testing.assert_allclose(np.asarray(fast), np.asarray(slow), rtol=1e-7)        # centred fp32 copy vs the raw rows
    ref = [RC.cluster_scores(X, r.labels.numpy())[1:] for r in res]
    np.testing.assert_allclose(np.asarray(fast), np.asarray(ref), rtol=1e-5)


def test_cell_sums_are_the_exact_fixed_point_sums(eng):
    """The M-step of the lock-step loop runs on exact 64-bit fixed-point sums: centres equal the fp64-sum path bit for bit
    on this input, and a labelling fitted alone equals the same labelling fitted next to the others."""
    X = blobs(1500, 128, kc=5, noise=1.0, seed=2)
    ks = [2, 3, 5]
    ctx, a = eng.kmeans_sweep(X, ks, 0)
    eng.lockstep = False
    _, bb = eng.kmeans_sweep(X, ks, 0, ctx=ctx)
    eng.lockstep = True
    for ra, rb in zip(a, bb):
        assert torch.equal(ra.labels, rb.labels) and torch.equal(ra.centers, rb.centers) and ra.n_iter == rb.n_iter
        assert ra.inertia == rb.inertia


def test_tolerance_stopped_fits_also_need_no_extra_pass(eng):
    """A fit that stops on the centre-shift tolerance (SK/cluster/_kmeans.py:731-738) labels the rows once more and its
    centres are NOT the means of those labels: its final sums come from one more cell-sums pass and it rides through the
    inertia pass twice (against its centres, against the means).  CH / DB then take no pass of their own either."""
    X = blobs(8000, 128, kc=8, noise=1.5, seed=3)
    ks = [2, 3, 4, 5, 6]
    ctx, res = eng.kmeans_sweep(X, ks, 0)
    assert [r.strict for r in res] == [True, False, True, True, True]
    assert all(r.csums is not None and r.disp is not None for r in res)
    calls = []
    K = eng.K
    orig_disp, orig_sums = K.cluster_dispersion_multi, K.cluster_sums_multi
    K.cluster_dispersion_multi = lambda *a, **k: (calls.append('disp'), orig_disp(*a, **k))[1]
    K.cluster_sums_multi = lambda *a, **k: (calls.append('sums'), orig_sums(*a, **k))[1]
    try:
        scores = eng.cluster_scores(ctx, res, force='exact')
    finally:
        K.cluster_dispersion_multi, K.cluster_sums_multi = orig_disp, orig_sums
    assert calls == []
    ref = RC.cluster_core(X, ks, 0)
    for r, lab, inertia, sc, rsc in zip(res, ref['labels'], ref['inertia'], scores, ref['scores']):
        assert np.array_equal(r.labels.numpy(), lab)
        np.testing.assert_allclose(r.inertia, inertia, rtol=1e-5)
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-5)


def test_encoder_hand_off_normalises_in_the_first_pass(eng):
    """SURVEY.md 8f item 2, second half: raw encoder outputs + normalize=True give what the reference gets from the
    normalised embeddings FaceNet emits (encoders/facenet.py:153)."""
    from oracle import restatement as RS
    raw = blobs(1200, 512, kc=4, noise=1.0, seed=5, normalize=False) * np.float32(3.7)
    Xn = RS.normalize_l2(raw).astype(np.float32)
    ctx_a, a = eng.kmeans_sweep(raw, [2, 3, 4], 0, normalize=True)
    ctx_b, b = eng.kmeans_sweep(Xn, [2, 3, 4], 0)
    for ra, rb in zip(a, b):
        assert torch.equal(ra.labels, rb.labels) and ra.n_iter == rb.n_iter
        np.testing.assert_allclose(ra.inertia, rb.inertia, rtol=1e-6)
    ref = RC.cluster_core(Xn, [2, 3, 4], 0)
    assert all(np.array_equal(r.labels.numpy(), l) for r, l in zip(a, ref['labels']))
