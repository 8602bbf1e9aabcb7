"""-m gpu: the wide (tensor-core) Lloyd path of csrc/lloyd_tc.cu against the fp32 kernels and the oracle.

Bars: labels identical to the fp32 E-step (the tensor-core scores only pre-select; every near tie is re-decided with
the fp32 kernel's own arithmetic), fixed-point sums bit-identical between the incremental and the from-scratch form,
whole fits identical to the lock-step fp32 path and to scikit-learn given the same initial centres."""
import numpy as np
import pytest
import torch

from oracle import reference_calls as RC
from videotofaces_b200.synth import blobs

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if dtype is None else t.to(dtype).cuda()


@pytest.fixture(scope='module')
def K(engine):
    return engine.K


@pytest.fixture()
def wide_engine(engine):
    old = engine.wide, engine.profile
    engine.wide, engine.profile = True, True
    yield engine
    engine.wide, engine.profile = old


def _fp32_labels(K, Xc, C_all, off, nL):
    """vtf_lloyd_assign_multi with k > 32 per labelling: the warp-per-row fp32 kernel (rowdot.cu)."""
    N = Xc.shape[0]
    labels = torch.full((nL, N), -1, dtype=torch.int32, device='cuda')
    counts = torch.zeros(int(off[-1]), dtype=torch.int64, device='cuda')
    changed = torch.zeros(nL, dtype=torch.int64, device='cuda')
    K.lloyd_assign_multi(Xc, C_all, off, (1 << nL) - 1, labels, 0, counts, changed)
    return labels, counts


@pytest.mark.parametrize('N,d,ks', [(5000, 512, [40, 33, 64]), (3001, 1024, [34, 100]), (2600, 192, [48, 40, 36, 35]),
                                    (700, 64, [33]), (4000, 256, [100, 100, 90]), (2304, 128, [100] * 21)])
def test_wide_estep_equals_fp32_kernel_on_near_ties(engine, K, N, d, ks):
    rng = np.random.default_rng(N + d)
    nL, ktot = len(ks), sum(ks)
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    C = rng.standard_normal((ktot, d), dtype=np.float32)
    C[off[1] - 1] = C[off[0]]                      # exact duplicate centres inside one labelling: first index must win
    X = rng.standard_normal((N, d), dtype=np.float32)
    # adversarial rows: on (or a few ulp off) the bisector of two centres of a labelling
    for i in range(0, N, 3):
        L = i % nL
        a, b = rng.choice(ks[L], 2, replace=False) + off[L]
        X[i] = 0.5 * (C[a] + C[b]) + (rng.standard_normal(d, dtype=np.float32) * np.float32(1e-7 if i % 2 else 0))
    Xc, Cd = dev(X), dev(C)

    class Ctx:
        pass
    ctx = Ctx()
    ctx.Xc, ctx.N, ctx.d, ctx.b, ctx.e, ctx.wide = Xc, N, d, 0, N, None
    labels_new = torch.zeros((nL, N), dtype=torch.int32, device='cuda')
    n_fix = torch.zeros(1, dtype=torch.int32, device='cuda')
    engine._wide_estep(ctx, Cd, off, ks, (1 << nL) - 1, labels_new, n_fix, {})
    n_flag = int(n_fix.item())
    want, _ = _fp32_labels(K, Xc, Cd, off, nL)
    assert n_flag >= N // 3 - 5                                         # the planted ties were all caught
    assert n_flag < N * nL // 2                                         # and ordinary rows were not
    assert int((labels_new < 0).sum().item()) == 0                      # every flag was cleared
    assert torch.equal(labels_new, want), 'mismatches: %d' % int((labels_new != want).sum().item())


def test_wide_estep_inactive_labellings_untouched(engine, K):
    N, d, ks = 4100, 256, [33, 34, 35]
    rng = np.random.default_rng(3)
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    Xc, Cd = dev(rng.standard_normal((N, d), dtype=np.float32)), dev(rng.standard_normal((sum(ks), d), dtype=np.float32))

    class Ctx:
        pass
    ctx = Ctx()
    ctx.Xc, ctx.N, ctx.d, ctx.b, ctx.e, ctx.wide = Xc, N, d, 0, N, None
    labels_new = torch.full((3, N), 77, dtype=torch.int32, device='cuda')
    n_fix = torch.zeros(1, dtype=torch.int32, device='cuda')
    engine._wide_estep(ctx, Cd, off, ks, 0b101, labels_new, n_fix, {})
    want, _ = _fp32_labels(K, Xc, Cd, off, 3)
    assert torch.equal(labels_new[0], want[0]) and torch.equal(labels_new[2], want[2])
    assert bool((labels_new[1] == 77).all())


@pytest.mark.parametrize('N,d,ks', [(20000, 512, [3, 40, 57]), (3000, 260, [100, 2]), (900, 64, [5, 6])])
def test_fixed_point_sums_delta_equals_full(K, N, d, ks):
    rng = np.random.default_rng(N)
    nL, ktot = len(ks), sum(ks)
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    X = (rng.standard_normal((N, d), dtype=np.float32) * 3).astype(np.float32)
    X[::7] *= np.float32(1e-4)
    Xd = dev(X)
    two_s = 2.0 ** 36
    old = np.stack([rng.integers(0, k, N) for k in ks]).astype(np.int32)
    new = old.copy()
    for L, k in enumerate(ks):
        moved = rng.choice(N, N // 40, replace=False)
        new[L, moved] = rng.integers(0, k, moved.size)
    full = (1 << nL) - 1
    ones = torch.ones(nL, dtype=torch.int64, device='cuda')

    def from_scratch(lab):
        acc = torch.zeros((ktot, d), dtype=torch.int64, device='cuda')
        for L0, L1 in [(0, 1), (1, nL)]:
            K.wide_sums_full(Xd, dev(lab), 0, off, full, L0, L1, ones, 0, two_s, acc)
        return acc
    acc_old, acc_new = from_scratch(old), from_scratch(new)
    # the fixed-point sums are the exact integer sums
    q = np.rint(X.astype(np.float64) * two_s).astype(np.int64)
    want = np.zeros((ktot, d), dtype=np.int64)
    for L in range(nL):
        np.add.at(want, off[L] + new[L], q)
    assert np.array_equal(acc_new.cpu().numpy(), want)
    # count + commit + delta from the old state
    lab_all, lab_new = dev(old), dev(new)
    counts = torch.zeros(ktot, dtype=torch.int64, device='cuda')
    changed = torch.zeros(nL, dtype=torch.int64, device='cuda')
    K.wide_count(lab_new, lab_all, 0, N, off, full, counts, changed)
    assert np.array_equal(changed.cpu().numpy(), (old != new).sum(1))
    assert np.array_equal(counts.cpu().numpy(), np.concatenate([np.bincount(new[L], minlength=k) for L, k in enumerate(ks)]))
    cap = N // 20
    items = torch.empty((nL * cap, 4), dtype=torch.int32, device='cuda')
    n_items = torch.zeros(1, dtype=torch.int32, device='cuda')
    K.wide_commit(lab_new, lab_all, 0, N, off, full, changed, cap, items, n_items)
    assert int(n_items.item()) == int((old != new).sum())
    assert torch.equal(lab_all, lab_new)
    K.wide_sums_delta(Xd, items, n_items, off, two_s, acc_old)
    assert torch.equal(acc_old, acc_new)                              # bit-identical
    # a labelling over the cap is left to the full pass: nothing listed for it
    lab_all = dev(old)
    K.wide_commit(lab_new, lab_all, 0, N, off, full, changed, 1, items, n_items)
    assert int(n_items.item()) == 0


@pytest.mark.parametrize('N,d,kc,ks,norm', [(20000, 512, 8, list(range(2, 13)), True), (9000, 768, 5, [2, 7, 16, 31], False),
                                            (4099, 1024, 6, [33, 9], False)])
def test_wide_fit_equals_lockstep_fp32_fit(wide_engine, N, d, kc, ks, norm):
    """Same seeds, same stopping rules.  For k up to the number of planted blobs the fits are stable and must agree
    exactly.  Over-clustered fits (k > blobs) split Gaussian blobs along arbitrary directions: a single row that sits
    within an ulp of a bisector (the two fp32 kernels sum in different orders) sends them down different paths, so
    there the first iterations must agree exactly and the final inertia closely."""
    eng = wide_engine
    X = blobs(N, d, kc=kc, noise=1.0, seed=N % 97, normalize=norm)
    Xd = eng.to_device(X)
    ctx, res_w = eng.kmeans_sweep(Xd, ks, 0)
    stats = dict(eng.wide_stats)
    assert stats['iters'] == max(r.n_iter for r in res_w)
    # CH / DB from the fixed-point Lloyd sums (centred rows) against the two-pass form on X: translation invariant
    labs, cnts = [r.labels for r in res_w], [r.counts for r in res_w]
    a = eng.ch_db_scores_multi(ctx, labs, ks, cnts, [r.csums for r in res_w])
    b = eng.ch_db_scores_multi(ctx, labs, ks, cnts)
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=1e-6)           # fp32 rounding of x - mean: ~1e-8 measured
    for r in res_w:                                                                  # the sums follow the final labels exactly
        want = torch.zeros_like(r.csums).index_add_(0, r.labels.to(torch.int64), ctx.Xc.to(torch.float64))
        np.testing.assert_allclose(r.csums.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=1e-6 * float(want.abs().max()))
    C0s = [eng.K.gather_rows(ctx.Xc, torch.from_numpy(r.seeds).cuda()) for r in res_w]
    two_w = eng.lloyd_multi(ctx, C0s, max_iter=2)
    eng.wide = False
    _, res_s = eng.kmeans_sweep(Xd, ks, 0)
    two_s = eng.lloyd_multi(ctx, C0s, max_iter=2)
    for rw, rs in zip(two_w, two_s):
        assert torch.equal(rw.labels, rs.labels) and torch.equal(rw.counts, rs.counts)
        np.testing.assert_allclose(rw.centers.cpu().numpy(), rs.centers.cpu().numpy(), rtol=0, atol=1e-6)
        assert abs(rw.inertia - rs.inertia) <= 1e-6 * abs(rs.inertia)
    for rw, rs in zip(res_w, res_s):
        assert np.array_equal(rw.seeds, rs.seeds)
        if rw.k <= kc:
            assert rw.n_iter == rs.n_iter and rw.strict == rs.strict
            assert torch.equal(rw.labels, rs.labels)
            assert torch.equal(rw.counts, rs.counts)
            np.testing.assert_allclose(rw.centers.cpu().numpy(), rs.centers.cpu().numpy(), rtol=0, atol=1e-6)
            assert abs(rw.inertia - rs.inertia) <= 1e-6 * abs(rs.inertia)
        else:
            assert abs(rw.inertia - rs.inertia) <= 2e-3 * abs(rs.inertia)
        assert int(rw.counts.sum().item()) == N
    # after the first iterations the M-step runs from the change lists
    assert any(v > 0 for v in stats['listed']) or stats['iters'] <= 2


def test_wide_fit_matches_sklearn_given_init(wide_engine):
    eng = wide_engine
    N, d, ks = 12000, 512, [4, 9, 24]
    X = blobs(N, d, kc=8, noise=1.0, seed=5, normalize=True)
    rng = np.random.default_rng(0)
    inits = [X[rng.choice(N, k, replace=False)].copy() for k in ks]
    ctx, res = eng.kmeans_sweep(eng.to_device(X), ks, 0, inits=inits)
    for r, k, init in zip(res, ks, inits):
        ref = RC.kmeans_fit(X, k, 0, init=init)
        agree = float((r.labels.cpu().numpy() == ref.labels_).mean())
        assert agree >= 0.9999, agree                                 # identical labels given the same initial centres
        assert abs(r.inertia - ref.inertia_) <= 1e-4 * ref.inertia_           # north_star tolerance
        assert r.n_iter == ref.n_iter_


def test_wide_fit_empty_cluster_relocation(wide_engine):
    """SK/cluster/tests/test_k_means.py:85-113 through the wide path on the device: a centre that attracts nothing gets the
    farthest point (_k_means_common.pyx:167-211); the fixed-point sums and the labels they belong to stay untouched."""
    eng = wide_engine
    X = np.zeros((10, 64), dtype=np.float32)
    X[:, 0] = [-10, -9.5, -9, -8.5, -8, -1, 1, 9, 9.5, 10]
    init = np.zeros((3, 64), dtype=np.float32)
    init[:, 0] = [-10, -10, -10]
    ctx, res = eng.kmeans_sweep(X, [3], 0, inits=[init])
    ref = RC.kmeans_fit(X, 3, 0, init=init)
    assert np.array_equal(res[0].labels.cpu().numpy(), ref.labels_) and res[0].n_iter == ref.n_iter_
    assert abs(res[0].inertia - ref.inertia_) <= 1e-5 * ref.inertia_
