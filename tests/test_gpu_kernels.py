"""-m gpu: every CUDA kernel, called through the C ABI (ctypes -> libvtf_b200.so), against the CPU
oracle on the same seeded inputs.  Bars: bit-exact / identical for index and label outputs (ties
reported), 1e-4 relative (BASELINE.json north_star) or tighter for floating point -- the tolerance is
written at each assert."""
import math

import numpy as np
import pytest
import torch

from oracle import reference_calls as RC
from oracle import restatement as RS
from videotofaces_b200.synth import blobs, plant_dupes

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if dtype is None else t.to(dtype).cuda()


@pytest.fixture(scope='module')
def K(engine):
    return engine.K


@pytest.mark.parametrize('N,d', [(1, 512), (37, 64), (1000, 512), (513, 768), (260, 1024), (100, 20)])
def test_normalize_and_norms(K, N, d):
    rng = np.random.default_rng(N + d)
    X = rng.standard_normal((N, d), dtype=np.float32) * 3
    if N > 2:
        X[1] = 0                                             # zero-norm guard (SK/preprocessing/_data.py:127-132)
    got = K.normalize_rows(dev(X)).cpu().numpy()
    np.testing.assert_allclose(got, RS.normalize_l2(X), rtol=1e-6, atol=1e-9)   # norm summed in another order: a few ulp
    np.testing.assert_allclose(K.row_sqnorms(dev(X)).cpu().numpy(), RS.row_norms_sq(X), rtol=2e-6)


@pytest.mark.parametrize('N,d', [(5, 512), (4097, 512), (3000, 768), (2049, 1024), (100, 36)])
def test_col_stats_center_gather(K, N, d):
    X = blobs(N, d, kc=3, seed=N, normalize=False) + 0.5
    s, q = K.col_stats(dev(X))
    X64 = X.astype(np.float64)
    np.testing.assert_allclose(s.cpu().numpy(), X64.sum(0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(q.cpu().numpy(), (X64 * X64).sum(0), rtol=1e-12)
    mean32 = (X64.sum(0) / N).astype(np.float32)
    assert np.array_equal(K.center_rows(dev(X), dev(mean32)).cpu().numpy(), X - mean32)      # bit-exact
    idx = np.random.default_rng(0).permutation(N)[:max(1, N // 3)].astype(np.int64)
    assert np.array_equal(K.gather_rows(dev(X), dev(idx)).cpu().numpy(), X[idx])


@pytest.mark.parametrize('N,d,T', [(3000, 512, 3), (1025, 768, 2), (700, 1024, 5), (50, 64, 1)])
def test_kpp_dist_matches_upcast_path(K, N, d, T):
    X = blobs(N, d, kc=4, seed=7, normalize=(d == 512))
    Xc = (X - X.mean(0)).astype(np.float32)
    cand = Xc[np.random.default_rng(1).integers(0, N, T)]
    want = RS.euclidean_sq_upcast(cand, Xc)
    out, pots = K.kpp_dist(dev(Xc), dev(cand), None)
    got = out.cpu().numpy()
    # fp64 arithmetic rounded once to fp32: identical except where the fp64 sums straddle a rounding boundary
    assert (got == want).mean() > 0.999
    # a candidate's distance to itself cancels to ~1e-12 * |x|^2 in fp64 either way: absolute floor
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=1e-9 * float(want.max()))
    np.testing.assert_allclose(pots.cpu().numpy(), want.sum(1, dtype=np.float64), rtol=1e-7)
    closest = want[0] * 0.5
    out2, pots2 = K.kpp_dist(dev(Xc), dev(cand), dev(closest))
    np.testing.assert_allclose(out2.cpu().numpy(), np.minimum(want, closest[None, :]), rtol=2e-7, atol=1e-9 * float(want.max()))


@pytest.mark.parametrize('N', [1, 100, 4096, 4097, 50000])
def test_kpp_pick_is_the_sequential_fp32_cumsum(K, N):
    rng = np.random.default_rng(N)
    closest = (rng.random(N, dtype=np.float32) * 2).astype(np.float32)
    cs = np.cumsum(closest)                                   # numpy accumulates fp32 sequentially
    thr = np.sort((rng.random(4) * float(cs[-1])).astype(np.float32))
    thr[-1] = np.float32(cs[-1] * 2)                          # never reached -> -1 (clipped to N-1 by the host)
    want = np.searchsorted(cs, thr, side='left').astype(np.int64)
    want[want >= N] = -1
    got = K.kpp_pick(dev(closest), dev(thr)).cpu().numpy()
    assert np.array_equal(got, want)                          # bit-exact: same additions in the same order


def _adversarial_sequences():
    rng = np.random.default_rng(0)
    out = {}
    out['uniform_1m'] = (rng.random(1_000_003, dtype=np.float32) * 2).astype(np.float32)
    # exact ties at many scales: multiples of 2^-k collide with the half-ulp of the running sum again and again
    out['ties'] = (rng.integers(0, 64, 300_000) * np.float32(2.0 ** -7)).astype(np.float32)
    out['ties_small'] = (rng.integers(0, 3, 200_000) * np.float32(0.5)).astype(np.float32)
    z = (rng.random(100_000, dtype=np.float32)).astype(np.float32)
    z[rng.random(100_000) < 0.7] = 0.0
    z[:5000] = 0.0                                           # leading zeros: the sum stays 0
    out['zeros'] = z
    big = rng.random(50_000, dtype=np.float32).astype(np.float32)
    big[::997] *= np.float32(1e6)                            # addends far larger than the running sum
    out['jumps'] = big
    out['wide_range'] = np.exp(rng.uniform(-60, 20, 120_000)).astype(np.float32)
    out['tiny'] = (rng.random(30_000, dtype=np.float32) * np.float32(1e-38)).astype(np.float32)
    out['kmeans_like'] = (rng.standard_normal(400_000).astype(np.float32) ** 2 * np.float32(0.8) + np.float32(0.3)).astype(np.float32)
    return out


@pytest.mark.parametrize('name', ['uniform_1m', 'ties', 'ties_small', 'zeros', 'jumps', 'wide_range', 'tiny', 'kmeans_like'])
def test_kpp_pick_parallel_scan_is_bit_exact(K, name):
    """The block-scan pick must reproduce numpy's sequential fp32 cumsum bit for bit -- final sum and every
    searchsorted position -- on sequences built to hit ties, binade changes, zeros and oversized addends."""
    v = _adversarial_sequences()[name]
    cs = np.cumsum(v)                                        # sequential fp32
    rng = np.random.default_rng(1)
    thr = np.sort((rng.random(7) * float(cs[-1])).astype(np.float32))
    thr[0] = cs[len(cs) // 3]                                # exactly equal to a prefix value: side='left'
    thr = np.sort(thr)
    want = np.searchsorted(cs, thr, side='left').astype(np.int64)
    want[want >= v.size] = -1
    for mode in ('serial', 'scan', 'fast'):                  # one-thread chain, block scan over every tile, tile skipping
        idx, total = K.kpp_pick(dev(v), dev(thr), mode=mode, want_total=True)
        assert np.array_equal(idx.cpu().numpy(), want), (mode, name)
        assert total.cpu().numpy()[0].tobytes() == cs[-1].tobytes(), (mode, name)
    # a shard of a longer chain: the running sum enters as an offset (and with it the binade the tiles were prepared for)
    h = v.size // 2
    cs2 = np.cumsum(np.concatenate([[cs[h - 1]], v[h:]]).astype(np.float32))[1:]
    want2 = np.searchsorted(cs2, thr, side='left').astype(np.int64)
    want2[thr <= cs[h - 1]] = -1                             # met by the earlier shard
    want2[want2 >= v.size - h] = -1
    for mode in ('scan', 'fast'):
        idx, total = K.kpp_pick(dev(v[h:]), dev(thr), mode=mode, want_total=True, offset=dev(cs[h - 1:h]))
        assert np.array_equal(idx.cpu().numpy(), want2), (mode, name, 'offset')
        assert total.cpu().numpy()[0].tobytes() == cs[-1].tobytes(), (mode, name, 'offset')


def test_kpp_pick_fast_on_many_seeds(K):
    """Tile skipping against the sequential chain on k-means++-like distance vectors of 1-2 M elements, several seeds and
    threshold sets (incl. thresholds just above / below tile boundaries of the running sum)."""
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        N = 1_000_000 + 8192 * seed + seed
        v = (rng.standard_normal(N).astype(np.float32) ** 2 * np.float32(0.01 + 0.3 * seed) + np.float32(0.5)).astype(np.float32)
        if seed % 2:
            v[rng.random(N) < 0.3] = 0.0
        cs = np.cumsum(v)
        thr = np.sort(np.concatenate([(rng.random(3) * float(cs[-1])).astype(np.float32),
                                      cs[[8191, 8192, 16 * 8192 - 1]], np.nextafter(cs[[40 * 8192]], np.float32(np.inf))]).astype(np.float32))
        want = np.searchsorted(cs, thr, side='left').astype(np.int64)
        want[want >= N] = -1
        idx, total = K.kpp_pick(dev(v), dev(thr), mode='fast', want_total=True)
        assert np.array_equal(idx.cpu().numpy(), want), seed
        assert total.cpu().numpy()[0].tobytes() == cs[-1].tobytes(), seed


@pytest.mark.parametrize('T', [1, 2, 3, 8])
def test_kpp_round_glue_equals_the_host_arithmetic(K, T):
    """argmin of the fp32-rounded potentials, thresholds = smallest fp32 >= u * pot (stable ascending order), and the
    unsort + clip that follows the pick -- the device glue must equal numpy on the same numbers (SK _kmeans.py:249-266)."""
    rng = np.random.default_rng(T)
    for rep in range(20):
        pots = rng.random(T) * 1e5 + 1.0
        if T > 1 and rep % 3 == 0:
            pots[1] = pots[0] = pots.min() * 0.5                 # tie after fp32 rounding too: first minimum wins
        u = rng.random(T)
        if T > 2 and rep % 4 == 0:
            u[2] = u[0]                                          # equal thresholds: stable order
        p32 = pots.astype(np.float32)
        best = int(np.argmin(p32))
        v = u * np.float64(p32[best])
        t = v.astype(np.float32)
        low = t.astype(np.float64) < v
        t[low] = np.nextafter(t[low], np.float32(np.inf))
        order = np.argsort(t, kind='stable')
        thr_d, order_d, best_d = K.kpp_round_prep(dev(pots), dev(u))
        assert int(best_d.item()) == best
        assert thr_d.cpu().numpy().tobytes() == t[order].tobytes()
        assert np.array_equal(order_d.cpu().numpy(), order)
        picked = rng.integers(-1, 1000, T).astype(np.int64)
        want = np.empty(T, np.int64)
        want[order] = np.where(picked < 0, 999, picked)
        assert np.array_equal(K.kpp_unsort(dev(picked), order_d, 1000).cpu().numpy(), want)


@pytest.mark.parametrize('N,d,k', [(5000, 512, 6), (3001, 768, 3), (1000, 1024, 32), (33, 64, 2), (2000, 20, 4), (600, 1024, 60)])
def test_lloyd_assign_and_sums(K, N, d, k):
    X = blobs(N, d, kc=min(k, 8), seed=11, normalize=(d == 512))
    Xc = (X - X.mean(0)).astype(np.float32)
    C = Xc[np.random.default_rng(2).choice(N, k, replace=False)]
    want = RS.lloyd_assign(Xc, C)
    labels = torch.full((N,), -1, dtype=torch.int32, device='cuda')
    counts = torch.zeros(k, dtype=torch.int64, device='cuda')
    nch = torch.zeros(1, dtype=torch.int64, device='cuda')
    K.lloyd_assign(dev(Xc), dev(C), labels, counts, nch)
    got = labels.cpu().numpy()
    assert (got == want).mean() >= 0.9995, 'label flips beyond fp32 ties'
    assert int(nch.item()) == N and np.array_equal(counts.cpu().numpy(), np.bincount(got, minlength=k))
    counts.zero_(); nch.zero_()
    K.lloyd_assign(dev(Xc), dev(C), labels, counts, nch)
    assert int(nch.item()) == 0                               # second pass: nothing changes
    sums = K.cluster_sums(dev(Xc), labels, k).cpu().numpy()
    want_s = np.zeros((k, d)); np.add.at(want_s, got, Xc.astype(np.float64))
    np.testing.assert_allclose(sums, want_s, rtol=1e-12, atol=1e-10)
    C64 = np.zeros((k, d)); nz = np.bincount(got, minlength=k) > 0
    C64[nz] = want_s[nz] / np.bincount(got, minlength=k)[nz][:, None]
    sq, rt, far = K.cluster_dispersion(dev(Xc), labels, dev(C64), want_far=True)
    d2 = ((Xc.astype(np.float64) - C64[got]) ** 2).sum(1)
    np.testing.assert_allclose(sq.cpu().numpy(), np.bincount(got, weights=d2, minlength=k), rtol=1e-11)
    np.testing.assert_allclose(rt.cpu().numpy(), np.bincount(got, weights=np.sqrt(d2), minlength=k), rtol=1e-11)
    np.testing.assert_allclose(far.cpu().numpy(), d2.astype(np.float32), rtol=1e-6)
    C_new, C64d, stats = K.lloyd_update(dev(want_s), counts, nch, dev(C))
    st = stats.cpu().numpy()
    np.testing.assert_allclose(C_new.cpu().numpy()[nz], C64[nz].astype(np.float32), rtol=1e-6, atol=1e-9)
    assert np.array_equal(st[2:], np.bincount(got, minlength=k).astype(np.float64)) and st[1] == 0


@pytest.mark.parametrize('N,d,ks', [(5000, 512, [2, 3, 4, 5, 6]), (2001, 768, [2, 5]), (1500, 1024, [3, 7, 8]), (300, 64, [2])])
def test_lockstep_kernels_equal_single_labelling_kernels(K, N, d, ks):
    """One pass for several labellings (assign / sums / dispersion / update) must reproduce the per-k kernels."""
    X = blobs(N, d, kc=6, seed=13, normalize=(d == 512))
    Xc = dev((X - X.mean(0)).astype(np.float32))
    rng = np.random.default_rng(3)
    Cs = [Xc[torch.from_numpy(rng.choice(N, k, replace=False)).cuda()].contiguous() for k in ks]
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    nL, ktot = len(ks), int(off[-1])
    C_all = torch.cat(Cs).contiguous()
    labels_all = torch.full((nL, N), -1, dtype=torch.int32, device='cuda')
    counts = torch.zeros(ktot, dtype=torch.int64, device='cuda')
    changed = torch.zeros(nL, dtype=torch.int64, device='cuda')
    active = (1 << nL) - 1
    if nL > 1:
        active &= ~2                                          # labelling 1 frozen: must stay untouched
    K.lloyd_assign_multi(Xc, C_all, off, active, labels_all, 0, counts, changed)
    sums = K.cluster_sums_multi(Xc, labels_all, 0, off, active)
    for L, (k, C) in enumerate(zip(ks, Cs)):
        lo, hi = int(off[L]), int(off[L + 1])
        if not (active >> L) & 1:
            assert (labels_all[L] == -1).all() and int(counts[lo:hi].sum()) == 0 and int(changed[L]) == 0
            continue
        lab = torch.full((N,), -1, dtype=torch.int32, device='cuda')
        cnt = torch.zeros(k, dtype=torch.int64, device='cuda')
        ch = torch.zeros(1, dtype=torch.int64, device='cuda')
        K.lloyd_assign(Xc, C, lab, cnt, ch)
        assert torch.equal(labels_all[L], lab) and torch.equal(counts[lo:hi], cnt) and int(changed[L]) == int(ch[0]) == N
        assert torch.equal(sums[lo:hi], K.cluster_sums(Xc, lab, k))          # deterministic: bit-equal
    full = (1 << nL) - 1
    labels_all[labels_all < 0] = 0
    C64 = C_all.double().contiguous()
    sq, rt = K.cluster_dispersion_multi(Xc, labels_all, 0, off, full, C64)
    for L, k in enumerate(ks):
        lo, hi = int(off[L]), int(off[L + 1])
        s1, r1, _ = K.cluster_dispersion(Xc, labels_all[L].contiguous(), C64[lo:hi].contiguous())
        assert torch.allclose(sq[lo:hi], s1, rtol=1e-13) and torch.allclose(rt[lo:hi], r1, rtol=1e-13)
    sums = K.cluster_sums_multi(Xc, labels_all, 0, off, full)
    counts = torch.cat([torch.bincount(labels_all[L].long(), minlength=k) for L, k in enumerate(ks)])
    changed = torch.arange(nL, device='cuda', dtype=torch.int64)
    C_new, C64n, stats = K.lloyd_update_multi(sums, counts, changed, off, full, C_all, 2 + max(ks))
    for L, k in enumerate(ks):
        lo, hi = int(off[L]), int(off[L + 1])
        cn, _, st = K.lloyd_update(sums[lo:hi].contiguous(), counts[lo:hi].contiguous(), changed[L:L + 1].contiguous(), Cs[L])
        assert torch.equal(C_new[lo:hi], cn) and torch.allclose(stats[L, :2 + k], st, rtol=1e-12)


@pytest.mark.parametrize('N,d,R,thr', [(2000, 512, 32, 0.55), (777, 768, 5, 0.9), (100, 1024, 40, -1.0), (64, 64, 1, 0.3)])
def test_classify_kernel(K, N, d, R, thr):
    B = blobs(N + R, d, kc=8, seed=0)
    X, Rf = np.ascontiguousarray(B[:N]), np.ascontiguousarray(B[N:])
    wdist, wlab = RC.classify_core(X, Rf, R, thr)
    dist, lab = K.cosine_classify(K.normalize_rows(dev(X)), K.normalize_rows(dev(Rf)), thr)
    assert np.array_equal(lab.cpu().numpy(), wlab)            # identical labels (north_star)
    np.testing.assert_allclose(dist.cpu().numpy(), wdist, rtol=1e-5, atol=3e-7)


@pytest.mark.parametrize('N,d,thr', [(1500, 512, 0.2), (700, 768, 0.25), (65, 1024, 0.2), (2, 512, 0.2), (1, 512, 0.2)])
def test_dedup_kernel(engine, N, d, thr):
    X = plant_dupes(blobs(N, d, kc=4, seed=5, normalize=(d == 512)), frac=20, seed=1, normalize=(d == 512))[0]
    mins, inds, idx = engine.dedup(X, thr)
    rmins, rinds, ridx = RC.dedup_core(X, thr)
    ties = np.abs(rmins - np.float32(thr)) <= 1e-6
    assert np.array_equal(np.setdiff1d(idx, np.nonzero(ties)[0]), np.setdiff1d(ridx, np.nonzero(ties)[0]))
    if ties.any():
        print('threshold ties at rows', np.nonzero(ties)[0])
    np.testing.assert_allclose(mins, rmins, rtol=1e-4, atol=3e-7)
    assert (inds == rinds).mean() > 0.99 and inds[0] == 0 and mins[0] == np.float32(10000.0)


@pytest.mark.parametrize('N,d,thr', [(3000, 512, 0.2), (2500, 768, 0.25), (1000, 1024, 0.2), (300, 64, 0.2), (9000, 512, 0.2), (2, 512, 0.2)])
def test_dedup_tensor_core_kernel(engine, N, d, thr):
    """tcgen05 dedup (split fp16, three MMAs per K step) against the scikit-learn oracle: exact duplicate set
    (ties within 1e-6 of the threshold reported), logged distances as accurate as an fp32 GEMM."""
    X = plant_dupes(blobs(N, d, kc=4, seed=5, normalize=(d == 512)), frac=20, seed=1, normalize=(d == 512))[0]
    mins, inds, idx = engine.dedup(X, thr, force='tc')
    rmins, rinds, ridx = RC.dedup_core(X, thr)
    ties = np.abs(rmins - np.float32(thr)) <= 1e-6
    assert np.array_equal(np.setdiff1d(idx, np.nonzero(ties)[0]), np.setdiff1d(ridx, np.nonzero(ties)[0]))
    np.testing.assert_allclose(mins, rmins, rtol=2e-6, atol=4e-7)          # fp32-GEMM accuracy, not fp16
    assert (inds == rinds).mean() > 0.99 and inds[0] == 0 and mins[0] == np.float32(10000.0)
    smins, sinds, sidx = engine.dedup(X, thr, force='simt')
    assert np.array_equal(sidx, idx) or ties.any()


@pytest.mark.parametrize('N,d,k', [(500, 512, 3), (1000, 768, 5), (130, 64, 2)])
def test_silhouette_exact_kernel(engine, N, d, k):
    X = blobs(N, d, kc=k, seed=3, normalize=(d == 512))
    lab = RC.kmeans_fit(X, k, 0).labels_
    ctx = engine.prepare(X)
    got = engine.silhouette_scores(ctx, [dev(lab, torch.int32)], [k], force='exact')[0]
    want = RC.cluster_scores(X, lab)[0]
    assert abs(got - want) <= 1e-5 * abs(want) + 1e-8


def _tc_cluster_sums(engine, Xc_t, labels_t, ks):
    """Run pack + tcgen05 sweep and fold the per-cell sums into per-cluster sums [N, k] per labelling."""
    K = engine.K
    perm, chunk_cell, c2c, G, P = engine.build_cells(labels_t, ks)
    scale, shift, pad_m = engine.sil_scale(Xc_t)
    Xh, aug = K.sil_pack(Xc_t, perm, P, scale, shift, pad_m)
    cellsum = K.zeros((G, P), torch.float32)
    K.sil_sweep(Xh, aug, scale, chunk_cell, G, 0, P // 128, cellsum)
    torch.cuda.synchronize()
    cs, pm = cellsum.cpu().numpy().astype(np.float64), perm.cpu().numpy()
    out = []
    for j, k in enumerate(ks):
        S = np.zeros((Xc_t.shape[0], k))
        m = c2c[j].cpu().numpy()
        live = pm >= 0
        for g in range(G):
            S[pm[live], m[g]] += cs[g, live]
        out.append(S)
    return out


@pytest.mark.parametrize('N,d,ks', [(300, 64, [2]), (1000, 128, [2, 3]), (2500, 512, [2, 3, 4, 5, 6]),
                                    (1200, 768, [2, 4]), (900, 1024, [3]), (5000, 512, [2, 6])])
def test_tensor_core_sweep_matches_exact_rowsums(engine, N, d, ks):
    """The tcgen05 kernel against the fp64-computed per-row cluster sums of the oracle."""
    X = blobs(N, d, kc=4, seed=9, normalize=(d == 512))
    ctx = engine.prepare(X)
    rng = np.random.default_rng(0)
    labels = [rng.integers(0, k, N).astype(np.int32) for k in ks]
    got = _tc_cluster_sums(engine, ctx.Xc, [dev(l) for l in labels], ks)
    for S, lab, k in zip(got, labels, ks):
        want = RS.silhouette_rowsums(X, lab, k).astype(np.float64)
        # fp16 operands: each distance carries ~2^-12 relative rounding noise, sums average it out
        np.testing.assert_allclose(S, want, rtol=2e-4, atol=1e-3)
        rel = np.abs(S.sum(0) - want.sum(0)) / want.sum(0)
        assert rel.max() < 2e-5, rel                              # no systematic bias


@pytest.mark.parametrize('N,d,ks', [(700, 64, [2, 3]), (3000, 512, [2, 3, 4, 5, 6]), (1500, 768, [2, 4]), (1100, 1024, [3]), (10000, 512, [2, 6])])
def test_symmetric_sweep_matches_full_sweep(engine, N, d, ks):
    """Upper-triangle sweep (row sums + transposed column sums) against the full sweep and the oracle."""
    X = blobs(N, d, kc=4, seed=9, normalize=(d == 512))
    ctx = engine.prepare(X)
    rng = np.random.default_rng(0)
    labels = [dev(rng.integers(0, k, N).astype(np.int32)) for k in ks]
    engine.sym = False
    full = engine.silhouette_scores(ctx, labels, ks, force='tc')
    engine.sym = True
    sym = engine.silhouette_scores(ctx, labels, ks, force='tc')
    engine.sym = None
    np.testing.assert_allclose(sym, full, rtol=2e-5, atol=2e-6)
    if N <= 3000:
        want = [RC.cluster_scores(X, l.cpu().numpy())[0] for l in labels]
        np.testing.assert_allclose(sym, want, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize('N,d', [(40000, 512), (9000, 768), (300, 64)])
def test_dynamic_schedule_computes_every_tile_once(engine, N, d):
    """Work items claimed from the shared counter (the default) against the static round-robin list: same sums, and the
    launch reports exactly U(U+1)/2 tiles -- nothing skipped, nothing done twice."""
    K = engine.K
    X = blobs(N, d, kc=4, seed=5, normalize=(d == 512))
    ctx = engine.prepare(X)
    ks = [2, 5]
    rng = np.random.default_rng(1)
    labels = [dev(rng.integers(0, k, N).astype(np.int32)) for k in ks]
    perm, chunk_cell, c2c, G, P = engine.build_cells(labels, ks)
    Xc = engine._pad16(ctx.Xc, 64)
    scale, shift, pad_m = engine.sil_scale(Xc)
    Xh, aug = K.sil_pack(Xc, perm, P, scale, shift, pad_m)
    U = P // 256
    out = []
    for dynamic in (False, True, True):
        cellsum = K.zeros((G, P), torch.float32)
        tiles = K.zeros((1,), torch.int64)
        counter = K.zeros((1,), torch.int32)
        K.sil_sweep_sym(Xh, aug, scale, chunk_cell, G, 0, 1, cellsum, counter.data_ptr() if dynamic else None, tiles)
        assert int(tiles.item()) == U * (U + 1) // 2
        if dynamic:
            assert int(counter.item()) >= 1                      # items were claimed from the counter
        out.append(cellsum.double().cpu().numpy())
    np.testing.assert_allclose(out[1], out[0], rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(out[2], out[0], rtol=2e-5, atol=1e-3)
    engine.dynamic = False
    a = engine.silhouette_scores(ctx, labels, ks, force='tc')
    engine.dynamic = True
    b = engine.silhouette_scores(ctx, labels, ks, force='tc')
    np.testing.assert_allclose(a, b, rtol=2e-6, atol=1e-7)


def test_sweep_rejects_bad_shapes(K):
    from videotofaces_b200 import _native
    with pytest.raises(_native.NativeError):
        K.sil_sweep(torch.zeros((100, 64), dtype=torch.float16, device='cuda'), torch.zeros((2, 100, 16), dtype=torch.float16, device='cuda'), 1.0,
                    torch.zeros(4, dtype=torch.int32, device='cuda'), 1, 0, 1, torch.zeros((1, 100), device='cuda'))


@pytest.mark.parametrize('N,d', [(5, 512), (4097, 512), (3000, 768), (2049, 1024), (100, 36)])
def test_center_rows_with_stats(K, N, d):
    """The centring pass that also returns max |xc| and max |xc_row|^2: Xc bit-equal to X - mean, the maxima exact / 1 ulp."""
    X = blobs(N, d, kc=3, seed=N, normalize=False) + 0.5
    mean32 = (X.astype(np.float64).sum(0) / N).astype(np.float32)
    Xc, stats = K.center_rows_stats(dev(X), dev(mean32))
    want = X - mean32
    assert np.array_equal(Xc.cpu().numpy(), want)
    st = stats.cpu().numpy()
    assert st[0] == np.abs(want).max()
    np.testing.assert_allclose(st[1], (want.astype(np.float64) ** 2).sum(1).max(), rtol=1e-6)


@pytest.mark.parametrize('N,d,ks,active', [
    (5000, 512, [2, 3, 4, 5, 6], 0b11111),          # one group, 720 label tuples
    (3000, 768, [2, 3, 4, 5, 6], 0b10110),          # frozen labellings keep their sums; 2 column slices, the second half empty
    (1025, 1024, [7, 9, 11, 13], 0b1111),           # 9009 tuples: split into groups
    (777, 64, [48], 0b1),                           # the largest k one labelling may have
    (2048, 132, [3, 2], 0b11),                      # d not a multiple of 128
    (40000, 512, [2, 3, 4, 5, 6, 7, 8, 9, 10], 0b111111111),   # more than 8 labellings
])
def test_cell_sums_kernel(K, N, d, ks, active):
    """Cell-wise fixed-point cluster sums against a plain index_add of round(x * 2^S): integers, so bit-exact."""
    rng = np.random.default_rng(N)
    X = (rng.standard_normal((N, d), dtype=np.float32) * 0.3).astype(np.float32)
    # blob-like labels (few occupied cells) for the first labellings, random ones for the rest
    base = rng.integers(0, 6, N)
    labels = np.stack([(base % k if j < 2 else rng.integers(0, k, N)) for j, k in enumerate(ks)]).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    two_s = 2.0 ** 33
    acc = torch.full((int(off[-1]), d), 12345, dtype=torch.int64, device='cuda')
    K.cell_sums(dev(X), dev(labels), 0, off, active, two_s, acc)
    q = torch.from_numpy(np.rint(X.astype(np.float64) * two_s).astype(np.int64))
    got = acc.cpu()
    for L, k in enumerate(ks):
        lo, hi = int(off[L]), int(off[L + 1])
        if (active >> L) & 1:
            want = torch.zeros((k, d), dtype=torch.int64).index_add_(0, torch.from_numpy(labels[L].astype(np.int64)), q)
            assert torch.equal(got[lo:hi], want), (L, k)
        else:
            assert bool((got[lo:hi] == 12345).all())            # untouched
    # a row shard (b > 0) of a wider label matrix, as the ranks of a multi-GPU run call it
    b = N // 3
    acc2 = torch.zeros((int(off[-1]), d), dtype=torch.int64, device='cuda')
    K.cell_sums(dev(X[b:]), dev(labels), b, off, active, two_s, acc2)
    L = int(np.log2(active & -active))
    want = torch.zeros((ks[L], d), dtype=torch.int64).index_add_(0, torch.from_numpy(labels[L, b:].astype(np.int64)), q[b:])
    assert torch.equal(acc2.cpu()[int(off[L]):int(off[L + 1])], want)


@pytest.mark.parametrize('N,ks', [(700, [2, 3, 5]), (100_000, [2, 3, 4, 5, 6]), (33, [4]), (5000, [16, 16, 16]), (257, [2, 2])])
def test_build_cells_kernel(engine, N, ks):
    """Cell layout on the device: every row once, every 32-position chunk inside one cell, cells in key order and padded to
    32, P padded to 256, cell -> cluster tables right; same cell count as torch.unique of the label tuples."""
    K = engine.K
    rng = np.random.default_rng(N)
    base = rng.integers(0, 7, N)
    labels = np.stack([(base * (j + 1) + (rng.random(N) < 0.05)) % k for j, k in enumerate(ks)]).astype(np.int32)
    perm, chunk_cell, c2c, G, P = K.build_cells(dev(labels), ks)
    pm, cc, cm = perm.cpu().numpy(), chunk_cell.cpu().numpy(), c2c.cpu().numpy()
    assert P % 256 == 0 and pm.shape == (P,) and cc.shape == (P // 32,) and cm.shape == (len(ks), G)
    assert sorted(pm[pm >= 0].tolist()) == list(range(N))
    assert G == len(np.unique(labels.T, axis=0))
    cell_of_pos = np.repeat(cc, 32)
    live = pm >= 0
    for j in range(len(ks)):
        assert np.array_equal(cm[j][cell_of_pos[live]], labels[j][pm[live]])
    # cells are contiguous, in ascending order, each a multiple of 32 positions; padding only at the end of a cell
    used = cell_of_pos[:np.nonzero(live)[0].max() + 1]
    assert np.all(np.diff(used) >= 0)
    sizes = np.bincount(cell_of_pos[live], minlength=G)
    starts = np.concatenate([[0], np.cumsum((sizes + 31) // 32 * 32)])
    for g in range(G):
        seg = pm[starts[g]:starts[g + 1]]
        assert (seg[:sizes[g]] >= 0).all() and (seg[sizes[g]:] < 0).all()
    # the engine takes this path for narrow k-ranges and the sort-based one otherwise: same silhouette either way
    lab_t = [dev(l) for l in labels]
    p2 = engine.build_cells(lab_t, ks)
    assert p2[3] == G and p2[4] == P


@pytest.mark.parametrize('N,d,ks', [(5000, 512, [2, 3, 4, 5, 6]), (3001, 768, [2, 5]), (1500, 1024, [3, 4, 3]), (700, 128, [48]),
                                    (20000, 512, [2, 3, 4, 5, 2, 3])])
def test_cell_dispersion_kernel(K, N, d, ks):
    """Cell-wise dispersion (fp32 differences against hi + lo centres, exact fixed-point accumulation over the rows) against
    numpy fp64 and the all-fp64 warp-per-row kernel: per-row rounding ~1e-6, averaged down in the sums."""
    rng = np.random.default_rng(N)
    X = blobs(N, d, kc=5, seed=N, normalize=False)
    base = rng.integers(0, 6, N)
    labels = np.stack([(base + (rng.random(N) < 0.1)) % k for k in ks]).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    C64 = rng.standard_normal((int(off[-1]), d)) * 0.5
    bound = 4.5 * float((X.astype(np.float64) ** 2).sum(1).max() + (C64 ** 2).sum(1).max())
    sq_unit = 2.0 ** math.floor(math.log2(2.0 ** 62 / (N * bound)))
    rt_unit = 2.0 ** math.floor(math.log2(2.0 ** 62 / (N * math.sqrt(bound))))
    out = K.cell_dispersion(dev(X), dev(labels), 0, off, dev(C64), sq_unit, rt_unit).cpu().numpy()
    ktot = int(off[-1])
    sq, rt = out[:ktot] / sq_unit, out[ktot:] / rt_unit
    wsq, wrt = np.zeros(ktot), np.zeros(ktot)
    X64 = X.astype(np.float64)
    for L, k in enumerate(ks):
        d2 = ((X64 - C64[off[L] + labels[L]]) ** 2).sum(1)
        np.add.at(wsq, off[L] + labels[L], d2)
        np.add.at(wrt, off[L] + labels[L], np.sqrt(d2))
    np.testing.assert_allclose(sq, wsq, rtol=2e-6)
    np.testing.assert_allclose(rt, wrt, rtol=2e-6)
    osq, ort = K.cluster_dispersion_multi(dev(X), dev(labels), 0, off, (1 << len(ks)) - 1, dev(C64))
    np.testing.assert_allclose(sq, osq.cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(rt, ort.cpu().numpy(), rtol=2e-6)
    # integers: a second run gives the same bits, whatever order the CTAs claimed their chunks in
    out2 = K.cell_dispersion(dev(X), dev(labels), 0, off, dev(C64), sq_unit, rt_unit).cpu().numpy()
    assert np.array_equal(out, out2)
    # a labelling measured twice -- against its centres and against another set (the means of a tolerance-stopped fit) --
    # shares its label row: the cell key does not grow
    if len(ks) <= 11:
        C2 = np.concatenate([C64, C64[off[0]:off[1]] * 0.9 + 0.01])
        off2 = np.concatenate([off, [off[-1] + ks[0]]]).astype(np.int32)
        rows = list(range(len(ks))) + [0]
        assert K.cell_dispersion_ok(d, ks + [ks[0]], rows) or len(ks) >= 7
        if K.cell_dispersion_ok(d, ks + [ks[0]], rows):
            o3 = K.cell_dispersion(dev(X), dev(labels), 0, off2, dev(C2), sq_unit, rt_unit, rows).cpu().numpy()
            k2 = int(off2[-1])
            assert np.array_equal(o3[:ktot], out[:ktot]) and np.array_equal(o3[k2:k2 + ktot], out[ktot:])
            d2 = ((X64 - C2[off2[-2] + labels[0]]) ** 2).sum(1)
            want = np.zeros(ks[0]); np.add.at(want, labels[0], d2)
            np.testing.assert_allclose(o3[ktot:k2] / sq_unit, want, rtol=2e-6)


@pytest.mark.parametrize('N,d', [(9000, 512), (3000, 768), (700, 64)])
def test_dedup_screen_equals_split_sweep(engine, N, d):
    """Two-stage dedup (single-MMA screen + exact re-evaluation of the kept candidates) against the split-fp16 sweep and the
    scikit-learn oracle: same duplicate set, same logged distances."""
    K = engine.K
    X = plant_dupes(blobs(N, d, kc=4, seed=5, normalize=(d == 512)), frac=20, seed=1, normalize=(d == 512))[0]
    Xn = K.normalize_rows(dev(X))
    thr = np.float32(0.2)

    def decode(key):
        k = key.cpu().numpy()
        return (k >> 32).astype(np.uint32).view(np.float32), (k & 0xffffffff).astype(np.int64)
    m_scr, i_scr = decode(K.cosine_dedup_tc(Xn, 0, 1, float(thr), mode='screen'))
    assert K.dedup_stats['mode'] == 'screen'
    m_spl, i_spl = decode(K.cosine_dedup_tc(Xn, 0, 1, mode='split'))
    rmins, rinds, ridx = RC.dedup_core(X, thr)
    ties = np.abs(rmins - thr) <= 1e-6
    for m in (m_scr, m_spl):
        got = np.nonzero(m <= thr)[0]
        assert np.array_equal(np.setdiff1d(got, np.nonzero(ties)[0]), np.setdiff1d(ridx, np.nonzero(ties)[0]))
    # exact fp32 dots of the kept candidates: what is left against the oracle's SGEMM is the two summation orders (d = 768:
    # a few 1e-7 on a distance of 0.1), far inside the 1e-4 of the contract
    np.testing.assert_allclose(m_scr, rmins, rtol=1e-5, atol=4e-7)
    np.testing.assert_allclose(m_scr, m_spl, rtol=1e-5, atol=4e-7)
    assert (i_scr == rinds).mean() > 0.99 and (i_scr == i_spl).mean() > 0.99


@pytest.mark.parametrize('N,d,chunk', [(3000, 512, 2), (5000, 768, 3), (1500, 512, 1)])
def test_dedup_screen_column_chunks_continue_the_candidate_lists(engine, N, d, chunk, monkeypatch):
    """Large inputs run the screen as one launch per column chunk; a row's candidate lists travel from chunk to chunk through
    `cand`.  Forced here on small inputs (VTF_SCREEN_CHUNK tiles per launch): the keys must be BIT-IDENTICAL to the
    single-launch screen -- the same elements enter the same lists in the same order."""
    K = engine.K
    X = plant_dupes(blobs(N, d, kc=4, seed=9, normalize=(d == 512)), frac=20, seed=2, normalize=(d == 512))[0]
    Xn = K.normalize_rows(dev(X))
    monkeypatch.setenv('VTF_SCREEN_CHUNK', '100000')
    one = K.cosine_dedup_tc(Xn, 0, 1, 0.2, mode='screen').cpu().numpy()
    assert K.dedup_stats['screen_launches'] == 1
    monkeypatch.setenv('VTF_SCREEN_CHUNK', str(chunk))
    many = K.cosine_dedup_tc(Xn, 0, 1, 0.2, mode='screen').cpu().numpy()
    units = (N + 255) // 256
    assert K.dedup_stats['screen_launches'] == (units + chunk - 1) // chunk > 1
    assert np.array_equal(one, many)
    # and two ranks' shares (unit-cyclic) cover every row exactly as the single rank does
    k0 = K.cosine_dedup_tc(Xn, 0, 2, 0.2, mode='screen').cpu().numpy()
    k1 = K.cosine_dedup_tc(Xn, 1, 2, 0.2, mode='screen').cpu().numpy()
    assert np.array_equal(np.minimum(k0.view(np.uint64), k1.view(np.uint64)), one.view(np.uint64))


@pytest.mark.parametrize('N', [2, 3, 255, 256, 257, 513, 1024])
def test_dedup_screen_on_tiny_and_ragged_inputs(engine, N):
    """The tensor-core screen forced onto inputs of less than one or just over a whole number of 256-row units (padding rows
    and columns, a single work item, items of one tile): same decisions and distances as the fp32 kernel."""
    d = 512
    X = blobs(max(N, 8), d, kc=2, seed=N)[:N].copy()
    if N >= 3:
        X[N - 1] = X[0]                                       # an exact duplicate of the first row in the last row
    m_tc, i_tc, idx_tc = engine.dedup(X, 0.2, force='tc')
    m_fp, i_fp, idx_fp = engine.dedup(X, 0.2, force='simt')
    assert np.array_equal(idx_tc, idx_fp)
    np.testing.assert_allclose(m_tc[1:], m_fp[1:], rtol=1e-5, atol=4e-7)
    if N >= 3:
        assert idx_tc.size >= 1 and idx_tc[-1] == N - 1 and i_tc[N - 1] == 0 and m_tc[N - 1] <= 1e-6


def test_dedup_screen_flags_rows_it_cannot_decide(engine):
    """The screen keeps the four best candidates per column half and evaluates them exactly; a row is undecidable only if its
    exact minimum is above the threshold while a FIFTH column of a half could still be below -- i.e. five or more earlier
    rows inside the screen's error band (1e-3) right above the threshold.  Built on purpose here: the units of such rows are
    re-swept with the split operands, and the decisions equal the oracle's."""
    K = engine.K
    rng = np.random.default_rng(3)
    N, d, thr = 4096, 512, np.float32(0.2)
    X = blobs(N, d, kc=4, seed=8)
    for r in range(10):                                       # row i: six earlier rows (same column half) at cos-dist thr + {1..6}e-4
        i = 2000 + 50 * r
        xi = X[i].astype(np.float64)
        for m in range(6):
            j = 256 * m + 3 * r + m
            c = 1.0 - (float(thr) + (m + 1) * 1e-4)
            g = rng.standard_normal(d)
            g -= g.dot(xi) * xi
            g /= np.linalg.norm(g)
            X[j] = (c * xi + np.sqrt(1 - c * c) * g).astype(np.float32)
    mins, inds, idx = engine.dedup(X, float(thr), force='tc')
    assert K.dedup_stats['mode'] == 'screen' and K.dedup_stats['reswept_units'] > 0
    rmins, rinds, ridx = RC.dedup_core(X, thr)
    ties = np.abs(rmins - thr) <= 1e-6
    assert np.array_equal(np.setdiff1d(idx, np.nonzero(ties)[0]), np.setdiff1d(ridx, np.nonzero(ties)[0]))
    np.testing.assert_allclose(mins, rmins, rtol=1e-4, atol=4e-7)
    # ordinary data: nothing to re-sweep
    X2 = plant_dupes(blobs(9000, 512, kc=4, seed=5), frac=20, seed=1)[0]
    engine.dedup(X2, float(thr), force='tc')
    assert K.dedup_stats['reswept_units'] == 0


@pytest.mark.parametrize('N,d', [(5000, 512), (3001, 768), (1000, 1024), (77, 128), (300, 100)])
def test_normalize_col_stats_fused(K, N, d):
    """Encoder hand-off: normalise + column statistics in one pass = the two separate kernels (rows bit-equal)."""
    X = blobs(N, d, kc=3, seed=N, normalize=False) * np.float32(2.5) + np.float32(0.3)
    Xn, s, q = K.normalize_col_stats(dev(X))
    want = K.normalize_rows(dev(X))
    assert torch.equal(Xn, want)
    ws, wq = K.col_stats(want)
    np.testing.assert_allclose(s.cpu().numpy(), ws.cpu().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(q.cpu().numpy(), wq.cpu().numpy(), rtol=1e-12)
