"""-m gpu: the tensor-core silhouette sweep on inputs built to break fp16 operands / truncating fp32 accumulation
(VERDICT r1 'what's weak' 1-3): exact duplicate rows, tight clusters, two clusters that nearly coincide, a far outlier,
un-centred data -- each against sklearn.metrics.silhouette_score (/root/reference/src/videotofaces/grouping.py:105)
at the north_star tolerance 1e-4 relative, in the full and in the symmetric (benchmarked) form; plus the symmetric sweep
against the oracle with real KMeans labels at a size where its score is far from 0."""
import numpy as np
import pytest
import torch

from oracle import reference_calls as RC
from videotofaces_b200.synth import adversarial, blobs

pytestmark = pytest.mark.gpu

RTOL = 1e-4        # BASELINE.json north_star: "silhouette scores fall within 1e-4 relative"


def _labels_from_sklearn(X, ks):
    return [RC.kmeans_fit(X, k, 0).labels_.astype(np.int32) for k in ks]


@pytest.mark.parametrize('kind,ks', [('dupes', [2, 6]), ('tight', [3, 6]), ('close', [3, 4]), ('outlier', [2, 4]),
                                     ('offset', [2, 6])])
@pytest.mark.parametrize('sym', [False, True])
def test_sweep_on_adversarial_inputs(engine, kind, ks, sym):
    N, d = 8192, 512
    X = adversarial(kind, N, d)
    labels = _labels_from_sklearn(X, ks)
    want = [RC.cluster_scores(X, l)[0] for l in labels]
    ctx = engine.prepare(X)
    dl = [torch.from_numpy(l).cuda() for l in labels]
    engine.sym = sym
    try:
        got = engine.silhouette_scores(ctx, dl, ks, force='tc')
    finally:
        engine.sym = None
    exact = engine.silhouette_scores(ctx, dl, ks, force='exact')
    print(kind, 'sym' if sym else 'full', 'tc', got, 'exact', exact, 'sklearn', want,
          'rel', [abs(g - w) / abs(w) for g, w in zip(got, want)])
    np.testing.assert_allclose(exact, want, rtol=RTOL, atol=0)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)


def test_symmetric_sweep_vs_oracle_real_labels_20k(engine):
    """The benchmarked form (symmetric, claimed from the work counter) against sklearn with real KMeans labels at 20k
    rows: scores are O(0.05), so rtol=1e-4 with no atol really tests four digits."""
    N, d, ks = 20480, 512, [2, 4, 6]
    X = blobs(N, d, kc=8, noise=1.0, seed=3)
    engine.sym = True
    try:
        ctx, res = engine.kmeans_sweep(X, ks, 0)
        scores = engine.cluster_scores(ctx, res, force='tc')
    finally:
        engine.sym = None
    inits = [X[r.seeds] for r in res]
    ref = RC.cluster_core(X, ks, 0, inits=inits)
    for r, lab, sc, rsc, k in zip(res, ref['labels'], scores, ref['scores'], ks):
        assert np.array_equal(r.labels.cpu().numpy(), lab)
        assert np.array_equal(r.seeds, RC.kmeanspp_indices(X, k, 0))          # k-means++ seeds equal sklearn's
        assert abs(rsc[1]) > 1e-3
        np.testing.assert_allclose(sc[1:], rsc[1:], rtol=RTOL, atol=0)


def test_duplicate_rows_cancel_in_the_sweep(engine):
    """The augmented operand carries the tensor core's OWN squared norms (vtf_sil_diag), so T_pq = g_pq - m_p - m_q of two
    identical rows is whatever T_pp is: report it (0 = exact cancellation) against the old fp32-FMA norms, whose T_pp shows the
    accumulation bias (~3e-6 m) that sat under every small distance."""
    K = engine.K
    X = adversarial('dupes', 8192, 512)
    ctx = engine.prepare(X)
    lab = torch.zeros(8192, dtype=torch.int32, device='cuda')
    perm, chunk_cell, c2c, G, P = engine.build_cells([lab], [1])
    Xc = engine._pad16(ctx.Xc, 64)
    scale, shift, pad_m = engine.sil_scale(Xc)
    out = {}
    for tc_norms in (False, True):
        K.tc_norms = tc_norms
        try:
            Xh, aug = K.sil_pack(Xc, perm, P, scale, shift, pad_m)
        finally:
            K.tc_norms = True
        m = K.sil_diag(Xh, torch.zeros_like(aug), 0.5)
        resid = K.sil_diag(Xh, aug, 1.0)
        live = perm >= 0
        rel = (resid[live].abs() / m[live]).cpu().numpy()
        out[tc_norms] = rel
        print('tc_norms', tc_norms, 'max |T_pp| / m', rel.max(), 'mean', rel.mean(), 'exact zeros', float((rel == 0).mean()))
    assert out[True].max() <= 2.0 ** -22          # at most a couple of ulps of the norm
    assert out[True].mean() < 0.1 * out[False].mean()
