"""bench.py's host-side pieces that need no GPU: the t = a*N + c*N^2 fit behind the labelled CPU extrapolation, the
algorithmic-work table of SURVEY.md 8(d), the parity records."""
import importlib.util
import os.path as osp

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
spec = importlib.util.spec_from_file_location('vtf_bench', osp.join(ROOT, 'bench.py'))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_fit_recovers_linear_and_quadratic_terms():
    pts = [(n, 3e-4 * n + 5e-8 * n * n) for n in (6144, 12288, 16384)]
    a, c = bench.fit_quadratic(pts)
    assert abs(a - 3e-4) < 1e-9 and abs(c - 5e-8) < 1e-13
    ex = bench.extrapolate(dict(rows=1_000_000), pts, 16)
    assert abs(ex['full_seconds_EXTRAPOLATED'] - (300 + 5e4)) < 1e-3 and 'extrapolation' in ex['note']
    # noisy points that would give a negative coefficient fall back to the pure quadratic
    a, c = bench.fit_quadratic([(1000, 1.0), (2000, 1.5)])
    assert a >= 0 and c > 0


def test_kmeanspp_pass_count_matches_the_survey():
    assert bench.kpp_passes(range(2, 7)) == 8            # 2-6 -> 2 + 6
    assert bench.kpp_passes(range(2, 33)) == 61          # 2-32 -> 61


def test_phase_table_bounds():
    cfg = bench.config_of('c3')
    pk = dict(hbm=6531.9, tf_sust=1390.0, tf_burst=1662.8, source='test')
    N, d = 1_000_000, 512
    phases = {'prepare': 1.3, 'kmeans++': 11.9, 'lloyd': 18.1, 'sil_cells+pack': 2.5, 'sil_sweep': 364.0, 'sil_finish': 0.14, 'ch_db': 6.7}
    t = bench.phase_table(cfg, N, d, phases, pk, 1, dict(ms_per_step=405.7, n_iter=[3, 4, 4, 3, 3], any_not_strict=True))
    rows = {r['phase']: r for r in t['phases']}
    assert rows['silhouette']['bound'] == 'tensor'
    np.testing.assert_allclose(rows['silhouette']['bound_ms'], 2 * N * N * d / 1390e12 * 1e3, rtol=1e-12)
    np.testing.assert_allclose(rows['prepare']['bound_ms'], 2 * N * d * 4 / 6531.9e9 * 1e3, rtol=1e-12)
    np.testing.assert_allclose(rows['kmeans++']['algorithmic_bytes'], 8 * N * d * 4)
    assert rows['silhouette']['frac'] > 1.9                  # the symmetric sweep skips half of the algorithmic flops
    assert 0 < rows['lloyd']['frac'] < 1 and abs(t['total']['ms'] - 405.7) < 1e-9


def test_parity_records():
    X = np.zeros((4, 2), np.float32)
    ref = dict(mins=np.asarray([1e4, .1, .3, .2], np.float32), inds=np.asarray([0, 0, 1, 2]), idx=np.asarray([1, 3]))
    p = bench.dedup_parity(ref['mins'].copy(), ref['inds'].copy(), ref['idx'].copy(), ref, 0.2)
    assert p['ok'] and p['ties'] == 1 and p['dup_set_equal']
    p = bench.dedup_parity(ref['mins'].copy(), ref['inds'].copy(), np.asarray([1]), ref, 0.2)
    assert p['ok']                                           # row 3 sits on the threshold: a reported tie, not a mismatch
    p = bench.dedup_parity(ref['mins'].copy(), ref['inds'].copy(), np.asarray([2]), ref, 0.2)
    assert not p['ok']
