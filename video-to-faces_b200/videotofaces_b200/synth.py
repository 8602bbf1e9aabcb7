"""Seeded synthetic embeddings for the grouping hot path.

The reference has no data generator: real inputs are FaceNet (512-d, unit norm,
/root/reference/src/videotofaces/encoders/facenet.py:153) or ViT (768/1024-d, LayerNorm output,
encoders/vit.py:100-101) embeddings.  No encoder weights exist offline, so every test and
bench line uses these seeded stand-ins; CPU oracle and GPU path read identical bytes.
"""
import numpy as np


def _normalize_rows(X):
    n = np.sqrt(np.einsum('ij,ij->i', X, X))
    n[n == 0] = 1.0
    return (X / n[:, None]).astype(np.float32)


def blobs(N, d, kc=8, noise=1.0, seed=0, normalize=True):
    """`kc` Gaussian blobs in `d` dims; rows L2-normalised when `normalize` (FaceNet-like)."""
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((kc, d), dtype=np.float32)
    X = C[rng.integers(0, kc, N)]
    # generate the noise in slabs so a 1M x 1024 matrix does not need a second full-size temp
    step = 1 << 16
    for s in range(0, N, step):
        e = min(N, s + step)
        X[s:e] += np.float32(noise) * rng.standard_normal((e - s, d), dtype=np.float32)
    if normalize:
        X = _normalize_rows(X)
    return np.ascontiguousarray(X, dtype=np.float32)


def blob_truth(N, d, kc=8, seed=0, normalize=True):
    """(centres [kc,d], blob index of every row [N]) of blobs(N, d, kc, ., seed): the first two draws of the same
    generator, for tests that need ground-truth classes (test_grouping's labels.txt and reference embeddings)."""
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((kc, d), dtype=np.float32)
    ids = rng.integers(0, kc, N)
    return (_normalize_rows(C) if normalize else C), ids.astype(np.int64)


def eval_case(exclude_other):
    """(X, R, labels.txt numbers) of the two test_grouping golden cases (oracle/make_golden.py:testmode_cases): five
    heavily overlapping blobs, reference embeddings = jittered blob centres, ground truth = blob number (1-based)."""
    C, ids = blob_truth(500, 512, kc=5, seed=4)
    X = blobs(500, 512, kc=5, noise=5.0, seed=4)
    R = (C + np.float32(0.15) * np.random.default_rng(9).standard_normal(C.shape, dtype=np.float32)).astype(np.float32)
    return X, (np.ascontiguousarray(R[:4]) if exclude_other else R), ids + 1


def classify_faces_case():
    """(X, R, class names) of the classify_faces / classification-mode golden cases (oracle/make_golden.py:
    classify_faces_cases): five overlapping blobs, reference embeddings = jittered blob centres, so that some rows fall
    to a wrong class and some beyond the 'other' threshold."""
    C, _ = blob_truth(300, 512, kc=5, seed=6)
    X = blobs(300, 512, kc=5, noise=3.0, seed=6)
    R = (C + np.float32(0.1) * np.random.default_rng(3).standard_normal(C.shape, dtype=np.float32)).astype(np.float32)
    return X, R, ['anna', 'bob', 'carol', 'dave', 'eve']


def plant_dupes(X, frac=50, seed=1, normalize=True):
    """Plant m = N/frac near-duplicates: sources in rows < N/2, targets in rows >= N/2.

    x_dst = normalize(x_src + eps * g / sqrt(d)), eps ~ U(0.3, 1.1) so that cosine
    distances straddle the 0.2 threshold (about 0.04 .. 0.33).  Returns (X, src, dst).
    """
    X = X.copy()
    N, d = X.shape
    m = max(N // frac, 1) if N >= 4 else 0
    rng = np.random.default_rng(seed)
    half = N // 2
    src = rng.choice(half, size=m, replace=False)
    dst = half + rng.choice(N - half, size=m, replace=False)
    eps = rng.uniform(0.3, 1.1, size=m).astype(np.float32)
    g = rng.standard_normal((m, d), dtype=np.float32)
    scale = np.sqrt(np.einsum('ij,ij->i', X[src], X[src]))[:, None]
    Y = X[src] + scale * eps[:, None] * g / np.float32(np.sqrt(d))
    if normalize:
        Y = _normalize_rows(Y)
    X[dst] = Y
    return X, src, dst


def adversarial(kind, N=8192, d=512, seed=0):
    """Inputs built to stress the fp16-operand silhouette sweep ('what's weak' item 2) -- the shapes the
    application really produces (consecutive frames of one face, `--enc-dup-thr -1`) rather than well-separated blobs:

      'dupes'    30 % of the rows are exact copies: half of them single twins of random rows, half long runs of one
                 frame (a handful of rows repeated hundreds of times)
      'tight'    clusters whose spread is 1 % of their separation (intra / inter = 1e-2), not normalised
      'close'    two of the clusters are 20x closer to each other than to the rest: a and b of their rows are both
                 small against the (centred) norms, which is where cancellation in |x|^2+|y|^2-2xy hurts
      'outlier'  one row with 1e4 times the typical norm (forces the power-of-two fp16 scale down for everyone)
      'offset'   un-centred ViT-like rows: a common offset 50x the spread in every component
    """
    rng = np.random.default_rng(1000 + seed)
    if kind == 'dupes':
        X = blobs(N, d, kc=6, noise=1.0, seed=seed)
        n_twin, n_run = int(0.15 * N), int(0.15 * N)
        dst = rng.choice(N, size=n_twin + n_run, replace=False)
        src_pool = np.setdiff1d(np.arange(N), dst)
        X[dst[:n_twin]] = X[rng.choice(src_pool, size=n_twin, replace=True)]
        runs = rng.choice(src_pool, size=8, replace=False)
        X[dst[n_twin:]] = X[runs[rng.integers(0, 8, n_run)]]
        return np.ascontiguousarray(X)
    if kind == 'tight':
        return blobs(N, d, kc=6, noise=0.01, seed=seed, normalize=False)
    if kind == 'close':
        C = rng.standard_normal((4, d)).astype(np.float32)
        C[1] = C[0] + np.float32(0.05) * rng.standard_normal(d).astype(np.float32)
        X = C[rng.integers(0, 4, N)] + np.float32(0.02) * rng.standard_normal((N, d), dtype=np.float32)
        return np.ascontiguousarray(X, dtype=np.float32)
    if kind == 'outlier':
        X = blobs(N, d, kc=6, noise=1.0, seed=seed)
        X[N // 3] *= np.float32(1e4)
        return X
    if kind == 'offset':
        return np.ascontiguousarray(blobs(N, d, kc=6, noise=1.0, seed=seed, normalize=False) + np.float32(50.0))
    raise ValueError(kind)


def config(name):
    """The BASELINE.json configs as (kwargs for blobs, extras)."""
    table = {
        'c1': dict(N=8, d=512, kc=2, noise=0.5, seed=0, normalize=True),
        'c2': dict(N=50_000, d=512, kc=8, noise=1.0, seed=0, normalize=True),
        'c3': dict(N=1_000_000, d=512, kc=8, noise=1.0, seed=0, normalize=True),
        'c4': dict(N=1_000_000, d=1024, kc=16, noise=1.0, seed=0, normalize=False),
        'c5': dict(N=4_000_000, d=768, kc=8, noise=1.0, seed=0, normalize=False),
    }
    return dict(table[name])
