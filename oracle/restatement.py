"""ORACLE (test infrastructure, not product code) -- numpy restatement of the arithmetic.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.

The grouping hot path of the reference delegates every number to scikit-learn (unpinned in
/root/reference/requirements.txt:4; the image holds 1.9.0).  This file restates the published
algorithms of those scikit-learn entry points in plain numpy so that the CUDA kernels can be
checked stage by stage (row norms, centred copy, k-means++ rounds, one Lloyd iteration,
per-row silhouette sums, ...), which the black-box calls in oracle/reference_calls.py cannot
offer.  `SK/` below is site-packages/sklearn of scikit-learn 1.9.0.

Pinned by tests/test_oracle_golden.py against (i) tests/golden/*.npz, produced by running
the real reference package (oracle/make_golden.py), and (ii) the scikit-learn known-answer
tests named in SURVEY.md section 4.
"""
import numpy as np

# ----------------------------------------------------------------------------- cosine


def row_norms_sq(X):
    """SK/utils/extmath.py:89 -- einsum('ij,ij->i') in the array dtype."""
    return np.einsum('ij,ij->i', X, X)


def normalize_l2(X):
    """SK/preprocessing/_data.py:2076-2081 with the zero-norm guard of :101-133."""
    X = np.asarray(X, dtype=np.float32)
    norms = np.sqrt(row_norms_sq(X))
    norms = norms.copy()
    norms[norms < 10 * np.finfo(np.float32).eps] = 1.0
    return X / norms[:, None]


def cosine_distances(X, Y=None):
    """SK/metrics/pairwise.py:1136-1182 (1 - cosine_similarity, clip to [0, 2], zero diagonal
    when Y is X) over SK/metrics/pairwise.py:1697-1752 (normalise both sides, SGEMM)."""
    Xn = normalize_l2(X)
    Yn = Xn if Y is None else normalize_l2(Y)
    S = Xn @ Yn.T
    S *= -1
    S += 1
    np.clip(S, 0, 2, out=S)
    if Y is None:
        np.fill_diagonal(S, 0.0)
    return S


def dedup_core(X, threshold):
    """dupes.py:60-65 without the N x N fp64 np.tri temporary: row i looks at columns j < i."""
    D = cosine_distances(X)
    N = X.shape[0]
    mins = np.empty(N, dtype=np.float32)
    inds = np.zeros(N, dtype=np.int64)
    # the reference adds 10000 to every masked entry (j >= i); row 0 is fully masked and
    # its min is D[0,0] + 10000 = 10000 at index 0
    mins[0] = np.float32(10000.0)
    for i in range(1, N):
        row = D[i, :i]
        j = int(row.argmin())
        mins[i] = row[j]
        inds[i] = j
    idx = (mins <= threshold).nonzero()[0]
    return mins, inds, idx


def hash_dedup_core(H, threshold):
    """dupes.py:55-65, 'hash' branch: Hamming distance between 64-bit hashes, each row against earlier rows."""
    Hb = np.asarray(H) != 0
    N = Hb.shape[0]
    mins = np.full(N, 10000, dtype=np.uint16)
    inds = np.zeros(N, dtype=np.int64)
    for i in range(1, N):
        dd = (Hb[:i] != Hb[i]).sum(axis=1)
        j = int(dd.argmin())
        mins[i], inds[i] = dd[j], j
    return mins, inds, (mins <= threshold).nonzero()[0]


def classify_core(X, R, n_classes, thr):
    """grouping.py:51-56."""
    dist = cosine_distances(X, R)
    inds = dist.argmin(axis=1)
    if thr and thr != -1:
        inds[dist.min(axis=1) >= thr] = n_classes
    return dist, inds


# ----------------------------------------------------------------------------- euclidean


def euclidean_sq_upcast(X, Y):
    """SK/metrics/pairwise.py:567-638 for fp32 inputs: fp64 compute, fp32 store, then the
    clamp of :413-416.  (The chunking of the original only bounds memory.)"""
    Xd = X.astype(np.float64)
    Yd = Y.astype(np.float64)
    d = -2.0 * (Xd @ Yd.T)
    d += np.einsum('ij,ij->i', Xd, Xd)[:, None]
    d += np.einsum('ij,ij->i', Yd, Yd)[None, :]
    d = d.astype(np.float32)
    np.maximum(d, 0, out=d)
    return d


# ----------------------------------------------------------------------------- k-means


def kmeans_tolerance(X, tol=1e-4):
    """SK/cluster/_kmeans.py:285-293."""
    return np.mean(np.var(X, axis=0)) * tol


def n_local_trials(k):
    """SK/cluster/_kmeans.py:228."""
    return 2 + int(np.log(k))


def kmeans_plusplus(Xc, k, rstate):
    """SK/cluster/_kmeans.py:180-278 on an already centred fp32 matrix, unit weights.

    Returns (indices int64 [k], trace) where trace lists, per round, the candidate ids and
    their potentials -- the quantities the GPU kernels are checked against.
    """
    rs = np.random.RandomState(rstate)
    N = Xc.shape[0]
    w = np.ones(N, dtype=np.float32)
    trials = n_local_trials(k)
    first = rs.choice(N, p=w / w.sum())                                   # :231
    idx = [int(first)]
    closest = euclidean_sq_upcast(Xc[first][None, :], Xc)                 # :240  [1, N] fp32
    pot = closest @ w                                                     # :243  fp32
    trace = []
    for _ in range(1, k):
        rand_vals = rs.uniform(size=trials) * pot                         # :249
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)         # :250  fp32 sequential cumsum
        np.clip(cand, None, N - 1, out=cand)                              # :254
        dc = euclidean_sq_upcast(Xc[cand], Xc)                            # :257
        np.minimum(closest, dc, out=dc)                                   # :262
        pots = dc @ w.reshape(-1, 1)                                      # :263
        b = int(np.argmin(pots))                                          # :266
        pot = pots[b]
        closest = dc[b][None, :]
        idx.append(int(cand[b]))
        trace.append(dict(cand=cand.copy(), pots=pots.ravel().copy(), best=b))
    return np.asarray(idx, dtype=np.int64), trace


def lloyd_assign(Xc, C):
    """E-step of SK/cluster/_k_means_lloyd.pyx:168-218: argmin_c (|c|^2 - 2 x.c) in fp32,
    first index wins ties (strict '<' at :210)."""
    c2 = row_norms_sq(C).astype(np.float32)
    labels = np.empty(Xc.shape[0], dtype=np.int32)
    for s in range(0, Xc.shape[0], 256):                                  # CHUNK_SIZE 256
        e = min(s + 256, Xc.shape[0])
        pd = c2[None, :] - np.float32(2.0) * (Xc[s:e] @ C.T)
        labels[s:e] = pd.argmin(axis=1)
    return labels


def lloyd_update(Xc, labels, C_old):
    """M-step: per-cluster sums and counts (lloyd.pyx:215-218), empty-cluster relocation
    (SK/cluster/_k_means_common.pyx:167-211), averaging (:274-295), centre shift (:298-311)."""
    k, d = C_old.shape
    sums = np.zeros((k, d), dtype=np.float32)
    np.add.at(sums, labels, Xc)
    wic = np.bincount(labels, minlength=k).astype(np.float32)
    empty = np.where(wic == 0)[0]
    if empty.size:
        dist = ((Xc - C_old[labels]) ** 2).sum(axis=1)
        if np.max(dist) != 0:
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for e_id, f in zip(empty, far):
                old = labels[f]
                sums[old] -= Xc[f]
                sums[e_id] = Xc[f]
                wic[e_id] = 1
                wic[old] -= 1
    C_new = sums.copy()
    amax = int(np.argmax(wic))
    for j in range(k):
        if wic[j] > 0:
            C_new[j] *= np.float32(1.0 / wic[j])
    for j in range(k):
        if not wic[j] > 0:
            C_new[j] = C_new[amax]
    shift = np.sqrt(((C_new - C_old) ** 2).sum(axis=1)).astype(np.float32)
    return C_new, wic, shift


def inertia(Xc, C, labels):
    """SK/cluster/_k_means_common.pyx:94-124: direct differences in fp32."""
    diff = Xc - C[labels]
    return float(np.einsum('ij,ij->', diff, diff, dtype=np.float32))


def kmeans_lloyd(X, init_centers=None, k=None, rstate=0, max_iter=300, tol=1e-4):
    """KMeans(n_clusters=k, random_state, n_init='auto').fit (SK/cluster/_kmeans.py:1436-1563)
    with algorithm='lloyd', one k-means++ init (n_init 'auto' -> 1, :885-887).

    `init_centers` (un-centred, [k, d]) replaces the seeding when given.
    Returns dict(labels, centers, inertia, n_iter, seeds).
    """
    X = np.ascontiguousarray(X, dtype=np.float32)
    tol_ = kmeans_tolerance(X, tol)                                       # :1473 / :882
    mean = X.mean(axis=0)                                                 # :1487
    Xc = X - mean
    seeds = None
    if init_centers is None:
        seeds, _ = kmeans_plusplus(Xc, k, rstate)
        C = Xc[seeds].copy()
    else:
        C = (np.asarray(init_centers, dtype=np.float32) - mean).astype(np.float32)
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    it = 0
    for it in range(max_iter):                                            # :705
        labels = lloyd_assign(Xc, C)
        C_new, _, shift = lloyd_update(Xc, labels, C)
        C = C_new                                                         # :721
        if np.array_equal(labels, labels_old):                            # :723
            strict = True
            break
        if (shift ** 2).sum() <= tol_:                                    # :731-738
            break
        labels_old = labels
    if not strict:
        labels = lloyd_assign(Xc, C)                                      # :742-754
    return dict(labels=labels, centers=C + mean, inertia=inertia(Xc, C, labels),
                n_iter=it + 1, seeds=seeds)


# ----------------------------------------------------------------------------- scores


def silhouette_rowsums(X, labels, k):
    """Per-row, per-cluster sums of euclidean distances: the quantity _silhouette_reduce
    accumulates (SK/metrics/cluster/_unsupervised.py:184-189) from the fp64-computed,
    fp32-stored, clamped, fp32-sqrt'ed distances of SK/metrics/pairwise.py:376-427 with a
    zeroed diagonal (:2285).  Returns fp32 [N, k]."""
    N = X.shape[0]
    out = np.zeros((N, k), dtype=np.float32)
    step = max(1, (1 << 24) // max(N, 1))
    for s in range(0, N, step):
        e = min(N, s + step)
        D = np.sqrt(euclidean_sq_upcast(X[s:e], X))
        D[np.arange(e - s), np.arange(s, e)] = 0.0
        for c in range(k):
            out[s:e, c] = D[:, labels == c].sum(axis=1, dtype=np.float64)
    return out


def silhouette_score(X, labels):
    """SK/metrics/cluster/_unsupervised.py:59-146, 211-323."""
    labels = np.unique(labels, return_inverse=True)[1]
    freqs = np.bincount(labels)
    k = freqs.size
    N = X.shape[0]
    if not 1 < k < N:
        raise ValueError('Number of labels is %d. Valid values are 2 to n_samples - 1 (inclusive)' % k)
    cd = silhouette_rowsums(X, labels, k)
    rows = np.arange(N)
    intra = cd[rows, labels].copy()
    cd[rows, labels] = np.inf
    cd /= freqs
    inter = cd.min(axis=1)
    denom = (freqs - 1).take(labels, mode='clip')
    with np.errstate(divide='ignore', invalid='ignore'):
        intra /= denom
        s = inter - intra
        s /= np.maximum(intra, inter)
    return float(np.mean(np.nan_to_num(s)))


def calinski_harabasz_score(X, labels):
    """SK/metrics/cluster/_unsupervised.py:333-403 (fp64 throughout)."""
    Xd = X.astype(np.float64)
    labels = np.unique(labels, return_inverse=True)[1]
    k = int(labels.max()) + 1
    N = X.shape[0]
    extra, intra = 0.0, 0.0
    mean = Xd.mean(axis=0)
    for c in range(k):
        ck = Xd[labels == c]
        mk = ck.mean(axis=0)
        extra += ck.shape[0] * np.sum((mk - mean) ** 2)
        intra += np.sum((ck - mk) ** 2)
    return float(1.0 if intra == 0.0 else extra * (N - k) / (intra * (k - 1.0)))


def davies_bouldin_score(X, labels):
    """SK/metrics/cluster/_unsupervised.py:413-487."""
    labels = np.unique(labels, return_inverse=True)[1]
    k = int(labels.max()) + 1
    intra = np.zeros(k)
    cent = np.zeros((k, X.shape[1]))
    for c in range(k):
        ck = X[labels == c]
        m = ck.mean(axis=0)
        cent[c] = m
        intra[c] = np.mean(np.sqrt(euclidean_sq_upcast(ck, m[None, :].astype(X.dtype))))
    cd = np.sqrt(np.maximum(((cent[:, None, :] - cent[None, :, :]) ** 2).sum(-1), 0))
    if np.allclose(intra, 0) or np.allclose(cd, 0):
        return 0.0
    cd[cd == 0] = np.inf
    comb = intra[:, None] + intra
    return float(np.mean(np.max(comb / cd, axis=1)))


def select_best(scores):
    """grouping.py:116 -- first maximum of the silhouette column."""
    return max(scores, key=lambda x: x[1])[0]


def rand_score(labels_true, labels_pred):
    """sklearn.metrics.rand_score (SK/metrics/cluster/_supervised.py:280-357) restated: contingency table (:106-192),
    pair confusion matrix (:195-271), (n_00 + n_11) / all pairs; 1.0 in the two degenerate cases."""
    a = np.unique(np.asarray(labels_true), return_inverse=True)[1]
    b = np.unique(np.asarray(labels_pred), return_inverse=True)[1]
    n = np.int64(a.shape[0])
    na, nb = int(a.max()) + 1 if n else 0, int(b.max()) + 1 if n else 0
    T = np.bincount(a.astype(np.int64) * nb + b, minlength=na * nb).reshape(na, nb).astype(np.int64)
    n_c, n_k = T.sum(axis=1), T.sum(axis=0)
    sum_squares = (T ** 2).sum()
    c11 = sum_squares - n
    c01 = T.dot(n_k).sum() - sum_squares
    c10 = T.T.dot(n_c).sum() - sum_squares
    c00 = n ** 2 - c01 - c10 - sum_squares
    num, den = c00 + c11, c00 + c01 + c10 + c11
    if num == den or den == 0:
        return 1.0
    return float(num / den)
