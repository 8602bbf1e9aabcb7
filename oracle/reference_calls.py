"""ORACLE (test infrastructure, not product code) -- the reference's own call sequence.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package (video-to-faces_b200/) never does.

The reference (sephirot-github/video-to-faces) authors no arithmetic on the grouping hot
path: src/videotofaces/dupes.py:60-68 and src/videotofaces/grouping.py:51-56, 99-120 are
thin call sites into scikit-learn, an un-vendored and UNPINNED dependency
(requirements.txt:4, setup.py:21).  This file restates those call sites verbatim in
behaviour (same sklearn entry points, same arguments, same masks / thresholds / tie rules)
without the file-system side effects, so it can run on the GPU box where /root/reference
does not exist.  The de-facto pin is the image's scikit-learn 1.9.0 / numpy 2.3.5 /
scipy 1.18.1 / OpenBLAS 0.3.30.

Parity status: the reference's own tests pin nothing on this path, so
the pin is "outputs of the reference itself run here": oracle/make_golden.py imports the
real package from /root/reference and writes tests/golden/*.npz; tests/test_oracle_golden.py
checks this module (and oracle/restatement.py) against those vectors and against the
scikit-learn known-answer tests listed in SURVEY.md section 4.
"""
import numpy as np
import sklearn.cluster
import sklearn.metrics
import sklearn.metrics.pairwise


def dedup_core(X, threshold):
    """remove_dupes_overall, 'enc' branch, arithmetic only (dupes.py:60-68).

    Returns (mins fp32 [N], inds int64 [N], idx int64 [n_dup]); row 0 is never a dupe because
    its whole distance row is masked with +10000.
    """
    D = sklearn.metrics.pairwise.cosine_distances(X)                      # dupes.py:60
    D += (1 - np.tri(X.shape[0], k=-1).astype(D.dtype)) * 10000          # dupes.py:62
    mins = D.min(axis=1)                                                  # dupes.py:63
    inds = D.argmin(axis=1)                                               # dupes.py:64
    idx = (mins <= threshold).nonzero()[0]                                # dupes.py:65
    return mins, inds, idx


def hash_dedup_core(H, threshold):
    """remove_dupes_overall, 'hash' branch (dupes.py:55-58, 62-65): H is [N,64] of 0/1 average hashes."""
    import sklearn.metrics
    D = sklearn.metrics.pairwise_distances(H, metric=lambda a, b: np.count_nonzero(a != b))   # dupes.py:56
    D = D.astype(np.uint16)                                                                   # dupes.py:58
    D += (1 - np.tri(H.shape[0], k=-1).astype(D.dtype)) * 10000
    mins = D.min(axis=1)
    inds = D.argmin(axis=1)
    idx = (mins <= threshold).nonzero()[0]
    return mins, inds, idx


def dedup_apply(X, filenames, idx):
    """dupes.py:66-68: surviving rows and names, order preserved."""
    keep = np.ones(X.shape[0], dtype=bool)
    keep[idx] = False
    goods = [fn for i, fn in enumerate(filenames) if keep[i]]
    return np.delete(X, idx, axis=0), goods


def classify_core(X, R, n_classes, thr):
    """classify, arithmetic only (grouping.py:51-56). Returns (dist fp32 [N,R], inds int64 [N])."""
    dist = sklearn.metrics.pairwise.cosine_distances(X, R)                # grouping.py:51
    inds = dist.argmin(axis=1)                                            # grouping.py:52
    if thr and thr != -1:                                                 # grouping.py:53
        mins = dist.min(axis=1)
        inds[mins >= thr] = n_classes                                     # grouping.py:55
    return dist, inds


def kmeans_fit(X, k, rstate, init=None):
    """One KMeans fit as the reference issues it (grouping.py:100); `init` overrides the
    k-means++ seeding with explicit centres (the "same initial centroids" parity contract)."""
    if init is None:
        cm = sklearn.cluster.KMeans(n_clusters=k, random_state=rstate, n_init='auto')
    else:
        cm = sklearn.cluster.KMeans(n_clusters=k, random_state=rstate, n_init=1, init=init)
    cm.fit(X)
    return cm


def cluster_scores(X, labels):
    """grouping.py:105-107."""
    s1 = sklearn.metrics.silhouette_score(X, labels)
    s2 = sklearn.metrics.calinski_harabasz_score(X, labels)
    s3 = sklearn.metrics.davies_bouldin_score(X, labels)
    return s1, s2, s3


def cluster_core(X, clusters, rstate, inits=None, with_scores=True):
    """cluster_faces, arithmetic only (grouping.py:95-120).

    Returns dict(clusters, labels, inertia, n_iter, centers, scores, best_k).
    """
    clusters = [c for c in clusters if c <= X.shape[0]]                   # grouping.py:95
    labels, inertia, n_iter, centers = [], [], [], []
    for j, k in enumerate(clusters):                                      # grouping.py:99-101
        cm = kmeans_fit(X, k, rstate, None if inits is None else inits[j])
        labels.append(cm.labels_)
        inertia.append(float(cm.inertia_))
        n_iter.append(int(cm.n_iter_))
        centers.append(cm.cluster_centers_)
    scores = []
    best_k = None
    if with_scores:
        for i in range(len(clusters)):                                    # grouping.py:104-108
            s1, s2, s3 = cluster_scores(X, labels[i])
            scores.append((clusters[i], s1, s2, s3))
        best_k = max(scores, key=lambda x: x[1])[0]                       # grouping.py:116
    return dict(clusters=clusters, labels=labels, inertia=inertia, n_iter=n_iter,
                centers=centers, scores=scores, best_k=best_k)


def kmeanspp_indices(X, k, rstate):
    """Seed rows chosen by the k-means++ that KMeans.fit runs internally.

    fit() centres X first (SK/cluster/_kmeans.py:1487-1490); distances are translation
    invariant up to rounding, and SURVEY.md section 8c verified that the public
    kmeans_plusplus on the raw X returns the same indices as the private call inside fit.
    To be exact we run it on the centred copy, as fit does.
    """
    Xc = X - X.mean(axis=0)
    _, idx = sklearn.cluster.kmeans_plusplus(Xc, k, random_state=rstate)
    return idx


def testmode_scores(X, R, gt, n_clusters, thr, rstate):
    """test_grouping, arithmetic only (grouping.py:146-151): (accuracy, rand score, silhouette score, labels)."""
    _, inds = classify_core(X, R, R.shape[0], thr)                        # grouping.py:146
    acc = np.count_nonzero(inds + 1 == gt) / gt.size                      # grouping.py:147
    cm = sklearn.cluster.KMeans(n_clusters=n_clusters, random_state=rstate).fit(X)      # grouping.py:149
    rand_scr = sklearn.metrics.rand_score(gt, cm.labels_)                 # grouping.py:150
    silh_scr = sklearn.metrics.silhouette_score(X, cm.labels_)            # grouping.py:151
    return acc, rand_scr, silh_scr, cm.labels_
