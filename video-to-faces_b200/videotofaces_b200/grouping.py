"""Classification and clustering of face embeddings -- drop-in for videotofaces.grouping
(/root/reference/src/videotofaces/grouping.py:50-137): `classify`, `classify_faces`, `cluster_faces`
with the reference's signatures, console output, folder layout and CSV logs.  The numbers come from
GroupingEngine (engine.py) instead of scikit-learn.
"""
import os
import os.path as osp

import numpy as np

from .fileops import copy_files, move_files, remove_files
from .runtime import get_engine, is_main, sync

TIE_EPS = 1e-6        # |distance - threshold| at or below this is reported as a tie (BASELINE.json north_star)


def _say(*a):
    """print on the rank that owns the console (every rank runs the arithmetic, rank 0 the side effects)."""
    if is_main():
        print(*a)


def classify(X, R, classes, thr, log, paths, out_dir):
    """grouping.py:50-66.  Appends 'other' to the CALLER's list when the threshold is active, as the
    reference does (grouping.py:56)."""
    dist, inds = get_engine().classify(X, R, thr)
    if thr and thr != -1:
        classes.append('other')
        ties = int(np.count_nonzero(np.abs(dist.min(axis=1) - np.float32(thr)) <= TIE_EPS)) if dist.size else 0
        if ties:                                                               # silent in the reference; north_star asks for it
            _say('NOTE: %u images have their nearest reference within %g of other_threshold=%s' % (ties, TIE_EPS, str(thr)))
    if log and is_main():
        names = [osp.basename(p) for p in paths]
        with open(osp.join(out_dir, 'faces', 'log_classification.csv'), 'w') as f:
            extra = '(other_threshold=%s)' % str(thr) if thr else ''
            f.write('file_name,' + ','.join('dist_' + c for c in classes if c != 'other') +
                    ',assigned_to_class' + extra + '\n')
            for i in range(dist.shape[0]):
                f.write('%s,' % names[i] + ','.join('%.4f' % v for v in dist[i]) + ',%s\n' % classes[inds[i]])
    return inds, classes


def encode_refs(refs, model):
    """grouping.py:42-46: the first image of every class goes through the (out-of-scope) encoder."""
    import cv2
    return model([cv2.imread(ps[0]) for (_, ps) in refs])


def classify_faces(paths, X, model, classif_params, R=None):
    """grouping.py:69-89.  `R` lets a caller hand over reference embeddings directly (the encoder
    networks are outside this package); otherwise `model` is called like the reference does."""
    refs, thr, log, out_dir = classif_params
    classes = [c for (c, _) in refs]
    _say('Found %u classes in ref_dir: %s' % (len(classes), ', '.join(classes)))
    _say('Extracting features from reference images')
    if R is None:
        R = encode_refs(refs, model)
    _say('Classifying images')
    inds, classes = classify(X, R, classes, thr, log, paths, out_dir)

    if is_main():
        img_dir = osp.dirname(osp.abspath(paths[0]))
        for c in classes:
            os.makedirs(osp.join(img_dir, c), exist_ok=True)
        move_files((p, osp.join(img_dir, classes[i], osp.basename(p))) for p, i in zip(paths, inds))   # grouping.py:79-80

    _say('Grouped %u images into %u folders:' % (len(paths), len(classes)))
    for i, c in enumerate(classes):
        _say(c + ': ' + str(np.count_nonzero(inds == i)))
    _say()
    sync()


def cluster_labels(X, clusters, rstate, inits=None, normalize=False):
    """The arithmetic of cluster_faces (grouping.py:95-120) without file system effects.
    Returns dict(clusters, labels [list of int32 arrays], scores [(k, sil, ch, db)], best_k,
    inertia, n_iter, centers).  normalize=True: X are un-normalised encoder outputs (a FaceNet without its closing
    F.normalize, encoders/facenet.py:153); the normalisation is fused into the first pass over X."""
    eng = get_engine()
    n = X.shape[0]
    clusters = [c for c in clusters if c <= n]                                 # grouping.py:95
    with eng.solo_if(eng.too_small_to_shard(n)):
        ctx, results = eng.kmeans_sweep(X, clusters, rstate, inits=inits, normalize=normalize)      # grouping.py:99-101
        # labels and centres are final here: their way to the host (side stream, helper thread) hides behind the sweep
        pending = eng.to_host_async([r.labels for r in results] + [r.centers for r in results])
        try:
            scores = eng.cluster_scores(ctx, results)                          # grouping.py:104-108
        except BaseException:
            pending.result()                                                   # the staging buffer is free again
            raise
    best_k = max(scores, key=lambda s: s[1])[0] if scores else None            # grouping.py:116 (first max)
    host = pending.result()
    nk = len(results)
    return dict(clusters=clusters, labels=host[:nk], scores=scores, best_k=best_k,
                inertia=[r.inertia for r in results], n_iter=[r.n_iter for r in results],
                centers=host[nk:], seeds=[r.seeds for r in results])


def cluster_faces(paths, X, cluster_params, normalize=False):
    """grouping.py:92-137.  `normalize` (not in the reference): see cluster_labels."""
    clusters, save_all, rstate, log, out_dir = cluster_params
    clusters = [c for c in clusters if c <= len(paths)]
    _say('Clustering images into %s groups' % ', '.join(str(cl) for cl in clusters))
    res = cluster_labels(X, clusters, rstate, normalize=normalize)
    labels, scores = res['labels'], res['scores']
    if log and is_main():
        with open(osp.join(out_dir, 'faces', 'log_clustering.csv'), 'w') as f:
            f.write('n_clusters,silhouette_score,calinski_harabasz_score,davies_bouldin_score\n')
            for score in scores:
                f.write('%u,%s,%s,%s\n' % score)

    if not save_all:
        best_k = res['best_k']
        i = clusters.index(best_k)
        clusters, labels = [clusters[i]], [labels[i]]
        _say('The number of groups chosen: %u' % best_k)

    _say('Grouped %u images into %s folders:' % (len(paths), '/'.join(str(cl) for cl in clusters)))
    img_dir = osp.dirname(osp.abspath(paths[0]))
    for k, lab in zip(clusters, labels):
        sub = 'G%u' % k if len(clusters) > 1 else ''
        if is_main():
            for j in range(k):
                os.makedirs(osp.join(img_dir, sub, str(j)), exist_ok=True)
            copy_files((p, osp.join(img_dir, sub, str(l), osp.basename(p))) for p, l in zip(paths, lab))  # grouping.py:129-130
        values, counts = np.unique(lab, return_counts=True)
        _say((sub + ': ' if sub else '') + ', '.join('%u: %u' % (v, c) for v, c in zip(values, counts)))
    _say()
    if is_main():
        remove_files(paths)                                                                        # grouping.py:136-137
    sync()


def get_ground_truths(paths, out_dir, exclude_other):
    """grouping.py:158-172: 1-based class numbers, one per line of <out_dir>/labels.txt; the highest number is the
    'other' class when `exclude_other` is set; the number of clusters is the highest remaining number."""
    try:
        with open(osp.join(out_dir, 'labels.txt')) as f:
            gt = np.asarray([int(x) for x in f.read().splitlines()])
    except Exception:  # noqa: BLE001  (the reference catches everything here, grouping.py:162)
        raise ValueError('Could not load ground truth labels for testing.'
                         'Expecting file "labels.txt" inside out_dir, filled with line-separated integers')
    keep = None
    if exclude_other:
        other_class = max(gt)
        keep = gt != other_class
        _say('Excluded %u images with "other" class' % np.count_nonzero(~keep))
        paths = [f for f, kp in zip(paths, keep) if kp]
        gt = gt[keep]
    return gt, paths, max(gt), keep


def grouping_test_scores(X, R, gt, n_clusters, thr, rstate, classes=None, log=False, paths=None, out_dir=None):
    """The three numbers of test_grouping (grouping.py:146-152) on embeddings that are already there:
    classification accuracy against 1-based ground truth, rand score and silhouette score of one KMeans fit."""
    eng = get_engine()
    classes = list(classes) if classes is not None else [str(i) for i in range(R.shape[0])]
    inds, _ = classify(X, R, classes, thr, log, paths, out_dir)                # grouping.py:146
    acc = np.count_nonzero(inds + 1 == gt) / gt.size                           # grouping.py:147
    with eng.solo_if(eng.too_small_to_shard(X.shape[0])):
        ctx, results = eng.kmeans_sweep(X, [int(n_clusters)], rstate)          # grouping.py:149
        rand_scr = eng.rand_score(gt, results[0].labels)                       # grouping.py:150
        silh_scr = eng.silhouette_scores(ctx, [results[0].labels], [results[0].k], [results[0].counts])[0]   # :151
    return acc, rand_scr, float(np.float32(silh_scr)), results[0]


def test_grouping(paths, refs, test_params, X=None, R=None):
    """grouping.py:140-155.  The encoder networks are outside this package: `X` / `R` hand the embeddings of `paths`
    (all of them, before the 'other' exclusion) and of the first image of every class over directly; without them
    the upstream encoders are used when that package is importable."""
    style, mname, device, out_dir, exclude_other, bs, area, thr, rstate = test_params
    gt, paths, n_clusters, keep = get_ground_truths(paths, out_dir, exclude_other)
    if X is None or R is None:
        from videotofaces.grouping import get_encoder_model, encode_faces    # the unmodified reference package
        model = get_encoder_model(style, mname, device)
        X = encode_faces(paths, model, bs, area)
        R = encode_refs(refs, model)
    elif keep is not None:
        X = np.asarray(X)[keep]
    acc, rand_scr, silh_scr, _ = grouping_test_scores(X, R, gt, n_clusters, None if exclude_other else thr, rstate,
                                                      classes=[c for (c, _) in refs], log=True, paths=paths, out_dir=out_dir)
    _say('%.4f / %.4f / %.4f' % (acc, rand_scr, silh_scr))
    _say('classification accuracy / rand score for clustering / silhouette score for clustering')
    sync()
test_grouping.__test__ = False                    # a reference entry point, not a pytest test
