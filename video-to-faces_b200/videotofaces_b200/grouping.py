"""Classification and clustering of face embeddings -- drop-in for videotofaces.grouping
(/root/reference/src/videotofaces/grouping.py:50-137): `classify`, `classify_faces`, `cluster_faces`
with the reference's signatures, console output, folder layout and CSV logs.  The numbers come from
GroupingEngine (engine.py) instead of scikit-learn.
"""
import os
import os.path as osp
import shutil

import numpy as np

from .runtime import get_engine


def classify(X, R, classes, thr, log, paths, out_dir):
    """grouping.py:50-66.  Appends 'other' to the CALLER's list when the threshold is active, as the
    reference does (grouping.py:56)."""
    dist, inds = get_engine().classify(X, R, thr)
    if thr and thr != -1:
        classes.append('other')
    if log:
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
    print('Found %u classes in ref_dir: %s' % (len(classes), ', '.join(classes)))
    print('Extracting features from reference images')
    if R is None:
        R = encode_refs(refs, model)
    print('Classifying images')
    inds, classes = classify(X, R, classes, thr, log, paths, out_dir)

    img_dir = osp.dirname(osp.abspath(paths[0]))
    for c in classes:
        os.makedirs(osp.join(img_dir, c), exist_ok=True)
    for p, i in zip(paths, inds):
        os.replace(p, osp.join(img_dir, classes[i], osp.basename(p)))

    print('Grouped %u images into %u folders:' % (len(paths), len(classes)))
    for i, c in enumerate(classes):
        print(c + ': ' + str(np.count_nonzero(inds == i)))
    print()


def cluster_labels(X, clusters, rstate, inits=None):
    """The arithmetic of cluster_faces (grouping.py:95-120) without file system effects.
    Returns dict(clusters, labels [list of int32 arrays], scores [(k, sil, ch, db)], best_k,
    inertia, n_iter, centers)."""
    eng = get_engine()
    n = X.shape[0]
    clusters = [c for c in clusters if c <= n]                                 # grouping.py:95
    ctx, results = eng.kmeans_sweep(X, clusters, rstate, inits=inits)          # grouping.py:99-101
    scores = eng.cluster_scores(ctx, results)                                  # grouping.py:104-108
    best_k = max(scores, key=lambda s: s[1])[0] if scores else None            # grouping.py:116 (first max)
    return dict(clusters=clusters, labels=[r.labels.cpu().numpy() for r in results], scores=scores, best_k=best_k,
                inertia=[r.inertia for r in results], n_iter=[r.n_iter for r in results],
                centers=[r.centers.cpu().numpy() for r in results],
                seeds=[r.seeds for r in results])


def cluster_faces(paths, X, cluster_params):
    """grouping.py:92-137."""
    clusters, save_all, rstate, log, out_dir = cluster_params
    clusters = [c for c in clusters if c <= len(paths)]
    print('Clustering images into %s groups' % ', '.join(str(cl) for cl in clusters))
    res = cluster_labels(X, clusters, rstate)
    labels, scores = res['labels'], res['scores']
    if log:
        with open(osp.join(out_dir, 'faces', 'log_clustering.csv'), 'w') as f:
            f.write('n_clusters,silhouette_score,calinski_harabasz_score,davies_bouldin_score\n')
            for score in scores:
                f.write('%u,%s,%s,%s\n' % score)

    if not save_all:
        best_k = res['best_k']
        i = clusters.index(best_k)
        clusters, labels = [clusters[i]], [labels[i]]
        print('The number of groups chosen: %u' % best_k)

    print('Grouped %u images into %s folders:' % (len(paths), '/'.join(str(cl) for cl in clusters)))
    img_dir = osp.dirname(osp.abspath(paths[0]))
    for k, lab in zip(clusters, labels):
        sub = 'G%u' % k if len(clusters) > 1 else ''
        for j in range(k):
            os.makedirs(osp.join(img_dir, sub, str(j)), exist_ok=True)
        for p, l in zip(paths, lab):
            shutil.copyfile(p, osp.join(img_dir, sub, str(l), osp.basename(p)))
        values, counts = np.unique(lab, return_counts=True)
        print((sub + ': ' if sub else '') + ', '.join('%u: %u' % (v, c) for v, c in zip(values, counts)))
    print()
    for p in paths:
        os.remove(p)
