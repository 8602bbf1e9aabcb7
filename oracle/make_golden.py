"""ORACLE tooling (test infrastructure): generate tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

It imports videotofaces from /root/reference/src (IPython stubbed: utils/gallery.py:7 needs it
and the image has none), runs the three hot-path entry points on seeded synthetic embeddings
with throw-away image files, and stores only the OUTPUTS (labels, index sets, logs, scores);
inputs are regenerated from the seeds by videotofaces_b200.synth.  Library versions are recorded
in tests/golden/MANIFEST.json because scikit-learn is unpinned by the reference.
"""
import io
import json
import os
import os.path as osp
import sys
import tempfile
import types
import contextlib

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, osp.join(ROOT, 'video-to-faces_b200'))
from videotofaces_b200.synth import blobs, blob_truth, classify_faces_case, plant_dupes  # noqa: E402
sys.path.insert(0, ROOT)
from oracle.make_hashes import make_hashes  # noqa: E402

for m in ('IPython', 'IPython.display'):
    mod = types.ModuleType(m)
    mod.display = lambda *a, **k: None
    mod.HTML = lambda *a, **k: None
    sys.modules[m] = mod
sys.path.insert(0, '/root/reference/src')
import videotofaces.dupes as ref_dupes        # noqa: E402
import videotofaces.grouping as ref_grouping  # noqa: E402

GOLD = osp.join(ROOT, 'tests', 'golden')


def fake_faces(out_dir, n):
    fdir = osp.join(out_dir, 'faces')
    os.makedirs(fdir, exist_ok=True)
    paths = []
    for i in range(n):
        p = osp.join(fdir, 'f%06d.jpg' % i)
        with open(p, 'wb') as f:
            f.write(b'x')
        paths.append(p)
    return paths


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def golden_dedup(name, X, thr):
    with tempfile.TemporaryDirectory() as td:
        paths = fake_faces(td, X.shape[0])
        os.makedirs(osp.join(td, 'intermediate'), exist_ok=True)
        (Xk, goods), out = quiet(ref_dupes.remove_dupes_overall, X, paths, ('enc', thr, True, td))
        with open(osp.join(td, 'intermediate', 'log_dupes3.csv')) as f:
            log = f.read().replace(td, '<OUT>')
        kept = np.asarray([int(osp.basename(g)[1:7]) for g in goods], dtype=np.int64)
        rows = [l.split(',') for l in log.strip().split('\n')[1:]]
        mins = np.asarray([np.float32(r[2]) for r in rows], dtype=np.float32)
        inds = np.asarray([int(osp.basename(r[1])[1:7]) for r in rows], dtype=np.int64)
        flag = np.asarray([int(r[3]) for r in rows], dtype=np.int8)
    np.savez_compressed(osp.join(GOLD, name + '.npz'), kept=kept, mins=mins, inds=inds, flag=flag,
                        n_out=np.int64(Xk.shape[0]), thr=np.float64(thr))
    with open(osp.join(GOLD, name + '.log_dupes3.csv'), 'w') as f:
        f.write(log)
    return dict(stdout=out, n_kept=int(Xk.shape[0]))


def golden_hash(name, H, thr):
    with tempfile.TemporaryDirectory() as td:
        paths = fake_faces(td, H.shape[0])
        names = [osp.basename(p) for p in paths]          # detection.py hands over bare file names
        os.makedirs(osp.join(td, 'intermediate'), exist_ok=True)
        (Hk, goods), out = quiet(ref_dupes.remove_dupes_overall, H, names, ('hash', thr, True, td))
        with open(osp.join(td, 'intermediate', 'log_dupes2.csv')) as f:
            log = f.read()
        kept = np.asarray([int(g[1:7]) for g in goods], dtype=np.int64)
        rows = [l.split(',') for l in log.strip().split('\n')[1:]]
        mins = np.asarray([int(r[2]) for r in rows], dtype=np.int64)
        inds = np.asarray([int(r[1][1:7]) for r in rows], dtype=np.int64)
        flag = np.asarray([int(r[3]) for r in rows], dtype=np.int8)
    np.savez_compressed(osp.join(GOLD, name + '.npz'), kept=kept, mins=mins, inds=inds, flag=flag,
                        n_out=np.int64(Hk.shape[0]), thr=np.int64(thr))
    with open(osp.join(GOLD, name + '.log_dupes2.csv'), 'w') as f:
        f.write(log)
    return dict(stdout=out, n_kept=int(Hk.shape[0]))


def golden_classify(name, X, R, thr):
    classes = ['c%02d' % i for i in range(R.shape[0])]
    with tempfile.TemporaryDirectory() as td:
        paths = fake_faces(td, X.shape[0])
        (inds, classes_out), out = quiet(ref_grouping.classify, X, R, list(classes), thr, True, paths, td)
        with open(osp.join(td, 'faces', 'log_classification.csv')) as f:
            log = f.read()
    np.savez_compressed(osp.join(GOLD, name + '.npz'), inds=inds.astype(np.int64), thr=np.float64(thr if thr else 0))
    with open(osp.join(GOLD, name + '.log_classification.csv'), 'w') as f:
        f.write(log)
    return dict(classes_out=classes_out)


def golden_cluster(name, X, clusters, rstate):
    """save_all=True so that every k leaves its labels behind as folder names."""
    with tempfile.TemporaryDirectory() as td:
        paths = fake_faces(td, X.shape[0])
        _, out = quiet(ref_grouping.cluster_faces, list(paths), X, (list(clusters), True, rstate, True, td))
        with open(osp.join(td, 'faces', 'log_clustering.csv')) as f:
            log = f.read()
        ks = [c for c in clusters if c <= X.shape[0]]
        labels = {}
        for k in ks:
            lab = np.full(X.shape[0], -1, dtype=np.int32)
            sub = 'G%u' % k if len(ks) > 1 else ''
            for j in range(k):
                for fn in os.listdir(osp.join(td, 'faces', sub, str(j))):
                    lab[int(fn[1:7])] = j
            labels['labels_k%d' % k] = lab
    # the same call with save_all=False tells which k the reference picks
    with tempfile.TemporaryDirectory() as td:
        paths = fake_faces(td, X.shape[0])
        _, out2 = quiet(ref_grouping.cluster_faces, list(paths), X, (list(clusters), False, rstate, False, td))
    np.savez_compressed(osp.join(GOLD, name + '.npz'), ks=np.asarray(ks), **labels)
    with open(osp.join(GOLD, name + '.log_clustering.csv'), 'w') as f:
        f.write(log)
    return dict(stdout_all=out, stdout_best=out2)


def golden_testmode(name, X, R, gt, thr, rstate, exclude_other):
    """test_grouping (grouping.py:140-155) of the real reference.  Its encoder calls are the one part that cannot run
    here (no weights offline) and are not on the path: they are replaced by look-ups into X / R; classify, KMeans,
    rand_score, silhouette_score and the printed line are the reference's own."""
    saved = ref_grouping.get_encoder_model, ref_grouping.encode_faces, ref_grouping.encode_refs
    ref_grouping.get_encoder_model = lambda style, mname, device: None
    ref_grouping.encode_faces = lambda paths, model, bs, area: np.ascontiguousarray(
        X[[int(osp.basename(p)[1:7]) for p in paths]])
    ref_grouping.encode_refs = lambda refs, model: R
    try:
        with tempfile.TemporaryDirectory() as td:
            paths = fake_faces(td, X.shape[0])
            with open(osp.join(td, 'labels.txt'), 'w') as f:
                f.write('\n'.join(str(int(g)) for g in gt) + '\n')
            refs = [('c%02d' % i, ['unused.jpg']) for i in range(R.shape[0])]
            params = ('live', 'facenet_vgg', 'cpu', td, exclude_other, 16, None, thr, rstate)
            _, out = quiet(ref_grouping.test_grouping, paths, refs, params)
            with open(osp.join(td, 'faces', 'log_classification.csv')) as f:
                log = f.read()
    finally:
        ref_grouping.get_encoder_model, ref_grouping.encode_faces, ref_grouping.encode_refs = saved
    with open(osp.join(GOLD, name + '.log_classification.csv'), 'w') as f:
        f.write(log)
    return dict(stdout=out)


def _tree(root, skip=('ref',)):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if not (d == root and x in skip))
        out.extend(osp.relpath(osp.join(d, f), root) for f in sorted(files))
    return sorted(out)


def golden_classify_faces(name, X, R, names, thr):
    """classify_faces (grouping.py:69-89) of the real reference: its encode_refs (the encoder network, off the path and
    without weights offline) is replaced by a look-up of R; classify, the folder moves and the console text are its own."""
    saved = ref_grouping.encode_refs
    ref_grouping.encode_refs = lambda refs, model: R
    try:
        with tempfile.TemporaryDirectory() as td:
            paths = fake_faces(td, X.shape[0])
            refs = [(c, ['unused.jpg']) for c in names]
            _, out = quiet(ref_grouping.classify_faces, list(paths), X, None, (refs, thr, True, td))
            tree = _tree(td)
            with open(osp.join(td, 'faces', 'log_classification.csv')) as f:
                log = f.read()
    finally:
        ref_grouping.encode_refs = saved
    with open(osp.join(GOLD, name + '.log_classification.csv'), 'w') as f:
        f.write(log)
    return dict(stdout=out, tree=tree)


def golden_classification_mode(name, X, R, names, dup_thr, oth_thr, save_dupes):
    """video_to_faces(mode='grouping', group_mode='classification') of the real reference (main.py:13-82): argument
    checks, ref_dir scan, face scan, dedup, classify_faces, 'Done'.  Only the two encoder calls are stubbed."""
    import videotofaces.main as ref_main
    saved = ref_main.get_encoder_model, ref_main.encode_faces, ref_grouping.encode_refs
    ref_main.get_encoder_model = lambda style, mname, device: None
    ref_main.encode_faces = lambda paths, model, bs, area: np.ascontiguousarray(X[[int(osp.basename(p)[1:7]) for p in paths]])
    ref_grouping.encode_refs = lambda refs, model: R
    try:
        with tempfile.TemporaryDirectory() as td:
            fake_faces(td, X.shape[0])
            for c in names + ['empty_class']:
                os.makedirs(osp.join(td, 'ref', c))
            for c in names:
                with open(osp.join(td, 'ref', c, 'a.jpg'), 'wb') as f:
                    f.write(b'x')
            _, out = quiet(ref_main.video_to_faces, mode='grouping', style='live', out_dir=td, group_mode='classification',
                           ref_dir=osp.join(td, 'ref'), enc_dup_thr=dup_thr, enc_oth_thr=oth_thr, save_dupes=save_dupes,
                           group_log=True, device='cpu')
            tree = _tree(td)
            with open(osp.join(td, 'faces', 'log_classification.csv')) as f:
                log = f.read()
            out = out.replace(td, '<OUT>')
    finally:
        ref_main.get_encoder_model, ref_main.encode_faces, ref_grouping.encode_refs = saved
    with open(osp.join(GOLD, name + '.log_classification.csv'), 'w') as f:
        f.write(log)
    return dict(stdout=out, tree=tree)


def classify_faces_cases(man):
    X, R, names = classify_faces_case()
    gen = ("X=blobs(300,512,kc=5,noise=3,seed=6); C,_=blob_truth(300,512,kc=5,seed=6); "
           "R=C+0.1*default_rng(3).standard_normal(C.shape,float32); classes anna..eve; ")
    man['cases']['classify_faces_300x512'] = dict(gen=gen + 'thr=0.9', **golden_classify_faces('classify_faces_300x512', X, R, names, 0.9))
    man['cases']['classify_faces_300x512_nothr'] = dict(gen=gen + 'thr=-1',
                                                        **golden_classify_faces('classify_faces_300x512_nothr', X, R, names, -1))
    Xd = plant_dupes(X, frac=10, seed=5)[0]
    man['cases']['classification_mode_300x512'] = dict(
        gen=gen + "X=plant_dupes(X,frac=10,seed=5)[0]; video_to_faces(mode='grouping', group_mode='classification', "
                  "enc_dup_thr=0.2, enc_oth_thr=0.9, save_dupes=True); ref_dir has an extra image-less class folder",
        **golden_classification_mode('classification_mode_300x512', Xd, R, names, 0.2, 0.9, True))


def testmode_cases(man):
    # noise 5: the blobs overlap enough that classification, clustering and the 'other' threshold all make mistakes
    C, ids = blob_truth(500, 512, kc=5, seed=4)
    X = blobs(500, 512, kc=5, noise=5.0, seed=4)
    R = (C + np.float32(0.15) * np.random.default_rng(9).standard_normal(C.shape, dtype=np.float32)).astype(np.float32)
    gen = ("X=blobs(500,512,kc=5,noise=5,seed=4); C,ids=blob_truth(500,512,kc=5,seed=4); labels.txt=ids+1; "
           "R=C+0.15*default_rng(9).standard_normal(C.shape,float32); ")
    man['cases']['testmode_500x512'] = dict(gen=gen + "thr=0.97; rstate=0; exclude_other=False",
                                            **golden_testmode('testmode_500x512', X, R, ids + 1, 0.97, 0, False))
    man['cases']['testmode_500x512_excl'] = dict(gen=gen + "R=R[:4]; thr=0.97; rstate=0; exclude_other=True",
                                                 **golden_testmode('testmode_500x512_excl', X, np.ascontiguousarray(R[:4]),
                                                                   ids + 1, 0.97, 0, True))


def main():
    if '--classify-faces-only' in sys.argv:    # add the classify_faces / classification-mode cases to an existing manifest
        with open(osp.join(GOLD, 'MANIFEST.json')) as f:
            man = json.load(f)
        classify_faces_cases(man)
        with open(osp.join(GOLD, 'MANIFEST.json'), 'w') as f:
            json.dump(man, f, indent=1, sort_keys=True)
        print(json.dumps({k: v for k, v in man['cases'].items() if k.startswith('classif')}, indent=1)[:4000])
        return
    if '--testmode-only' in sys.argv:          # add the test_grouping cases to an existing manifest
        with open(osp.join(GOLD, 'MANIFEST.json')) as f:
            man = json.load(f)
        testmode_cases(man)
        with open(osp.join(GOLD, 'MANIFEST.json'), 'w') as f:
            json.dump(man, f, indent=1, sort_keys=True)
        print(json.dumps({k: v for k, v in man['cases'].items() if k.startswith('testmode')}, indent=1))
        return
    import sklearn, scipy
    os.makedirs(GOLD, exist_ok=True)
    man = dict(sklearn=sklearn.__version__, numpy=np.__version__, scipy=scipy.__version__,
               reference='/root/reference (sephirot-github/video-to-faces, no git metadata)', cases={})

    X = plant_dupes(blobs(400, 512, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0]
    man['cases']['dedup_400x512'] = dict(gen="plant_dupes(blobs(400,512,kc=4,noise=1,seed=0),frac=10,seed=1)[0]",
                                         **golden_dedup('dedup_400x512', X, 0.2))
    X = plant_dupes(blobs(300, 768, kc=3, noise=1.0, seed=3, normalize=False), frac=10, seed=2, normalize=False)[0]
    man['cases']['dedup_300x768'] = dict(gen="plant_dupes(blobs(300,768,kc=3,noise=1,seed=3,normalize=False),frac=10,seed=2,normalize=False)[0]",
                                         **golden_dedup('dedup_300x768', X, 0.25))

    man['cases']['hash_500'] = dict(gen='oracle.make_hashes.make_hashes(500, seed=7); hash_thr=8', **golden_hash('hash_500', make_hashes(500, 7), 8))

    B = blobs(332, 512, kc=8, noise=1.0, seed=0)
    X, R = np.ascontiguousarray(B[:300]), np.ascontiguousarray(B[300:])
    man['cases']['classify_300x512_r32'] = dict(gen="B=blobs(332,512,kc=8,noise=1,seed=0); X=B[:300]; R=B[300:]; thr=0.55",
                                                **golden_classify('classify_300x512_r32', X, R, 0.55))
    man['cases']['classify_300x512_r32_nothr'] = dict(gen="same, thr=-1",
                                                      **golden_classify('classify_300x512_r32_nothr', X, R, -1))

    X = blobs(60, 512, kc=3, noise=0.5, seed=0)
    man['cases']['cluster_60x512'] = dict(gen="blobs(60,512,kc=3,noise=0.5,seed=0); clusters 2..6; rstate 0",
                                          **golden_cluster('cluster_60x512', X, range(2, 7), 0))
    X = blobs(300, 768, kc=5, noise=1.0, seed=2, normalize=False)
    man['cases']['cluster_300x768'] = dict(gen="blobs(300,768,kc=5,noise=1,seed=2,normalize=False); clusters 2..5; rstate 0",
                                           **golden_cluster('cluster_300x768', X, range(2, 6), 0))
    X = blobs(8, 512, kc=2, noise=0.5, seed=0)
    man['cases']['cluster_c1_8x512'] = dict(gen="config 1 stand-in: blobs(8,512,kc=2,noise=0.5,seed=0); clusters 2..6; rstate 0",
                                            **golden_cluster('cluster_c1_8x512', X, range(2, 7), 0))
    X = blobs(2000, 512, kc=8, noise=1.0, seed=0)
    man['cases']['cluster_2000x512'] = dict(gen="blobs(2000,512,kc=8,noise=1,seed=0); clusters 2..6; rstate 0",
                                            **golden_cluster('cluster_2000x512', X, range(2, 7), 0))

    # the reference's failure mode: 4 rows and '2-6' keeps k=2,3,4 (grouping.py:95) and k=4 dies in
    # silhouette_score (SK/metrics/cluster/_unsupervised.py:42-46)
    X = blobs(4, 512, kc=2, noise=0.5, seed=0)
    try:
        with tempfile.TemporaryDirectory() as td:
            paths = fake_faces(td, 4)
            quiet(ref_grouping.cluster_faces, paths, X, (list(range(2, 7)), False, 0, False, td))
        man['cases']['cluster_4rows_error'] = 'no error?!'
    except ValueError as e:
        man['cases']['cluster_4rows_error'] = str(e)

    testmode_cases(man)
    classify_faces_cases(man)

    with open(osp.join(GOLD, 'MANIFEST.json'), 'w') as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print(json.dumps(man, indent=1)[:3000])


if __name__ == '__main__':
    main()
