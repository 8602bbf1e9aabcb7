"""TEST INFRASTRUCTURE: the classify_faces / classification-mode golden checks, shared by the CPU suite (engine over
tests/fake_kernels.py) and the -m gpu suite (the real engine).  Golden outputs come from the REAL reference package
(oracle/make_golden.py: golden_classify_faces, golden_classification_mode; /root/reference/src/videotofaces/
grouping.py:69-89, main.py:13-82) with only its encoder calls stubbed."""
import json
import os
import os.path as osp

import numpy as np

from videotofaces_b200 import grouping as vgrouping
from videotofaces_b200.main import video_to_faces
from videotofaces_b200.synth import classify_faces_case, plant_dupes


def fake_faces(out_dir, n):
    fdir = osp.join(out_dir, 'faces')
    os.makedirs(fdir, exist_ok=True)
    paths = [osp.join(fdir, 'f%06d.jpg' % i) for i in range(n)]
    for p in paths:
        with open(p, 'wb') as f:
            f.write(b'x')
    return paths


def tree(root, skip=('ref',)):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if not (d == root and x in skip))
        out.extend(osp.relpath(osp.join(d, f), root) for f in sorted(files))
    return sorted(out)


def manifest(golden_dir):
    with open(osp.join(golden_dir, 'MANIFEST.json')) as f:
        return json.load(f)['cases']


def compare_classification_log(got_path, want_path):
    """Header, file names and assigned classes equal; the '%.4f' distances within one unit of the last printed digit
    (1e-4 relative on distances in [0, 2] can move the rounding of the fourth decimal)."""
    got = open(got_path).read().strip().split('\n')
    want = open(want_path).read().strip().split('\n')
    assert got[0] == want[0] and len(got) == len(want)
    for a, b in zip(got[1:], want[1:]):
        a, b = a.split(','), b.split(',')
        assert a[0] == b[0] and a[-1] == b[-1], (a, b)
        assert np.abs(np.asarray(a[1:-1], float) - np.asarray(b[1:-1], float)).max() <= 1.0001e-4


def check_classify_faces(tmp_path, golden_dir, capsys, name, thr):
    want = manifest(golden_dir)[name]
    X, R, names = classify_faces_case()
    paths = fake_faces(str(tmp_path), X.shape[0])
    refs = [(c, ['unused.jpg']) for c in names]
    capsys.readouterr()
    vgrouping.classify_faces(list(paths), X, None, (refs, thr, True, str(tmp_path)), R=R)
    out = capsys.readouterr().out
    assert out == want['stdout']                                        # console text, class counts included
    assert tree(str(tmp_path)) == want['tree']                          # every file in the reference's folder
    compare_classification_log(osp.join(str(tmp_path), 'faces', 'log_classification.csv'),
                               osp.join(golden_dir, name + '.log_classification.csv'))
    assert refs == [(c, ['unused.jpg']) for c in names]                 # the refs list itself is not mutated (grouping.py:71)


def check_classification_mode(tmp_path, golden_dir, capsys, name='classification_mode_300x512'):
    want = manifest(golden_dir)[name]
    X, R, names = classify_faces_case()
    X = plant_dupes(X, frac=10, seed=5)[0]
    td = str(tmp_path)
    fake_faces(td, X.shape[0])
    for c in names + ['empty_class']:
        os.makedirs(osp.join(td, 'ref', c))
    for c in names:
        with open(osp.join(td, 'ref', c, 'a.jpg'), 'wb') as f:
            f.write(b'x')
    capsys.readouterr()
    video_to_faces(mode='grouping', style='live', out_dir=td, group_mode='classification', ref_dir=osp.join(td, 'ref'),
                   enc_dup_thr=0.2, enc_oth_thr=0.9, save_dupes=True, group_log=True, device='cpu',
                   features=X, ref_features=R)
    out = capsys.readouterr().out.replace(td, '<OUT>')
    assert out == want['stdout']
    assert tree(td) == want['tree']
    compare_classification_log(osp.join(td, 'faces', 'log_classification.csv'),
                               osp.join(golden_dir, name + '.log_classification.csv'))
