"""The drop-in entry points (grouping.py / dupes.py / main.py of the package) on CPU: the host logic -- console text,
folder moves, logs, the 'other' threshold, tie reports -- driven by the stand-in kernels of tests/fake_kernels.py and
checked against golden outputs of the REAL reference (tests/golden/, oracle/make_golden.py).  The same checks run
against the CUDA engine in tests/test_gpu_parity.py."""
import os
import os.path as osp

import numpy as np
import pytest

import dropin_cases as DC
from fake_kernels import FakeKernels
from oracle import reference_calls as RC
from videotofaces_b200 import dupes as vdupes, grouping as vgrouping, runtime
from videotofaces_b200.engine import GroupingEngine
from videotofaces_b200.synth import blobs, classify_faces_case


@pytest.fixture()
def cpu_engine():
    runtime.reset_engine(GroupingEngine(FakeKernels()))
    yield runtime.get_engine()
    runtime.reset_engine(None)


@pytest.mark.parametrize('name,thr', [('classify_faces_300x512', 0.9), ('classify_faces_300x512_nothr', -1)])
def test_classify_faces_golden_host_logic(cpu_engine, tmp_path, golden_dir, capsys, name, thr):
    DC.check_classify_faces(tmp_path, golden_dir, capsys, name, thr)


def test_classification_mode_golden_host_logic(cpu_engine, tmp_path, golden_dir, capsys):
    DC.check_classification_mode(tmp_path, golden_dir, capsys)


@pytest.mark.parametrize('thr', [-0.5, 0.0, None, -1, 2.5])
def test_threshold_switch_matches_reference_rule(cpu_engine, thr):
    """grouping.py:53: `if thr and thr != -1` -- 0.0 / None / -1 switch the threshold off, any other value (a negative
    one included: every row goes to 'other') switches it on."""
    X, R, names = classify_faces_case()
    classes = list(names)
    inds, out = vgrouping.classify(X, R, classes, thr, False, None, None)
    _, want = RC.classify_core(X, R, len(names), thr)
    assert np.array_equal(inds, want)
    assert out is classes and (classes[-1] == 'other') == bool(thr and thr != -1)


def test_ties_at_the_thresholds_are_reported(cpu_engine, tmp_path, capsys):
    """north_star: 'ties at thresholds reported'.  A row whose nearest distance equals the threshold within 1e-6 is
    named on the console; nothing is printed otherwise (the reference prints nothing)."""
    X, R, names = classify_faces_case()
    dist, _ = cpu_engine.classify(X, R, None)
    thr = float(dist.min(axis=1)[7])                                 # row 7 sits exactly on the threshold
    vgrouping.classify(X, R, list(names), thr, False, None, None)
    assert 'within 1e-06 of other_threshold' in capsys.readouterr().out
    vgrouping.classify(X, R, list(names), 0.9, False, None, None)
    assert capsys.readouterr().out == ''
    paths = DC.fake_faces(str(tmp_path), X.shape[0])
    mins, _, _ = cpu_engine.dedup(X, 0.2)
    vdupes.remove_dupes_overall(X, paths, ('enc', float(mins[5]), True, str(tmp_path)))
    assert 'within 1e-06 of enc_dup_thr' in capsys.readouterr().out


def test_done_is_printed_like_the_reference(cpu_engine, tmp_path, capsys):
    """main.py:82."""
    from videotofaces_b200.main import video_to_faces
    X = blobs(40, 512, kc=2, noise=0.5, seed=0)
    DC.fake_faces(str(tmp_path), 40)
    video_to_faces(mode='grouping', style='live', out_dir=str(tmp_path), clusters='2-3', enc_dup_thr=-1, features=X)
    assert capsys.readouterr().out.endswith('\nDone\n')


def test_keep_on_device_hands_the_survivors_over_as_a_tensor(cpu_engine, tmp_path):
    """video_to_faces keeps the rows that survive dedup on the device for the grouping call (main.py:74-80 only passes them
    on); the public entry point returns numpy unless asked."""
    import torch
    X = blobs(300, 512, kc=3, noise=0.5, seed=1)
    X[200:220] = X[10:30]                                              # exact duplicates
    a = vdupes.dedup_rows(X, 0.2)
    b = vdupes.dedup_rows(X, 0.2, keep_on_device=True)
    assert isinstance(a['kept'], np.ndarray) and isinstance(b['kept'], torch.Tensor)
    assert np.array_equal(a['kept'], b['kept'].cpu().numpy()) and np.array_equal(a['idx'], b['idx'])
    assert a['idx'].size >= 20
    paths = DC.fake_faces(str(tmp_path), X.shape[0])
    Xk, goods = vdupes.remove_dupes_overall(X, paths, ('enc', 0.2, False, str(tmp_path)), keep_on_device=True)
    assert isinstance(Xk, torch.Tensor) and Xk.shape[0] == len(goods) == X.shape[0] - a['idx'].size
