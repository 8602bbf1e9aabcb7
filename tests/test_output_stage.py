"""the native batched output stage (vtf_file_ops) leaves the same directory tree as
the reference's per-file loops (grouping.py:78-84, 122-137; dupes.py:70-79), and embeddings that are still on the device
go through the drop-in entry points without a host copy and give the same results as host arrays."""
import os
import os.path as osp

import numpy as np
import pytest

from videotofaces_b200 import fileops


def _tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            p = osp.join(d, f)
            with open(p, 'rb') as fh:
                out[osp.relpath(p, root)] = fh.read()
    return out


def _make(root, n):
    os.makedirs(root)
    paths = []
    for i in range(n):
        p = osp.join(root, 'face_%06d.jpg' % i)
        with open(p, 'wb') as f:
            f.write(b'jpeg%d' % i)
        paths.append(p)
    return paths


@pytest.mark.parametrize('threads', ['0', '1', '8'])
def test_copy_move_remove_match_the_plain_loops(tmp_path, monkeypatch, threads):
    monkeypatch.setenv('VTF_IO_THREADS', threads)
    n = 1500                                   # above VTF_IO_MIN_BATCH: the native batch runs unless threads == 0
    paths = _make(str(tmp_path / 'a'), n)
    labels = np.random.default_rng(0).integers(0, 5, n)
    for j in range(5):
        os.makedirs(tmp_path / 'a' / str(j))
        os.makedirs(tmp_path / 'b' / str(j))
    fileops.copy_files((p, osp.join(str(tmp_path / 'a'), str(l), osp.basename(p))) for p, l in zip(paths, labels))
    fileops.move_files((p, osp.join(str(tmp_path / 'b'), str(l), osp.basename(p))) for p, l in zip(paths[:700], labels))
    fileops.remove_files(paths[700:])
    # the same with the reference's loops
    import shutil
    paths2 = _make(str(tmp_path / 'ref' / 'a'), n)
    for j in range(5):
        os.makedirs(tmp_path / 'ref' / 'a' / str(j))
        os.makedirs(tmp_path / 'ref' / 'b' / str(j))
    for p, l in zip(paths2, labels):
        shutil.copyfile(p, osp.join(str(tmp_path / 'ref' / 'a'), str(l), osp.basename(p)))
    for p, l in zip(paths2[:700], labels):
        os.replace(p, osp.join(str(tmp_path / 'ref' / 'b'), str(l), osp.basename(p)))
    for p in paths2[700:]:
        os.remove(p)
    got = {k: v for k, v in _tree(str(tmp_path)).items() if not k.startswith('ref' + os.sep)}
    assert got == _tree(str(tmp_path / 'ref'))
    assert len(got) == n + 700


def test_errors_surface_after_the_batch(tmp_path, monkeypatch):
    monkeypatch.setenv('VTF_IO_THREADS', '4')
    paths = _make(str(tmp_path / 'a'), 1200)
    os.remove(paths[900])
    with pytest.raises(FileNotFoundError):
        fileops.remove_files(paths)
    assert not any(osp.exists(p) for p in paths)          # everything else was still removed


def test_io_threads_env(monkeypatch):
    monkeypatch.setenv('VTF_IO_THREADS', '3')
    assert fileops.io_threads() == 3
    monkeypatch.setenv('VTF_IO_THREADS', '0')
    assert fileops.io_threads() == 0                      # the reference's plain Python loop
    monkeypatch.delenv('VTF_IO_THREADS')
    assert 1 <= fileops.io_threads() <= 8                 # default: native batch (measured 2.8x / 5.4x faster, profiles/r02_fileops_bench.json)


def test_native_batch_semantics(tmp_path, monkeypatch):
    """copyfile semantics of the native path: contents only, existing destinations are overwritten, empty files and
    non-ASCII names work, a failing move reports the first failing path."""
    monkeypatch.setenv('VTF_IO_THREADS', '4')
    monkeypatch.setattr(fileops, 'MIN_BATCH', 1)
    src = tmp_path / 's'
    dst = tmp_path / 'd'
    src.mkdir(); dst.mkdir()
    names = ['a.jpg', 'пример.png', 'empty.bin', 'big.bin']
    (src / names[0]).write_bytes(b'x' * 10)
    (src / names[1]).write_bytes(b'y' * 70000)
    (src / names[2]).write_bytes(b'')
    (src / names[3]).write_bytes(os.urandom(3 << 20))
    (dst / names[0]).write_bytes(b'old content that is longer than the new one')
    fileops.copy_files((str(src / n), str(dst / n)) for n in names)
    for n in names:
        assert (dst / n).read_bytes() == (src / n).read_bytes()
    with pytest.raises(FileNotFoundError) as ei:
        fileops.move_files([(str(src / 'a.jpg'), str(dst / 'moved.jpg')), (str(src / 'missing'), str(dst / 'x')), (str(src / 'big.bin'), str(dst / 'b2'))])
    assert ei.value.filename == str(src / 'missing')
    assert (dst / 'moved.jpg').exists() and (dst / 'b2').exists()      # the rest of the batch was still carried out


@pytest.mark.gpu
def test_device_resident_embeddings_go_through_the_drop_in_entry_points(tmp_path, engine, capsys):
    import torch
    from videotofaces_b200 import dupes, grouping
    from videotofaces_b200.synth import blobs, plant_dupes
    N = 3000
    X = plant_dupes(blobs(N, 512, kc=4, seed=3), frac=25, seed=2)[0]
    for sub in ('host', 'dev'):
        os.makedirs(tmp_path / sub / 'faces')
        for i in range(N):
            (tmp_path / sub / 'faces' / ('%05d.jpg' % i)).write_bytes(b'x')
    names = ['%05d.jpg' % i for i in range(N)]
    Xh, goods_h = dupes.remove_dupes_overall(X, [str(tmp_path / 'host' / 'faces' / n) for n in names],
                                             ('enc', 0.2, True, str(tmp_path / 'host')))
    Xd, goods_d = dupes.remove_dupes_overall(torch.from_numpy(X).cuda(), [str(tmp_path / 'dev' / 'faces' / n) for n in names],
                                             ('enc', 0.2, True, str(tmp_path / 'dev')))
    assert isinstance(Xd, torch.Tensor) and Xd.is_cuda                     # no host copy of the survivors
    assert np.array_equal(Xd.cpu().numpy(), Xh)
    # a HOST matrix with keep_on_device (what video_to_faces does): one upload, the survivors stay in HBM
    r = dupes.dedup_rows(X, 0.2, keep_on_device=True)
    assert isinstance(r['kept'], torch.Tensor) and r['kept'].is_cuda and np.array_equal(r['kept'].cpu().numpy(), Xh)
    assert [osp.basename(g) for g in goods_d] == [osp.basename(g) for g in goods_h]
    assert sorted(os.listdir(tmp_path / 'dev' / 'intermediate' / 'dupes3')) == sorted(os.listdir(tmp_path / 'host' / 'intermediate' / 'dupes3'))
    a = grouping.cluster_labels(Xh, [2, 3, 4], 0)
    b = grouping.cluster_labels(Xd, [2, 3, 4], 0)
    assert all(np.array_equal(x, y) for x, y in zip(a['labels'], b['labels']))
    np.testing.assert_allclose(np.asarray(a['scores']), np.asarray(b['scores']), rtol=1e-6)   # the sweep's fp32 RED order varies
    R = X[:7].copy()
    ih, _ = grouping.classify(Xh, R, [str(i) for i in range(7)], 0.8, False, None, None)
    idv, _ = grouping.classify(Xd, torch.from_numpy(R).cuda(), [str(i) for i in range(7)], 0.8, False, None, None)
    assert np.array_equal(ih, idv)


@pytest.mark.gpu
def test_pageable_upload_is_exact_and_used_for_plain_numpy_inputs(engine):
    """main.py:71 hands the grouping stage an ordinary (pageable) numpy array; large ones go to the device through
    vtf_upload_pageable (native threads + page-locked double buffers).  Byte-exact, any size, repeated calls reuse the buffers."""
    import torch
    K = engine.K
    rng = np.random.default_rng(0)
    for n_rows in (1, 4097, 40000):
        A = rng.standard_normal((n_rows, 516), dtype=np.float32)          # 516 columns: pieces end in the middle of rows
        for threads in (1, 3, 8):
            got = K.upload_pageable(A, torch.float32, threads)
            assert torch.equal(got.cpu(), torch.from_numpy(A))
    big = rng.standard_normal((20000, 512), dtype=np.float32)             # 41 MB: engine.to_device takes the native path
    t = engine.to_device(big)
    assert t.is_cuda and torch.equal(t.cpu(), torch.from_numpy(big))
    a = engine.dedup(big[:9000], 0.2)
    b = engine.dedup(torch.from_numpy(big[:9000]).cuda(), 0.2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.gpu
def test_results_that_travel_behind_the_sweep_equal_the_synchronous_copy(engine):
    """cluster_labels sends labels and centres to the host on a side stream while the sweep runs (engine.to_host_async); the
    arrays equal a plain synchronous copy, and a second call reuses the staging buffer without mixing results up."""
    import torch
    from videotofaces_b200 import grouping
    from videotofaces_b200.synth import blobs
    ts = [torch.arange(100000, device='cuda', dtype=torch.int32) * 3, torch.randn(7, 512, device='cuda'),
          torch.zeros(0, device='cuda', dtype=torch.int32)]
    p1 = engine.to_host_async(ts)
    a = p1.result()
    p2 = engine.to_host_async([t + 1 for t in ts[:2]])
    b = p2.result()
    for got, t in zip(a, ts):
        assert np.array_equal(got, t.cpu().numpy())
    assert np.array_equal(b[0], (ts[0] + 1).cpu().numpy()) and np.array_equal(b[1], (ts[1] + 1).cpu().numpy())
    X = blobs(9000, 512, kc=3, seed=4)
    r1 = grouping.cluster_labels(X, [2, 3, 4], 0)
    ctx, res = engine.kmeans_sweep(X, [2, 3, 4], 0)
    assert all(np.array_equal(l, r.labels.cpu().numpy()) for l, r in zip(r1['labels'], res))
    assert all(np.array_equal(c, r.centers.cpu().numpy()) for c, r in zip(r1['centers'], res))
