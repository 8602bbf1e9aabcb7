"""World-size-2 run of the sharded path on CPU (gloo): the exchange steps of dist.py / engine.py --
all-reduce of column statistics, candidate potentials, centroid sums and silhouette partial sums,
all-gather of closest distances and labels, min-reduce of the dedup keys -- must reproduce the
single-process result and the scikit-learn oracle.  The per-shard compute is the CPU stand-in of
tests/fake_kernels.py; the CUDA kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'video-to-faces_b200'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fake_kernels import FakeKernels
        from videotofaces_b200.engine import GroupingEngine
        from videotofaces_b200.dist import Comm
        from videotofaces_b200.synth import blobs, plant_dupes
        eng = GroupingEngine(FakeKernels(), Comm())
        assert eng.comm.world == world and eng.comm.rank == rank
        X = blobs(700, 128, kc=5, noise=1.0, seed=1)
        ks = [2, 3, 4]
        ctx, res = eng.kmeans_sweep(X, ks, 0)
        exact = eng.cluster_scores(ctx, res, force='exact')
        tc = eng.silhouette_scores(ctx, [r.labels for r in res], ks, [r.counts for r in res], force='tc')
        eng.sym = True
        tc_sym = eng.silhouette_scores(ctx, [r.labels for r in res], ks, [r.counts for r in res], force='tc')
        eng.sym = None
        assert np.allclose(tc_sym, tc, rtol=1e-6), (tc_sym, tc)
        # the wide (tensor-core E-step, fixed-point M-step) Lloyd path, forced: same fits through other exchange steps
        eng.wide = True
        _, wres = eng.kmeans_sweep(X, ks, 0, ctx=ctx)
        eng.wide = None
        for r, w in zip(res, wres):
            assert torch.equal(r.labels, w.labels) and r.n_iter == w.n_iter and torch.equal(r.counts, w.counts), r.k
            assert abs(r.inertia - w.inertia) <= 1e-9 * r.inertia
        wide_chdb = eng.ch_db_scores_multi(ctx, [w.labels for w in wres], ks, [w.counts for w in wres], [w.csums for w in wres])
        assert np.allclose(np.asarray(wide_chdb), np.asarray([e[2:] for e in exact]), rtol=1e-6)
        Xd = plant_dupes(X, frac=10, seed=1)[0]
        mins, inds, idx = eng.dedup(Xd, 0.2)
        dist_m, lab = eng.classify(X[:300], X[300:310], 0.5)
        if rank == 0:
            out.put(dict(labels=[r.labels.numpy() for r in res], inertia=[r.inertia for r in res],
                         n_iter=[r.n_iter for r in res], seeds=[r.seeds for r in res], scores=exact, tc=tc,
                         idx=idx, mins=mins, lab=lab))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_match_single_process_and_oracle():
    from oracle import reference_calls as RC
    from videotofaces_b200.synth import blobs, plant_dupes
    ctxm = mp.get_context('spawn')
    out = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = blobs(700, 128, kc=5, noise=1.0, seed=1)
    ks = [2, 3, 4]
    ref = RC.cluster_core(X, ks, 0)
    for j, k in enumerate(ks):
        assert np.array_equal(got['seeds'][j], RC.kmeanspp_indices(X, k, 0))
        assert np.array_equal(got['labels'][j], ref['labels'][j])
        assert got['n_iter'][j] == ref['n_iter'][j]
        np.testing.assert_allclose(got['inertia'][j], ref['inertia'][j], rtol=1e-4)
        np.testing.assert_allclose(got['scores'][j][1:], ref['scores'][j][1:], rtol=1e-4)
        np.testing.assert_allclose(got['tc'][j], ref['scores'][j][1], rtol=1e-4)
    Xd = plant_dupes(X, frac=10, seed=1)[0]
    rmins, _, ridx = RC.dedup_core(Xd, 0.2)
    assert np.array_equal(got['idx'], ridx)
    np.testing.assert_allclose(got['mins'], rmins, rtol=1e-4, atol=3e-7)
    _, rlab = RC.classify_core(X[:300], X[300:310], 10, 0.5)
    assert np.array_equal(got['lab'], rlab)


def _dropin_worker(rank, world, port, root, out):
    """Every rank calls the drop-in entry points on the SAME directory (what `torchrun -m videotofaces_b200` does):
    rank 0 owns files, logs and console; the others only compute their shards and wait at the closing barrier."""
    import contextlib
    import io
    for p in (os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'video-to-faces_b200'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fake_kernels import FakeKernels
        from videotofaces_b200 import dupes, grouping, runtime
        from videotofaces_b200.engine import GroupingEngine
        from videotofaces_b200.dist import Comm
        from videotofaces_b200.main import video_to_faces
        from videotofaces_b200.synth import blobs, classify_faces_case, plant_dupes
        runtime.reset_engine(GroupingEngine(FakeKernels(), Comm()))
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            X = plant_dupes(blobs(400, 128, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0]
            a = os.path.join(root, 'a')
            paths = sorted(os.path.join(a, 'faces', f) for f in os.listdir(os.path.join(a, 'faces')))
            Xk, goods = dupes.remove_dupes_overall(X, paths, ('enc', 0.2, True, a))
            grouping.cluster_faces(list(goods), Xk, ([2, 3, 4], True, 0, True, a))
            Xc, R, names = classify_faces_case()
            b = os.path.join(root, 'b')
            video_to_faces(mode='grouping', style='live', out_dir=b, group_mode='classification', ref_dir=os.path.join(b, 'ref'),
                           enc_dup_thr=0.2, enc_oth_thr=0.9, save_dupes=True, features=plant_dupes(Xc, frac=10, seed=5)[0],
                           ref_features=R)
            # fewer rows than ranks * 64: computed un-sharded on every rank (no empty shards)
            tiny = grouping.cluster_labels(blobs(9, 128, kc=2, noise=0.5, seed=0), [2, 3], 0)
        out.put((rank, buf.getvalue(), [l.tolist() for l in tiny['labels']]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_drop_in_entry_points_from_two_ranks(tmp_path, golden_dir):
    """cluster_faces / classify_faces / remove_dupes_overall called from several ranks must not race on the
    file system.  The tree and the console of rank 0 equal the single-process (= reference) result; rank 1 prints nothing."""
    import dropin_cases as DC
    from oracle import reference_calls as RC
    from videotofaces_b200.synth import blobs, classify_faces_case, plant_dupes
    X = plant_dupes(blobs(400, 128, kc=4, noise=1.0, seed=0), frac=10, seed=1)[0]
    DC.fake_faces(str(tmp_path / 'a'), 400)
    Xc, R, names = classify_faces_case()
    DC.fake_faces(str(tmp_path / 'b'), 300)
    for c in names + ['empty_class']:
        os.makedirs(tmp_path / 'b' / 'ref' / c)
    for c in names:
        (tmp_path / 'b' / 'ref' / c / 'a.jpg').write_bytes(b'x')
    ctxm = mp.get_context('spawn')
    out = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_dropin_worker, args=(r, 2, port, str(tmp_path), out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (txt, tiny)) for r, txt, tiny in (out.get(timeout=240) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1][0] == ''                                                    # only rank 0 talks
    # directory a: dedup + clustering, against the oracle
    _, _, ridx = RC.dedup_core(X, 0.2)
    ref = RC.cluster_core(np.delete(X, ridx, axis=0), [2, 3, 4], 0)
    kept = [i for i in range(400) if i not in set(ridx.tolist())]
    assert sorted(os.listdir(tmp_path / 'a' / 'intermediate' / 'dupes3')) == ['f%06d.jpg' % i for i in ridx]
    for k, lab in zip(ref['clusters'], ref['labels']):
        for j in range(k):
            assert sorted(os.listdir(tmp_path / 'a' / 'faces' / ('G%u' % k) / str(j))) == [
                'f%06d.jpg' % i for i, l in zip(kept, lab) if l == j]
    assert not [f for f in os.listdir(tmp_path / 'a' / 'faces') if f.endswith('.jpg')]     # originals removed once
    assert 'Removed %u near-duplicates' % ridx.size in got[0][0]
    # directory b: the classification-mode golden of the real reference
    want = DC.manifest(golden_dir)['classification_mode_300x512']
    assert DC.tree(str(tmp_path / 'b')) == want['tree']
    assert got[0][0].replace(str(tmp_path / 'b'), '<OUT>').endswith(want['stdout'])
    assert got[0][1] == got[1][1]                                             # the un-sharded tiny case: same labels everywhere
