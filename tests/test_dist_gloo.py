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
