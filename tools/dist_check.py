"""Multi-GPU parity check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank runs the sharded engine on the same seeded inputs; rank 0 compares with the scikit-learn
oracle (labels identical given the device-chosen seeds, inertia / scores within 1e-4, exact dedup set,
identical classification labels) and prints one JSON line."""
import json
import os
import os.path as osp
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np          # noqa: E402
import torch                # noqa: E402
import torch.distributed as dist   # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200.synth import blobs, plant_dupes
    from oracle import reference_calls as RC
    eng = get_engine()
    rank, world = eng.comm.rank, eng.comm.world
    report = dict(world=world, ok=True, checks=[])

    def check(name, cond, detail=None):
        report['checks'].append(dict(name=name, ok=bool(cond), detail=detail))
        report['ok'] = report['ok'] and bool(cond)

    cases = [(8193, 512, [2, 3, 4], True, None), (6000, 768, [2, 5], False, None), (3001, 512, [2, 3], True, None),
             (20000, 512, list(range(2, 13)), True, True)]      # last: the wide (tensor-core) Lloyd path, forced
    for (N, d, ks, normalize, wide) in cases:
        X = blobs(N, d, kc=6, noise=1.0, seed=1, normalize=normalize)
        eng.wide = wide
        ctx, res = eng.kmeans_sweep(X, ks, 0)
        eng.wide = None
        scores = eng.cluster_scores(ctx, res)
        if rank == 0:
            ref = RC.cluster_core(X, ks, 0, inits=[X[r.seeds] for r in res])
            for r, lab, inertia, n_iter, sc, rsc in zip(res, ref['labels'], ref['inertia'], ref['n_iter'], scores, ref['scores']):
                tag = 'N=%d k=%d%s' % (N, r.k, ' wide' if wide else '')
                if r.k <= 6:
                    agree = float((r.labels.cpu().numpy() == lab).mean())
                    check('labels ' + tag, agree >= 0.9999, agree)
                    check('n_iter ' + tag, r.n_iter == n_iter, (r.n_iter, n_iter))
                    check('inertia ' + tag, abs(r.inertia - inertia) <= 1e-4 * inertia)
                    check('scores ' + tag, np.allclose(sc[1:], rsc[1:], rtol=1e-4), (sc, rsc))
                else:
                    # more clusters than blobs: the split of a Gaussian blob is chaotic in the last ulp of a distance (any two
                    # fp32 summation orders part ways), so only the quality of the optimum is comparable
                    check('inertia ' + tag, abs(r.inertia - inertia) <= 2e-3 * inertia, (r.inertia, inertia))
            seeds_ok = [bool(np.array_equal(r.seeds, RC.kmeanspp_indices(X, r.k, 0))) for r in res]
            report.setdefault('seed_agreement', []).append(seeds_ok)
    X = plant_dupes(blobs(5000, 512, kc=4, seed=5), frac=20, seed=1)[0]
    mins, inds, idx = eng.dedup(X, 0.2)
    dist_m, lab = eng.classify(X[:3000], X[4000:4032], 0.55)
    X2 = plant_dupes(blobs(12000, 512, kc=4, seed=6), frac=20, seed=2)[0]
    mins2, inds2, idx2 = eng.dedup(X2, 0.2, force='tc')          # two-stage tensor-core dedup, units dealt over the ranks
    if rank == 0:
        rmins, rinds, ridx = RC.dedup_core(X, 0.2)
        check('dedup set', np.array_equal(idx, ridx))
        check('dedup mins', np.allclose(mins, rmins, rtol=1e-4, atol=3e-7))
        rmins2, rinds2, ridx2 = RC.dedup_core(X2, 0.2)
        check('tc dedup set', np.array_equal(idx2, ridx2))
        check('tc dedup mins', np.allclose(mins2, rmins2, rtol=1e-4, atol=3e-7))
        _, rlab = RC.classify_core(X[:3000], X[4000:4032], 32, 0.55)
        check('classify labels', np.array_equal(lab, rlab))
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not report['ok']:
        sys.exit(1)


if __name__ == '__main__':
    main()
