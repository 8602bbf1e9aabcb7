"""Run the other BASELINE.json configs (parity-test cases, not bench lines) on one GPU and print one JSON
line per config: timings by phase, shapes, and cheap size-independent checks.

    python tools/config_runs.py c2 c4 c5 [--scale 1.0]
"""
import argparse
import json
import os.path as osp
import sys
import time

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='+')
    ap.add_argument('--scale', type=float, default=1.0)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200.synth import blobs, plant_dupes
    from oracle import restatement as RS
    eng = get_engine()
    eng.profile = True
    for name in args.configs:
        eng.events = []
        if name == 'c2':      # 50k x 512: dedup (0.2) + classify vs 32 references
            N = int(50_000 * args.scale)
            B = blobs(N + 32, 512, kc=8, noise=1.0, seed=0)
            X, src, dst = plant_dupes(np.ascontiguousarray(B[:N]), frac=50, seed=1)
            R = np.ascontiguousarray(B[N:])
            Xd = eng.to_device(X)
            eng.dedup(Xd, 0.2)   # warm: same shape (module load, attribute and tensor-map set-up)
            (mins, inds, idx), t_dedup = timed(lambda: eng.dedup(Xd, 0.2))
            (smins, _, sidx), t_simt = timed(lambda: eng.dedup(Xd, 0.2, force='simt'))
            (dist, lab), t_cls = timed(lambda: eng.classify(Xd, R, 0.55))
            sub = np.sort(np.random.default_rng(0).choice(N, 64, replace=False))
            D = RS.cosine_distances(X[sub], X)
            ok_rows = all(abs(D[r, :i].min() - mins[i]) <= 1e-6 for r, i in enumerate(sub) if i > 0)
            Dr = RS.cosine_distances(X[sub], R)
            ok_cls = bool(np.array_equal(np.where(Dr.min(1) >= np.float32(0.55), 32, Dr.argmin(1)), lab[sub]))
            print(json.dumps(dict(config='c2', rows=N, dim=512, dedup_s=t_dedup, classify_s=t_cls, n_dupes=int(idx.size),
                                  emb_per_s=N / (t_dedup + t_cls), dedup_rows_match_oracle=bool(ok_rows),
                                  classify_rows_match_oracle=ok_cls, dedup_tflops=N * N * 512.0 / t_dedup / 1e12, dedup_simt_s=t_simt,
                                  tc_equals_simt_set=bool(np.array_equal(idx, sidx)), max_abs_min_diff=float(np.abs(mins - smins).max()))))
        elif name in ('c4', 'c5', 'c3'):
            if name == 'c4':   # 1M x 1024, k = 2..32, clusters-save-all
                N, d, ks, kc, norm = int(1_000_000 * args.scale), 1024, list(range(2, 33)), 16, False
            elif name == 'c5':  # 4M x 768 (per-GPU share at 8 GPUs: run scaled), dedup + silhouette k = 2..6
                N, d, ks, kc, norm = int(4_000_000 * args.scale), 768, [2, 3, 4, 5, 6], 8, False
            else:
                N, d, ks, kc, norm = int(1_000_000 * args.scale), 512, [2, 3, 4, 5, 6], 8, True
            X = blobs(N, d, kc=kc, noise=1.0, seed=0, normalize=norm)
            Xd = eng.to_device(X)
            out = dict(config=name, rows=N, dim=d, ks=[ks[0], ks[-1]])
            if name == 'c5':
                nd = min(N, 200_000)
                (_, _, idx), t_dedup = timed(lambda: eng.dedup(Xd[:nd].contiguous(), 0.2))
                out.update(dedup_rows=nd, dedup_s=t_dedup, dedup_tflops=nd * nd * float(d) / t_dedup / 1e12)
            (ctx, res), t_km = timed(lambda: eng.kmeans_sweep(Xd, ks, 0))
            scores, t_sc = timed(lambda: eng.cluster_scores(ctx, res))
            phases = {}
            for nm, a, b, _ in eng.events:
                phases[nm] = phases.get(nm, 0.0) + a.elapsed_time(b)
            sweeps = [(a.elapsed_time(b), f) for nm, a, b, f in eng.events if nm == 'sil_sweep']
            inert = [r.inertia for r in res]
            out.update(kmeans_s=t_km, scores_s=t_sc, emb_per_s=N / (t_km + t_sc), n_iter=[r.n_iter for r in res],
                       inertia_decreasing=bool(all(a > b for a, b in zip(inert, inert[1:]))),
                       best_k=max(scores, key=lambda s: s[1])[0], silhouette=[s[1] for s in scores], phases_ms=phases,
                       sweep_tflops=[(f() if callable(f) else f) / (t * 1e-3) / 1e12 for t, f in sweeps], n_sweeps=len(sweeps),
                       mem_gb=torch.cuda.max_memory_allocated() / 1e9)
            print(json.dumps(out))
        sys.stdout.flush()


if __name__ == '__main__':
    main()
