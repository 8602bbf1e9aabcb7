"""Per-kernel device-time totals of one config (CUPTI through torch.profiler: no replay, no serialisation).

    python tools/kernel_totals.py c4 --scale 0.2 [--top 30]

Prints one JSON line: wall time of the warm second run, phase spans, and the kernels sorted by total device time.
"""
import argparse
import json
import os.path as osp
import sys
import time

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np   # noqa: E402,F401
import torch         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('config')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--top', type=int, default=30)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200.synth import blobs
    eng = get_engine()
    if args.config == 'c4':
        N, d, ks, kc, norm = int(1_000_000 * args.scale), 1024, list(range(2, 33)), 16, False
    elif args.config == 'c5':
        N, d, ks, kc, norm = int(4_000_000 * args.scale), 768, [2, 3, 4, 5, 6], 8, False
    else:
        N, d, ks, kc, norm = int(1_000_000 * args.scale), 512, [2, 3, 4, 5, 6], 8, True
    X = blobs(N, d, kc=kc, noise=1.0, seed=0, normalize=norm)
    Xd = eng.to_device(X)

    def step():
        ctx, res = eng.kmeans_sweep(Xd, ks, 0)
        return res, eng.cluster_scores(ctx, res)

    step()                                   # warm: module load, attributes, tensor maps
    eng.profile = True
    eng.events = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, scores = step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    phases = {}
    for nm, a, b, _ in eng.events:
        phases[nm] = phases.get(nm, 0.0) + a.elapsed_time(b)
    eng.profile = False
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, 'device_time_total', None)
        if t is None:
            t = getattr(ev, 'cuda_time_total', 0.0)
        if t > 0:
            rows.append((ev.key[:90], int(ev.count), t / 1e3))
    rows.sort(key=lambda r: -r[2])
    # GPU idle time between consecutive device activities (kernels, copies, memsets) of the profiled step
    acts = []
    for ev in prof.events():
        tr = getattr(ev, 'time_range', None)
        if tr is not None and str(getattr(ev, 'device_type', '')).endswith('CUDA'):
            acts.append((tr.start, tr.end, ev.name[:70]))
    acts.sort()
    gaps, idle, last_end, last_name = [], 0.0, None, None
    for a, b, nm in acts:
        if last_end is not None and a > last_end:
            idle += a - last_end
            gaps.append((a - last_end, last_name, nm))
        if last_end is None or b > last_end:
            last_end, last_name = b, nm
    gaps.sort(key=lambda g: -g[0])
    print(json.dumps(dict(config=args.config, rows=N, dim=d, ks=[ks[0], ks[-1]], wall_ms=wall * 1e3, phases_ms=phases,
                          n_iter=[r.n_iter for r in res], launches=sum(r[1] for r in rows),
                          device_ms_total=sum(r[2] for r in rows), idle_ms=idle / 1e3, n_gaps=len(gaps),
                          gaps_over_20us=sum(1 for g in gaps if g[0] > 20), idle_in_gaps_over_20us_ms=sum(g[0] for g in gaps if g[0] > 20) / 1e3,
                          top_gaps=[dict(us=round(g[0], 1), after=g[1], before=g[2]) for g in gaps[:25]],
                          kernels=[dict(name=n, count=c, ms=round(t, 3)) for n, c, t in rows[:args.top]])))


if __name__ == '__main__':
    main()
