"""Cross-check of the sweep's event-bracketed time: CUDA events around the launch vs the kernel's own %globaltimer stamps
(first CTA entry -> last CTA exit, VTF_SWEEP_TIMING=1) for the symmetric sweep of config 3.

    VTF_SWEEP_TIMING=1 python tools/sweep_timing.py [--rows 1000000] [--reps 3]
"""
import argparse
import ctypes
import json
import os
import os.path as osp
import sys

os.environ.setdefault('VTF_SWEEP_TIMING', '1')
ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--reps', type=int, default=3)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200 import _native as nv
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200.synth import blobs
    eng = get_engine()
    eng.profile = True
    X = torch.from_numpy(blobs(args.rows, args.dim, kc=8, noise=1.0, seed=0))
    ctx, results = eng.kmeans_sweep(X, [2, 3, 4, 5, 6], 0)
    out = []
    for r in range(args.reps + 1):
        eng.events = []
        eng.cluster_scores(ctx, results)
        torch.cuda.synchronize()
        ev = {n: a.elapsed_time(b) for n, a, b, _ in eng.events}
        t = (ctypes.c_uint64 * 2)()
        nv.call('vtf_debug_sweep_timing', t)
        rec = dict(events_ms=ev.get('sil_sweep'), kernel_globaltimer_ms=(t[1] - t[0]) / 1e6)
        if r:
            out.append(rec)
            print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
