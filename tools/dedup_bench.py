"""Two-stage dedup at several sizes: time per call (CUDA events), executed TFLOP/s (N^2*d), screen launches, re-swept units.

    python tools/dedup_bench.py --sizes 50000x512,500000x768 [--check]

--check compares the keys with the split-fp16 sweep (the round-1 path) on the same rows.
VTF_SCREEN_CHUNK=<tiles> overrides the column-chunk length (a huge value = one launch, the un-chunked order).
"""
import argparse
import json
import os.path as osp
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', default='50000x512')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--check', action='store_true')
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    K = get_engine().K
    out = []
    for sz in args.sizes.split(','):
        N, d = (int(v) for v in sz.split('x'))
        g = torch.Generator(device='cuda').manual_seed(1)
        C = torch.randn(8, d, device='cuda', generator=g)
        X = C[torch.randint(0, 8, (N,), device='cuda', generator=g)] + torch.randn(N, d, device='cuda', generator=g)
        Xn = K.normalize_rows(X.contiguous())
        del X
        ts = []
        for r in range(args.reps + 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); key = K.cosine_dedup_tc(Xn, 0, 1, 0.2); b.record()
            torch.cuda.synchronize()
            if r:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        rec = dict(rows=N, dim=d, ms=ms, tflops=float(N) * N * d / ms / 1e9, stats=dict(K.dedup_stats))
        if args.check:
            ref = K.cosine_dedup_tc(Xn, 0, 1, None, mode='split')
            kh, rh = key.cpu().numpy(), ref.cpu().numpy()
            same_idx = float(np.mean((kh & 0xffffffff)[1:] == (rh & 0xffffffff)[1:]))
            dk = (kh >> 32).astype(np.uint32).view(np.float32)[1:]
            dr = (rh >> 32).astype(np.uint32).view(np.float32)[1:]
            rec['check'] = dict(argmin_agree=same_idx, max_abs_dist_diff=float(np.abs(dk - dr).max()),
                                dup_set_equal=bool(np.array_equal(dk <= 0.2, dr <= 0.2)))
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del Xn, key


if __name__ == '__main__':
    main()
