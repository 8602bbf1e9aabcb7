"""Host -> device upload of a pageable matrix: torch's pageable copy vs vtf_upload_pageable (native threads + pinned double
buffers) vs the same matrix in pinned memory.   python tools/upload_bench.py [--mb 2048]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--mb', type=int, default=2048)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    K = get_engine().K
    n = args.mb * (1 << 20) // 4
    A = np.random.default_rng(0).random(n, dtype=np.float32).reshape(-1, 512)
    out = {}

    def timed(fn, reps=3):
        ts = []
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return min(ts[1:]), r
    dt, ref = timed(lambda: torch.from_numpy(A).to('cuda'))
    out['torch_pageable_GBps'] = A.nbytes / dt / 1e9
    for th in (1, 2, 4, 8, 16):
        dt, got = timed(lambda: K.upload_pageable(A, torch.float32, th))
        out['native_%d_threads_GBps' % th] = A.nbytes / dt / 1e9
        assert torch.equal(got, ref)
    P = torch.from_numpy(A).pin_memory()
    dt, _ = timed(lambda: P.to('cuda', non_blocking=True))
    out['torch_pinned_GBps'] = A.nbytes / dt / 1e9
    out['mb'] = args.mb
    print(json.dumps(out))


if __name__ == '__main__':
    main()
