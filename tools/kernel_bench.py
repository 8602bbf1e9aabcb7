"""Per-kernel timings (CUDA events, after warm-up) against the HBM / tensor rooflines of MEASURED_PEAKS.json.

    python tools/kernel_bench.py [--rows 1000000] [--dim 512] > gpurun_out/kernels.json
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


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--k', type=int, default=6)
    ap.add_argument('--only', default='', help='comma list of substrings: run only matching kernels (for ncu)')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--warm', type=int, default=2)
    args = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    eng = get_engine()
    K = eng.K
    N, d, k = args.rows, args.dim, args.k
    try:
        pk = json.load(open(osp.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pk = dict(hbm_gbs=6650.0, bf16_tflops_sustained=1400.0)
    X = torch.randn(N, d, device='cuda')
    X = X / X.norm(dim=1, keepdim=True)
    mean = X.mean(0).contiguous()
    C = X[torch.randint(0, N, (k,), device='cuda')].contiguous()
    labels = torch.randint(0, k, (N,), device='cuda', dtype=torch.int32)
    counts = torch.zeros(k, dtype=torch.int64, device='cuda')
    nch = torch.zeros(1, dtype=torch.int64, device='cuda')
    closest = torch.rand(N, device='cuda')
    thr = torch.tensor([0.25, 0.5, 0.75], device='cuda') * float(closest.sum().item())
    C64 = C.double()
    R = X[:32].contiguous()
    rows = []

    only = [o for o in args.only.split(',') if o]

    def add(name, fn, nbytes=None, flops=None, reps=None, warm=None):
        if only and not any(o in name for o in only):
            return
        ms = timeit(fn, reps=reps or args.reps, warm=args.warm if warm is None else warm)
        r = dict(kernel=name, ms=ms)
        if nbytes:
            r.update(gbs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / pk['hbm_gbs'], bytes=nbytes)
        if flops:
            r.update(tflops=flops / ms / 1e9, flop=flops)
        rows.append(r)

    B = N * d * 4
    add('normalize_rows', lambda: K.normalize_rows(X), 2 * B)
    add('row_sqnorms', lambda: K.row_sqnorms(X), B)
    add('col_stats', lambda: K.col_stats(X), B)
    add('center_rows', lambda: K.center_rows(X, mean), 2 * B)
    add('kpp_dist T=3', lambda: K.kpp_dist(X, X[:3].contiguous(), closest), B + 3 * N * 4 + N * 4)
    xsq = K.kpp_rownorm64(X)
    add('kpp_rownorm64 (once per fit)', lambda: K.kpp_rownorm64(X), B + N * 8)
    add('kpp_dist T=3 with precomputed |x|^2', lambda: K.kpp_dist(X, X[:3].contiguous(), closest, xsq), B + 3 * N * 4 + N * 12)
    add('kpp_dist T=2 with precomputed |x|^2', lambda: K.kpp_dist(X, X[:2].contiguous(), closest, xsq), B + 2 * N * 4 + N * 12)
    add('kpp_pick (exact block scan)', lambda: K.kpp_pick(closest, thr, mode='scan'), N * 4)
    add('kpp_pick fast (tile skipping)', lambda: K.kpp_pick(closest, thr, mode='fast'), 3 * N * 4)
    add('kpp_pick_serial (one-thread chain)', lambda: K.kpp_pick(closest, thr, serial=True), N * 4)
    add('lloyd_assign k=%d' % k, lambda: K.lloyd_assign(X, C, labels, counts, nch), B + N * 8, 2.0 * N * k * d)
    add('cluster_sums k=%d' % k, lambda: K.cluster_sums(X, labels, k), B + N * 4)
    add('cluster_dispersion k=%d' % k, lambda: K.cluster_dispersion(X, labels, C64), B + N * 4)
    ks = [2, 3, 4, 5, 6]
    off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    C_all = X[torch.randint(0, N, (int(off[-1]),), device='cuda')].contiguous()
    labels_all = torch.stack([torch.randint(0, kk, (N,), device='cuda', dtype=torch.int32) for kk in ks]).contiguous()
    cnt_all = torch.zeros(int(off[-1]), dtype=torch.int64, device='cuda')
    ch_all = torch.zeros(len(ks), dtype=torch.int64, device='cuda')
    add('lloyd_assign_multi k=2..6', lambda: K.lloyd_assign_multi(X, C_all, off, 31, labels_all, 0, cnt_all, ch_all), B + 5 * N * 8, 2.0 * N * 20 * d)
    add('cluster_sums_multi k=2..6', lambda: K.cluster_sums_multi(X, labels_all, 0, off, 31), B + 5 * N * 4)
    acc = torch.zeros((int(off[-1]), d), dtype=torch.int64, device='cuda')
    # realistic labellings: nested blobs give 8-20 occupied cells, random labels 720 -- time both
    add('cell_sums k=2..6 (random labels: 720 cells)', lambda: K.cell_sums(X, labels_all, 0, off, 31, 2.0 ** 30, acc), B + 5 * N * 4)
    proj = (X[:, :5] > 0).to(torch.int32)
    labels_blob = torch.stack([(proj[:, :3] * torch.tensor([1, 2, 4], device='cuda', dtype=torch.int32)).sum(1) % kk for kk in ks]).to(torch.int32).contiguous()
    add('cell_sums k=2..6 (structured labels: <= 8 cells)', lambda: K.cell_sums(X, labels_blob, 0, off, 31, 2.0 ** 30, acc), B + 5 * N * 4)
    add('cell_sums k=6 alone', lambda: K.cell_sums(X, labels_all, 0, off, 16, 2.0 ** 30, acc), B + N * 4)
    if d % 128 == 0 and d <= 1024:
        C64a = C_all.double()
        add('cell_dispersion k=2..6 (structured labels)', lambda: K.cell_dispersion(X, labels_blob, 0, off, C64a, 2.0 ** 30, 2.0 ** 30), B + 5 * N * 4)
        add('cell_dispersion k=2..6 (random labels: 720 cells)', lambda: K.cell_dispersion(X, labels_all, 0, off, C64a, 2.0 ** 30, 2.0 ** 30), B + 5 * N * 4)
    add('cluster_dispersion_multi k=2..6', lambda: K.cluster_dispersion_multi(X, labels_all, 0, off, 31, C_all.double()), B + 5 * N * 4)
    # the same E-step as one tcgen05 contraction (csrc/lloyd_tc.cu): centre pack + split-fp16 assign + exact refine of near ties
    class _Ctx:
        pass
    wctx = _Ctx()
    wctx.Xc, wctx.N, wctx.d, wctx.b, wctx.e, wctx.wide = X, N, d, 0, N, None
    if d % 64 == 0:
        lab_new = torch.zeros((len(ks), N), dtype=torch.int32, device='cuda')
        n_fix = torch.zeros(1, dtype=torch.int32, device='cuda')
        cache = {}
        add('wide E-step k=2..6 (tcgen05: pack + assign + refine)',
            lambda: eng._wide_estep(wctx, C_all, off, ks, 31, lab_new, n_fix, cache), B + 5 * N * 4, 2.0 * N * 20 * d)
    add('cosine_classify R=32', lambda: K.cosine_classify(X, R, 0.9), B + N * 32 * 4 + N * 8, 2.0 * N * 32 * d)
    nd = min(N, 50_000)
    Xd = X[:nd].contiguous()
    add('cosine_dedup N=%d (fp32 FMA)' % nd, lambda: K.cosine_dedup(Xd, 0, 1), None, float(nd) * nd * d)
    add('cosine_dedup_tc N=%d (tcgen05, split fp16 + refine)' % nd, lambda: K.cosine_dedup_tc(Xd, 0, 1), None, float(nd) * nd * d)
    perm = torch.arange(N, device='cuda')
    P = ((N + 255) // 256) * 256
    perm = torch.cat([perm, torch.full((P - N,), -1, device='cuda', dtype=torch.int64)])
    add('sil_pack', lambda: K.sil_pack(X, perm, P, 1024.0, 6, 4.0 * 0.5 * 1024.0 ** 2 * 1.02), B + N * d * 2)
    json.dump(dict(rows=args.rows, dim=d, peaks=dict(hbm_gbs=pk['hbm_gbs'], bf16_tflops_sustained=pk.get('bf16_tflops_sustained')),
                   kernels=rows), sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main()
