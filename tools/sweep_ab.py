"""Same-box A/B of sweep-kernel builds: one process, one set of packed operands, every library variant timed in
turn (interleaved, so the power/thermal state of the box is shared).

    python tools/sweep_ab.py --rows 1000000 --reps 3 gpurun_ab/libvtf_a.so gpurun_ab/libvtf_b.so ...
"""
import argparse
import ctypes
import json
import os.path as osp
import subprocess
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def smi_clock():
    try:
        out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-i', '0'],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(',')
        return float(out[0]), float(out[1])
    except Exception:  # noqa: BLE001
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('libs', nargs='+')
    ap.add_argument('--rows', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--sym', action='store_true', help='time vtf_sil_sweep_sym (upper triangle) instead of the full sweep')
    ap.add_argument('--modes', default='static', help='--sym only: comma list of static,dynamic (work counter) per library')
    args = ap.parse_args()
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200._native import c_f32, c_i64, c_int, c_vp
    eng = get_engine()
    K = eng.K
    N, d = args.rows, args.dim
    g = torch.Generator(device='cuda').manual_seed(0)
    C = torch.randn(8, d, device='cuda', generator=g)
    X = C[torch.randint(0, 8, (N,), device='cuda', generator=g)] + torch.randn(N, d, device='cuda', generator=g)
    X = (X / X.norm(dim=1, keepdim=True)).contiguous()
    Xc = (X - X.mean(0)).contiguous()
    ks = [2, 3, 4, 5, 6]
    labels = [torch.randint(0, k, (N,), device='cuda', dtype=torch.int32, generator=g) if False else
              (torch.arange(N, device='cuda') % k).to(torch.int32) for k in ks]
    perm, chunk_cell, c2c, G, P = eng.build_cells(labels, ks)
    scale, shift, pad_m = eng.sil_scale(Xc)
    Xh, aug = K.sil_pack(Xc, perm, P, scale, shift, pad_m)
    U = P // 256
    flops = 2.0 * P * P * d if not args.sym else 2.0 * 256 * 256 * d * (U * (U + 1) / 2)
    libs = []
    for path in args.libs:
        lib = ctypes.CDLL(osp.abspath(path))
        lib.vtf_sil_sweep.restype = c_int
        lib.vtf_sil_sweep.argtypes = [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_int, c_i64, c_i64, c_vp, c_vp]
        if args.sym:
            lib.vtf_sil_sweep_sym.restype = c_int
            lib.vtf_sil_sweep_sym.argtypes = [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]
        for mode in (args.modes.split(',') if args.sym else ['static']):
            libs.append((osp.basename(path) + ':' + mode, lib))
    cellsum = torch.zeros((G, P), dtype=torch.float32, device='cuda')
    counter = torch.zeros(4, dtype=torch.int32, device='cuda')
    tiles = torch.zeros(1, dtype=torch.int64, device='cuda')
    stream = c_vp(torch.cuda.current_stream().cuda_stream)
    results = {n: [] for n, _ in libs}
    ref = None
    for rep in range(args.reps + 1):
        for name, lib in libs:
            cellsum.zero_()
            counter.zero_()
            tiles.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if args.sym:
                rc = lib.vtf_sil_sweep_sym(Xh.data_ptr(), aug.data_ptr(), P, d, scale, chunk_cell.data_ptr(), G, 0, 1,
                                           cellsum.data_ptr(), counter.data_ptr() if name.endswith(':dynamic') else None,
                                           tiles.data_ptr(), stream)
            else:
                rc = lib.vtf_sil_sweep(Xh.data_ptr(), aug.data_ptr(), P, d, scale, chunk_cell.data_ptr(), G, 0, P // 128,
                                       cellsum.data_ptr(), stream)
            b.record()
            torch.cuda.synchronize()
            assert rc == 0, name
            clk = smi_clock()
            chk = float(cellsum.double().sum().item())
            if ref is None:
                ref = chk
            if rep > 0:
                assert not args.sym or int(tiles.item()) == U * (U + 1) // 2, (name, int(tiles.item()), U * (U + 1) // 2)
                results[name].append(dict(ms=a.elapsed_time(b), tflops=flops / a.elapsed_time(b) / 1e9, clk_after=clk,
                                          checksum_rel=abs(chk - ref) / abs(ref)))
    for name, r in results.items():
        ms = [x['ms'] for x in r]
        print(json.dumps(dict(lib=name, ms_median=float(np.median(ms)), ms=ms, tflops_median=flops / float(np.median(ms)) / 1e9,
                              checksum_rel_max=max(x['checksum_rel'] for x in r), clk=[x['clk_after'] for x in r])))


if __name__ == '__main__':
    main()
