"""BASELINE.json configs 3 / 4 / 5 at full size on 1..8 GPUs (one rank per GPU under torchrun, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/config_dist.py c5 [--scale 1.0] [--reps 1]

The matrix is generated ON THE DEVICE (same Philox seed on every rank, so every rank holds the same bytes; 4M x 768
would otherwise need 12 GB of host memory per rank) in the shape of synth.blobs / synth.plant_dupes.  Timed region:
X resident in HBM -> labels, scores and the dedup index set on the host; wall clock between barriers, max over ranks.
Checks that do not depend on the size: planted duplicates found, inertia decreasing in k, silhouette in [-1, 1],
every rank holds the same labels.  Rank 0 prints one JSON line per config.
"""
import argparse
import json
import os
import os.path as osp
import sys
import time

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    sys.path.insert(0, p)

import torch                        # noqa: E402
import torch.distributed as dist    # noqa: E402


def device_blobs(N, d, kc, noise, seed, normalize, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    C = torch.randn((kc, d), generator=g, device=dev)
    ids = torch.randint(0, kc, (N,), generator=g, device=dev)
    X = C[ids]
    step = 1 << 18
    for s in range(0, N, step):
        e = min(N, s + step)
        X[s:e] += noise * torch.randn((e - s, d), generator=g, device=dev)
    if normalize:
        X /= X.norm(dim=1, keepdim=True).clamp_min(1e-30)
    return X.contiguous()


def device_plant_dupes(X, frac, seed):
    """synth.plant_dupes on the device: m = N/frac targets in the upper half copy a source of the lower half plus noise
    that puts the cosine distance around the 0.2 threshold (about 0.04 .. 0.33)."""
    N, d = X.shape
    g = torch.Generator(device=X.device)
    g.manual_seed(seed)
    m, half = N // frac, N // 2
    src = torch.randperm(half, generator=g, device=X.device)[:m]
    dst = half + torch.randperm(N - half, generator=g, device=X.device)[:m]
    eps = 0.3 + 0.8 * torch.rand((m, 1), generator=g, device=X.device)
    Y = X[src]
    Y = Y + Y.norm(dim=1, keepdim=True) * eps * torch.randn((m, d), generator=g, device=X.device) / (d ** 0.5)
    X[dst] = Y
    cos = 1.0 - torch.nn.functional.cosine_similarity(X[src], X[dst], dim=1)
    return src, dst, cos


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('configs', nargs='+')
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--reps', type=int, default=1)
    args = ap.parse_args()
    multi = 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if multi:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200.runtime import get_engine
    eng = get_engine()
    rank, world = eng.comm.rank, eng.comm.world
    dev = eng.device

    def sync():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()

    def timed(fn):
        sync()
        t = time.perf_counter()
        out = fn()
        sync()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return out, float(dt.item())

    for name in args.configs:
        if name == 'c4':
            N, d, ks, kc, norm, dedup = int(1_000_000 * args.scale), 1024, list(range(2, 33)), 16, False, False
        elif name == 'c5':
            N, d, ks, kc, norm, dedup = int(4_000_000 * args.scale), 768, [2, 3, 4, 5, 6], 8, False, True
        else:
            N, d, ks, kc, norm, dedup = int(1_000_000 * args.scale), 512, [2, 3, 4, 5, 6], 8, True, False
        X = device_blobs(N, d, kc, 1.0, 0, norm, dev)
        out = dict(config=name, rows=N, dim=d, ks=[ks[0], ks[-1]], n_gpus=world, data='synthetic, generated on the device')
        if dedup:
            src, dst, cos = device_plant_dupes(X, 50, 1)
            # warm-up on a slice of the same width (module load, attributes, tensor maps, NCCL channels)
            eng.dedup(X[:min(N, 65536)].contiguous(), 0.2)
            (mins, inds, idx), t_dedup = timed(lambda: eng.dedup(X, 0.2))
            dst_h, cos_h = dst.cpu().numpy(), cos.cpu().numpy()
            import numpy as np
            found = np.isin(dst_h, idx)
            clear_dup, clear_not = cos_h < 0.19, cos_h > 0.21
            out.update(dedup_s=t_dedup, dedup_tflops=float(N) * N * d / t_dedup / 1e12, n_dupes=int(idx.size), planted=int(dst_h.size),
                       planted_below_thr_found=bool(found[clear_dup].all()), planted_above_thr_kept=bool((~found[clear_not]).all()),
                       other_rows_marked=int(idx.size - found.sum()))
            X2, t_del = timed(lambda: eng.delete_rows(X, idx))
            out.update(delete_s=t_del, rows_after_dedup=int(X2.shape[0]))
            del X
            X = X2
        # warm-up of the clustering path on a slice
        wctx, wres = eng.kmeans_sweep(X[:min(X.shape[0], 32768)].contiguous(), ks[:3], 0)
        eng.cluster_scores(wctx, wres)
        del wctx, wres
        best = None
        for _ in range(args.reps):
            eng.profile, eng.events = True, []
            (ctx, res), t_km = timed(lambda: eng.kmeans_sweep(X, ks, 0))
            scores, t_sc = timed(lambda: eng.cluster_scores(ctx, res))
            eng.profile = False
            phases = {}
            for nm, a, b, _f in eng.events:
                phases[nm] = phases.get(nm, 0.0) + a.elapsed_time(b)
            cur = dict(kmeans_s=t_km, scores_s=t_sc, phases_ms=phases)
            if best is None or t_km + t_sc < best['kmeans_s'] + best['scores_s']:
                best = cur
        inert = [r.inertia for r in res]
        sil = [s[1] for s in scores]
        same = True
        if multi:
            h = torch.stack([r.labels.to(torch.int64).sum() for r in res])
            lo, hi = h.clone(), h.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo, hi))
        total = best['kmeans_s'] + best['scores_s'] + out.get('dedup_s', 0.0) + out.get('delete_s', 0.0)
        out.update(best)
        out.update(total_s=total, emb_per_s=N / total, n_iter=[r.n_iter for r in res], inertia_decreasing=bool(all(a > b for a, b in zip(inert, inert[1:]))),
                   silhouette=sil, silhouette_in_range=bool(all(-1.0 <= s <= 1.0 for s in sil)), best_k=max(scores, key=lambda s: s[1])[0],
                   labels_equal_on_all_ranks=same, wide=dict((k, v) for k, v in eng.wide_stats.items() if k in ('iters', 'rows', 'labellings', 'sum_k')),
                   mem_gb=torch.cuda.max_memory_allocated() / 1e9)
        if rank == 0:
            print(json.dumps(out))
            sys.stdout.flush()
        del X, ctx, res
        torch.cuda.empty_cache()
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
