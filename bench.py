"""bench.py -- the BASELINE.json metric: embeddings/sec for the k=2..6 KMeans + silhouette sweep.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows 1000000] [--dim 512] [--impl reference]

A step = one pass of the hot path over the synthetic embedding matrix of BASELINE config 3
(`blobs(1_000_000, 512, kc=8, noise=1, seed=0)`): KMeans.fit prologue, k-means++ and Lloyd for
k = 2..6, silhouette + Calinski-Harabasz + Davies-Bouldin for every k, selection of the best k.

  value   X already resident in HBM, CUDA events around the step, max over ranks
  e2e     the public API call on a pinned HOST matrix: H2D of X and D2H of labels/scores inside the
          timed region (videotofaces_b200.grouping.cluster_labels)
  roofline  the tcgen05 silhouette sweep: the MMA flops it executes per launch (upper triangle of the Gram
          matrix: ~P^2*d, the distance matrix is symmetric) / its CUDA-event time, against the measured dense
          bf16 peak of MEASURED_PEAKS.json (kind::f16 runs at the bf16 rate)
  cpu_baseline  the reference's own sklearn call sequence (oracle/reference_calls.py) on a bounded
          row sample, timed on this box's host cores (rank 0, N=1 only)

--impl reference times that CPU path as the reference arm.  N > 1 (torchrun): strong scaling, the
1M rows are sharded over the ranks, each rank holds the full matrix.
"""
import argparse
import json
import os
import os.path as osp
import subprocess
import sys
import threading
import time

ROOT = osp.dirname(osp.abspath(__file__))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

KS = [2, 3, 4, 5, 6]
CPU_SAMPLE_ROWS = 16384


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--rows', type=int, default=1_000_000)
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-rows', type=int, default=CPU_SAMPLE_ROWS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def make_data(rows, dim):
    from videotofaces_b200.synth import blobs
    return blobs(rows, dim, kc=8 if dim != 1024 else 16, noise=1.0, seed=0, normalize=(dim == 512))


def peaks():
    try:
        with open(osp.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='MEASURED_PEAKS.json')
    except Exception:  # noqa: BLE001
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == 'Active' for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(self.rows[0][1]), reasons=reasons, samples=len(sm))


def cpu_reference_step(Xs):
    """The reference's CPU path on a row sample: grouping.py:99-120 over scikit-learn, all host threads."""
    from oracle import reference_calls as RC
    t0 = time.perf_counter()
    res = RC.cluster_core(Xs, KS, 0)
    return time.perf_counter() - t0, res


def cpu_info():
    import sklearn
    try:
        from threadpoolctl import threadpool_info
        tp = [{k: i.get(k) for k in ('user_api', 'internal_api', 'num_threads', 'version')} for i in threadpool_info()]
    except Exception:  # noqa: BLE001
        tp = None
    model = ''
    try:
        with open('/proc/cpuinfo') as f:
            model = [l.split(':', 1)[1].strip() for l in f if l.startswith('model name')][0]
    except Exception:  # noqa: BLE001
        pass
    return dict(cores=os.cpu_count(), cpu_model=model, sklearn=sklearn.__version__, numpy=np.__version__, threadpools=tp)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    Xs = make_data(args.cpu_rows, args.dim)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(Xs[:2048])
    times = []
    for _ in range(max(1, min(args.steps, 3))):
        dt, _ = cpu_reference_step(Xs)
        times.append(dt)
    t = float(np.mean(times))
    value = args.cpu_rows / t
    info = cpu_info()
    sample = ('sklearn KMeans(k=2..6)+silhouette+CH+DB on %d of the %d rows (same generator); the silhouette term is O(N^2), '
              'so the full-size CPU throughput is lower still (about %.1f emb/s extrapolated with t ~ N^2)'
              % (args.cpu_rows, args.rows, args.rows / (t * (args.rows / args.cpu_rows) ** 2)))
    line = dict(impl='reference', metric='embeddings/sec, k=2-6 KMeans+silhouette sweep', value=value, unit='embeddings/s',
                n_gpus=args.gpus, steps=len(times), warmup=min(args.warmup, 1), ms_per_step=t * 1e3, higher_is_better=True,
                scaling='strong', vs_baseline=None, dtype='f64 (sklearn upcast path)', data='synthetic',
                config=dict(workload='config3: 1Mx512 k=2..6 KMeans + silhouette/CH/DB sweep (CPU arm: %d-row sample)' % args.cpu_rows,
                            rows=args.cpu_rows, dim=args.dim, ks=KS),
                cpu_baseline=dict(value=value, unit='embeddings/s', cores=info['cores'], kind='port', sample=sample, info=info),
                e2e=dict(value=value, unit='embeddings/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner from C at
    any NCCL_DEBUG level >= VERSION, which includes WARN), so file descriptor 1 is pointed at stderr for the whole run
    and the JSON line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    args = parse()
    _claim_stdout()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200 import _native
    from videotofaces_b200.runtime import get_engine
    from videotofaces_b200.grouping import cluster_labels
    eng = get_engine()

    X = make_data(args.rows, args.dim)
    Xpin = torch.from_numpy(X).pin_memory()
    Xd = Xpin.cuda(non_blocking=True)
    torch.cuda.synchronize()
    N, d = X.shape

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ctx = eng.prepare(Xd)
        ctx, res = eng.kmeans_sweep(None, KS, 0, ctx=ctx)
        scores = eng.cluster_scores(ctx, res)
        return res, scores

    eng.profile = True
    for _ in range(args.warmup):
        res, scores = step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _native.launch_count()
    eng.events = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res, scores = step()
    ev1.record()
    barrier()
    launches = _native.launch_count() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # dominant kernel: the tcgen05 silhouette sweep, timed by CUDA events on its launch stream
    sweeps = [(a.elapsed_time(b), flops() if callable(flops) else flops) for (name, a, b, flops) in eng.events
              if name == 'sil_sweep']
    pk = peaks()
    roof = None
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the sweep from the committed ncu --set full capture
        with open(osp.join(ROOT, 'profiles', 'r01_sweep_ncu_summary.json')) as f:
            ncu = json.load(f)
        if ncu['rows'] == N and ncu['dim'] == d and world == 1:
            traffic = ncu['traffic_bytes_per_launch']
    except Exception:  # noqa: BLE001
        pass
    if sweeps:
        t_ms = float(np.mean([t for t, _ in sweeps]))
        flops = float(np.mean([f for _, f in sweeps]))
        ach = flops / (t_ms * 1e-3) / 1e12
        full_gram = 2.0 * float(((N + 255) // 256) * 256) ** 2 * d
        roof = dict(bound='tensor', kernel='sil_sweep_kernel (tcgen05 cta_group::2 kind::f16, fp32 accumulate in TMEM)', achieved=ach,
                    peak=pk['tf_sust'], unit='TFLOP/s', frac=ach / pk['tf_sust'], traffic=traffic, ms_per_launch=t_ms,
                    flop_per_launch=flops, share_of_step=t_ms * len(sweeps) / args.steps / ms_step,
                    note=('flop_per_launch counts the MMAs the kernel EXECUTES: D is symmetric, so only tiles on or above the diagonal '
                          'are computed (%.3f of the full Gram %.4g flop that SURVEY 8d lists); achieved/peak is the tensor-pipe figure'
                          % (flops * world / full_gram, full_gram)) if flops * world < 0.75 * full_gram else 'full Gram sweep',
                    peak_source=pk['source'] + ' bf16_tflops_sustained (kernel timed inside a long step)')
    phases = {}
    for name, a, b, _ in eng.events:
        phases[name] = phases.get(name, 0.0) + a.elapsed_time(b) / args.steps

    # end to end through the public API, host buffers
    e2e = None
    if not args.no_e2e:
        out = cluster_labels(Xpin.numpy(), KS, 0)       # warm
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            out = cluster_labels(Xpin.numpy(), KS, 0)
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / n_e2e], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e = dict(value=N / float(t_e2e.item()), unit='embeddings/s', h2d_bytes_per_step=int(X.nbytes),
                   d2h_bytes_per_step=int(sum(l.nbytes for l in out['labels']) + sum(c.nbytes for c in out['centers']) + 8 * 4 * len(KS)),
                   ms_per_step=float(t_e2e.item()) * 1e3, api='videotofaces_b200.grouping.cluster_labels(X_host, [2..6], 0)')

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Xs = np.ascontiguousarray(X[:args.cpu_rows])
        dt, ref = cpu_reference_step(Xs)
        info = cpu_info()
        cpu = dict(value=args.cpu_rows / dt, unit='embeddings/s', cores=info['cores'], kind='port',
                   sample='reference call sequence (grouping.py:99-120 over scikit-learn %s) on the first %d rows: %.1f s; '
                          'O(N^2) silhouette => about %.1f emb/s extrapolated to %d rows'
                          % (info['sklearn'], args.cpu_rows, dt, N / (dt * (N / args.cpu_rows) ** 2), N),
                   info=info)

    if rank == 0:
        line = dict(metric='embeddings/sec, k=2-6 KMeans+silhouette sweep', value=N / (ms_step * 1e-3), unit='embeddings/s',
                    n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True,
                    scaling='strong', vs_baseline=None, dtype='f16 operands / f32 accumulate (Gram), f32+f64 elsewhere',
                    data='synthetic',
                    config=dict(workload='config3: blobs(%d,%d,kc=8,noise=1,seed=0) KMeans k=2..6 (k-means++, Lloyd) + silhouette/CH/DB sweep' % (N, d),
                                rows=N, dim=d, ks=KS, parallelism='rows sharded over %d GPU(s), X replicated' % world,
                                l2='inputs (%.1f GB) exceed the 126 MB L2; no explicit flush' % (X.nbytes / 1e9)),
                    clocks=sampler.summary(), gpu_launches=int(launches), e2e=e2e, roofline=roof, cpu_baseline=cpu,
                    phases_ms=phases, n_iter=[r.n_iter for r in res], scores=[list(s) for s in scores],
                    best_k=max(scores, key=lambda s: s[1])[0])
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
