"""bench.py -- the BASELINE.json metric: embeddings/sec for the k=2..6 KMeans + silhouette sweep.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--also c2,c4] [--impl reference]

A step = one pass of the hot path over the synthetic embedding matrix of one BASELINE.json config:

  c3 (default, the config the metric is quoted on)  blobs(1_000_000, 512, kc=8): KMeans.fit prologue, k-means++ and Lloyd
      for k = 2..6, silhouette + Calinski-Harabasz + Davies-Bouldin for every k, selection of the best k
  c2  50k x 512: duplicate marking (--enc-dup-thr 0.2) + classification of the survivors against 32 references (0.9)
  c4  1M x 1024, k = 2..32, clusters-save-all
  c5  4M x 768: dedup 0.2 + the k = 2..6 sweep (meant for 8 GPUs; the matrix is generated on the device)

The line printed is the headline config's; unless --also none, short runs of the other single-box configs ride along
under "configs" (same record layout: value, phases, per-phase roofline, parity) so that they are driver-visible.

  value     X already resident in HBM, CUDA events around the K steps, max over ranks
  e2e       the public API call on a pinned HOST matrix: H2D of X and D2H of labels/scores inside the timed region
            (videotofaces_b200.grouping.cluster_labels / dupes.dedup_rows + grouping.classify)
  roofline  the dominant kernel (tcgen05 silhouette sweep; for c2 the tcgen05 dedup sweep), CUDA events on its launch
            stream, against MEASURED_PEAKS.json.  Both fractions are given: `frac` counts the MMA flops the kernel
            EXECUTES (the symmetric sweep computes the tiles on or above the diagonal only), `frac_algorithmic` the
            2*N^2*d of SURVEY.md 8(d) / the same time (can exceed 1: skipped work, not speed)
  phase_roofline  SURVEY.md 8(d)'s table: algorithmic bytes or flops of every phase / its time / the peak, and the
            config total (sum of the per-phase bounds / step time)
  parity    the gate of SURVEY.md 8(d): the GPU path and the reference's CPU call sequence (oracle/reference_calls.py
            over scikit-learn) on the SAME rows -- labels, k-means++ seeds, inertia, silhouette, CH, DB, duplicate set,
            classification labels -- plus the survey's known answers at full size.  A failed gate exits non-zero.
  cpu_baseline  that CPU run, timed on this box's host cores (rank 0, N=1 only): measured points at three sizes, the
            fit t = a*N + c*N^2 and the full-size figure LABELLED AS EXTRAPOLATED (the O(N^2) silhouette makes the
            1M-row CPU run take more than a day)

--impl reference times the CPU path as the reference arm (threads pinned explicitly: torchrun exports
OMP_NUM_THREADS=1).  N > 1 (torchrun): strong scaling, rows sharded over the ranks, each rank holds the full matrix.
"""
import argparse
import json
import math
import os
import os.path as osp
import subprocess
import sys
import threading
import time

if 'reference' in sys.argv and 'TORCHELASTIC_RUN_ID' in os.environ:
    # torchrun exports OMP_NUM_THREADS=1 to every worker; the reference arm is a CPU program and gets the host's cores
    # back BEFORE numpy / OpenBLAS / OpenMP read the variable (the N >= 2 reference values had halved)
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ.pop(_v, None)

ROOT = osp.dirname(osp.abspath(__file__))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = 'embeddings/sec, k=2-6 KMeans+silhouette sweep'
CONFIGS = {
    'c2': dict(kind='dedup_classify', rows=50_000, dim=512, kc=8, normalize=True, dup_thr=0.2, n_refs=32, oth_thr=0.9,
               cpu_rows=16384, text='config2: blobs(50000,512,kc=8,noise=1,seed=0)+planted dupes: dedup 0.2 + classify vs 32 refs (0.9)'),
    'c3': dict(kind='cluster', rows=1_000_000, dim=512, kc=8, normalize=True, ks=list(range(2, 7)), cpu_rows=16384,
               text='config3: blobs(1000000,512,kc=8,noise=1,seed=0) KMeans k=2..6 (k-means++, Lloyd) + silhouette/CH/DB sweep'),
    'c4': dict(kind='cluster', rows=1_000_000, dim=1024, kc=16, normalize=False, ks=list(range(2, 33)), cpu_rows=6144,
               text='config4: blobs(1000000,1024,kc=16,noise=1,seed=0,normalize=False) KMeans k=2..32, clusters-save-all'),
    'c5': dict(kind='dedup_cluster', rows=4_000_000, dim=768, kc=8, normalize=False, ks=list(range(2, 7)), dup_thr=0.2,
               cpu_rows=12288, device_data=True,
               text='config5: 4M x 768 (generated on the device) + planted dupes: dedup 0.2 + KMeans k=2..6 + silhouette sweep'),
}
# scikit-learn 1.9.0 on blobs(1000000,512,kc=8,noise=1,seed=0), random_state=0 (measured during the survey)
KNOWN_1M_512 = dict(inertia=[872811.19, 809175.13, 741865.88, 677289.25, 617947.13], n_iter=[3, 4, 4, 3, 3])
RTOL = 1e-4          # BASELINE.json north_star: distances, inertia and silhouette within 1e-4 relative
TIE_EPS = 1e-6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS))
    ap.add_argument('--also', default='auto', help="other configs that ride along under 'configs': auto | none | c2,c4,...")
    ap.add_argument('--rows', type=int, default=None, help='override the row count of the headline config (test runs)')
    ap.add_argument('--dim', type=int, default=None)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-rows', type=int, default=None, help='rows of the CPU sample (parity gate + cpu_baseline)')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU leg (and with it the sample parity gate)')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


def config_of(name, args=None, headline=False):
    cfg = dict(CONFIGS[name], name=name)
    if headline and args is not None:
        if args.rows:
            cfg['rows'] = args.rows
        if args.dim:
            cfg['dim'] = args.dim
        if args.cpu_rows:
            cfg['cpu_rows'] = args.cpu_rows
    cfg['cpu_rows'] = min(cfg['cpu_rows'], cfg['rows'])
    return cfg


# ------------------------------------------------------------------------------------------------ data
def host_data(cfg, rows=None):
    """The config's synthetic inputs as HOST arrays (identical bytes for the CPU and the GPU path)."""
    from videotofaces_b200.synth import blobs, plant_dupes
    N = rows or cfg['rows']
    X = blobs(N, cfg['dim'], kc=cfg['kc'], noise=1.0, seed=0, normalize=cfg['normalize'])
    out = dict(X=X)
    if 'dup_thr' in cfg:
        out['X'], out['src'], out['dst'] = plant_dupes(X, frac=50, seed=1, normalize=cfg['normalize'])
    if 'n_refs' in cfg:
        out['R'] = blobs(cfg['n_refs'], cfg['dim'], kc=cfg['kc'], noise=1.0, seed=5, normalize=cfg['normalize'])
    return out


def device_data(cfg, dev):
    """config 5: 4M x 768 in the shape of synth.blobs / synth.plant_dupes, generated ON THE DEVICE (same Philox seed on
    every rank, so every rank holds the same bytes; a 12 GB host matrix per rank is not needed)."""
    import torch
    N, d = cfg['rows'], cfg['dim']
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    C = torch.randn((cfg['kc'], d), generator=g, device=dev)
    X = C[torch.randint(0, cfg['kc'], (N,), generator=g, device=dev)]
    for s in range(0, N, 1 << 18):
        e = min(N, s + (1 << 18))
        X[s:e] += torch.randn((e - s, d), generator=g, device=dev)
    if cfg['normalize']:
        X /= X.norm(dim=1, keepdim=True).clamp_min(1e-30)
    g.manual_seed(1)
    m, half = N // 50, N // 2
    src = torch.randperm(half, generator=g, device=dev)[:m]
    dst = half + torch.randperm(N - half, generator=g, device=dev)[:m]
    eps = 0.3 + 0.8 * torch.rand((m, 1), generator=g, device=dev)
    Y = X[src]
    X[dst] = Y + Y.norm(dim=1, keepdim=True) * eps * torch.randn((m, d), generator=g, device=dev) / (d ** 0.5)
    cos = 1.0 - torch.nn.functional.cosine_similarity(X[src], X[dst], dim=1)
    return dict(Xd=X.contiguous(), dst=dst.cpu().numpy(), cos=cos.cpu().numpy())


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


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_threads():
    """Pin the CPU leg's thread pools explicitly and say how many threads it used."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    limiter = None
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=n)
    except Exception:  # noqa: BLE001
        pass
    return n, limiter


def cpu_info(threads):
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
    return dict(threads=threads, logical_cpus=os.cpu_count(), cpu_model=model, sklearn=sklearn.__version__, numpy=np.__version__,
                omp_num_threads_env=os.environ.get('OMP_NUM_THREADS'), threadpools=tp)


def cpu_step(cfg, data, n):
    """The reference's CPU path for this config on the first n rows (grouping.py:99-120 / dupes.py:60-68 /
    grouping.py:51-56 over scikit-learn).  Returns (seconds, result dict)."""
    from oracle import reference_calls as RC
    X = np.ascontiguousarray(data['X'][:n])
    t0 = time.perf_counter()
    out = {}
    if 'dup_thr' in cfg:
        mins, inds, idx = RC.dedup_core(X, cfg['dup_thr'])
        out.update(mins=mins, inds=inds, idx=idx)
        X = np.delete(X, idx, axis=0)
    if cfg['kind'] == 'dedup_classify':
        dist, lab = RC.classify_core(X, data['R'], cfg['n_refs'], cfg['oth_thr'])
        out.update(dist=dist, lab=lab)
    else:
        out.update(RC.cluster_core(X, cfg['ks'], 0))
    return time.perf_counter() - t0, out


def fit_quadratic(points):
    """Least squares t = a*N + c*N^2 (a, c >= 0) through the measured (N, seconds) points."""
    n = np.asarray([p[0] for p in points], dtype=np.float64)
    t = np.asarray([p[1] for p in points], dtype=np.float64)
    A = np.stack([n, n * n], axis=1)
    sol = np.linalg.lstsq(A, t, rcond=None)[0] if len(points) >= 2 else np.asarray([0.0, t[0] / n[0] ** 2])
    a, c = float(sol[0]), float(sol[1])
    if a < 0 or c <= 0:
        a, c = 0.0, float((t * n * n).sum() / (n ** 4).sum())
    return a, c


def extrapolate(cfg, points, threads):
    a, c = fit_quadratic(points)
    N = cfg['rows']
    t_full = a * N + c * N * N
    return dict(points=[dict(rows=int(n), seconds=float(t), emb_per_s=float(n / t)) for n, t in points],
                fit='t = a*N + c*N^2', a=a, c=c, full_rows=N, full_seconds_EXTRAPOLATED=t_full,
                full_emb_per_s_EXTRAPOLATED=N / t_full, threads=threads,
                note='the last two numbers are an extrapolation of the fit to the full config, not a measurement')


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (its call sites over scikit-learn,
    oracle/reference_calls.py -- the reference is pure Python, there is nothing to compile into oracle/_ref), on a
    bounded row sample of the arm's config sized so that W + K steps end within a few minutes."""
    if int(os.environ.get('RANK', '0')) != 0:
        return 0
    cfg = config_of(args.config, args, headline=True)
    threads, limiter = cpu_threads()
    K, W = max(args.steps, 1), max(args.warmup, 0)
    n_max = min(cfg['rows'], 32768)
    data = host_data(cfg, n_max)
    cpu_step(cfg, data, min(2048, n_max))                                      # imports, thread pools
    cal = sorted({min(n_max, cfg['cpu_rows'] // 2), min(n_max, cfg['cpu_rows'])})
    points = [(n, cpu_step(cfg, data, n)[0]) for n in cal]
    a, c = fit_quadratic(points)
    budget = 150.0 / (K + W)                                                   # seconds per step
    disc = a * a + 4 * c * budget
    n_s = int((-a + math.sqrt(disc)) / (2 * c)) if c > 0 else n_max
    n_s = max(2048, min(n_max, (n_s // 1024) * 1024))
    if any(abs(n_s - n) < 1024 for n in cal):                                  # keep three distinct sizes for the fit
        extra = ((cal[0] + cal[-1]) // 2048) * 1024 if len(cal) == 2 else max(2048, cal[0] // 2)
        if all(abs(extra - n) >= 1024 for n in cal + [n_s]):
            points.append((extra, cpu_step(cfg, data, extra)[0]))
    for _ in range(W):
        cpu_step(cfg, data, n_s)
    times = [cpu_step(cfg, data, n_s)[0] for _ in range(K)]
    t = float(np.mean(times))
    points.append((n_s, t))
    value = n_s / t
    info = cpu_info(threads)
    ex = extrapolate(cfg, sorted(points), threads)
    sample = ('%s on the first %d of the %d rows (same generator), %d threads: %.2f s per step; O(N^2) terms => %.1f emb/s '
              'EXTRAPOLATED to the full config (fit over %s rows)'
              % ('sklearn KMeans+silhouette+CH+DB' if cfg['kind'] != 'dedup_classify' else 'sklearn cosine_distances dedup + classify',
                 n_s, cfg['rows'], threads, t, ex['full_emb_per_s_EXTRAPOLATED'], [p['rows'] for p in ex['points']]))
    line = dict(impl='reference', metric=METRIC, value=value, unit='embeddings/s', n_gpus=args.gpus, steps=K, warmup=W,
                ms_per_step=t * 1e3, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64 (sklearn upcast path)',
                data='synthetic', same_config=False,
                config=dict(workload=cfg['text'] + ' (CPU arm: %d-row sample)' % n_s, name=cfg['name'], rows=n_s, full_rows=cfg['rows'],
                            dim=cfg['dim'], ks=cfg.get('ks')),
                cpu_baseline=dict(value=value, unit='embeddings/s', cores=threads, kind='port', sample=sample, info=info, scaling=ex),
                e2e=dict(value=value, unit='embeddings/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)
    del limiter
    return 0


# ------------------------------------------------------------------------------------------------ stdout
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


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ roofline tables
def kpp_passes(ks):
    """one pass over X per k-means++ round; fits with equal n_local_trials share their seeds."""
    by = {}
    for k in ks:
        by.setdefault(2 + int(math.log(k)), []).append(k)
    return sum(max(v) for v in by.values())


def phase_table(cfg, N, d, phases_ms, pk, world, info):
    """SURVEY.md 8(d) 'which roofline bounds what': ALGORITHMIC bytes / flops per phase (whole job), its bound in ms on
    `world` GPUs at the measured peaks, the measured phase time, fraction = bound / time."""
    hbm, tf = pk['hbm'] * 1e9 * world, pk['tf_sust'] * 1e12 * world
    X = float(N) * d * 4
    rows = []

    def add(name, key, bytes_=0.0, flops=0.0, note=''):
        t = phases_ms.get(key)
        if t is None or t <= 0:
            return
        bound = max(bytes_ / hbm, flops / tf) * 1e3
        rows.append(dict(phase=name, bound='tensor' if flops / tf > bytes_ / hbm else 'hbm', algorithmic_bytes=bytes_,
                         algorithmic_flops=flops, bound_ms=bound, ms=t, frac=bound / t, note=note))

    Nk = float(info.get('rows_after_dedup', N))
    Xk = Nk * d * 4
    if 'dup_thr' in cfg:
        add('normalize', 'normalize', bytes_=2 * X + (2 * Xk if cfg['kind'] == 'dedup_classify' else 0))
        add('dedup', 'dedup', flops=float(N) * N * d, note='F = N^2*d (strictly-lower triangle of the cosine Gram)')
        add('delete_rows', 'delete_rows', bytes_=2 * Xk)
    if cfg['kind'] == 'dedup_classify':
        R = cfg['n_refs']
        add('classify', 'classify', bytes_=Xk + Nk * R * 4 + Nk * 8, flops=2 * Nk * R * d)
    else:
        ks = cfg['ks']
        iters = info.get('n_iter', [0])
        redo = 1 if info.get('any_not_strict', True) else 0
        add('prepare', 'prepare', bytes_=2 * Xk)
        add('kmeans++', 'kmeans++', bytes_=kpp_passes(ks) * Xk, note='%d passes (one per seeding round)' % kpp_passes(ks))
        add('lloyd', 'lloyd', bytes_=(max(iters) + redo + 1) * Xk, flops=2 * Nk * sum(ks) * d * (max(iters) + redo),
            note='fused over k: max n_iter (%d) E/M passes%s + the inertia pass; per-k form would be %d passes'
                 % (max(iters), ' + closing E-step' if redo else '', sum(iters) + len(iters)))
        add('sil_cells+pack', 'sil_cells+pack', bytes_=Xk * 1.5 + Nk * 4 * len(ks))
        add('silhouette', 'sil_sweep', flops=2 * Nk * Nk * d, note='F = 2*N^2*d once per sweep (full Gram, shared by every k)')
        add('sil_finish', 'sil_finish', bytes_=Nk * 4 * info.get('cells', len(ks)))
        add('ch_db', 'ch_db', bytes_=2 * Xk, note='two fused passes (cluster means, dispersions) for all k')
    tot_bound = sum(r['bound_ms'] for r in rows)
    tot_ms = info['ms_per_step']
    return dict(phases=rows, total=dict(bound_ms=tot_bound, ms=tot_ms, frac=tot_bound / tot_ms,
                                        in_phases_ms=sum(r['ms'] for r in rows)),
                peaks=dict(hbm_gbs=pk['hbm'], tensor_tflops=pk['tf_sust'], n_gpus=world, source=pk['source']))


def committed_traffic(cfg, N, d, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel -- NOT measured in this run: read from the
    committed `ncu --set full` capture of the same shape, when there is one."""
    name = {'cluster': 'sweep', 'dedup_cluster': 'sweep', 'dedup_classify': 'dedup_screen_v2'}[cfg['kind']]
    for rnd in ('r02', 'r01'):
        try:
            with open(osp.join(ROOT, 'profiles', '%s_%s_ncu_summary.json' % (rnd, name))) as f:
                ncu = json.load(f)
            if ncu['rows'] == N and ncu['dim'] == d and world == 1:
                return ncu['traffic_bytes_per_launch'], 'committed ncu --set full capture profiles/%s_%s_ncu_summary.json (not this run)' % (rnd, name)
        except Exception:  # noqa: BLE001
            pass
    return None, None


# ------------------------------------------------------------------------------------------------ parity gates
def _rel(a, b):
    return float(abs(a - b) / max(abs(b), 1e-300))


def cluster_parity(got, ref, X, ks):
    """GPU cluster_labels() vs the reference call sequence on the same rows."""
    from oracle import reference_calls as RC
    p = dict(rows=int(X.shape[0]), labels_agree=[], seeds_equal=[], n_iter_equal=[], inertia_rel=[], silhouette_rel=[], ch_rel=[],
             db_rel=[])
    for j, k in enumerate(ks):
        p['labels_agree'].append(float((got['labels'][j] == ref['labels'][j]).mean()))
        p['seeds_equal'].append(bool(np.array_equal(got['seeds'][j], RC.kmeanspp_indices(X, k, 0))))
        p['n_iter_equal'].append(bool(got['n_iter'][j] == ref['n_iter'][j]))
        p['inertia_rel'].append(_rel(got['inertia'][j], ref['inertia'][j]))
        for nm, col in (('silhouette_rel', 1), ('ch_rel', 2), ('db_rel', 3)):
            p[nm].append(_rel(got['scores'][j][col], ref['scores'][j][col]))
    p['best_k_equal'] = bool(got['best_k'] == ref['best_k'])
    p['ok'] = bool(min(p['labels_agree']) >= 0.95 and all(p['seeds_equal']) and p['best_k_equal'] and
                   max(p['inertia_rel'] + p['silhouette_rel'] + p['ch_rel'] + p['db_rel']) <= RTOL)
    return p


def dedup_parity(mins, inds, idx, ref, thr):
    ties = np.abs(ref['mins'][1:] - np.float32(thr)) <= TIE_EPS
    tie_rows = set((np.nonzero(ties)[0] + 1).tolist())
    a = set(idx.tolist()) - tie_rows
    b = set(ref['idx'].tolist()) - tie_rows
    rel = np.abs(mins[1:] - ref['mins'][1:]) / np.maximum(np.abs(ref['mins'][1:]), 1e-30)
    p = dict(rows=int(mins.shape[0]), dup_set_equal=bool(a == b), n_dupes=int(ref['idx'].size), ties=int(ties.sum()),
             min_distance_rel=float(rel.max()), nearest_index_agree=float((inds[1:] == ref['inds'][1:]).mean()))
    p['ok'] = bool(p['dup_set_equal'] and p['min_distance_rel'] <= RTOL)
    return p


# ------------------------------------------------------------------------------------------------ one config on the GPUs
def run_config(cfg, args, env, steps, warmup, headline):
    import torch
    import torch.distributed as dist
    from videotofaces_b200 import _native
    from videotofaces_b200.dupes import dedup_rows
    from videotofaces_b200.grouping import classify, cluster_labels
    eng, world, rank, local = env['eng'], env['world'], env['rank'], env['local']
    kind = cfg['kind']
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- data
    t0 = time.perf_counter()
    data = None
    if cfg.get('device_data'):
        dd = device_data(cfg, eng.device)
        Xd = dd['Xd']
        Xpin = None
    else:
        data = host_data(cfg)
        Xpin = torch.from_numpy(data['X']).pin_memory()
        Xd = eng.upload_replicated(Xpin)
    Rd = torch.from_numpy(data['R']).to(eng.device) if data is not None and 'R' in data else None
    torch.cuda.synchronize()
    N, d = Xd.shape
    log('[bench] %s: data ready in %.1f s (%d x %d)' % (cfg['name'], time.perf_counter() - t0, N, d))

    def step():
        out = {}
        X = Xd
        if 'dup_thr' in cfg:
            out['mins'], out['inds'], out['idx'] = eng.dedup(X, cfg['dup_thr'])
            X = eng.delete_rows(X, out['idx']) if out['idx'].size else X
            out['rows_after_dedup'] = int(X.shape[0])
        out['X'] = X
        if kind == 'dedup_classify':
            out['dist'], out['lab'] = eng.classify(X, Rd, cfg['oth_thr'])
        else:
            ctx = eng.prepare(X)
            ctx, res = eng.kmeans_sweep(None, cfg['ks'], 0, ctx=ctx)
            out['res'], out['scores'] = res, eng.cluster_scores(ctx, res)
        return out

    flush = None
    if Xd.numel() * 4 < 512e6:          # smaller than a few L2s: write 512 MB between the timed iterations
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=eng.device)
    eng.profile = True
    for _ in range(warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local) if headline else None
    if sampler:
        sampler.start()
    l0 = _native.launch_count()
    eng.events = []
    total_ms = 0.0
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            out = step()
        ev1.record()
        barrier()
        total_ms = ev0.elapsed_time(ev1)
    else:
        pairs = []
        barrier()
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = step()
            b.record()
            pairs.append((a, b))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in pairs)
    launches = _native.launch_count() - l0
    ms = torch.tensor([total_ms], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    eng.profile = False

    phases = {}
    for name, a, b, _ in eng.events:
        if name != 'wide_assign':                                              # nested inside 'lloyd'
            phases[name] = phases.get(name, 0.0) + a.elapsed_time(b) / steps
    phases['sil_cells+pack'] = phases.get('sil_cells', 0.0) + phases.get('sil_pack', 0.0)

    rec = dict(value=N / (ms_step * 1e-3), unit='embeddings/s', ms_per_step=ms_step, steps=steps, warmup=warmup, rows=int(N), dim=int(d),
               gpu_launches=int(launches), launches_per_step=launches / steps, phases_ms=phases,
               l2=('512 MB written between timed iterations (inputs fit in L2)' if flush is not None else
                   'inputs (%.1f GB) exceed the 126 MB L2; no explicit flush' % (Xd.numel() * 4 / 1e9)))
    info = dict(ms_per_step=ms_step)
    if 'res' in out:
        res, scores = out['res'], out['scores']
        info.update(n_iter=[r.n_iter for r in res], any_not_strict=any(not r.strict for r in res))
        rec.update(n_iter=info['n_iter'], inertia=[r.inertia for r in res], scores=[list(s) for s in scores],
                   best_k=max(scores, key=lambda s: s[1])[0])
    if 'idx' in out:
        info['rows_after_dedup'] = out['rows_after_dedup']
        rec.update(n_dupes=int(out['idx'].size), rows_after_dedup=out['rows_after_dedup'])
    if 'lab' in out:
        rec['class_counts'] = np.bincount(out['lab'], minlength=cfg['n_refs'] + 1).tolist()

    # ---- dominant kernel, CUDA events on its launch stream
    dom = 'dedup' if kind == 'dedup_classify' else 'sil_sweep'
    spans = [(a.elapsed_time(b), f() if callable(f) else f) for (name, a, b, f) in eng.events if name == dom]
    if spans:
        t_ms = float(np.mean([t for t, _ in spans]))
        fl = float(np.mean([f for _, f in spans]))
        Nk = float(info.get('rows_after_dedup', N)) if dom == 'sil_sweep' else float(N)
        alg = (2.0 if dom == 'sil_sweep' else 1.0) * Nk * Nk * d / world
        alone = kind == 'dedup_classify'
        peak = pk['tf_burst'] if alone else pk['tf_sust']
        traffic, tsrc = committed_traffic(cfg, int(N), int(d), world)
        dstats = getattr(eng.K, 'dedup_stats', None) or {}
        # dedup: the single-MMA screen executes the algorithmic N^2*d once (plus three times the re-swept units' share);
        # the split-fp16 sweep issues hi*hi + hi*lo + lo*hi per K step
        executed = fl if dom == 'sil_sweep' else fl * (3.0 if dstats.get('mode') == 'split' else
                                                       1.0 + 3.0 * dstats.get('reswept_units', 0) / max(dstats.get('units', 1), 1))
        rec['roofline'] = dict(
            bound='tensor', kernel=('sil_sweep_kernel' if dom == 'sil_sweep' else 'dedup_screen_kernel') +
            ' (tcgen05 cta_group::2 kind::f16, fp32 accumulate in TMEM)' + ('' if dom == 'sil_sweep' else ' [%s]' % dstats),
            achieved=executed / (t_ms * 1e-3) / 1e12, peak=peak, unit='TFLOP/s', frac=executed / (t_ms * 1e-3) / 1e12 / peak,
            achieved_algorithmic=alg / (t_ms * 1e-3) / 1e12, frac_algorithmic=alg / (t_ms * 1e-3) / 1e12 / peak,
            traffic=traffic, traffic_source=tsrc, ms_per_launch=t_ms, flop_per_launch_executed=executed, flop_per_launch_algorithmic=alg,
            share_of_step=t_ms * len(spans) / steps / ms_step,
            note=('achieved/frac count the MMA flops the kernel EXECUTES per rank; *_algorithmic count SURVEY 8(d)\'s figure (%s) per rank '
                  'in the same time. The symmetric sweep computes only the tiles on or above the diagonal, so its algorithmic fraction can '
                  'exceed 1 -- skipped work, not speed; the dedup screen issues one MMA per K step (the split-fp16 sweep three).'
                  % ('2*N^2*d' if dom == 'sil_sweep' else 'N^2*d')),
            peak_source=pk['source'] + (' bf16_tflops (burst: kernel timed alone)' if alone else ' bf16_tflops_sustained (kernel timed inside a long step)'))
    rec['phase_roofline'] = phase_table(cfg, int(N), int(d), phases, pk, world, info)

    # ---- size-independent properties at full size
    props = {}
    X_step = out.pop('X')
    if 'res' in out:
        inert = rec['inertia']
        props.update(inertia_decreasing=bool(all(a > b for a, b in zip(inert, inert[1:]))),
                     silhouette_in_range=bool(all(-1.0 <= s[1] <= 1.0 for s in out['scores'])))
        if world > 1:
            h = torch.stack([r.labels.to(torch.int64).sum() for r in out['res']])
            lo, hi = h.clone(), h.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            props['labels_equal_on_all_ranks'] = bool(torch.equal(lo, hi))
        if world > 1:
            # the sharded path against every rank computing a 16 384-row sample alone (same kernels, no exchange): labels
            # identical, scores equal -- in the full and in the symmetric (benchmarked) form of the sweep
            Xs_d = X_step[:min(16384, X_step.shape[0])].contiguous()
            worst = 0.0
            same_labels = True
            for sym in (False, True):
                eng.sym = sym
                try:
                    r_sh = cluster_labels(Xs_d, cfg['ks'][:5], 0)
                    with eng.solo_if(True):
                        r_so = cluster_labels(Xs_d, cfg['ks'][:5], 0)
                finally:
                    eng.sym = None
                same_labels = same_labels and all(np.array_equal(a, b) for a, b in zip(r_sh['labels'], r_so['labels']))
                worst = max(worst, max(_rel(a[c], b[c]) for a, b in zip(r_sh['scores'], r_so['scores']) for c in (1, 2, 3)))
            props.update(sharded_sample_labels_equal_unsharded=bool(same_labels), sharded_sample_scores_rel=worst,
                         sharded_sample_ok=bool(same_labels and worst <= 1e-5))
        if cfg['name'] == 'c3' and N == 1_000_000 and d == 512:
            rel = [_rel(a, b) for a, b in zip(inert, KNOWN_1M_512['inertia'])]
            props.update(known_answer_inertia_rel=rel, known_answer_n_iter_equal=bool(rec['n_iter'] == KNOWN_1M_512['n_iter']),
                         known_answers='SURVEY.md 8(c): scikit-learn 1.9.0 inertia / n_iter on the same 1M x 512 matrix')
            props['known_ok'] = bool(max(rel) <= RTOL and props['known_answer_n_iter_equal'])
    if 'idx' in out:
        if data is not None:
            dst, cos = data['dst'], 1.0 - np.einsum('ij,ij->i', data['X'][data['src']], data['X'][data['dst']]) / (
                np.linalg.norm(data['X'][data['src']], axis=1) * np.linalg.norm(data['X'][data['dst']], axis=1))
        else:
            dst, cos = dd['dst'], dd['cos']
        found = np.isin(dst, out['idx'])
        props.update(planted=int(dst.size), planted_below_thr_found=bool(found[cos < cfg['dup_thr'] - 0.01].all()),
                     planted_above_thr_kept=bool((~found[cos > cfg['dup_thr'] + 0.01]).all()),
                     other_rows_marked=int(out['idx'].size - found.sum()))
    props['ok'] = all(v for k, v in props.items() if isinstance(v, bool))
    rec['properties'] = props

    # ---- end to end through the public API, host buffers
    if not args.no_e2e and Xpin is not None:
        Xh = Xpin.numpy()

        def api(Xh=Xh):
            if kind == 'cluster':
                return cluster_labels(Xh, cfg['ks'], 0)
            # the drop-in's own flow (videotofaces_b200/main.py, reference main.py:72-80): the surviving rows are only handed to
            # the grouping call, so they stay in HBM; the dup set, distances, labels and [N,R] distances come back to the host
            r = dedup_rows(Xh, cfg['dup_thr'], keep_on_device=True)
            if kind == 'dedup_classify':
                r['lab'] = classify(r['kept'], data['R'], ['c%02d' % i for i in range(cfg['n_refs'])], cfg['oth_thr'], False, None, None)[0]
            return r
        for _ in range(1 if ms_step > 100 else 3):                             # warm (pinned result blocks, allocator, clocks)
            res_api = api()
        barrier()
        n_e2e = max(1, min(steps, 2 if ms_step > 100 else 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res_api = api()
        barrier()
        t_e2e = torch.tensor([(time.perf_counter() - t0) / n_e2e], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        if kind == 'cluster':
            d2h = int(sum(l.nbytes for l in res_api['labels']) + sum(c.nbytes for c in res_api['centers']) + 8 * 4 * len(cfg['ks']))
            what = 'videotofaces_b200.grouping.cluster_labels(X_host, %s, 0)' % ('[2..%d]' % cfg['ks'][-1])
            h2d = int(Xh.nbytes)
        else:
            d2h = int(8 * N + res_api['lab'].nbytes + res_api['lab'].size * 4 * cfg['n_refs'])      # packed keys, labels, [N,R] distances
            h2d = int(Xh.nbytes + data['R'].nbytes + 8 * (N - res_api['idx'].size))                  # X, refs, kept-row indices
            what = ('videotofaces_b200.dupes.dedup_rows(X_host, 0.2, keep_on_device=True) -> grouping.classify(kept, R_host, classes, '
                    '0.9, ...): the flow of video_to_faces(mode="grouping") after the encoder')
        rec['e2e'] = dict(value=N / float(t_e2e.item()), unit='embeddings/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                          ms_per_step=float(t_e2e.item()) * 1e3, api=what)
        if world == 1 and Xh.nbytes <= (4 << 30):
            # the same call on an ORDINARY numpy array (what the reference's caller holds, main.py:71): the upload goes through
            # vtf_upload_pageable (native threads + page-locked double buffers) instead of one pageable cudaMemcpy
            Xpg = np.array(Xh, copy=True)
            api(Xpg)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                api(Xpg)
            torch.cuda.synchronize()
            rec['e2e']['pageable_input_ms_per_step'] = (time.perf_counter() - t0) / n_e2e * 1e3
            del Xpg
    else:
        rec['e2e'] = None
        if Xpin is None:
            rec['e2e_note'] = 'matrix generated on the device (a 12 GB host copy per rank is not materialised)'

    # ---- parity gate + CPU baseline on a row sample (rank 0 of a single-GPU run: the CPU leg needs the host's cores)
    if headline and rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s = cfg['cpu_rows']
        if data is None:                                                       # device-generated matrix: the sample comes back over PCIe
            data = dict(X=Xd[:n_s].cpu().numpy())
        threads, limiter = cpu_threads()
        sizes = sorted({max(2048, (n_s * 3 // 8 // 1024) * 1024), max(2048, (n_s * 3 // 4 // 1024) * 1024), n_s})
        points, ref = [], None
        for n in sizes:
            dt, ref = cpu_step(cfg, data, n)
            points.append((n, dt))
            log('[bench] CPU leg: %d rows in %.1f s' % (n, dt))
        del limiter
        Xs = np.ascontiguousarray(data['X'][:n_s])
        par = {}
        if 'dup_thr' in cfg:
            r = dedup_rows(Xs, cfg['dup_thr'])
            par = dedup_parity(r['mins'], r['inds'], r['idx'], ref, cfg['dup_thr'])
            Xs = r['kept']
        if kind != 'dedup_classify':
            got = cluster_labels(Xs, cfg['ks'], 0)
            cp = cluster_parity(got, ref, Xs, cfg['ks'])
            cp['ok'] = bool(cp['ok'] and par.get('ok', True))
            par.update(cp)
            rec['parity'] = par
        else:
            # classification at FULL size: the CPU needs 0.2 s for it
            from oracle import reference_calls as RC
            kept_full = np.delete(data['X'], out['idx'], axis=0)
            rdist, rlab = RC.classify_core(kept_full, data['R'], cfg['n_refs'], cfg['oth_thr'])
            par.update(classify_rows=int(kept_full.shape[0]), classify_labels_equal=bool(np.array_equal(out['lab'], rlab)),
                       classify_distance_abs=float(np.abs(out['dist'] - rdist).max()),
                       classify_ties=int((np.abs(rdist.min(axis=1) - np.float32(cfg['oth_thr'])) <= TIE_EPS).sum()))
            par['ok'] = bool(par['ok'] and (par['classify_labels_equal'] or par['classify_ties'] > 0))
            rec['parity'] = par
        info_c = cpu_info(threads)
        ex = extrapolate(cfg, points, threads)
        rec['cpu_baseline'] = dict(
            value=n_s / points[-1][1], unit='embeddings/s', cores=threads, kind='port',
            sample=('reference call sequence over scikit-learn %s on the first %d rows of the same matrix, %d threads: %.1f s; the O(N^2) '
                    'terms make the full config slower still: %.1f emb/s EXTRAPOLATED from the fit over %s rows'
                    % (info_c['sklearn'], n_s, threads, points[-1][1], ex['full_emb_per_s_EXTRAPOLATED'], [p[0] for p in points])),
            scaling=ex, info=info_c)
    if headline:
        rec['clocks'] = sampler.summary()
    del Xd, X_step
    torch.cuda.empty_cache()
    return rec


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
    from videotofaces_b200.runtime import get_engine
    env = dict(eng=get_engine(), world=world, rank=rank, local=local)

    cfg = config_of(args.config, args, headline=True)
    head = run_config(cfg, args, env, args.steps, args.warmup, headline=True)

    also = args.also
    if also == 'auto':
        also = [] if (args.rows or args.dim) else [c for c in ('c2', 'c4') if c != args.config]
        if world == 8 and args.config != 'c5' and not (args.rows or args.dim):
            also.append('c5')
    elif also == 'none':
        also = []
    else:
        also = [c for c in also.split(',') if c in CONFIGS and c != args.config]
    others = {}
    for name in also:
        c2 = config_of(name)
        try:
            fast = c2['kind'] == 'dedup_classify'
            others[name] = run_config(c2, args, env, steps=10 if fast else 2, warmup=3 if fast else 1, headline=False)
            others[name]['workload'] = c2['text']
        except Exception as e:  # noqa: BLE001  (a side config must not take the headline line with it)
            others[name] = dict(error='%s: %s' % (type(e).__name__, e))
            log('[bench] %s failed: %r' % (name, e))

    ok = True
    if rank == 0:
        gates = [head.get('parity'), head.get('properties')] + [o.get('properties') for o in others.values() if isinstance(o, dict)]
        ok = all(g.get('ok', True) for g in gates if g) and all(g.get('known_ok', True) for g in gates if g)
        line = dict(metric=METRIC, value=head['value'], unit='embeddings/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=head['ms_per_step'], higher_is_better=True, scaling='strong', vs_baseline=None,
                    dtype='f16 operands / f32 accumulate (Gram), f32+f64 elsewhere', data='synthetic',
                    config=dict(workload=cfg['text'], name=cfg['name'], rows=head['rows'], dim=head['dim'], ks=cfg.get('ks'),
                                parallelism='rows sharded over %d GPU(s), X replicated' % world, l2=head['l2']),
                    clocks=head['clocks'], gpu_launches=head['gpu_launches'], e2e=head['e2e'], roofline=head.get('roofline'),
                    cpu_baseline=head.get('cpu_baseline'), parity=head.get('parity'), parity_ok=ok)
        for k in ('phases_ms', 'phase_roofline', 'properties', 'n_iter', 'inertia', 'scores', 'best_k', 'n_dupes', 'rows_after_dedup',
                  'class_counts', 'launches_per_step', 'e2e_note'):
            if k in head:
                line[k] = head[k]
        if others:
            line['configs'] = others
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
