"""Host orchestration of the grouping hot path on B200 (one process per GPU).

Mirrors, stage by stage, what the reference delegates to scikit-learn 1.9.0 ("SK/" below is
site-packages/sklearn; reference lines are /root/reference/src/videotofaces/...):

  dedup()           dupes.py:60-65        cosine_distances + masked min/argmin + threshold
  classify()        grouping.py:51-56     cosine_distances + argmin + 'other' threshold
  kmeans_sweep()    grouping.py:99-101    KMeans(k, random_state, n_init='auto').fit for every k
  cluster_scores()  grouping.py:104-108   silhouette / calinski-harabasz / davies-bouldin per k

Only O(k), O(d) and O(trials) scalars ever come back to the host inside these loops (convergence
flags, candidate potentials, RNG draws); all O(N) work is in the kernels of csrc/.
"""
import math
import os

import numpy as np
import torch

from .dist import Comm, shard_bounds

TC_MIN_ROWS = 4096        # below this the exact fp32 silhouette kernel is used (and is fast enough)
TC_SYM_MIN_UNITS = 148    # symmetric (upper-triangle) sweep once there are at least 2 units of 256 rows per CTA pair
TC_DEDUP_MIN_ROWS = 8192  # below this the fp32 FMA dedup kernel is used
TC_MAX_CELLSUM_BYTES = 32 << 30
WIDE_MIN_SUM_K = int(os.environ.get('VTF_WIDE_MIN_SUM_K', '33'))   # sum(k) of a k-range from which the E-step runs as one tensor-core contraction (lloyd_tc.cu);
                          # below it a single register-tiled pass (rowtile.cu) is HBM-bound and cheaper
WIDE_MIN_ROWS = 8192
WIDE_MAX_COLS = 2048      # centre columns per launch (the kernel keeps their table in shared memory)
SOLO_ROWS_PER_RANK = 64   # fewer rows per rank than this: no sharding, every rank computes the whole input


class KMeansResult:
    __slots__ = ('k', 'labels', 'centers', 'inertia', 'n_iter', 'seeds', 'counts', 'strict', 'csums', 'disp')

    def __init__(self, **kw):
        for n in self.__slots__:
            setattr(self, n, kw.get(n))


class FitContext:
    """What KMeans.fit derives from X before seeding (SK/cluster/_kmeans.py:1473-1496), shared by every k."""
    __slots__ = ('X', 'Xc', 'N', 'd', 'd_orig', 'bounds', 'b', 'e', 'col_sum', 'mean64', 'mean32', 'tol', 'wide', 'amax', 'max_sqn', 'xsq', 'kpp_first')


def n_local_trials(k):
    return 2 + int(math.log(k))                                            # SK/cluster/_kmeans.py:228


def _ceil_to_f32(v64):
    """Smallest fp32 >= v (searchsorted compares the fp32 cumsum with fp64 thresholds, _kmeans.py:250)."""
    t = np.asarray(v64, dtype=np.float64).astype(np.float32)
    low = t.astype(np.float64) < v64
    t[low] = np.nextafter(t[low], np.float32(np.inf))
    return t


class _Span:
    """CUDA-event bracket on the current stream (bench.py reads engine.events); a no-op unless
    engine.profile is set."""

    def __init__(self, eng, name, flops=0.0):
        self.eng, self.name, self.flops = eng, name, flops
        self.on = bool(eng.profile) and eng.device.type == 'cuda'

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            self.eng.events.append((self.name, self.a, self.b, self.flops))
        return False


class _Solo:
    def __init__(self, eng, on):
        self.eng, self.on = eng, on

    def __enter__(self):
        if self.on:
            self.saved = self.eng.comm
            self.eng.comm = Comm.single()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.eng.comm = self.saved
        return False


class GroupingEngine:
    def __init__(self, kernels, comm=None):
        self.K = kernels
        self.comm = comm if comm is not None else Comm()
        self.device = kernels.device
        self.profile = False
        self.events = []
        self.lockstep = True     # all fits of a k-range share each pass over X (False: one Lloyd loop per k)
        self.sym = None          # None: symmetric silhouette sweep for large inputs; True / False forces it
        self.wide = None         # None: tensor-core E-step for wide k-ranges on large inputs; True / False forces it
        self.wide_eps_scale = float(os.environ.get('VTF_WIDE_EPS_SCALE', '1'))
        self.wide_stats = {}     # iterations / fix-list and change-list sizes of the last wide run (tests, tools)
        # symmetric sweep: claim work items from a counter shared by all ranks (NVLink peer memory) instead of dealing
        # them out by rank -- a slower GPU then takes fewer items.  VTF_SWEEP_DYNAMIC=0 keeps the static split.
        self.dynamic = os.environ.get('VTF_SWEEP_DYNAMIC', '1') != '0'
        self._wc = None          # (device address of 2 uint32 counters on rank 0, keep-alive object), False = unavailable
        self._wc_calls = 0

    def span(self, name, flops=0.0):
        return _Span(self, name, flops)

    def solo_if(self, cond):
        """Context manager: while active (and `cond` holds) the engine runs un-sharded -- every rank computes the whole
        (tiny) problem redundantly with the same deterministic kernels, so all ranks hold the same result without an
        exchange.  Used for inputs with fewer than SOLO_ROWS_PER_RANK rows per rank, where shards would be empty or
        smaller than one tile.  `cond` must be derived from global quantities (every rank takes the same branch)."""
        return _Solo(self, bool(cond) and self.comm.active)

    def too_small_to_shard(self, n):
        return n < SOLO_ROWS_PER_RANK * self.comm.world

    # ------------------------------------------------------------------------------ utilities
    def to_device(self, A, dtype=torch.float32):
        if isinstance(A, torch.Tensor):
            return A.to(device=self.device, dtype=dtype).contiguous()
        A = np.ascontiguousarray(A, dtype={torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32,
                                           torch.int64: np.int64}[dtype])
        t = torch.from_numpy(A)
        if self.device.type == 'cuda' and A.nbytes >= (16 << 20) and hasattr(self.K, 'upload_pageable') and not t.is_pinned():
            # an ordinary numpy array (what the reference hands over, main.py:71): native threads stage it through page-locked
            # buffers in parallel instead of one pageable cudaMemcpy (6-13 GB/s)
            threads = int(os.environ.get('VTF_UPLOAD_THREADS', str(min(8, os.cpu_count() or 1))))
            if threads > 0:
                return self.K.upload_pageable(A, dtype, threads)
        return t.to(self.device, non_blocking=False)

    def to_host(self, tensors):
        """Device tensors -> numpy arrays with ONE synchronisation and NO host-side copy: every tensor is copied asynchronously
        into page-locked memory that torch's caching host allocator hands out (after the first calls a cached block: no
        cudaHostAlloc), the stream is waited for once, and the arrays returned are views of those blocks -- they own them until
        they are garbage collected, when the blocks go back to the cache.  (`.cpu()` per tensor stages through pageable memory and
        synchronises every time; a shared staging buffer costs a second pass over the bytes on the host: 0.6 ms for the
        [50k, 32] distance matrix of config 2.)"""
        if self.device.type != 'cuda':
            return [t.detach().cpu().numpy() for t in tensors]
        outs = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        for o, t in zip(outs, tensors):
            if t.numel():
                o.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [o.numpy() for o in outs]

    def to_host_async(self, tensors):
        """to_host that does not wait: the copies are queued on a side stream behind what the current stream has been given so
        far, and a helper thread waits for them and copies the results out of a persistent page-locked staging buffer (numpy
        releases the GIL for the copies) -- while the caller goes on launching work.  (Handing out freshly pinned blocks as in
        to_host was measured slower here: with tens of MB per call the allocator's cudaHostAlloc calls land in the timed calls.)  `.result()` joins the thread and returns the numpy
        arrays.  Used for the labels and centres of a k-range, which are final before the silhouette sweep starts: their
        trip to the host hides behind the sweep instead of following it."""
        if self.device.type != 'cuda':
            res = [t.detach().cpu().numpy() for t in tensors]
            return type('Done', (), {'result': staticmethod(lambda: res)})()
        import threading
        sizes = [((t.numel() * t.element_size() + 255) // 256) * 256 for t in tensors]
        total = sum(sizes)
        if getattr(self, '_stage_async', None) is None or self._stage_async.numel() < total:
            self._stage_async = torch.empty(max(total, 1 << 20), dtype=torch.uint8, pin_memory=True)
        if getattr(self, '_d2h_stream', None) is None:
            self._d2h_stream = torch.cuda.Stream()
        side = self._d2h_stream
        side.wait_stream(torch.cuda.current_stream())
        views, at = [], 0
        with torch.cuda.stream(side):
            for t, sz in zip(tensors, sizes):
                v = self._stage_async[at:at + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
                src = t.contiguous()
                src.record_stream(side)
                v.copy_(src, non_blocking=True)
                views.append(v)
                at += sz
            done = torch.cuda.Event()
            done.record(side)
        box = {}

        def finish():
            done.synchronize()
            box['out'] = [v.numpy().copy() for v in views]
        th = threading.Thread(target=finish, daemon=True)
        th.start()

        class _Pending:
            @staticmethod
            def result():
                th.join()
                return box['out']
        return _Pending()

    def upload_replicated(self, A):
        """Host matrix -> the same full matrix on every rank.  With several ranks each one uploads only its row
        shard over its own PCIe link and the shards are all-gathered over NVLink (2 GB: ~3 ms at 8 GPUs) instead
        of every rank pulling the whole matrix through the host ('NCCL scatter')."""
        if isinstance(A, torch.Tensor) and A.is_cuda:
            return A.to(device=self.device, dtype=torch.float32).contiguous()
        if not self.comm.active:
            return self.to_device(A)
        A = A if isinstance(A, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(A, dtype=np.float32))
        n = A.shape[0]
        bounds = self._bounds(n)
        b, e = bounds[self.comm.rank], bounds[self.comm.rank + 1]
        full = torch.empty(A.shape, dtype=torch.float32, device=self.device)
        sizes = {bounds[r + 1] - bounds[r] for r in range(self.comm.world)}
        chunks = int(os.environ.get('VTF_UPLOAD_CHUNKS', '4'))
        if (self.device.type == 'cuda' and len(sizes) == 1 and chunks > 1 and A.is_pinned() and
                (e - b) * A.shape[1] * 4 >= (32 << 20) * chunks):
            # Pipelined: the shard crosses PCIe in `chunks` pieces on a copy stream while the pieces that have arrived are
            # already being exchanged over NVLink (piece c of every rank; the pieces of one exchange are not adjacent in
            # `full`, so the collective takes a list of views).  Hides the all-gather behind the host-to-device copy.
            m = (e - b + chunks - 1) // chunks
            main = torch.cuda.current_stream()
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream()
            cs = self._copy_stream
            cs.wait_stream(main)                                   # `full` was allocated on the main stream
            full.record_stream(cs)
            evs = []
            with torch.cuda.stream(cs):
                for c in range(chunks):
                    lo, hi = b + c * m, min(e, b + (c + 1) * m)
                    if hi > lo:
                        full[lo:hi].copy_(A[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    evs.append(ev)
            for c in range(chunks):
                main.wait_event(evs[c])
                lo_c, hi_c = c * m, min(e - b, (c + 1) * m)
                if hi_c > lo_c:
                    self.comm.allgather_pieces(full, bounds, lo_c, hi_c)
            return full
        if e > b and not A.is_pinned() and hasattr(self.K, 'upload_pageable') and (e - b) * A.shape[1] * 4 >= (16 << 20):
            threads = int(os.environ.get('VTF_UPLOAD_THREADS', str(min(8, max(1, (os.cpu_count() or 1) // self.comm.world)))))
            if threads > 0:
                self.K.upload_pageable(A[b:e].numpy(), torch.float32, threads, out=full[b:e])
                self.comm.allgather_rows(full, bounds)
                return full
        full[b:e].copy_(A[b:e], non_blocking=True)
        self.comm.allgather_rows(full, bounds)
        return full

    @staticmethod
    def _pad16(X, mult=16):
        """Zero columns up to a multiple of 16 (64 for the tensor-core kernels): dot products, norms and distances
        are unchanged, and every kernel can rely on 128-bit loads / whole k-blocks (FaceNet 512, ViT 768/1024 are
        already aligned)."""
        d = X.shape[1]
        pad = (-d) % mult
        return X if pad == 0 else torch.nn.functional.pad(X, (0, pad)).contiguous()

    def _bounds(self, n, align=1):
        return shard_bounds(n, self.comm.world, align)

    # ------------------------------------------------------------------------------ a1 dedup
    def dedup(self, X, threshold, force=None):
        """dupes.py:60-65.  Returns (mins fp32 [N], inds int64 [N], idx int64 [n_dup]) as numpy.
        Large inputs run on the tensor cores (split-fp16, fp32-GEMM accuracy), small ones on the fp32 pipe;
        `force` = 'tc' | 'simt' overrides the choice."""
        X = self.to_device(X)
        N = X.shape[0]
        if N == 0:
            return np.zeros(0, np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64)
        with self.solo_if(self.too_small_to_shard(N)):
            return self._dedup(X, threshold, force)

    def _dedup(self, X, threshold, force):
        N = X.shape[0]
        tc_ok = X.shape[1] <= 4096 and hasattr(self.K, 'cosine_dedup_tc')
        use_tc = (tc_ok and N >= TC_DEDUP_MIN_ROWS) if force is None else (force == 'tc')
        if use_tc and not tc_ok:
            raise ValueError('tensor-core dedup needs d <= 4096')
        X = self._pad16(X, 64 if use_tc else 16)
        d = X.shape[1]
        with self.span('normalize'):
            Xn = self.K.normalize_rows(X)
        with self.span('dedup', flops=float(N) * N * d):
            key_h = None
            if use_tc and not self.comm.active and getattr(self.K, 'dedup_defer', False):
                # one rank: the screen's "units to re-sweep" count comes down WITH the keys (one synchronisation, not two)
                key, n_flag, resweep = self.K.cosine_dedup_tc(Xn, 0, 1, float(np.float32(threshold)), defer=True)
                if n_flag is None:
                    key_h = self.to_host([key])[0]
                else:
                    key_h, nf = self.to_host([key, n_flag])
                    if int(nf[0]) > 0:
                        resweep()
                        key_h = self.to_host([key])[0]
            elif use_tc:
                key = self.K.cosine_dedup_tc(Xn, self.comm.rank, self.comm.world, float(np.float32(threshold)))
            else:
                key = self.K.cosine_dedup(Xn, self.comm.rank, self.comm.world)
        if key_h is None:
            self.comm.allreduce_min(key)
            key_h = self.to_host([key])[0]
        mins = (key_h >> 32).astype(np.uint32).view(np.float32)
        inds = (key_h & 0xffffffff).astype(np.int64)
        idx = (mins <= np.float32(threshold)).nonzero()[0]                    # dupes.py:65, fp32 compare
        return mins, inds, idx

    def dedup_hash(self, H, threshold):
        """dupes.py:55-65, 'hash' branch: H is [N,64] of 0/1 (np.stack of ahash vectors).  Returns
        (mins uint16 [N], inds int64 [N], idx int64 [n_dup]) -- D is cast to uint16 by the reference (:58)."""
        H = np.ascontiguousarray(np.asarray(H) != 0, dtype=np.int64)
        if H.ndim != 2 or H.shape[1] != 64:
            raise ValueError('hash dedup expects [N,64] average hashes (dupes.py:11-15)')
        if H.shape[0] == 0:
            return np.zeros(0, np.uint16), np.zeros(0, np.int64), np.zeros(0, np.int64)
        with self.solo_if(self.too_small_to_shard(H.shape[0])):
            key = self.K.hash_dedup(torch.from_numpy(H).to(self.device), self.comm.rank, self.comm.world)
            self.comm.allreduce_min(key)
        key_h = key.cpu().numpy()
        mins = (key_h >> 32).astype(np.uint16)
        inds = (key_h & 0xffffffff).astype(np.int64)
        idx = (mins <= threshold).nonzero()[0]
        return mins, inds, idx

    def delete_rows(self, X, idx):
        """np.delete(X, idx, axis=0) (dupes.py:68) on the device; returns a device tensor."""
        X = self.to_device(X)
        d = X.shape[1]
        keep = np.ones(X.shape[0], dtype=bool)
        keep[idx] = False
        kept = torch.from_numpy(np.nonzero(keep)[0].astype(np.int64)).to(self.device)
        with self.span('delete_rows'):
            out = self.K.gather_rows(self._pad16(X), kept)
        return out if out.shape[1] == d else out[:, :d].contiguous()

    # ------------------------------------------------------------------------------ a3 classify
    def classify(self, X, R, thr):
        """grouping.py:51-56.  Returns (dist fp32 [N,R], inds int64 [N]) as numpy."""
        X = self._pad16(self.to_device(X))
        R = self._pad16(self.to_device(R))
        with self.solo_if(self.too_small_to_shard(X.shape[0])):
            dist, label = self._classify(X, R, thr)
        on = bool(thr) and thr != -1                                          # grouping.py:53
        if on and thr < 0:
            # the kernel encodes "threshold off" as a negative value; a negative threshold other than -1 is ON in the
            # reference and sends every row to 'other' (cosine distances are clipped to [0, 2], so min >= thr always)
            label[:] = R.shape[0]                                             # grouping.py:55
        return dist, label

    def _classify(self, X, R, thr):
        with self.span('normalize'):
            Xn = self.K.normalize_rows(X)
            Rn = self.K.normalize_rows(R)
        b, e = self._my_rows(X.shape[0])
        on = bool(thr) and thr != -1 and thr >= 0
        nR = R.shape[0]
        dist = self.K.zeros((X.shape[0], nR), torch.float32) if self.comm.active else None
        label = self.K.zeros((X.shape[0],), torch.int64) if self.comm.active else None
        with self.span('classify'):
            d_loc, l_loc = self.K.cosine_classify(Xn[b:e], Rn, float(thr) if on else -1.0)
        if self.comm.active:
            dist[b:e] = d_loc
            label[b:e] = l_loc
            bounds = self._bounds(X.shape[0])
            self.comm.allgather_rows(dist, bounds)
            self.comm.allgather_rows(label, bounds)
        else:
            dist, label = d_loc, l_loc
        dist_h, label_h = self.to_host([dist, label])
        return dist_h, label_h

    def _my_rows(self, n, align=1):
        bounds = self._bounds(n, align)
        return bounds[self.comm.rank], bounds[self.comm.rank + 1]

    # ------------------------------------------------------------------------------ a6 prologue
    def prepare(self, X, normalize=False):
        """KMeans.fit's set-up (SK/cluster/_kmeans.py:1473-1496).  normalize=True: X are un-normalised encoder outputs and
        the encoder's closing F.normalize (encoders/facenet.py:153) is applied here, fused with the column statistics."""
        X = self.upload_replicated(X)
        d_orig = X.shape[1]
        with self.span('prepare'):
            ctx = self._prepare(self._pad16(X, 64 if X.shape[0] >= TC_MIN_ROWS and X.shape[1] <= 1024 else 16), normalize)
        ctx.d_orig = d_orig
        ctx.tol /= d_orig                  # mean over the real columns: zero padding adds no variance
        return ctx

    def _prepare(self, X, normalize=False):
        ctx = FitContext()
        ctx.N, ctx.d = X.shape[0], X.shape[1]
        ctx.bounds = self._bounds(ctx.N)
        ctx.b, ctx.e = ctx.bounds[self.comm.rank], ctx.bounds[self.comm.rank + 1]
        if normalize:
            # one pass: Xn = x / |x| and the column sums of Xn for this rank's rows; the other ranks' rows are gathered
            Xn_loc, s, q = self.K.normalize_col_stats(X[ctx.b:ctx.e])
            if self.comm.active:
                X = torch.empty_like(X)
                X[ctx.b:ctx.e] = Xn_loc
                self.comm.allgather_rows(X, ctx.bounds)
            else:
                X = Xn_loc
        else:
            s, q = self.K.col_stats(X[ctx.b:ctx.e])
        ctx.X = X
        sq = torch.cat([s, q])
        self.comm.allreduce_sum(sq)
        s, q = sq[:ctx.d], sq[ctx.d:]
        ctx.col_sum = s
        ctx.mean64 = s / ctx.N
        ctx.mean32 = ctx.mean64.to(torch.float32)
        var = q / ctx.N - ctx.mean64 * ctx.mean64                              # np.var(X, axis=0), _kmeans.py:292
        tol = var.clamp_min(0).sum() * 1e-4                                    # _kmeans.py:293 (mean taken in prepare())
        if hasattr(self.K, 'center_rows_stats'):
            # centring pass (_kmeans.py:1487-1490) that also yields max |xc| and max |xc_row|^2; ONE small D2H for the three
            # scalars the host needs (tolerance, fixed-point unit / fp16 scales, range of the sweep's norm pieces)
            ctx.Xc, stats = self.K.center_rows_stats(X, ctx.mean32)
            h = torch.cat([tol.reshape(1), stats.to(torch.float64)]).cpu().numpy()
            ctx.tol, ctx.amax, ctx.max_sqn = float(h[0]), float(h[1]), float(h[2])
        else:
            ctx.tol = float(tol.item())
            ctx.Xc = self.K.center_rows(X, ctx.mean32)
        return ctx

    # ------------------------------------------------------------------------------ a7 k-means++
    def kmeans_plusplus(self, ctx, k, rstate):
        """SK/cluster/_kmeans.py:180-278 on the centred matrix; RNG draws come from numpy's MT19937
        on the host in the reference's order, all O(N) arithmetic runs on the device."""
        K, comm = self.K, self.comm
        N, b, e = ctx.N, ctx.b, ctx.e
        rs = np.random.RandomState(rstate)
        trials = n_local_trials(k)
        # rs.choice(N, p=uniform) consumes one random_sample() and picks searchsorted(cdf, u, 'right')
        first = min(int(rs.random_sample() * N), N - 1)                        # _kmeans.py:231
        # the uniform draws of every round (_kmeans.py:249) do not depend on the data: draw them all now, in the
        # reference's order, and keep the rounds themselves on the device -- no host round trip until the end
        U = self.to_device(np.stack([rs.uniform(size=trials) for _ in range(1, k)]).reshape(max(k - 1, 0), trials),
                           torch.float64) if k > 1 else None
        closest = K.empty((N,), torch.float32)
        xsq = None
        if hasattr(K, 'kpp_rownorm64'):                                        # |x|^2 (fp64) of this rank's rows, once per fit
            if getattr(ctx, 'xsq', None) is None:
                ctx.xsq = K.kpp_rownorm64(ctx.Xc[b:e])
            xsq = ctx.xsq
        # distances to the first centre (_kmeans.py:240): every fit of the sweep starts a fresh RandomState(rstate), so all
        # seedings of a k-range pick the SAME first row -- one pass serves them all
        cache = getattr(ctx, 'kpp_first', None)
        if cache is not None and cache[0] == first:
            out, pots = cache[1].clone(), cache[2].clone()
        else:
            out, pots = K.kpp_dist(ctx.Xc[b:e], ctx.Xc[first:first + 1], None, xsq)
            ctx.kpp_first = (first, out.clone(), pots.clone())
        chosen = []
        for r in range(k - 1):
            comm.allreduce_sum(pots)                                           # _kmeans.py:243 / :263
            thr, order, best = K.kpp_round_prep(pots, U[r])                    # :266 of the last round, :249 of this one
            closest[b:e] = out.index_select(0, best)[0]                        # :272
            comm.allgather_rows(closest, ctx.bounds)
            if r > 0:
                chosen.append(cand.index_select(0, best))                      # :270
            cand = K.kpp_unsort(K.kpp_pick(closest, thr), order, N)            # :250-254
            out, pots = K.kpp_dist(ctx.Xc[b:e], K.gather_rows(ctx.Xc, cand), closest[b:e], xsq)      # :257-263
        seeds = [first]
        if k > 1:
            comm.allreduce_sum(pots)
            best = K.kpp_round_prep(pots, U[0])[2]                             # :266 (the last centre's distances are not needed)
            chosen.append(cand.index_select(0, best))
            seeds.extend(int(v) for v in torch.cat(chosen).cpu().numpy())      # the one D2H of the seeding
        return np.asarray(seeds, dtype=np.int64)

    # ------------------------------------------------------------------------------ a6 Lloyd
    def lloyd(self, ctx, C0, max_iter=300):
        """SK/cluster/_kmeans.py:629-758 (single Lloyd run) from centred initial centres C0 [k,d]: the lock-step loop
        with one labelling, so that a fit gives the same bits whether it runs alone or next to the other k."""
        return self._lloyd_group(ctx, [C0], max_iter)[0]

    def _amax(self, ctx):
        """max |x| of the centred matrix, once per fit context (fixed-point unit of the M-step sums, fp16 scales)."""
        if getattr(ctx, 'amax', None) is None:
            ctx.amax = float(self.K.absmax(ctx.Xc).item())                     # every rank holds the whole matrix
        return ctx.amax if ctx.amax > 0.0 else 1.0

    def _fixed_point(self, ctx):
        """(2^S, 2^-S): the unit of the exact 64-bit fixed-point cluster sums, |sum| <= N * max|x| * 2^S < 2^62."""
        S = 61 - math.ceil(math.log2(max(ctx.N, 2) * self._amax(ctx)))
        return 2.0 ** S, 2.0 ** -S

    @staticmethod
    def _pack_groups(ks, d):
        """Labellings that share one pass: at most 32 per call, sum(k) <= 96 and all centres within ~160 KB of
        shared memory (k = 2..6 at d = 512 is one group; k = 2..32 at d = 1024 becomes several)."""
        groups, cur, tot = [], [], 0
        for j, k in enumerate(ks):
            if cur and (len(cur) >= 32 or tot + k > 96 or (tot + k) * d * 4 > 160 * 1024):
                groups.append(cur)
                cur, tot = [], 0
            cur.append(j)
            tot += k
        if cur:
            groups.append(cur)
        return groups

    def lloyd_multi(self, ctx, C0_list, max_iter=300):
        """Every fit of the k-range in lock step: one read of X per iteration serves all k that are still iterating
        (each fit keeps its own stopping rule and freezes when it is done, so results equal the separate fits)."""
        out = [None] * len(C0_list)
        ks = [int(c.shape[0]) for c in C0_list]
        if self._use_wide(ctx, ks):
            for g0 in range(0, len(ks), 32):
                grp = list(range(g0, min(g0 + 32, len(ks))))
                for j, r in zip(grp, self._lloyd_wide(ctx, [C0_list[j] for j in grp], max_iter)):
                    out[j] = r
            return out
        for grp in self._pack_groups(ks, ctx.d):
            for j, r in zip(grp, self._lloyd_group(ctx, [C0_list[j] for j in grp], max_iter)):
                out[j] = r
        return out

    def _lloyd_group(self, ctx, C0s, max_iter):
        """Lock-step Lloyd for a group of labellings: per iteration ONE E-step pass over X (register-tiled fp32 kernel,
        all centres of the group) and ONE M-step pass (cell-wise exact fixed-point sums, csrc/cells.cu), then the centre
        update and the per-labelling stopping rules of SK/cluster/_kmeans.py:700-758."""
        K, comm = self.K, self.comm
        N, d, b, e = ctx.N, ctx.d, ctx.b, ctx.e
        nL = len(C0s)
        ks = [int(c.shape[0]) for c in C0s]
        off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
        ktot = int(off[-1])
        stride = 2 + max(ks)
        C_all = torch.cat([c.contiguous() for c in C0s]).contiguous()
        labels_all = torch.full((nL, N), -1, dtype=torch.int32, device=self.device)
        Xs = ctx.Xc[b:e]
        cells = hasattr(K, 'cell_sums') and max(ks) <= K.CELL_MAX_K and d % 2 == 0
        if cells:
            two_s, inv_two_s = self._fixed_point(ctx)
            acc = K.zeros((ktot, d), torch.int64)
        active = (1 << nL) - 1
        n_iter, strict, counts_fin, csums = [0] * nL, [False] * nL, [None] * nL, [None] * nL
        for it in range(max_iter):
            cc = K.zeros((ktot + nL,), torch.int64)                            # counts | changed: one memset, one all-reduce
            counts, changed = cc[:ktot], cc[ktot:]
            K.lloyd_assign_multi(Xs, C_all, off, active, labels_all, b, counts, changed)
            if cells:
                K.cell_sums(Xs, labels_all, b, off, active, two_s, acc)
                acc_g = acc
                if comm.active:
                    acc_g = acc.clone()
                    comm.allreduce_sum(acc_g)                                  # integers: exact in any order
                    comm.allreduce_sum(cc)
                C_new, stats = K.wide_update(acc_g, inv_two_s, counts, changed, off, active, C_all, stride)
            else:
                sums = K.cluster_sums_multi(Xs, labels_all, b, off, active)
                comm.allreduce_sum(sums)
                comm.allreduce_sum(cc)
                C_new, _, stats = K.lloyd_update_multi(sums, counts, changed, off, active, C_all, stride)
            st = stats.cpu().numpy()                                           # the one D2H of the iteration
            for L in range(nL):
                if not (active >> L) & 1:
                    continue
                lo, hi = int(off[L]), int(off[L + 1])
                stL = st[L, :2 + hi - lo]
                sumsL = None
                if (stL[2:] == 0).any():                                       # empty cluster: relocate (rare)
                    sumsL = (acc_g[lo:hi].to(torch.float64) * inv_two_s) if cells else sums[lo:hi].contiguous()
                    cn, _, stL = self._relocate_empty(ctx, labels_all[L], C_all[lo:hi].contiguous(), sumsL,
                                                      counts[lo:hi].contiguous(), changed[L:L + 1].contiguous(), stL)
                    C_new[lo:hi] = cn
                n_iter[L] = it + 1
                if stL[1] == 0:                                                # labels == labels_old, _kmeans.py:723
                    strict[L] = True
                    active &= ~(1 << L)
                    counts_fin[L] = counts[lo:hi].clone()
                    if sumsL is None:      # the sums of this pass ARE the per-cluster sums of the final labels: CH / DB reuse them
                        csums[L] = (acc_g[lo:hi].to(torch.float64) * inv_two_s) if cells else sums[lo:hi].clone()
                elif stL[0] <= ctx.tol:                                        # _kmeans.py:731-738
                    active &= ~(1 << L)
            C_all = C_new                                                      # _kmeans.py:721
            if active == 0:
                break
        redo = sum(1 << L for L in range(nL) if not strict[L])
        if redo:                                                               # _kmeans.py:742-754
            counts = K.zeros((ktot,), torch.int64)
            changed = K.zeros((nL,), torch.int64)
            K.lloyd_assign_multi(Xs, C_all, off, redo, labels_all, b, counts, changed)
            comm.allreduce_sum(counts)
            for L in range(nL):
                if not strict[L]:
                    counts_fin[L] = counts[int(off[L]):int(off[L + 1])].clone()
            if cells:
                # the per-cluster sums of the FINAL labels of the tolerance-stopped fits: CH / DB need their means
                K.cell_sums(Xs, labels_all, b, off, redo, two_s, acc)
                acc_g = acc
                if comm.active:
                    acc_g = acc.clone()
                    comm.allreduce_sum(acc_g)
                for L in range(nL):
                    if not strict[L]:
                        csums[L] = acc_g[int(off[L]):int(off[L + 1])].to(torch.float64) * inv_two_s
        # One pass measures every row against the centre of its cluster: sum(sq) is the inertia (_k_means_common.pyx:94-124).
        # For a fit that stopped on 'labels unchanged' the centres are the fp32 means of its final labels, so the same numbers
        # are CH's intra dispersion and DB's distance sums; a tolerance-stopped fit rides along a second time, measured
        # against the fp64 means of its final labels (a "virtual" labelling with the same label row).
        ext = [L for L in range(nL) if not strict[L] and csums[L] is not None]
        C64 = C_all.to(torch.float64)
        off_x = off
        if ext:
            means = [csums[L] / counts_fin[L].to(torch.float64).clamp_min(1).unsqueeze(1) for L in ext]
            C64 = torch.cat([C64] + means).contiguous()
            off_x = np.concatenate([off, off[-1] + np.cumsum([ks[L] for L in ext])]).astype(np.int32)
        kx = int(off_x[-1])
        sq, rt = self._dispersion(ctx, Xs, labels_all, b, off_x, [ks[L] for L in range(nL)] + [ks[L] for L in ext], C64,
                                  list(range(nL)) + ext)
        sq_h = sq.cpu().numpy()
        res = []
        for L in range(nL):
            lo, hi = int(off[L]), int(off[L + 1])
            lab = labels_all[L]
            comm.allgather_rows(lab, ctx.bounds)
            centers = (C_all[lo:hi] + ctx.mean32)[:, :ctx.d_orig].contiguous()  # _kmeans.py:1543-1546
            disp = None
            if strict[L]:
                disp = (sq[lo:hi], rt[lo:hi])
            elif L in ext:
                xlo = int(off_x[nL + ext.index(L)])
                disp = (sq[xlo:xlo + ks[L]], rt[xlo:xlo + ks[L]])
            res.append(KMeansResult(k=hi - lo, labels=lab, centers=centers, inertia=float(sq_h[lo:hi].sum()),
                                    n_iter=n_iter[L], counts=counts_fin[L], strict=strict[L], csums=csums[L], disp=disp))
        return res

    # ------------------------------------------------------------------------------ a6, wide k-ranges on the tensor cores
    def _use_wide(self, ctx, ks):
        # decided from global quantities only: every rank must take the same path (the paths differ in their collectives)
        ok = (getattr(self.K, 'wide_lloyd', False) and ctx.d % 64 == 0 and ctx.d <= 4096 and max(ks) <= 100 and
              ctx.N >= self.comm.world)
        if self.wide is not None:
            if self.wide and not ok:
                raise ValueError('wide Lloyd needs d % 64 == 0, d <= 4096 and k <= 100')
            return bool(self.wide)
        return ok and sum(ks) >= WIDE_MIN_SUM_K and ctx.N >= WIDE_MIN_ROWS * self.comm.world

    def _wide_operand(self, ctx):
        """Row operand of the wide E-step, built once per fit context: fp16 hi / lo pieces of scale * Xc for this
        rank's rows, row norms for the error bound, and the fixed-point unit 2^S of the M-step sums."""
        if getattr(ctx, 'wide', None) is None:
            K = self.K
            Xs = ctx.Xc[ctx.b:ctx.e]
            n = Xs.shape[0]
            P = ((n + 255) // 256) * 256
            amax = self._amax(ctx)
            scale = 2.0 ** min(60, math.floor(math.log2(2048.0 / amax)))   # hi keeps 11 bits, lo the next 11; 1/scale^2 stays a normal fp32
            Xhi, Xlo = K.split_fp16(Xs, P, scale)
            rown = K.zeros((P,), torch.float32)
            rown[:n] = torch.sqrt(K.row_sqnorms(Xs))
            two_s, inv_two_s = self._fixed_point(ctx)
            ctx.wide = dict(Xhi=Xhi, Xlo=Xlo, rown=rown, scale=scale, two_s=two_s, inv_two_s=inv_two_s, n=n)
        return ctx.wide

    @staticmethod
    def _wide_columns(off, ks, batch):
        """Column table of one launch: the centres of the labellings in `batch`, concatenated and padded to whole
        256-column tiles.  info = flags | local index << 4 | labelling << 20 (flag 2: last centre of its labelling)."""
        ncol = sum(ks[L] for L in batch)
        Kp = ((ncol + 255) // 256) * 256
        src = np.full(Kp, -1, dtype=np.int32)
        info = np.zeros(Kp, dtype=np.uint32)
        col0 = np.full(len(ks), -1, dtype=np.int32)
        c = 0
        for L in batch:
            k = ks[L]
            col0[L] = c
            src[c:c + k] = np.arange(off[L], off[L] + k, dtype=np.int32)
            info[c:c + k] = (np.arange(k, dtype=np.uint32) << 4) | np.uint32(L << 20)
            info[c + k - 1] |= 2
            c += k
        return src, info, col0

    def _wide_estep(self, ctx, C_all, off, ks, active, labels_new, n_fix, cache):
        """labels_new[L] = argmin_c |x - c|^2 for every active labelling L (SK/cluster/_k_means_lloyd.pyx:196-213):
        tensor-core scores, then exact fp32 re-evaluation of the rows whose two best centres are closer than the
        error bound of the tensor-core sum."""
        K = self.K
        wx = self._wide_operand(ctx)
        # Tolerance BETWEEN two scores (the candidate test compares a score with the running best): each score is off by at
        # most one ulp per accumulating MMA (3 MMAs per K = 16 step, truncating fp32 accumulation) plus the dropped lo.lo
        # product and the residual of the two-piece split (3 * 2^-22), all relative to |x| |c|; two scores -> twice that.
        eps = 2.0 * (3.0 * (ctx.d // 16) * 2.0 ** -23 + 3.0 * 2.0 ** -22) * self.wide_eps_scale
        plan = cache.get(active)
        if plan is None:
            batches, cur, ncol = [], [], 0
            for L in range(len(ks)):
                if not (active >> L) & 1:
                    continue
                if cur and ncol + ks[L] > WIDE_MAX_COLS:
                    batches.append(cur)
                    cur, ncol = [], 0
                cur.append(L)
                ncol += ks[L]
            if cur:
                batches.append(cur)
            plan = []
            for batch in batches:
                src, info, col0 = self._wide_columns(off, ks, batch)
                plan.append((self.to_device(src, torch.int32), torch.from_numpy(info.view(np.int32)).to(self.device), col0,
                             sum(ks[L] for L in batch)))
            cache.clear()                  # the active set only ever shrinks: older plans are dead
            cache[active] = plan
        Xs = ctx.Xc[ctx.b:ctx.e]
        n_fix.zero_()
        for src, info, col0, n_cols in plan:
            Bhi, Blo, tab, cmax = K.wide_pack_centres(C_all, src, info, wx['scale'])
            with self.span('wide_assign', flops=2.0 * wx['n'] * n_cols * ctx.d):
                K.wide_assign(wx['Xhi'], wx['Xlo'], wx['n'], Bhi, Blo, n_cols, tab, wx['rown'], cmax, wx['scale'], eps, labels_new)
            K.wide_refine(Xs, C_all, tab, off, col0, labels_new, wx['n'], n_fix)
        return n_fix.clone() if self.profile else None

    def _lloyd_wide(self, ctx, C0s, max_iter):
        """_lloyd_group with the E-step on the tensor cores and the M-step in exact fixed point, updated from the rows
        that changed (csrc/lloyd_tc.cu).  Same stopping rules per labelling, same results."""
        K, comm = self.K, self.comm
        N, d, b, e = ctx.N, ctx.d, ctx.b, ctx.e
        n = e - b
        nL = len(C0s)
        ks = [int(c.shape[0]) for c in C0s]
        off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
        ktot = int(off[-1])
        stride = 2 + max(ks)
        wx = self._wide_operand(ctx)
        two_s, inv_two_s = wx['two_s'], wx['inv_two_s']
        C_all = torch.cat([c.contiguous() for c in C0s]).contiguous()
        labels_all = torch.full((nL, N), -1, dtype=torch.int32, device=self.device)
        labels_new = torch.zeros((nL, max(n, 1)), dtype=torch.int32, device=self.device)
        Xs = ctx.Xc[b:e]
        acc = K.zeros((ktot, d), torch.int64)
        n_fix = K.zeros((1,), torch.int32)
        cap = max(1024, n // 4)                    # more changes than this: recompute the labelling's sums from the labels
        items = K.empty((nL * cap, 4), torch.int32)
        n_items = K.zeros((1,), torch.int32)
        sum_groups = self._pack_groups(ks, 0)
        cache = {}
        active = (1 << nL) - 1
        n_iter, strict, counts_fin = [0] * nL, [False] * nL, [None] * nL
        log = dict(iters=0, fixed=[], listed=[])
        kmax = max(ks)
        valid = np.arange(kmax)[None, :] < np.asarray(ks)[:, None]              # [nL, kmax]: real cluster slots of a stats row

        def m_step(which, counts, changed):
            # cluster sizes / change counts of labels_new, labels_all = labels_new, and the sums brought up to date: from the
            # change list where few rows moved, from scratch otherwise (decided per labelling on the device)
            K.wide_count(labels_new, labels_all, b, n, off, which, counts, changed)
            K.wide_commit(labels_new, labels_all, b, n, off, which, changed, cap, items, n_items)
            K.wide_sums_delta(Xs, items, n_items, off, two_s, acc)
            for grp in sum_groups:
                if any((which >> L) & 1 for L in grp):
                    K.wide_sums_full(Xs, labels_all, b, off, which, grp[0], grp[-1] + 1, changed, cap, two_s, acc)
        for it in range(max_iter):
            fx = self._wide_estep(ctx, C_all, off, ks, active, labels_new, n_fix, cache)
            cc = K.zeros((ktot + nL,), torch.int64)                            # counts | changed
            counts, changed = cc[:ktot], cc[ktot:]
            m_step(active, counts, changed)
            if self.profile:
                log['fixed'].append(fx)
                log['listed'].append(n_items.clone())
            acc_g = acc
            if comm.active:
                acc_g = acc.clone()
                comm.allreduce_sum(acc_g)                                      # integers: exact in any order
                comm.allreduce_sum(cc)
            C_new, stats = K.wide_update(acc_g, inv_two_s, counts, changed, off, active, C_all, stride)
            st = stats.cpu().numpy()                                           # the one D2H of the iteration
            # the per-labelling decisions for all labellings at once (31 of them, 227 iterations in config 4: a Python loop
            # over every active labelling was 0.15 ms of GPU idle per iteration); only labellings where something happens
            # -- an empty cluster, a stop -- are visited one by one
            act = np.array([(active >> L) & 1 for L in range(nL)], dtype=bool)
            empty = ((st[:, 2:2 + kmax] == 0) & valid).any(axis=1) & act
            for L in np.nonzero(act)[0]:
                n_iter[L] = it + 1
            for L in np.nonzero(act & (empty | (st[:, 1] == 0) | (st[:, 0] <= ctx.tol)))[0]:
                L = int(L)
                lo, hi = int(off[L]), int(off[L + 1])
                stL = st[L, :2 + hi - lo]
                if empty[L]:                                                   # empty cluster: relocate (rare)
                    sums = acc_g[lo:hi].to(torch.float64) * inv_two_s
                    cn, _, stL = self._relocate_empty(ctx, labels_all[L], C_all[lo:hi].contiguous(), sums,
                                                      counts[lo:hi].contiguous(), changed[L:L + 1].contiguous(), stL)
                    C_new[lo:hi] = cn
                if stL[1] == 0:                                                # labels == labels_old, _kmeans.py:723
                    strict[L] = True
                    active &= ~(1 << L)
                    counts_fin[L] = counts[lo:hi].clone()
                elif stL[0] <= ctx.tol:                                        # _kmeans.py:731-738
                    active &= ~(1 << L)
            C_all = C_new                                                      # _kmeans.py:721
            log['iters'] = it + 1
            if active == 0:
                break
        redo = sum(1 << L for L in range(nL) if not strict[L])
        if redo:                                                               # _kmeans.py:742-754
            self._wide_estep(ctx, C_all, off, ks, redo, labels_new, n_fix, {})
            counts = K.zeros((ktot,), torch.int64)
            changed = K.zeros((nL,), torch.int64)
            m_step(redo, counts, changed)              # the sums follow the final labels: CH / DB reuse them
            comm.allreduce_sum(counts)
            for L in range(nL):
                if not strict[L]:
                    counts_fin[L] = counts[int(off[L]):int(off[L + 1])].clone()
        acc_g = acc
        if comm.active:
            acc_g = acc.clone()
            comm.allreduce_sum(acc_g)
        csums = acc_g.to(torch.float64) * inv_two_s     # per-cluster sums of the CENTRED rows, consistent with the labels
        if self.profile:
            log['fixed'] = [int(v.item()) for v in log['fixed']]
            log['listed'] = [int(v.item()) for v in log['listed']]
            log.update(rows=n, labellings=nL, sum_k=ktot, cap=cap)
            self.wide_stats = log
        res = [None] * nL
        C64_all = C_all.to(torch.float64)
        for grp in sum_groups:                                                 # _k_means_common.pyx:94-124
            g0, g1 = grp[0], grp[-1] + 1
            lo0, hi0 = int(off[g0]), int(off[g1])
            # every labelling against its centres (inertia; for 'labels unchanged' fits also CH's intra dispersion and DB's
            # distance sums); tolerance-stopped fits a second time against the fp64 means of their final labels
            ext = [L for L in grp if not strict[L]][:32 - len(grp)]
            C64 = C64_all[lo0:hi0]
            off_x = (off[g0:g1 + 1] - off[g0]).astype(np.int32)
            if ext:
                means = [csums[int(off[L]):int(off[L + 1])] / counts_fin[L].to(torch.float64).clamp_min(1).unsqueeze(1) for L in ext]
                C64 = torch.cat([C64] + means).contiguous()
                off_x = np.concatenate([off_x, off_x[-1] + np.cumsum([ks[L] for L in ext])]).astype(np.int32)
            sq, rt = self._dispersion(ctx, Xs, labels_all, b, off_x, [ks[L] for L in grp] + [ks[L] for L in ext], C64.contiguous(),
                                      list(grp) + ext)
            sq_h = sq.cpu().numpy()
            for L in grp:
                lo, hi = int(off[L]) - lo0, int(off[L + 1]) - lo0
                lab = labels_all[L]
                comm.allgather_rows(lab, ctx.bounds)
                centers = (C_all[lo0 + lo:lo0 + hi] + ctx.mean32)[:, :ctx.d_orig].contiguous()   # _kmeans.py:1543-1546
                disp = None
                if strict[L]:
                    disp = (sq[lo:hi], rt[lo:hi])
                elif L in ext:
                    xlo = int(off_x[len(grp) + ext.index(L)])
                    disp = (sq[xlo:xlo + ks[L]], rt[xlo:xlo + ks[L]])
                res[L] = KMeansResult(k=hi - lo, labels=lab, centers=centers, inertia=float(sq_h[lo:hi].sum()),
                                      n_iter=n_iter[L], counts=counts_fin[L], strict=strict[L],
                                      csums=csums[lo0 + lo:lo0 + hi], disp=disp)
        return res

    def _dispersion(self, ctx, Xs, labels_all, b, off_x, ks_x, C64, row_of):
        """(sq, rt) fp64 [sum k], all-reduced: per centre slot the sums of |x - c|^2 and |x - c| over this fit context's rows;
        labelling q measures against centre slots [off_x[q], off_x[q+1]) with label row row_of[q] of labels_all (rows may be
        shared).  Narrow k-ranges run cell-wise with exact fixed-point accumulation (csrc/cells.cu); the units follow from
        the largest squared row norm of the centred matrix (centres are means of rows, so |x - c|^2 <= 4 max |x|^2)."""
        K, comm = self.K, self.comm
        kx = int(off_x[-1])
        max_sqn = getattr(ctx, 'max_sqn', None)
        if max_sqn and max_sqn > 0.0 and hasattr(K, 'cell_dispersion') and K.cell_dispersion_ok(ctx.d, ks_x, row_of):
            bound = 4.5 * max_sqn
            sq_unit = 2.0 ** math.floor(math.log2(2.0 ** 62 / (max(ctx.N, 2) * bound)))
            rt_unit = 2.0 ** math.floor(math.log2(2.0 ** 62 / (max(ctx.N, 2) * math.sqrt(bound))))
            out = K.cell_dispersion(Xs, labels_all, b, off_x, C64, sq_unit, rt_unit, row_of)
            comm.allreduce_sum(out)                                            # integers: exact in any order
            return out[:kx].to(torch.float64) / sq_unit, out[kx:].to(torch.float64) / rt_unit
        labels_x = labels_all[row_of].contiguous()
        sq, rt = K.cluster_dispersion_multi(Xs, labels_x, b, off_x, (1 << labels_x.shape[0]) - 1, C64)
        both = torch.cat([sq, rt])
        comm.allreduce_sum(both)
        return both[:kx], both[kx:]

    def _relocate_empty(self, ctx, labels, C_old, sums, counts, n_changed, st):
        """SK/cluster/_k_means_common.pyx:167-211: give every empty cluster the point farthest from
        its own (old) centre.  Rare; O(k) host steps around one extra pass over X."""
        K, comm = self.K, self.comm
        b, e = ctx.b, ctx.e
        empty = np.nonzero(st[2:] == 0)[0]
        far = K.empty((ctx.N,), torch.float32)
        _, _, far_loc = K.cluster_dispersion(ctx.Xc[b:e], labels[b:e], C_old.to(torch.float64), want_far=True)
        far[b:e] = far_loc
        comm.allgather_rows(far, ctx.bounds)
        vals, idxs = torch.topk(far, len(empty))
        if float(vals[0].item()) == 0.0:                                       # :190-193
            return K.lloyd_update(sums, counts, n_changed, C_old)[:2] + (st,)
        comm.allgather_rows(labels, ctx.bounds)
        sums = sums.clone()
        counts = counts.clone()
        for new_id, far_idx in zip(empty.tolist(), idxs.tolist()):
            old_id = int(labels[far_idx].item())
            x = ctx.Xc[far_idx].to(torch.float64)
            sums[old_id] -= x
            sums[new_id] = x
            counts[new_id] = 1
            counts[old_id] -= 1
        C_new, C64, stats = K.lloyd_update(sums, counts, n_changed, C_old)
        return C_new, C64, stats.cpu().numpy()

    def kmeans_sweep(self, X, clusters, rstate, inits=None, ctx=None, normalize=False):
        """grouping.py:99-101 for every k.  `inits[j]` (un-centred [k,d]) replaces k-means++ for k_j.

        Seeds are prefix-shared: every fit starts a fresh RandomState(rstate), so for equal
        n_local_trials the seed list of k is a prefix of that of k+1."""
        ctx = ctx if ctx is not None else self.prepare(X, normalize)
        results = []
        seed_cache = {}
        if inits is None:
            by_trials = {}
            for k in clusters:
                by_trials.setdefault(n_local_trials(k), []).append(k)
            for trials, ks in by_trials.items():
                with self.span('kmeans++'):
                    seed_cache[trials] = self.kmeans_plusplus(ctx, max(ks), rstate)
        C0s, seeds_all = [], []
        for j, k in enumerate(clusters):
            if inits is None:
                seeds = seed_cache[n_local_trials(k)][:k]
                C0 = self.K.gather_rows(ctx.Xc, torch.from_numpy(seeds).to(self.device))
            else:
                seeds = None
                C0 = (self._pad16(self.to_device(inits[j]), ctx.d) - ctx.mean32).contiguous()
            C0s.append(C0)
            seeds_all.append(seeds)
        with self.span('lloyd'):
            results = self.lloyd_multi(ctx, C0s) if self.lockstep else [self.lloyd(ctx, c) for c in C0s]
        for r, sd in zip(results, seeds_all):
            r.seeds = sd
        return ctx, results

    # ------------------------------------------------------------------------------ a9 / a10
    def ch_db_scores(self, ctx, labels, k, counts=None):
        """calinski_harabasz_score and davies_bouldin_score for one labelling (see ch_db_scores_multi)."""
        return self.ch_db_scores_multi(ctx, [labels], [k], [counts])[0]

    def ch_db_scores_multi(self, ctx, labels_list, ks, counts_list=None, csums_list=None, disp_list=None):
        """calinski_harabasz_score and davies_bouldin_score (SK/metrics/cluster/_unsupervised.py:333-487) for every
        labelling of the k-range with TWO passes over X in total: one for all per-cluster sums, one for all
        distances to the cluster means (the reference makes both passes once per k)."""
        K, comm = self.K, self.comm
        b, e = ctx.b, ctx.e
        if counts_list is None:
            counts_list = [None] * len(ks)
        counts_list = [c if c is not None else torch.bincount(l.to(torch.int64), minlength=k)
                       for c, l, k in zip(counts_list, labels_list, ks)]
        out = [None] * len(ks)
        # Both scores are translation invariant.  When the fits left their per-cluster sums of the centred rows (the wide Lloyd
        # path keeps them exact and consistent with the final labels) the first pass over X is not needed: work on Xc.
        centred = csums_list is not None and all(c is not None for c in csums_list)
        # A fit that stopped on 'labels unchanged' (SK/cluster/_kmeans.py:723) has centres = fp32(means of its final labels), and
        # its inertia pass already measured every row against them: sq = CH's intra dispersion and rt = DB's distance sums up
        # to the fp32 rounding of the means (a second-order effect, ~1e-14 relative), so NO pass over X is left for the scores.
        reuse = centred and disp_list is not None and all(dd is not None for dd in disp_list)
        Xsrc = ctx.Xc if centred else ctx.X
        gmean = (ctx.col_sum / ctx.N - ctx.mean32.to(torch.float64)) if centred else ctx.mean64
        for grp in self._pack_groups(ks, 0):
            gk = [ks[j] for j in grp]
            off = np.concatenate([[0], np.cumsum(gk)]).astype(np.int32)
            nL = len(grp)
            labels_all = None if reuse else torch.stack([labels_list[j] for j in grp]).contiguous()
            full = (1 << nL) - 1
            if centred:
                sums = torch.cat([csums_list[j] for j in grp]).contiguous()
            else:
                sums = K.cluster_sums_multi(Xsrc[b:e], labels_all, b, off, full)
                comm.allreduce_sum(sums)
            cnt_all = torch.cat([counts_list[j] for j in grp]).to(torch.float64)
            means = (sums / cnt_all.clamp_min(1).unsqueeze(1)).contiguous()
            if reuse:
                both = torch.cat([disp_list[j][0] for j in grp] + [disp_list[j][1] for j in grp])
            else:
                sq, rt = K.cluster_dispersion_multi(Xsrc[b:e], labels_all, b, off, full, means)
                both = torch.cat([sq, rt])
                comm.allreduce_sum(both)
            ktot = int(off[-1])
            d_ = means.shape[1]
            flat = torch.cat([both.to(torch.float64), cnt_all, gmean.to(torch.float64), means.reshape(-1)]).cpu().numpy()   # one D2H
            sq_h, rt_h, cnt_h = flat[:ktot], flat[ktot:2 * ktot], flat[2 * ktot:3 * ktot]
            mean_h, means_h = flat[3 * ktot:3 * ktot + d_], flat[3 * ktot + d_:].reshape(ktot, d_)
            for L, j in enumerate(grp):
                lo, hi = int(off[L]), int(off[L + 1])
                present = cnt_h[lo:hi] > 0
                n_labels = int(present.sum())
                self._check_n_labels(n_labels, ctx.N)
                m, cnt, sqL, rtL = means_h[lo:hi][present], cnt_h[lo:hi][present], sq_h[lo:hi][present], rt_h[lo:hi][present]
                extra = float((cnt * ((m - mean_h) ** 2).sum(1)).sum())          # CH (:394-403)
                intra = float(sqL.sum())
                ch = 1.0 if intra == 0.0 else extra * (ctx.N - n_labels) / (intra * (n_labels - 1.0))
                intra_d = rtL / cnt                                               # DB (:467-487)
                # centroid distances as sklearn takes them (pairwise_distances on the fp64 centroids = the Gram form of
                # SK/metrics/pairwise.py:376-427, exact zeros on the diagonal): k x d GEMM instead of a k x k x d temporary
                # (k = 2..32 at d = 1024: 94 MB of broadcast temporaries, 60 ms of host time)
                mm = (m * m).sum(1)
                cd = mm[:, None] + mm[None, :] - 2.0 * (m @ m.T)
                np.fill_diagonal(cd, 0.0)
                cd = np.sqrt(np.maximum(cd, 0.0))
                if np.allclose(intra_d, 0) or np.allclose(cd, 0):
                    db = 0.0
                else:
                    cd[cd == 0] = np.inf
                    comb = intra_d[:, None] + intra_d
                    db = float(np.mean(np.max(comb / cd, axis=1)))
                out[j] = (float(ch), float(db))
        return out

    @staticmethod
    def _check_n_labels(n_labels, n_samples):
        if not 1 < n_labels < n_samples:                                       # _unsupervised.py:31-46
            raise ValueError('Number of labels is %d. Valid values are 2 to n_samples - 1 (inclusive)' % n_labels)

    # ------------------------------------------------------------------------------ a8 silhouette
    def silhouette_scores(self, ctx, labels_list, ks, counts_list=None, force=None):
        """silhouette_score for several labellings of the same X with ONE sweep over the Gram matrix
        (the reference recomputes all N^2 distances per k, grouping.py:104-105)."""
        if counts_list is None:
            counts_list = [torch.bincount(l.to(torch.int64), minlength=k) for l, k in zip(labels_list, ks)]
        for c in counts_list:
            self._check_n_labels(int((c > 0).sum().item()), ctx.N)
        d = ctx.d
        tc_ok = d <= 1024
        use_tc = (tc_ok and ctx.N >= TC_MIN_ROWS) if force is None else (force == 'tc')
        if use_tc and not tc_ok:
            raise ValueError('tensor-core silhouette needs d <= 1024')
        if use_tc:
            return self._silhouette_tc(ctx, labels_list, ks, counts_list)
        return self._silhouette_exact(ctx, labels_list, ks, counts_list)

    def _silhouette_exact(self, ctx, labels_list, ks, counts_list):
        K, comm = self.K, self.comm
        N = ctx.N
        bounds = self._bounds(N, 64)
        b, e = bounds[comm.rank], bounds[comm.rank + 1]
        sqn = K.row_sqnorms(ctx.Xc)
        out = K.zeros((len(ks),), torch.float64)
        for j, (lab, k, cnt) in enumerate(zip(labels_list, ks, counts_list)):
            rowsum = K.zeros((N, k), torch.float64)
            K.sil_rowsums_exact(ctx.Xc, sqn, lab, k, b, e, rowsum)
            K.sil_finish_rows(rowsum, lab, cnt, b, e, out[j:j + 1])
        comm.allreduce_sum(out)
        return (out / N).cpu().numpy().tolist()

    def build_cells(self, labels_list, ks):
        """Group rows by their joint label tuple over all labellings ("cells"), pad every cell to a
        multiple of 32 columns and lay the rows out cell by cell.  Returns
        (perm int64 [P], chunk_cell int32 [P/32], cell_to_cluster int32 [nL, G], G, P)."""
        dev = self.device
        N = labels_list[0].numel()
        tuples = 1
        for k in ks:
            tuples = min(tuples * int(k), 1 << 40)
        if hasattr(self.K, 'build_cells') and tuples <= self.K.BUILD_CELLS_MAX_KEYS and len(ks) <= 32:
            # narrow k-ranges (k = 2..6: 720 tuples): keys, histogram, layout and scatter in three launches, no sort
            perm, chunk_cell, c2c, G, P = self.K.build_cells(torch.stack([l.to(torch.int32) for l in labels_list]).contiguous(),
                                                            [int(k) for k in ks])
            # G, P, chunk_cell and c2c are functions of the label histogram (the same on every rank); the ORDER of the rows
            # inside a cell comes from atomics and differs from rank to rank -- and the ranks must agree on it: each one sweeps
            # a share of the positions and their cell sums are added position by position.  Rank 0's layout is the layout.
            self.comm.broadcast(perm, src=0)
            return perm, chunk_cell, c2c, G, P
        key = torch.zeros(N, dtype=torch.int64, device=dev)
        span = 1
        for lab, k in zip(labels_list, ks):
            if span * k >= (1 << 62):
                _, key = torch.unique(key, return_inverse=True)
                span = int(key.max().item()) + 1
            key = key * k + lab.to(torch.int64)
            span *= k
        _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
        G = cnt.numel()
        order = torch.argsort(inv, stable=True)
        pc = ((cnt + 31) // 32) * 32
        poff = torch.cumsum(pc, 0) - pc
        start = torch.cumsum(cnt, 0) - cnt
        total = int(pc.sum().item())
        P = ((total + 255) // 256) * 256
        g_sorted = inv[order]
        dest = poff[g_sorted] + (torch.arange(N, device=dev) - start[g_sorted])
        perm = torch.full((P,), -1, dtype=torch.int64, device=dev)
        perm[dest] = order
        chunk_cell = torch.zeros(P // 32, dtype=torch.int32, device=dev)
        chunk_cell[:total // 32] = torch.repeat_interleave(torch.arange(G, device=dev, dtype=torch.int32), pc // 32)
        rep = torch.empty(G, dtype=torch.int64, device=dev)
        rep[inv] = torch.arange(N, device=dev)
        c2c = torch.stack([lab[rep].to(torch.int32) for lab in labels_list]).contiguous()
        return perm, chunk_cell, c2c, G, P

    def _work_counter(self):
        """Device address of the zeroed work counter for the next symmetric sweep, or None (static schedule).
        Two counters alternate: rank 0 zeroes the one for the NEXT sweep right before it launches this one -- every rank's
        previous use of it ended before the all-reduce that closed that sweep completed anywhere, so the memset cannot
        race with a claim.  Every rank must call this the same number of times (it does: once per sweep)."""
        K, comm = self.K, self.comm
        if not self.dynamic or not getattr(K, 'dynamic_schedule', False):
            return None
        if not comm.active:
            # one process, or a rank running un-sharded (solo_if): a private counter -- the ranks' shared one would hand
            # this rank's items to the others, which are not part of this sweep
            if getattr(self, '_wc_local', None) is None:
                self._wc_local = K.zeros((16,), torch.int32)
            K.shared_zero(self._wc_local.data_ptr(), 4)        # stream order: after the previous sweep, before this one
            return self._wc_local.data_ptr()
        if self._wc is None:
            try:
                ok, addr = 1, 0
                box = [None]
                if comm.rank == 0:
                    try:
                        addr, box[0] = K.shared_alloc(64)
                    except Exception:  # noqa: BLE001
                        box[0] = None
                comm.broadcast_object(box, src=0)
                if box[0] is None:
                    ok = 0
                elif comm.rank != 0:
                    try:
                        addr = K.shared_open(box[0])
                    except Exception:  # noqa: BLE001
                        ok = 0
                flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
                comm.allreduce_min(flag)                       # all ranks or none
                self._wc = (addr, None) if int(flag.item()) == 1 else False
            except Exception:  # noqa: BLE001
                self._wc = False
        if self._wc is False:
            return None
        base = self._wc[0]
        n = self._wc_calls
        self._wc_calls += 1
        if comm.rank == 0:
            K.shared_zero(base + 4 * ((n + 1) % 2), 4)
        return base + 4 * (n % 2)

    def sil_scale(self, Xc, ctx=None):
        """Power-of-two scale s for the fp16 operand and the split of the half squared norms m = s^2 |x|^2 / 2
        that ride in the MMA as fp16 pieces m * 2^-shift against the constant 2^shift.
        Constraints: max |x| s <= 4096 (fp16 keeps 11 significant bits), pad_m = 4 max(m) * 2^-shift <= 32768
        with shift <= 15.  Returns (scale, shift, pad_m)."""
        if ctx is not None and getattr(ctx, 'amax', None) is not None and getattr(ctx, 'max_sqn', None) is not None:
            amax, max_sqn = ctx.amax, ctx.max_sqn                          # from the centring pass: no kernel, no sync
        else:
            amax = float(self.K.absmax(Xc).item())
            max_sqn = float(self.K.row_sqnorms(Xc).max().item())
        if amax <= 0.0 or max_sqn <= 0.0:
            return 1.0, 0, 1.0
        scale = 2.0 ** math.floor(math.log2(min(4096.0 / amax, math.sqrt(2.0 ** 28 / max_sqn))))
        bound = 0.5 * scale * scale * max_sqn * 1.01 + 1.0
        pad_m = 4.0 * bound
        shift = max(0, math.ceil(math.log2(pad_m / 32768.0)))
        return scale, int(shift), float(pad_m)

    def _silhouette_tc(self, ctx, labels_list, ks, counts_list):
        K, comm = self.K, self.comm
        with self.span('sil_cells'):
            perm, chunk_cell, c2c, G, P = self.build_cells(labels_list, ks)
        if G * P * 4 > TC_MAX_CELLSUM_BYTES and len(ks) > 1:
            h = len(ks) // 2                                                   # too many cells: sweep twice
            return (self._silhouette_tc(ctx, labels_list[:h], ks[:h], counts_list[:h]) +
                    self._silhouette_tc(ctx, labels_list[h:], ks[h:], counts_list[h:]))
        with self.span('sil_pack'):
            Xc = self._pad16(ctx.Xc, 64)                                      # whole 64-column k-blocks
            scale, shift, pad_m = self.sil_scale(Xc, ctx)
            Xh, aug = K.sil_pack(Xc, perm, P, scale, shift, pad_m)
            cellsum = K.zeros((G, P), torch.float32)
        bounds = self._bounds(P // 128, 2)       # even row-block ranges: CTA pairs share one 256-row MMA
        rb0, rb1 = bounds[comm.rank], bounds[comm.rank + 1]
        U = P // 256
        sym = (U >= TC_SYM_MIN_UNITS * comm.world) if self.sym is None else bool(self.sym)
        if sym:
            # D is symmetric: sweep the tiles on or above the diagonal only; executed MMA flops = d * P^2 * (1 + 1/U)
            counter = self._work_counter()
            tiles_done = K.zeros((1,), torch.int64)
            per_tile = 2.0 * 256 * 256 * Xc.shape[1]
            with self.span('sil_sweep', flops=lambda: per_tile * float(tiles_done.item())):
                K.sil_sweep_sym(Xh, aug, scale, chunk_cell, G, comm.rank, comm.world, cellsum, counter, tiles_done)
            with self.span('sil_allreduce'):
                comm.allreduce_sum(cellsum)
        else:
            with self.span('sil_sweep', flops=2.0 * (rb1 - rb0) * 128 * P * Xc.shape[1]):
                K.sil_sweep(Xh, aug, scale, chunk_cell, G, rb0, rb1, cellsum)
        out = K.zeros((len(ks),), torch.float64)
        with self.span('sil_finish'):
            for j, (lab, k, cnt) in enumerate(zip(labels_list, ks, counts_list)):
                K.sil_finish_cells(cellsum, c2c[j], k, perm, lab, cnt, rb0 * 128, rb1 * 128, out[j:j + 1])
            comm.allreduce_sum(out)
        return (out / ctx.N).cpu().numpy().tolist()

    # ------------------------------------------------------------------------------ f3 rand score
    def rand_score(self, labels_true, labels_pred):
        """sklearn.metrics.rand_score (SK/metrics/cluster/_supervised.py:280-357) from the contingency table
        (:106-192) and the pair confusion matrix (:195-271).  `labels_true` is a host vector of arbitrary integers
        (labels.txt of test_grouping); `labels_pred` a label vector on the device or host.  The O(N) histogram runs
        on the device, row-sharded; the k x k integer arithmetic that follows stays on the host, in Python integers
        (the reference's int64 would overflow beyond 3 x 10^9 rows; below that the results are identical)."""
        K, comm = self.K, self.comm
        gt = np.asarray(labels_true)
        if gt.ndim != 1:
            raise ValueError('labels_true must be 1D: shape is %r' % (gt.shape,))          # check_clusterings, :34-77
        classes, inv = np.unique(gt, return_inverse=True)
        pred = labels_pred if isinstance(labels_pred, torch.Tensor) else torch.from_numpy(np.asarray(labels_pred))
        if pred.dim() != 1:
            raise ValueError('labels_pred must be 1D: shape is %r' % (tuple(pred.shape),))
        N = gt.shape[0]
        if pred.numel() != N:
            raise ValueError('Found input variables with inconsistent numbers of samples: %r' % [N, pred.numel()])
        pred = pred.to(self.device)
        if N == 0:
            return 1.0
        # np.unique of the predicted labels: KMeans labels are 0..k-1 but some may be unused; the table only needs
        # an upper bound, empty columns contribute nothing to any sum
        lo, hi = int(pred.min().item()), int(pred.max().item())
        pred32 = (pred - lo).to(torch.int32).contiguous()
        a = self.to_device(inv.astype(np.int32), torch.int32)
        b, e = self._my_rows(N)
        table, bad = K.contingency(a[b:e], pred32[b:e], len(classes), hi - lo + 1)
        both = torch.cat([table.view(-1), bad])
        comm.allreduce_sum(both)
        h = both.cpu().numpy()
        if h[-1] != 0:
            raise RuntimeError('contingency: %d labels out of range' % int(h[-1]))
        T = [[int(v) for v in row] for row in h[:-1].reshape(len(classes), hi - lo + 1)]
        n_c = [sum(row) for row in T]
        n_k = [sum(col) for col in zip(*T)]
        sum_squares = sum(v * v for row in T for v in row)
        c11 = sum_squares - N
        c01 = sum(v * n_k[j] for row in T for j, v in enumerate(row)) - sum_squares
        c10 = sum(v * n_c[i] for i, row in enumerate(T) for v in row) - sum_squares
        c00 = N * N - c01 - c10 - sum_squares
        num, den = c00 + c11, c00 + c01 + c10 + c11
        if num == den or den == 0:
            return 1.0
        return float(np.int64(num) / np.int64(den)) if den < (1 << 63) else num / den

    def cluster_scores(self, ctx, results, force=None):
        """grouping.py:104-108: [(k, silhouette, CH, DB)] for the KMeans results of a sweep."""
        ks = [r.k for r in results]
        labels = [r.labels for r in results]
        counts = [r.counts for r in results]
        sil = self.silhouette_scores(ctx, labels, ks, counts, force=force)
        with self.span('ch_db'):
            chdb = self.ch_db_scores_multi(ctx, labels, ks, counts, [getattr(r, 'csums', None) for r in results],
                                           [getattr(r, 'disp', None) for r in results])
        return [(r.k, float(s), ch, db) for r, s, (ch, db) in zip(results, sil, chdb)]
