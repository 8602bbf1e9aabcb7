"""Tensor-level wrappers over the C ABI: allocate outputs/workspaces with torch (device memory and
streams are torch's job here, nothing else) and launch the hand-written kernels.

Every method takes and returns torch CUDA tensors.  engine.py is written against this interface
only, which lets the CPU tests drive the sharding logic with a stand-in (tests/fake_kernels.py).
"""
import torch

from . import _native as nv
from ._native import c_f32, c_i64, c_int, ptr, stream


class CudaKernels:
    name = 'cuda-sm100a'

    def __init__(self, device):
        nv.require_gpu()
        self.device = torch.device(device)
        self._ws = None

    # ------------------------------------------------------------------ helpers
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    @staticmethod
    def _chk(X):
        assert X.is_cuda and X.dtype == torch.float32 and X.is_contiguous() and X.dim() == 2, 'need contiguous fp32 CUDA [N,d]'
        assert X.shape[1] % 4 == 0, 'feature dimension must be a multiple of 4'
        return X.shape[0], X.shape[1]

    # ------------------------------------------------------------------ a2 / prologue
    def normalize_rows(self, X):
        N, d = self._chk(X)
        out = torch.empty_like(X)
        nv.call('vtf_normalize_rows', ptr(X), c_i64(N), c_int(d), ptr(out), stream())
        return out

    def row_sqnorms(self, X):
        N, d = self._chk(X)
        out = self.empty((N,), torch.float32)
        nv.call('vtf_row_sqnorms', ptr(X), c_i64(N), c_int(d), ptr(out), stream())
        return out

    def col_stats(self, X):
        N, d = self._chk(X)
        s = self.empty((d,), torch.float64)
        q = self.empty((d,), torch.float64)
        ws = self._workspace(nv.load().vtf_col_stats_ws_bytes(d))
        nv.call('vtf_col_stats', ptr(X), c_i64(N), c_int(d), ptr(s), ptr(q), ptr(ws), stream())
        return s, q

    def center_rows(self, X, mean32):
        N, d = self._chk(X)
        out = torch.empty_like(X)
        nv.call('vtf_center_rows', ptr(X), c_i64(N), c_int(d), ptr(mean32), ptr(out), stream())
        return out

    def normalize_col_stats(self, X):
        """(Xn, col_sum, col_sumsq): L2-normalised rows and their fp64 column statistics in one pass (d % 128 == 0, d <= 1024;
        other widths: two passes)."""
        N, d = self._chk(X)
        if d % 128 != 0 or d > 1024:
            Xn = self.normalize_rows(X)
            return (Xn,) + tuple(self.col_stats(Xn))
        out = torch.empty_like(X)
        s = self.empty((d,), torch.float64)
        q = self.empty((d,), torch.float64)
        ws = self._workspace(nv.load().vtf_normalize_col_stats_ws_bytes(d))
        nv.call('vtf_normalize_col_stats', ptr(X), c_i64(N), c_int(d), ptr(out), ptr(s), ptr(q), ptr(ws), stream())
        return out, s, q

    def center_rows_stats(self, X, mean32):
        """(Xc, stats): stats = fp32 [max |xc|, max row |xc|^2], computed in the centring pass."""
        N, d = self._chk(X)
        out = torch.empty_like(X)
        stats = self.empty((2,), torch.float32)
        nv.call('vtf_center_rows_stats', ptr(X), c_i64(N), c_int(d), ptr(mean32), ptr(out), None, ptr(stats), stream())
        return out, stats

    def upload_pageable(self, A, dtype, threads, out=None):
        """C-contiguous numpy array in pageable memory -> device tensor (native threads + page-locked double buffers)."""
        t = out if out is not None else torch.empty(A.shape, dtype=dtype, device=self.device)
        assert A.flags['C_CONTIGUOUS'] and t.is_contiguous() and t.numel() * t.element_size() == A.nbytes
        nv.call('vtf_upload_pageable', nv.c_vp(A.ctypes.data), ptr(t), nv.c_sz(A.nbytes), c_int(threads), stream())
        return t

    def gather_rows(self, X, idx):
        N, d = self._chk(X)
        out = self.empty((idx.numel(), d), torch.float32)
        nv.call('vtf_gather_rows', ptr(X), c_int(d), ptr(idx), c_i64(idx.numel()), ptr(out), stream())
        return out

    # ------------------------------------------------------------------ a7
    def kpp_rownorm64(self, Xc):
        """|x_row|^2 in fp64, in kpp_dist's own summation order (once per fit instead of once per seeding round)."""
        N, d = self._chk(Xc)
        out = self.empty((N,), torch.float64)
        nv.call('vtf_kpp_rownorm64', ptr(Xc), c_i64(N), c_int(d), ptr(out), stream())
        return out

    def kpp_dist(self, Xc, cand, closest, xsq=None):
        N, d = self._chk(Xc)
        T = cand.shape[0]
        out = self.empty((T, N), torch.float32)
        pots = self.empty((T,), torch.float64)
        ws = self._workspace(nv.load().vtf_kpp_ws_bytes(T))
        nv.call('vtf_kpp_dist', ptr(Xc), c_i64(N), c_int(d), ptr(cand), c_int(T), ptr(closest), ptr(xsq), ptr(out), ptr(pots),
                ptr(ws), stream())
        return out, pots

    def kpp_pick(self, closest, thr32, serial=False, want_total=False, mode=None, offset=None):
        """mode: 'fast' (default: tile skipping), 'scan' (every tile through the block scan), 'serial' (one-thread chain);
        all three give the same bits.  VTF_PICK_MODE overrides the default."""
        import os
        T = thr32.numel()
        idx = self.empty((T,), torch.int64)
        total = self.empty((1,), torch.float32) if want_total else None
        mode = 'serial' if serial else (mode or os.environ.get('VTF_PICK_MODE', 'fast'))
        if mode == 'fast':
            ws = self._workspace(nv.load().vtf_kpp_pick_ws_bytes(closest.numel()))
            nv.call('vtf_kpp_pick_fast', ptr(closest), c_i64(closest.numel()), ptr(thr32), c_int(T), ptr(offset), ptr(total), ptr(idx),
                    ptr(ws), stream())
        else:
            nv.call('vtf_kpp_pick_serial' if mode == 'serial' else 'vtf_kpp_pick', ptr(closest), c_i64(closest.numel()), ptr(thr32),
                    c_int(T), ptr(offset), ptr(total), ptr(idx), stream())
        return (idx, total) if want_total else idx

    def kpp_round_prep(self, pots, u):
        """(thr sorted fp32 [T], order int32 [T], best int64 [1]) from the previous potentials and this round's draws."""
        T = u.numel()
        thr = self.empty((T,), torch.float32)
        order = self.empty((T,), torch.int32)
        best = self.empty((1,), torch.int64)
        nv.call('vtf_kpp_round_prep', ptr(pots), c_int(pots.numel()), ptr(u), c_int(T), ptr(thr), ptr(order), ptr(best),
                stream())
        return thr, order, best

    def kpp_unsort(self, picked, order, N):
        cand = self.empty((picked.numel(),), torch.int64)
        nv.call('vtf_kpp_unsort', ptr(picked), ptr(order), c_int(picked.numel()), c_i64(N), ptr(cand), stream())
        return cand

    # ------------------------------------------------------------------ f3
    def contingency(self, a, b, na, nb):
        """(table int64 [na, nb], bad int64 [1]) for two int32 label vectors of equal length."""
        table = self.empty((na, nb), torch.int64)
        bad = self.empty((1,), torch.int64)
        nv.call('vtf_contingency', ptr(a), ptr(b), c_i64(a.numel()), c_int(na), c_int(nb), ptr(table), ptr(bad), stream())
        return table, bad

    # ------------------------------------------------------------------ a6
    def lloyd_assign(self, Xc, C, labels, counts, n_changed):
        N, d = self._chk(Xc)
        nv.call('vtf_lloyd_assign', ptr(Xc), c_i64(N), c_int(d), ptr(C), c_int(C.shape[0]), ptr(labels), ptr(counts),
                ptr(n_changed), stream())

    def cluster_sums(self, X, labels, k):
        N, d = self._chk(X)
        sums = self.empty((k, d), torch.float64)
        ws = self._workspace(nv.load().vtf_cluster_sums_ws_bytes(k, d))
        nv.call('vtf_cluster_sums', ptr(X), c_i64(N), c_int(d), ptr(labels), c_int(k), ptr(sums), ptr(ws), stream())
        return sums

    def cluster_dispersion(self, X, labels, C64, want_far=False):
        N, d = self._chk(X)
        k = C64.shape[0]
        sq = self.empty((k,), torch.float64)
        rt = self.empty((k,), torch.float64)
        far = self.empty((N,), torch.float32) if want_far else None
        ws = self._workspace(nv.load().vtf_cluster_sums_ws_bytes(k, d))
        nv.call('vtf_cluster_dispersion', ptr(X), c_i64(N), c_int(d), ptr(labels), c_int(k), ptr(C64), ptr(sq), ptr(rt),
                ptr(far), ptr(ws), stream())
        return sq, rt, far

    def lloyd_update(self, sums, counts, n_changed, C_old):
        k, d = C_old.shape
        C_new = torch.empty_like(C_old)
        C64 = self.empty((k, d), torch.float64)
        stats = self.empty((2 + k,), torch.float64)
        nv.call('vtf_lloyd_update', ptr(sums), ptr(counts), ptr(n_changed), c_int(k), c_int(d), ptr(C_old), ptr(C_new),
                ptr(C64), ptr(stats), stream())
        return C_new, C64, stats

    # ---- lock-step forms: labels_all is int32 [n_lab, N_total]; the kernels see rows [b, e) of every labelling
    @staticmethod
    def _koff(k_off):
        import ctypes
        import numpy as np
        a = np.ascontiguousarray(k_off, dtype=np.int32)
        return a, a.ctypes.data_as(ctypes.c_void_p)

    def lloyd_assign_multi(self, Xs, C_all, k_off, active, labels_all, b, counts, n_changed):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        lab_ptr = nv.c_vp(labels_all.data_ptr() + 4 * b)
        nv.call('vtf_lloyd_assign_multi', ptr(Xs), c_i64(n), c_int(d), ptr(C_all), c_int(len(k_off) - 1), koff, active, lab_ptr,
                c_i64(labels_all.shape[1]), ptr(counts), ptr(n_changed), stream())

    def cluster_sums_multi(self, Xs, labels_all, b, k_off, active):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        ktot = int(k_off[-1])
        sums = self.zeros((ktot, d), torch.float64)
        ws = self._workspace(nv.load().vtf_cluster_sums_ws_bytes(ktot, d))
        nv.call('vtf_cluster_sums_multi', ptr(Xs), c_i64(n), c_int(d), nv.c_vp(labels_all.data_ptr() + 4 * b),
                c_i64(labels_all.shape[1]), c_int(len(k_off) - 1), koff, active, ptr(sums), ptr(ws), stream())
        return sums

    def cluster_dispersion_multi(self, Xs, labels_all, b, k_off, active, C64_all):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        ktot = int(k_off[-1])
        sq = self.zeros((ktot,), torch.float64)
        rt = self.zeros((ktot,), torch.float64)
        ws = self._workspace(nv.load().vtf_cluster_sums_ws_bytes(ktot, d))
        nv.call('vtf_cluster_dispersion_multi', ptr(Xs), c_i64(n), c_int(d), nv.c_vp(labels_all.data_ptr() + 4 * b),
                c_i64(labels_all.shape[1]), c_int(len(k_off) - 1), koff, active, ptr(C64_all), ptr(sq), ptr(rt), ptr(ws), stream())
        return sq, rt

    def lloyd_update_multi(self, sums, counts, n_changed, k_off, active, C_old, stats_stride):
        ktot, d = C_old.shape
        keep, koff = self._koff(k_off)
        C_new = C_old.clone()                          # frozen labellings keep their centres
        C64 = C_old.to(torch.float64)
        stats = self.zeros((len(k_off) - 1, stats_stride), torch.float64)
        nv.call('vtf_lloyd_update_multi', ptr(sums), ptr(counts), ptr(n_changed), c_int(len(k_off) - 1), koff, active, c_int(d),
                ptr(C_old), ptr(C_new), ptr(C64), ptr(stats), c_int(stats_stride), stream())
        return C_new, C64, stats

    # ------------------------------------------------------------------ a6, wide form (csrc/lloyd_tc.cu)
    wide_lloyd = True                  # the E-step of a whole k-range as one tcgen05 contraction

    def split_fp16(self, X, n_pad, scale):
        """scale * X = hi + lo in two fp16 matrices [n_pad, d] (rows past N are zero)."""
        N, d = self._chk(X)
        hi = self.empty((n_pad, d), torch.float16)
        lo = self.empty((n_pad, d), torch.float16)
        nv.call('vtf_dedup_pack', ptr(X), c_i64(N), c_int(d), c_i64(n_pad), c_f32(scale), ptr(hi), ptr(lo), stream())
        return hi, lo

    def wide_pack_centres(self, C_all, col_src, col_info, scale):
        Kp, d = col_src.numel(), C_all.shape[1]
        Bhi = self.empty((Kp, d), torch.float16)
        Blo = self.empty((Kp, d), torch.float16)
        tab = self.empty((Kp, 2), torch.float32)
        cmax = self.empty((1,), torch.int32)
        nv.call('vtf_wide_pack_centres', ptr(C_all), c_int(d), ptr(col_src), ptr(col_info), c_int(Kp), c_f32(scale), ptr(Bhi),
                ptr(Blo), ptr(tab), ptr(cmax), stream())
        return Bhi, Blo, tab, cmax

    def wide_assign(self, Xhi, Xlo, n_rows, Bhi, Blo, n_cols, tab, rown, cmax, scale, eps, labels_new):
        P, d = Xhi.shape
        nv.call('vtf_wide_assign', ptr(Xhi), ptr(Xlo), c_i64(n_rows), c_i64(P), c_int(d), ptr(Bhi), ptr(Blo), c_int(Bhi.shape[0]),
                c_int(n_cols), ptr(tab), ptr(rown), ptr(cmax), c_f32(scale), c_f32(eps), ptr(labels_new), c_i64(labels_new.shape[1]),
                stream())

    def wide_refine(self, Xs, C_all, tab, k_off, col0, labels_new, n_rows, n_fix):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        keep2, c0 = self._koff(col0)
        nv.call('vtf_wide_refine', ptr(Xs), c_int(d), ptr(C_all), ptr(tab), c_int(len(k_off) - 1), koff, c0, ptr(labels_new),
                c_i64(labels_new.shape[1]), c_i64(n_rows), ptr(n_fix), stream())

    def wide_count(self, labels_new, labels_all, b, n, k_off, active, counts, changed):
        keep, koff = self._koff(k_off)
        nv.call('vtf_wide_count', ptr(labels_new), c_i64(labels_new.shape[1]), nv.c_vp(labels_all.data_ptr() + 4 * b),
                c_i64(labels_all.shape[1]), c_i64(n), c_int(len(k_off) - 1), koff, active, ptr(counts), ptr(changed), stream())

    def wide_commit(self, labels_new, labels_all, b, n, k_off, active, changed, cap, items, n_items):
        keep, koff = self._koff(k_off)
        nv.call('vtf_wide_commit', ptr(labels_new), c_i64(labels_new.shape[1]), nv.c_vp(labels_all.data_ptr() + 4 * b),
                c_i64(labels_all.shape[1]), c_i64(n), c_int(len(k_off) - 1), koff, active, ptr(changed), c_i64(cap), ptr(items),
                ptr(n_items), items.shape[0], stream())

    def wide_sums_delta(self, Xs, items, n_items, k_off, two_s, acc):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        nv.call('vtf_wide_sums_delta', ptr(Xs), c_int(d), ptr(items), ptr(n_items), items.shape[0], c_int(len(k_off) - 1), koff,
                two_s, ptr(acc), stream())

    def wide_sums_full(self, Xs, labels_all, b, k_off, active, L0, L1, changed, cap, two_s, acc):
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        ws = self._workspace(nv.load().vtf_wide_sums_ws_bytes(int(k_off[L1] - k_off[L0]), d))
        nv.call('vtf_wide_sums_full', ptr(Xs), c_i64(n), c_int(d), nv.c_vp(labels_all.data_ptr() + 4 * b), c_i64(labels_all.shape[1]),
                c_int(len(k_off) - 1), koff, active, c_int(L0), c_int(L1), ptr(changed), c_i64(cap), two_s, ptr(acc), ptr(ws),
                stream())

    def wide_update(self, acc, inv_two_s, counts, changed, k_off, active, C_old, stats_stride):
        ktot, d = C_old.shape
        keep, koff = self._koff(k_off)
        C_new = C_old.clone()                          # frozen labellings keep their centres
        stats = self.zeros((len(k_off) - 1, stats_stride), torch.float64)
        nv.call('vtf_wide_update', ptr(acc), inv_two_s, ptr(counts), ptr(changed), c_int(len(k_off) - 1), koff, active, c_int(d),
                ptr(C_old), ptr(C_new), ptr(stats), c_int(stats_stride), stream())
        return C_new, stats

    # ------------------------------------------------------------------ a6, narrow k-ranges: sums by cells (csrc/cells.cu)
    CELL_MAX_K = 48                    # clusters per labelling the cell kernel takes

    def cell_sums(self, Xs, labels_all, b, k_off, active, two_s, acc):
        """acc[slot] = fixed-point sums (round(x * two_s), int64) of the rows of every ACTIVE labelling's clusters; frozen
        labellings keep theirs.  Same values as wide_sums_full."""
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        ws = self._workspace(64)
        nv.call('vtf_cell_sums', ptr(Xs), c_i64(n), c_int(d), nv.c_vp(labels_all.data_ptr() + 4 * b), c_i64(labels_all.shape[1]),
                c_int(len(k_off) - 1), koff, active, two_s, ptr(acc), None, ptr(ws), stream())

    def cell_dispersion_ok(self, d, ks, row_of=None):
        """One launch takes all labellings: at most 12 of them over at most 8 distinct label rows with at most 1024 joint
        label tuples, k <= 48, centres within 160 KB of shared memory.  Beyond that the warp-per-row kernel is used."""
        rows = list(range(len(ks))) if row_of is None else list(row_of)
        seen, tuples = {}, 1
        for k, r in zip(ks, rows):
            if r not in seen:
                seen[r] = k
                tuples *= int(k)
        return (d % 128 == 0 and d <= 1024 and max(ks) <= self.CELL_MAX_K and len(ks) <= 12 and len(seen) <= 8 and tuples <= 1024 and
                sum(ks) * d * 8 <= 160 * 1024)

    def cell_dispersion(self, Xs, labels_all, b, k_off, C64_all, sq_unit, rt_unit, row_of=None):
        """int64 [2 * ktot]: per centre slot sum |x - c|^2 * sq_unit, then sum |x - c| * rt_unit (exact integer sums of per-row
        values: add across ranks as integers, then divide by the units).  row_of[q] = label row labelling q reads."""
        import numpy as np
        n, d = self._chk(Xs)
        keep, koff = self._koff(k_off)
        keep2, rof = (None, None) if row_of is None else self._koff(np.asarray(row_of, dtype=np.int32))
        ktot = int(k_off[-1])
        out = self.empty((2 * ktot,), torch.int64)
        ws = self._workspace(64)
        nv.call('vtf_cell_dispersion', ptr(Xs), c_i64(n), c_int(d), nv.c_vp(labels_all.data_ptr() + 4 * b), c_i64(labels_all.shape[1]),
                c_int(len(k_off) - 1), koff, rof, ptr(C64_all), sq_unit, rt_unit, ptr(out), None, ptr(ws), stream())
        return out

    BUILD_CELLS_MAX_KEYS = 4096

    def build_cells(self, labels_all, ks):
        """Cell layout of the silhouette sweep for labels_all int32 [n_lab, N] (prod(ks) <= 4096): returns
        (perm int64 [P], chunk_cell int32 [P/32], cell_to_cluster int32 [n_lab, G], G, P).  One small D2H (G, P)."""
        import numpy as np
        nL, N = labels_all.shape
        tuples = 1
        for k in ks:
            tuples *= int(k)
        g_max = min(tuples, N)
        cap = ((N + 31 * g_max + 255) // 256) * 256
        perm = self.empty((cap,), torch.int64)
        chunk_cell = self.empty((cap // 32,), torch.int32)
        c2c = self.empty((nL, g_max), torch.int32)
        meta = self.empty((3,), torch.int64)
        ws = self._workspace(nv.load().vtf_build_cells_ws_bytes(N))
        keep, kh = self._koff(np.asarray(ks, dtype=np.int32))
        nv.call('vtf_build_cells', ptr(labels_all), c_i64(labels_all.stride(0)), c_i64(N), c_int(nL), kh, ptr(perm), c_i64(cap),
                ptr(chunk_cell), ptr(c2c), c_int(g_max), ptr(meta), ptr(ws), stream())
        G, P, _ = (int(v) for v in meta.cpu().tolist())
        return perm[:P], chunk_cell[:P // 32], c2c[:, :G].contiguous(), G, P

    # ------------------------------------------------------------------ a3 / a1
    def cosine_classify(self, Xn, Rn, thr):
        N, d = self._chk(Xn)
        R = Rn.shape[0]
        dist = self.empty((N, R), torch.float32)
        label = self.empty((N,), torch.int64)
        nv.call('vtf_cosine_classify', ptr(Xn), c_i64(N), c_int(d), ptr(Rn), c_int(R), c_f32(thr), ptr(dist), ptr(label),
                stream())
        return dist, label

    def cosine_dedup(self, Xn, rank, world):
        N, d = self._chk(Xn)
        key = self.empty((N,), torch.int64)
        nv.call('vtf_dedup_init', ptr(key), c_i64(N), stream())
        nv.call('vtf_cosine_dedup', ptr(Xn), c_i64(N), c_int(d), c_int(rank), c_int(world), ptr(key), stream())
        return key

    DEDUP_SCREEN_EPS = 1.0e-3          # bound of the single-MMA screen's error on a cosine similarity of unit rows

    dedup_defer = True                 # cosine_dedup_tc(defer=True): the re-sweep decision is left to the caller

    def cosine_dedup_tc(self, Xn, rank, world, thr=None, mode=None, defer=False):
        """Tensor-core dedup.  mode 'screen' (default when a threshold is given): single-MMA fp16 screen that keeps the best
        candidates per row + exact fp32 re-evaluation, with the units whose decision at `thr` the screen cannot guarantee
        re-swept by the split-fp16 kernel; mode 'split': split-fp16 operands (hi + lo), three MMAs per K step, everywhere.
        `self.dedup_stats` = units re-swept by the last call."""
        import os
        N, d = self._chk(Xn)
        P = ((N + 255) // 256) * 256
        scale = 4096.0                      # |xn| <= 1: hi keeps 11 bits, lo the next 11
        mode = mode or os.environ.get('VTF_DEDUP_MODE') or ('screen' if thr is not None else 'split')
        hi = self.empty((P, d), torch.float16)
        lo = self.empty((P, d), torch.float16)
        key = self.empty((N,), torch.int64)
        nv.call('vtf_dedup_pack', ptr(Xn), c_i64(N), c_int(d), c_i64(P), c_f32(scale), ptr(hi), ptr(lo), stream())
        nv.call('vtf_dedup_init', ptr(key), c_i64(N), stream())
        self.dedup_stats = dict(mode=mode, units=P // 256, reswept_units=0)
        if mode == 'split':
            nv.call('vtf_cosine_dedup_tc', ptr(hi), ptr(lo), c_i64(N), c_i64(P), c_int(d), c_f32(scale), c_int(rank), c_int(world),
                    ptr(key), stream())
            nv.call('vtf_dedup_refine', ptr(Xn), c_i64(N), c_int(d), c_int(rank), c_int(world), ptr(key), stream())
            return (key, None, None) if defer else key
        cand = self.empty((N, 8, 2), torch.int32)
        flags = self.zeros((P // 256 + 1,), torch.int32)         # [units] flags + [1] count
        n_launch = int(nv.load().vtf_dedup_screen_launches(N, d))
        counters = self.zeros((max(n_launch, 1),), torch.int32)   # one claim counter per column chunk
        nv.call('vtf_dedup_screen', ptr(hi), c_i64(N), c_i64(P), c_int(d), c_int(rank), c_int(world), ptr(cand), ptr(counters),
                c_int(n_launch), stream())
        self.dedup_stats['screen_launches'] = n_launch
        nv.call('vtf_dedup_refine4', ptr(Xn), c_i64(N), c_int(d), c_int(rank), c_int(world), ptr(cand), c_f32(scale),
                c_f32(float('nan') if thr is None else float(thr)), c_f32(self.DEDUP_SCREEN_EPS), ptr(key), ptr(flags),
                nv.c_vp(flags.data_ptr() + 4 * (P // 256)), stream())
        def resweep():
            # rare: rows with five neighbours of a column half in the band right above thr -- their units go through the split sweep
            units = torch.nonzero(flags[:-1]).flatten().to(torch.int32)
            units = torch.flip(units, [0]).contiguous()          # largest units first
            self.dedup_stats['reswept_units'] = int(units.numel())
            nv.call('vtf_cosine_dedup_tc_units', ptr(hi), ptr(lo), c_i64(N), c_i64(P), c_int(d), c_f32(scale), ptr(units),
                    c_i64(units.numel()), ptr(key), stream())
            nv.call('vtf_dedup_refine', ptr(Xn), c_i64(N), c_int(d), c_int(rank), c_int(world), ptr(key), stream())
        if defer:
            # the caller downloads `flags[-1]` together with the keys (one synchronisation instead of two) and calls resweep()
            # -- then downloads the keys again -- only if it is non-zero
            return key, flags[-1:], resweep
        if thr is not None and int(flags[-1].item()) > 0:
            resweep()
        return key

    def hash_dedup(self, bits, rank, world):
        """bits: int64 [N,64] of 0/1 -> packed (distance, argmin) keys, int64 [N]."""
        N = bits.shape[0]
        packed = self.empty((N,), torch.int64)
        key = self.empty((N,), torch.int64)
        nv.call('vtf_hash_pack', ptr(bits), c_i64(N), ptr(packed), stream())
        nv.call('vtf_hash_dedup', ptr(packed), c_i64(N), c_int(rank), c_int(world), ptr(key), stream())
        return key

    # ------------------------------------------------------------------ a8
    def sil_rowsums_exact(self, X, sqn, labels, k, row_begin, row_end, rowsum):
        N, d = self._chk(X)
        nv.call('vtf_silhouette_rowsums_exact', ptr(X), ptr(sqn), c_i64(N), c_int(d), ptr(labels), c_int(k),
                c_i64(row_begin), c_i64(row_end), ptr(rowsum), stream())

    def sil_finish_rows(self, rowsum, labels, counts, row_begin, row_end, out_slot):
        N, k = rowsum.shape
        nv.call('vtf_sil_finish_rows', ptr(rowsum), c_i64(N), c_int(k), ptr(labels), ptr(counts), c_i64(row_begin),
                c_i64(row_end), ptr(out_slot), stream())

    def absmax(self, X):
        N, d = self._chk(X)
        out = self.empty((1,), torch.float32)
        nv.call('vtf_absmax', ptr(X), c_i64(N), c_int(d), ptr(out), stream())
        return out

    tc_norms = True        # augmented operand from the tensor core's own squared norms (vtf_sil_diag); False: fp32-FMA norms

    def sil_pack(self, X, perm, P, scale, shift, pad_m):
        """fp16 operand [P,d] + augmented norm columns [2,P,16] (A and B flavour) for the tensor-core sweep.  The norms
        that ride in the augmented columns are taken from a diagonal pass of the sweep kernel itself, so that they carry
        the same accumulation rounding as the Gram entries they are subtracted from."""
        N, d = self._chk(X)
        Xh = self.empty((P, d), torch.float16)
        aug = self.empty((2, P, 16), torch.float16)
        nv.call('vtf_sil_pack', ptr(X), c_i64(N), c_int(d), ptr(perm), c_i64(P), c_f32(scale), c_int(shift), c_f32(pad_m),
                ptr(Xh), ptr(aug), None, stream())
        if self.tc_norms and P % 256 == 0 and d % 64 == 0 and 64 <= d <= 1024:
            m = self.sil_diag(Xh, torch.zeros_like(aug), 0.5)
            nv.call('vtf_sil_aug', ptr(m), ptr(perm), c_i64(N), c_i64(P), c_int(shift), ptr(aug), stream())
        return Xh, aug

    def sil_diag(self, Xh, aug, out_scale):
        """out[p] = out_scale * T_pp as the sweep accumulates it (zero aug: the tensor core's squared norms)."""
        P, d = Xh.shape
        out = self.empty((P,), torch.float32)
        nv.call('vtf_sil_diag', ptr(Xh), ptr(aug), c_i64(P), c_int(d), c_f32(out_scale), ptr(out), stream())
        return out

    def sil_sweep(self, Xh, aug, scale, chunk_cell, n_cells, rb_begin, rb_end, cellsum):
        P, d = Xh.shape
        nv.call('vtf_sil_sweep', ptr(Xh), ptr(aug), c_i64(P), c_int(d), c_f32(scale), ptr(chunk_cell), c_int(n_cells),
                c_i64(rb_begin), c_i64(rb_end), ptr(cellsum), stream())

    def sil_sweep_sym(self, Xh, aug, scale, chunk_cell, n_cells, rank, world, cellsum, counter=None, tiles_done=None):
        """`counter`: device address (int) of the zeroed uint32 work counter shared by all ranks -> dynamic schedule;
        None -> static round-robin by rank.  `tiles_done`: int64 [1] tensor, += tiles computed by this launch."""
        P, d = Xh.shape
        nv.call('vtf_sil_sweep_sym', ptr(Xh), ptr(aug), c_i64(P), c_int(d), c_f32(scale), ptr(chunk_cell), c_int(n_cells),
                c_int(rank), c_int(world), ptr(cellsum), None if counter is None else nv.c_vp(counter), ptr(tiles_done),
                stream())

    # ------------------------------------------------------------------ memory shared between ranks (CUDA IPC)
    dynamic_schedule = True            # the sweep can claim work items from a shared counter

    def shared_alloc(self, nbytes):
        """(device address, 64-byte handle) of a zeroed allocation other ranks can map."""
        import ctypes
        p = nv.c_vp()
        h = (ctypes.c_ubyte * 64)()
        nv.call('vtf_shared_alloc', nv.c_sz(nbytes), ctypes.byref(p), h)
        return int(p.value), bytes(h)

    def shared_open(self, handle):
        import ctypes
        p = nv.c_vp()
        h = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        nv.call('vtf_shared_open', h, ctypes.byref(p))
        return int(p.value)

    def shared_zero(self, addr, nbytes):
        nv.call('vtf_shared_zero', nv.c_vp(addr), nv.c_sz(nbytes), stream())

    def sil_finish_cells(self, cellsum, cell_to_cluster, k, perm, labels, counts, p_begin, p_end, out_slot):
        n_cells, P = cellsum.shape
        nv.call('vtf_sil_finish_cells', ptr(cellsum), c_i64(P), c_int(n_cells), ptr(cell_to_cluster), c_int(k), ptr(perm),
                ptr(labels), ptr(counts), c_i64(p_begin), c_i64(p_end), ptr(out_slot), stream())
