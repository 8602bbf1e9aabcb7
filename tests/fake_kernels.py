"""TEST INFRASTRUCTURE: a CPU stand-in with the interface of videotofaces_b200.kernels.CudaKernels.

It lets the CPU suite (no GPU in the build container) drive engine.py -- the k-means++ host
protocol, the Lloyd loop, the cell layout of the silhouette sweep and the multi-rank exchange steps
(gloo, world_size 2) -- end to end.  Every method restates the contract of the C entry point it
replaces (include/vtf_b200.h) with numpy / oracle.restatement; it is never imported by the product.
"""
import numpy as np
import torch

from oracle import restatement as R


def _np(t):
    return t.detach().cpu().numpy()


class FakeKernels:
    name = 'fake-cpu (tests only)'

    def __init__(self, device='cpu'):
        self.device = torch.device(device)
        self.launches = 0

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def normalize_rows(self, X):
        return torch.from_numpy(R.normalize_l2(_np(X)).astype(np.float32))

    def row_sqnorms(self, X):
        return torch.from_numpy(R.row_norms_sq(_np(X)).astype(np.float32))

    def col_stats(self, X):
        x = _np(X).astype(np.float64)
        return torch.from_numpy(x.sum(0)), torch.from_numpy((x * x).sum(0))

    def center_rows(self, X, mean32):
        return (X - mean32).contiguous()

    def gather_rows(self, X, idx):
        return X[idx].contiguous()

    def kpp_dist(self, Xc, cand, closest):
        d = R.euclidean_sq_upcast(_np(cand), _np(Xc))
        if closest is not None:
            d = np.minimum(d, _np(closest)[None, :])
        return torch.from_numpy(d), torch.from_numpy(d.sum(axis=1, dtype=np.float64))

    def kpp_pick(self, closest, thr32):
        cs = np.cumsum(_np(closest).astype(np.float32))          # sequential fp32, as the kernel walks it
        idx = np.searchsorted(cs, _np(thr32).astype(np.float32), side='left').astype(np.int64)
        idx[idx >= cs.size] = -1
        return torch.from_numpy(idx)

    def kpp_round_prep(self, pots, u):
        p32 = _np(pots).astype(np.float32)
        best = int(np.argmin(p32))
        v = _np(u) * np.float64(p32[best])
        t = v.astype(np.float32)
        low = t.astype(np.float64) < v
        t[low] = np.nextafter(t[low], np.float32(np.inf))
        order = np.argsort(t, kind='stable')
        return (torch.from_numpy(t[order]), torch.from_numpy(order.astype(np.int32)),
                torch.tensor([best], dtype=torch.int64))

    def kpp_unsort(self, picked, order, N):
        cand = torch.empty_like(picked)
        cand[order.to(torch.int64)] = torch.where(picked < 0, torch.full_like(picked, N - 1), picked)
        return cand

    def contingency(self, a, b, na, nb):
        a64, b64 = a.to(torch.int64), b.to(torch.int64)
        ok = (a64 >= 0) & (a64 < na) & (b64 >= 0) & (b64 < nb)
        table = torch.bincount(a64[ok] * nb + b64[ok], minlength=na * nb).view(na, nb)
        return table, torch.tensor([int((~ok).sum())], dtype=torch.int64)

    def lloyd_assign(self, Xc, C, labels, counts, n_changed):
        new = torch.from_numpy(R.lloyd_assign(_np(Xc), _np(C)))
        n_changed += int((new != labels).sum())
        labels.copy_(new)
        counts += torch.bincount(new.to(torch.int64), minlength=C.shape[0])

    def cluster_sums(self, X, labels, k):
        out = torch.zeros((k, X.shape[1]), dtype=torch.float64)
        out.index_add_(0, labels.to(torch.int64), X.to(torch.float64))
        return out

    def cluster_dispersion(self, X, labels, C64, want_far=False):
        k = C64.shape[0]
        diff = X.to(torch.float64) - C64[labels.to(torch.int64)]
        d2 = (diff * diff).sum(1)
        sq = torch.zeros(k, dtype=torch.float64).index_add_(0, labels.to(torch.int64), d2)
        rt = torch.zeros(k, dtype=torch.float64).index_add_(0, labels.to(torch.int64), d2.sqrt())
        return sq, rt, (d2.to(torch.float32) if want_far else None)

    def lloyd_update(self, sums, counts, n_changed, C_old):
        k = C_old.shape[0]
        amax = int(torch.argmax(counts))
        C_new = C_old.clone()
        for c in range(k):
            src = c if counts[c] > 0 else amax
            if counts[src] > 0:
                C_new[c] = (sums[src] / counts[src].to(torch.float64)).to(torch.float32)
        stats = torch.zeros(2 + k, dtype=torch.float64)
        stats[0] = ((C_new.to(torch.float64) - C_old.to(torch.float64)) ** 2).sum()
        stats[1] = float(n_changed[0])
        stats[2:] = counts.to(torch.float64)
        return C_new, C_new.to(torch.float64), stats

    # ---- lock-step forms, restated through the single-labelling ones
    def lloyd_assign_multi(self, Xs, C_all, k_off, active, labels_all, b, counts, n_changed):
        n = Xs.shape[0]
        for L in range(len(k_off) - 1):
            if not (active >> L) & 1:
                continue
            lo, hi = int(k_off[L]), int(k_off[L + 1])
            lab = labels_all[L, b:b + n].clone()
            cnt = torch.zeros(hi - lo, dtype=torch.int64)
            ch = torch.zeros(1, dtype=torch.int64)
            self.lloyd_assign(Xs, C_all[lo:hi].contiguous(), lab, cnt, ch)
            labels_all[L, b:b + n] = lab
            counts[lo:hi] += cnt
            n_changed[L] += ch[0]

    def cluster_sums_multi(self, Xs, labels_all, b, k_off, active):
        n = Xs.shape[0]
        out = torch.zeros((int(k_off[-1]), Xs.shape[1]), dtype=torch.float64)
        for L in range(len(k_off) - 1):
            if (active >> L) & 1:
                lo, hi = int(k_off[L]), int(k_off[L + 1])
                out[lo:hi] = self.cluster_sums(Xs, labels_all[L, b:b + n], hi - lo)
        return out

    def cluster_dispersion_multi(self, Xs, labels_all, b, k_off, active, C64_all):
        n = Xs.shape[0]
        sq = torch.zeros(int(k_off[-1]), dtype=torch.float64)
        rt = torch.zeros(int(k_off[-1]), dtype=torch.float64)
        for L in range(len(k_off) - 1):
            if (active >> L) & 1:
                lo, hi = int(k_off[L]), int(k_off[L + 1])
                sq[lo:hi], rt[lo:hi], _ = self.cluster_dispersion(Xs, labels_all[L, b:b + n], C64_all[lo:hi])
        return sq, rt

    def lloyd_update_multi(self, sums, counts, n_changed, k_off, active, C_old, stats_stride):
        C_new = C_old.clone()
        stats = torch.zeros((len(k_off) - 1, stats_stride), dtype=torch.float64)
        for L in range(len(k_off) - 1):
            if (active >> L) & 1:
                lo, hi = int(k_off[L]), int(k_off[L + 1])
                cn, _, st = self.lloyd_update(sums[lo:hi], counts[lo:hi], n_changed[L:L + 1], C_old[lo:hi])
                C_new[lo:hi] = cn
                stats[L, :2 + hi - lo] = st
        return C_new, C_new.to(torch.float64), stats

    def cosine_classify(self, Xn, Rn, thr):
        dist = np.clip(1.0 - _np(Xn) @ _np(Rn).T, 0, 2).astype(np.float32)
        lab = dist.argmin(1).astype(np.int64)
        if thr >= 0:
            lab[dist.min(1) >= np.float32(thr)] = Rn.shape[0]
        return torch.from_numpy(dist), torch.from_numpy(lab)

    def cosine_dedup(self, Xn, rank, world):
        x = _np(Xn)
        N = x.shape[0]
        key = np.full(N, np.int64(np.float32(10000.0).view(np.uint32)) << 32, dtype=np.int64)
        D = np.clip(1.0 - x @ x.T, 0, 2).astype(np.float32)
        for i in range(1, N):
            if (i // 64) % world != rank:
                continue
            j = int(D[i, :i].argmin())
            key[i] = (np.int64(D[i, j].view(np.uint32)) << 32) | j
        return torch.from_numpy(key)

    def hash_dedup(self, bits, rank, world):
        b = _np(bits) != 0
        N = b.shape[0]
        key = np.full(N, np.int64(10000) << 32, dtype=np.int64)
        for i in range(1, N):
            if (i // 256) % world != rank:
                continue
            dd = (b[:i] != b[i]).sum(1)
            j = int(dd.argmin())
            key[i] = (np.int64(dd[j]) << 32) | j
        return torch.from_numpy(key)

    def _dist_matrix(self, X, sqn, rows):
        x = _np(X)
        g = x[rows] @ x.T
        d2 = np.maximum(_np(sqn)[rows][:, None] + _np(sqn)[None, :] - 2 * g, 0)
        D = np.sqrt(d2)
        D[np.arange(len(rows)), rows] = 0
        return D

    def sil_rowsums_exact(self, X, sqn, labels, k, row_begin, row_end, rowsum):
        rows = np.arange(row_begin, row_end)
        if rows.size == 0:
            return
        D = self._dist_matrix(X, sqn, rows).astype(np.float64)
        lab = _np(labels)
        for c in range(k):
            rowsum[row_begin:row_end, c] += torch.from_numpy(D[:, lab == c].sum(1))

    @staticmethod
    def _sil(a_sum, b, n_own):
        if n_own <= 1:
            return 0.0
        a = a_sum / (n_own - 1)
        m = max(a, b)
        return 0.0 if not (m > 0 and np.isfinite(b)) else (b - a) / m

    def sil_finish_rows(self, rowsum, labels, counts, row_begin, row_end, out_slot):
        rs, lab, cnt = _np(rowsum), _np(labels), _np(counts)
        tot = 0.0
        for i in range(row_begin, row_end):
            own = lab[i]
            b = min([rs[i, c] / cnt[c] for c in range(len(cnt)) if c != own and cnt[c] > 0], default=np.inf)
            tot += self._sil(rs[i, own], b, cnt[own])
        out_slot += tot

    def absmax(self, X):
        return X.abs().max().reshape(1).to(torch.float32)

    def sil_pack(self, X, perm, P, scale, shift, pad_m):
        """aug stands in for the augmented operand: the half squared norms m of the ROUNDED rows."""
        assert 0 <= shift <= 15 and pad_m * 2.0 ** -shift <= 32768.0
        x = _np(X)
        pm = _np(perm)
        xh = np.zeros((P, x.shape[1]), dtype=np.float16)
        live = pm >= 0
        xh[live] = (x[pm[live]] * np.float32(scale)).astype(np.float16)
        f = xh.astype(np.float32)
        m = np.full(P, -pad_m, dtype=np.float32)
        m[live] = 0.5 * (f[live] * f[live]).sum(1)
        assert m.max() * 4 <= pad_m
        return torch.from_numpy(xh), torch.from_numpy(m)

    def sil_sweep(self, Xh, aug, scale, chunk_cell, n_cells, rb_begin, rb_end, cellsum):
        f = _np(Xh).astype(np.float32)
        m = _np(aug)
        cc = np.repeat(_np(chunk_cell), 32)
        rows = np.arange(rb_begin * 128, rb_end * 128)
        if rows.size == 0:
            return
        T = f[rows] @ f.T - m[rows][:, None] - m[None, :]
        D = (np.sqrt(np.maximum(-T, 0)) * np.float32(np.sqrt(2.0) / scale)).astype(np.float32)
        D[np.arange(rows.size), rows] = 0
        for c in range(n_cells):
            cellsum[c, rows[0]:rows[-1] + 1] += torch.from_numpy(D[:, cc == c].sum(1, dtype=np.float32))

    def sil_sweep_sym(self, Xh, aug, scale, chunk_cell, n_cells, rank, world, cellsum, counter=None, tiles_done=None):
        """Upper-triangle sweep: unit pairs (q, U-1-q) of 256 rows; off-diagonal tiles feed rows and columns."""
        f = _np(Xh).astype(np.float32)
        m = _np(aug)
        cc = np.repeat(_np(chunk_cell), 32)
        P = f.shape[0]
        U = P // 256
        out = np.zeros((n_cells, P), dtype=np.float64)
        for q in range(rank, (U + 1) // 2, world):
            for u in sorted({q, U - 1 - q}):
                rows = np.arange(u * 256, (u + 1) * 256)
                for t in range(u, U):
                    cols = np.arange(t * 256, (t + 1) * 256)
                    T = f[rows] @ f[cols].T - m[rows][:, None] - m[cols][None, :]
                    D = np.sqrt(np.maximum(-T, 0)) * (np.sqrt(2.0) / scale)
                    if t == u:
                        np.fill_diagonal(D, 0.0)
                    for c in range(n_cells):
                        out[c, rows] += D[:, cc[cols] == c].sum(1)
                        if t != u:
                            out[c, cols] += D[cc[rows] == c, :].sum(0)
        cellsum += torch.from_numpy(out.astype(np.float32))

    def sil_finish_cells(self, cellsum, cell_to_cluster, k, perm, labels, counts, p_begin, p_end, out_slot):
        cs, c2c, pm, lab, cnt = _np(cellsum), _np(cell_to_cluster), _np(perm), _np(labels), _np(counts)
        tot = 0.0
        for p in range(p_begin, p_end):
            if pm[p] < 0:
                continue
            S = np.zeros(k)
            np.add.at(S, c2c, cs[:, p].astype(np.float64))
            own = lab[pm[p]]
            b = min([S[c] / cnt[c] for c in range(k) if c != own and cnt[c] > 0], default=np.inf)
            tot += self._sil(S[own], b, cnt[own])
        out_slot += tot
