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

    def normalize_col_stats(self, X):
        Xn = self.normalize_rows(X)
        return (Xn,) + tuple(self.col_stats(Xn))

    def center_rows(self, X, mean32):
        return (X - mean32).contiguous()

    def gather_rows(self, X, idx):
        return X[idx].contiguous()

    def kpp_dist(self, Xc, cand, closest, xsq=None):
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

    # ---- wide Lloyd path (csrc/lloyd_tc.cu): the contracts of the vtf_wide_* entry points in numpy
    wide_lloyd = True

    def split_fp16(self, X, n_pad, scale):
        x = _np(X) * np.float32(scale)
        hi = np.zeros((n_pad, x.shape[1]), dtype=np.float16)
        lo = np.zeros_like(hi)
        hi[:x.shape[0]] = x.astype(np.float16)
        lo[:x.shape[0]] = (x - hi[:x.shape[0]].astype(np.float32)).astype(np.float16)
        return torch.from_numpy(hi), torch.from_numpy(lo)

    def wide_pack_centres(self, C_all, col_src, col_info, scale):
        src, info = _np(col_src), _np(col_info).view(np.uint32)
        c = _np(C_all)
        Kp, d = src.size, c.shape[1]
        rows = np.zeros((Kp, d), dtype=np.float32)
        rows[src >= 0] = c[src[src >= 0]]
        hi, lo = self.split_fp16(torch.from_numpy(rows), Kp, scale)
        tab = np.zeros((Kp, 2), dtype=np.float32)
        tab[:, 0] = np.where(src >= 0, 0.5 * (rows * rows).sum(1, dtype=np.float32), np.inf)
        tab[:, 1] = np.where(src >= 0, info, info | 1).astype(np.uint32).view(np.float32)
        cmax = np.array([tab[src >= 0, 0].max()], dtype=np.float32).view(np.int32)
        return hi, lo, torch.from_numpy(tab), torch.from_numpy(cmax.copy())

    def wide_assign(self, Xhi, Xlo, n_rows, Bhi, Blo, n_cols, tab, rown, cmax, scale, eps, labels_new):
        xh, xl = _np(Xhi).astype(np.float64)[:n_rows], _np(Xlo).astype(np.float64)[:n_rows]
        ch, cl = _np(Bhi).astype(np.float64), _np(Blo).astype(np.float64)
        g = (xh @ ch.T + xh @ cl.T + xl @ ch.T).astype(np.float32)              # hi.hi + hi.lo + lo.hi
        t = _np(tab)
        info = t[:, 1].copy().view(np.uint32)
        s_all = (t[None, :, 0] - g * np.float32(1.0 / (scale * scale))).astype(np.float32)
        tol = (np.float32(eps) * _np(rown)[:n_rows] * np.sqrt(np.float32(2.0) * _np(cmax).view(np.float32)[0])).astype(np.float32)
        out = _np(labels_new)
        inf = np.float32(np.inf)
        best = np.full(n_rows, inf, np.float32)
        lo_, hi_ = best.copy(), best.copy()
        arg = np.zeros(n_rows, np.uint32)
        mask = np.zeros(n_rows, np.uint32)
        multi = np.zeros(n_rows, bool)
        last_tile = (n_cols - 1) // 256 * 256
        for j in range(min(t.shape[0], last_tile + ((n_cols - last_tile + 31) // 32) * 32)):   # the last tile runs ceil32 wide
            s = s_all[:, j]
            local = np.uint32((int(info[j]) >> 4) & 0xffff)
            bit = np.uint32(0x100 << (int(local) & 15))
            p_clear, p_cand, p_new = s < lo_, s <= hi_, s < best
            mask = np.where(p_clear, bit, np.where(p_cand, mask | bit, mask)).astype(np.uint32)
            multi = np.where(p_clear, False, multi | p_cand)
            best = np.where(p_new, s, best)
            lo_ = np.where(p_new, s - tol, lo_)
            hi_ = np.where(p_new, s + tol, hi_)
            arg = np.where(p_new, local, arg).astype(np.uint32)
            if (int(info[j]) >> 1) & 1:
                L = int(info[j]) >> 20
                word = np.where(multi, arg | mask | np.uint32(0x80000000), arg).astype(np.uint32)
                out[L, :n_rows] = word.view(np.int32)
                best[:] = inf; lo_[:] = inf; hi_[:] = inf
                mask[:] = 0; multi[:] = False

    def wide_refine(self, Xs, C_all, tab, k_off, col0, labels_new, n_rows, n_fix):
        x, c, t, lab = _np(Xs), _np(C_all), _np(tab), _np(labels_new)
        for L in range(len(k_off) - 1):
            if col0[L] < 0:
                continue
            k = int(k_off[L + 1] - k_off[L])
            for r in np.nonzero(lab[L, :n_rows] < 0)[0]:
                m = (int(lab[L, r]) >> 8) & 0xffff
                cands = [cc for cc in range(k) if (m >> (cc & 15)) & 1]
                s = [np.float32(t[col0[L] + cc, 0] - np.float32(np.dot(x[r], c[k_off[L] + cc]))) for cc in cands]
                lab[L, r] = cands[int(np.argmin(s))]
                n_fix += 1

    def wide_count(self, labels_new, labels_all, b, n, k_off, active, counts, changed):
        for L in range(len(k_off) - 1):
            if (active >> L) & 1:
                new = labels_new[L, :n].to(torch.int64)
                counts[int(k_off[L]):int(k_off[L + 1])] += torch.bincount(new, minlength=int(k_off[L + 1] - k_off[L]))
                changed[L] += int((new != labels_all[L, b:b + n].to(torch.int64)).sum())

    def wide_commit(self, labels_new, labels_all, b, n, k_off, active, changed, cap, items, n_items):
        n_items.zero_()
        for L in range(len(k_off) - 1):
            if not (active >> L) & 1:
                continue
            new, old = labels_new[L, :n], labels_all[L, b:b + n].clone()
            rows = torch.nonzero(new != old).flatten()
            if int(changed[L]) <= cap:
                at = int(n_items[0])
                for i, r in enumerate(rows.tolist()):
                    items[at + i] = torch.tensor([r, L, int(old[r]), int(new[r])], dtype=torch.int32)
                n_items += len(rows)
            labels_all[L, b:b + n] = new

    @staticmethod
    def _fixed(x, two_s):
        return torch.from_numpy(np.rint(_np(x).astype(np.float64) * two_s).astype(np.int64))

    def wide_sums_delta(self, Xs, items, n_items, k_off, two_s, acc):
        q = self._fixed(Xs, two_s)
        for r, L, old, new in items[:int(n_items[0])].tolist():
            acc[int(k_off[L]) + new] += q[r]
            if old >= 0:
                acc[int(k_off[L]) + old] -= q[r]

    def wide_sums_full(self, Xs, labels_all, b, k_off, active, L0, L1, changed, cap, two_s, acc):
        q = self._fixed(Xs, two_s)
        n = Xs.shape[0]
        for L in range(L0, L1):
            if (active >> L) & 1 and int(changed[L]) > max(cap, 0):
                lo, hi = int(k_off[L]), int(k_off[L + 1])
                acc[lo:hi] = torch.zeros((hi - lo, q.shape[1]), dtype=torch.int64).index_add_(0, labels_all[L, b:b + n].to(torch.int64), q)

    CELL_MAX_K = 48

    def cell_sums(self, Xs, labels_all, b, k_off, active, two_s, acc):
        q = self._fixed(Xs, two_s)
        n = Xs.shape[0]
        for L in range(len(k_off) - 1):
            if (active >> L) & 1:
                lo, hi = int(k_off[L]), int(k_off[L + 1])
                acc[lo:hi] = torch.zeros((hi - lo, q.shape[1]), dtype=torch.int64).index_add_(0, labels_all[L, b:b + n].to(torch.int64), q)

    def cell_dispersion_ok(self, d, ks, row_of=None):
        return max(ks) <= self.CELL_MAX_K and len(ks) <= 12

    def cell_dispersion(self, Xs, labels_all, b, k_off, C64_all, sq_unit, rt_unit, row_of=None):
        rows = list(range(len(k_off) - 1)) if row_of is None else list(row_of)
        lab = labels_all[rows].contiguous()
        sq, rt = self.cluster_dispersion_multi(Xs, lab, b, k_off, (1 << (len(k_off) - 1)) - 1, C64_all)
        return torch.cat([torch.round(sq * sq_unit), torch.round(rt * rt_unit)]).to(torch.int64)

    def wide_update(self, acc, inv_two_s, counts, changed, k_off, active, C_old, stats_stride):
        sums = acc.to(torch.float64) * inv_two_s
        C_new, _, stats = self.lloyd_update_multi(sums, counts, changed, k_off, active, C_old, stats_stride)
        return C_new, stats
