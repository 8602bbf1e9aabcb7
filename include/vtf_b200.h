/*
 * vtf_b200.h -- C ABI of the B200-native grouping hot path of video-to-faces.
 *
 * The reference (sephirot-github/video-to-faces) has no FFI: its grouping stage
 * (src/videotofaces/dupes.py:51-92, src/videotofaces/grouping.py:50-137) calls scikit-learn,
 * whose native layer is OpenBLAS SGEMM/DGEMM plus the Cython Lloyd kernels.  Each entry point
 * below replaces one of those call sites; the citation names the reference line (and the
 * scikit-learn 1.9.0 routine behind it, "SK/" = site-packages/sklearn) it stands in for.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _h (host);
 *   - matrices are row-major, C-contiguous; X is [N, d] float32 with d % 4 == 0;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     asynchronous on it and nothing allocates: the caller owns outputs and workspaces;
 *   - return value: 0 = ok, negative = error (vtf_last_error() gives the text).  There is no
 *     CPU fallback: without an sm_100 device every compute call fails with VTF_ERR_DEVICE.
 */
#ifndef VTF_B200_H
#define VTF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTF_OK             0
#define VTF_ERR_ARG       -1
#define VTF_ERR_CUDA      -2
#define VTF_ERR_DEVICE    -3
#define VTF_ERR_WORKSPACE -4

int         vtf_version(void);
const char* vtf_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t     vtf_launch_count(void);
/* 0 when the current device is sm_100 (B200); VTF_ERR_DEVICE otherwise. */
int         vtf_device_check(void);

/* ---- a2: cosine normalisation ------------------------------------------------------------
 * sklearn.preprocessing.normalize inside cosine_distances (SK/metrics/pairwise.py:1744-1748,
 * SK/preprocessing/_data.py:2076-2081, zero-norm guard :127-132), called from dupes.py:60 and
 * grouping.py:51.  fp32 norm, sqrt, divide; norms < 10*eps become 1.                       */
int vtf_normalize_rows(const float* X, int64_t N, int d, float* Xn, void* stream);

/* ---- a6 prologue: KMeans.fit set-up (SK/cluster/_kmeans.py:1473-1496) -------------------
 * col_stats: per-column sum and sum of squares in fp64 over the rows given (a rank's shard;
 * partial results add across ranks).  ws: vtf_col_stats_ws_bytes(d).                      */
size_t vtf_col_stats_ws_bytes(int d);
int vtf_col_stats(const float* X, int64_t N, int d, double* col_sum, double* col_sumsq,
                  void* ws, void* stream);
/* Encoder hand-off: Xn = normalize(X) (as vtf_normalize_rows; FaceNet's own last step,
 * encoders/facenet.py:153) AND the column sums / sums of squares of Xn (fp64, as vtf_col_stats) in one pass over
 * un-normalised embeddings.  d % 128 == 0, d <= 1024.  ws: vtf_normalize_col_stats_ws_bytes(d).                */
size_t vtf_normalize_col_stats_ws_bytes(int d);
int vtf_normalize_col_stats(const float* X, int64_t N, int d, float* Xn, double* col_sum, double* col_sumsq, void* ws,
                            void* stream);
/* Xc = X - mean (fp32), SK/cluster/_kmeans.py:1487-1490.                                   */
int vtf_center_rows(const float* X, int64_t N, int d, const float* mean, float* Xc, void* stream);
/* The same, plus in the same pass what the later stages ask of the centred matrix: stats[0] = max |xc| (fp16 scales of the
 * tensor-core operands, unit of the fixed-point M-step sums), stats[1] = max_row |xc_row|^2 (range of the norm pieces of
 * the sweep's augmented operand), sqn[N] (nullable) = |xc_row|^2, all fp32.                                   */
int vtf_center_rows_stats(const float* X, int64_t N, int d, const float* mean, float* Xc, float* sqn, float* stats,
                          void* stream);

/* ---- a7: k-means++ (SK/cluster/_kmeans.py:180-278) ----------------------------------------
 * kpp_dist: for T candidate centres cand[T,d], out[t,i] = min(closest[i], ||x_i - cand_t||^2)
 * with the squared distance computed in fp64 and stored as fp32, clamped at 0
 * (SK/metrics/pairwise.py:567-638, 413-416); closest == NULL means "no previous centre".
 * pots[t] = sum_i out[t,i] accumulated in fp64 (:263).  ws: vtf_kpp_ws_bytes(T).           */
size_t vtf_kpp_ws_bytes(int T);
int vtf_kpp_dist(const float* Xc, int64_t N, int d, const float* cand, int T,
                 const float* closest, const double* xsq, float* out, double* pots, void* ws, void* stream);
/* xsq[i] = |x_i|^2 in fp64, accumulated in the order vtf_kpp_dist uses: computed once per fit and passed to every
 * round (nullable there: NULL recomputes it in the pass, same bits).                                          */
int vtf_kpp_rownorm64(const float* Xc, int64_t N, int d, double* xsq, void* stream);
/* kpp_pick: idx[t] = searchsorted(cumsum_fp32_sequential(closest), thr[t]) clipped to N-1
 * (:250-254).  thr must be ascending.  `offset_in`/`total_out` chain shards: the running
 * fp32 sum enters as *offset_in (NULL = 0) and leaves in *total_out (nullable); idx[t] = -1
 * when thr[t] is not reached inside this shard.                                            */
int vtf_kpp_pick(const float* closest, int64_t N, const float* thr, int T,
                 const float* offset_in, float* total_out, int64_t* idx, void* stream);
/* The same result again, with most of the array skipped: two multi-CTA passes compute, per tile of 8192 elements, the
 * binade the running sum is expected to enter it with and the exact integer total of the rounded addends under that
 * binade; the walk then jumps over every tile whose true incoming sum is in that binade, that does not carry it out of
 * the binade and that holds no threshold, and scans only the rest (binade crossings, ties, threshold tiles: ~15 of 122 at
 * 1M rows).  Bit-identical to vtf_kpp_pick / vtf_kpp_pick_serial.  ws: vtf_kpp_pick_ws_bytes(N).                  */
size_t vtf_kpp_pick_ws_bytes(int64_t N);
int vtf_kpp_pick_fast(const float* closest, int64_t N, const float* thr, int T, const float* offset_in,
                      float* total_out, int64_t* idx, void* ws, void* stream);
/* Round glue on the device (no host round trip per k-means++ round).  round_prep: best = argmin of the fp32-rounded
 * potentials of the previous round (:266, first minimum), pot = that potential; thr[t] = smallest fp32 >= u[t] * pot
 * (:249; u = the round's uniform draws, made in advance on the host by numpy's MT19937 -- they do not depend on the
 * data), sorted ascending, order[t] = draw index of thr[t].  unsort: cand[order[t]] = picked[t], -1 -> N-1 (:254).  */
int vtf_kpp_round_prep(const double* pots, int n_pots, const double* u, int T, float* thr, int32_t* order,
                       int64_t* best, void* stream);
int vtf_kpp_unsort(const int64_t* picked, const int32_t* order, int T, int64_t N, int64_t* cand, void* stream);
/* The same result from the plain one-thread chain (N dependent fp32 adds).  vtf_kpp_pick reproduces that chain
 * bit for bit with block-wide integer prefix sums inside each binade of the running sum and falls back to single
 * steps only at ties, binade changes and oversized addends; this entry point is the cross-check.  */
int vtf_kpp_pick_serial(const float* closest, int64_t N, const float* thr, int T,
                        const float* offset_in, float* total_out, int64_t* idx, void* stream);

/* ---- a6: Lloyd iteration (SK/cluster/_k_means_lloyd.pyx:23-218) ---------------------------
 * assign: labels[i] = argmin_c (|c|^2 - 2 x_i.c), fp32, first index on ties (:205-213);
 * counts[k] += rows per cluster (int64), *n_changed += rows whose label differs from the
 * previous content of labels[] (strict-convergence test, SK/cluster/_kmeans.py:723).
 * counts and n_changed are accumulated into (the caller zeroes them).                      */
int vtf_lloyd_assign(const float* Xc, int64_t N, int d, const float* C, int k,
                     int32_t* labels, int64_t* counts, int64_t* n_changed, void* stream);
/* cluster_sums: sums[c,:] = sum of rows with label c, fp64, deterministic (block partials
 * reduced in a fixed order) -- the M-step accumulation of :215-218 without its thread-order
 * nondeterminism.  ws: vtf_cluster_sums_ws_bytes(k, d).                                    */
size_t vtf_cluster_sums_ws_bytes(int k, int d);
int vtf_cluster_sums(const float* X, int64_t N, int d, const int32_t* labels, int k,
                     double* sums, void* ws, void* stream);
/* inertia / dispersion: per cluster c, sq[c] = sum_{i in c} ||x_i - C64[c]||^2 and
 * rt[c] = sum_{i in c} ||x_i - C64[c]|| (fp64).  sum(sq) is KMeans.inertia_
 * (SK/cluster/_k_means_common.pyx:94-124) and the intra term of calinski_harabasz_score
 * (SK/metrics/cluster/_unsupervised.py:394-399); rt/n_c is davies_bouldin's intra_dists
 * (:467-474).  Also far_d2[i] (nullable) = ||x_i - C[label_i]||^2 fp32 for the
 * empty-cluster relocation (SK/cluster/_k_means_common.pyx:186).  ws: as cluster_sums.    */
int vtf_cluster_dispersion(const float* X, int64_t N, int d, const int32_t* labels, int k,
                           const double* C64, double* sq, double* rt, float* far_d2,
                           void* ws, void* stream);
/* M-step finish (SK/cluster/_k_means_common.pyx:274-311): C_new[c] = sums[c]/counts[c] (fp32;
 * empty clusters take the centre of the biggest one), stats[0] = sum_c |C_new[c]-C_old[c]|^2,
 * stats[1] = *n_changed, stats[2+c] = counts[c] (doubles: one small D2H per iteration decides
 * convergence, SK/cluster/_kmeans.py:723-738).  Also writes C64 = double(C_new) when given.  */
int vtf_lloyd_update(const double* sums, const int64_t* counts, const int64_t* n_changed, int k,
                     int d, const float* C_old, float* C_new, double* C64, double* stats,
                     void* stream);

/* Lock-step forms: ONE pass over X serves every k of the sweep (grouping.py:99-101 runs one fit per k; the fits
 * are independent, so their iterations can share the reads).  Labelling L owns centre / cluster slots
 * [k_off_h[L], k_off_h[L+1]) of the concatenated arrays (k_off_h is a HOST array of n_lab+1 ints, k_off_h[0] = 0)
 * and the label vector labels[L * label_stride + row]; bit L of `active` clear = that fit has converged and is
 * skipped.  At most 32 labellings and sum(k) <= 100 per call.  Semantics per labelling as the single forms.   */
int vtf_lloyd_assign_multi(const float* Xc, int64_t N, int d, const float* C_all, int n_lab,
                           const int32_t* k_off_h, uint32_t active, int32_t* labels, int64_t label_stride,
                           int64_t* counts, int64_t* n_changed, void* stream);
int vtf_cluster_sums_multi(const float* X, int64_t N, int d, const int32_t* labels, int64_t label_stride,
                           int n_lab, const int32_t* k_off_h, uint32_t active, double* sums, void* ws,
                           void* stream);
int vtf_cluster_dispersion_multi(const float* X, int64_t N, int d, const int32_t* labels,
                                 int64_t label_stride, int n_lab, const int32_t* k_off_h, uint32_t active,
                                 const double* C64, double* sq, double* rt, void* ws, void* stream);
/* stats[L * stats_stride + {0: shift^2, 1: n_changed, 2+c: counts}] for every active labelling.               */
int vtf_lloyd_update_multi(const double* sums, const int64_t* counts, const int64_t* n_changed, int n_lab,
                           const int32_t* k_off_h, uint32_t active, int d, const float* C_old, float* C_new,
                           double* C64, double* stats, int stats_stride, void* stream);

/* ---- a6, wide form: the E-step of a whole k-range as one contraction X.C^T on the tensor cores -----------------
 * For k-ranges such as --clusters '2-32' (grouping.py:99-101 fits KMeans once per k; 527 centres in total) the
 * E-step SGEMM of SK/cluster/_k_means_lloyd.pyx:196-218 is compute-bound: 2*N*sum(k)*d flop per lock-step
 * iteration.  These entry points run it on tcgen05 and keep the M-step proportional to the rows that moved.
 *
 * pack_centres: the column operand of one launch.  Column c takes centre row col_src[c] of C_all (-1 = padding) and
 * the info word col_info[c] = flags | local_index << 4 | labelling << 20 (flag 2 = last centre of its labelling);
 * writes fp16 hi / lo pieces of scale * c [Kp, d], tab[c] = {|c|^2 / 2 (fp32, +inf for padding), info} and
 * *cmax2h = bits of max |c|^2 / 2.  Kp % 256 == 0.  The row operand comes from vtf_dedup_pack (same split).
 * assign: labels_new[L * label_stride + row] = argmin_c (|c|^2/2 - x.c) over the columns of labelling L, first
 * index on ties; scores from three fp16 MMAs (hi.hi + hi.lo + lo.hi), fp32 accumulate in TMEM.  Every (row, L)
 * with another centre within eps * |x_row| * max|c| of the best score is FLAGGED: the sign bit of its label word is
 * set and bits [8, 24) hold the candidate mask (bit (index & 15) set for every candidate; k <= 256 per labelling).
 * refine: decides the flagged (row, L) of the labellings with col0_h[L] >= 0 by exact fp32 dot products (the
 * arithmetic of vtf_lloyd_assign) and clears the flag, so the result equals the fp32 kernels'; *n_fix (nullable)
 * += pairs decided.  rown[P] = |x| per row (0 for padding).  Kp <= 2048 per launch; n_cols = columns in use (the
 * last 256-column tile runs only ceil32 of what it holds: MMA N and epilogue shrink with it).                 */
int vtf_wide_pack_centres(const float* C_all, int d, const int32_t* col_src, const uint32_t* col_info, int Kp,
                          float scale, void* Bhi, void* Blo, float* tab, uint32_t* cmax2h, void* stream);
int vtf_wide_assign(const void* Xhi, const void* Xlo, int64_t n_rows, int64_t P, int d, const void* Bhi,
                    const void* Blo, int Kp, int n_cols, const float* tab, const float* rown, const uint32_t* cmax2h,
                    float scale, float eps, int32_t* labels_new, int64_t label_stride, void* stream);
int vtf_wide_refine(const float* X, int d, const float* C_all, const float* tab, int n_lab, const int32_t* k_off_h,
                    const int32_t* col0_h, int32_t* labels_new, int64_t label_stride, int64_t n, uint32_t* n_fix,
                    void* stream);
/* count: counts[k_off[L] + c] += rows of labelling L with new label c, changed[L] += rows whose new label differs
 * from the old one (SK/cluster/_kmeans.py:723), for the active labellings.  commit: labels_old = labels_new and,
 * for labellings with changed[L] <= cap, items[] += {row, L, old, new} (*n_items zeroed here).                 */
int vtf_wide_count(const int32_t* lab_new, int64_t new_stride, const int32_t* lab_old, int64_t old_stride,
                   int64_t n, int n_lab, const int32_t* k_off_h, uint32_t active, int64_t* counts,
                   int64_t* changed, void* stream);
int vtf_wide_commit(const int32_t* lab_new, int64_t new_stride, int32_t* lab_old, int64_t old_stride, int64_t n,
                    int n_lab, const int32_t* k_off_h, uint32_t active, const int64_t* changed, int64_t cap,
                    void* items, uint32_t* n_items, uint32_t items_cap, void* stream);
/* M-step sums (SK/cluster/_k_means_lloyd.pyx:215-218) in 64-bit fixed point, acc[slot, j] = sum round(x_j * two_s):
 * integer adds are exact and order-independent, so delta (+x to the new slot, -x from the old one, for the listed
 * rows) and full (recompute the labellings of group [L0, L1) with changed[L] > cap from the labels) give
 * bit-identical results, on one GPU or summed over ranks.  ws: vtf_wide_sums_ws_bytes(slots in the group, d);
 * at most 100 slots per group.  update: vtf_lloyd_update_multi on acc / two_s, one CTA per labelling.          */
int vtf_wide_sums_delta(const float* X, int d, const void* items, const uint32_t* n_items, uint32_t items_cap,
                        int n_lab, const int32_t* k_off_h, double two_s, int64_t* acc, void* stream);
size_t vtf_wide_sums_ws_bytes(int k_group, int d);
int vtf_wide_sums_full(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                       const int32_t* k_off_h, uint32_t active, int L0, int L1, const int64_t* changed, int64_t cap,
                       double two_s, int64_t* acc, void* ws, void* stream);
int vtf_wide_update(const int64_t* acc, double inv_two_s, const int64_t* counts, const int64_t* changed, int n_lab,
                    const int32_t* k_off_h, uint32_t active, int d, const float* C_old, float* C_new, double* stats,
                    int stats_stride, void* stream);

/* The same from-scratch sums for NARROW k-ranges (k = 2..6), organised by cells: rows that share their label tuple over
 * the active labellings are grouped inside each 1024-row chunk (shared-memory counting sort), summed in registers -- one
 * add per element instead of one read-modify-write per element and labelling -- and folded into the per-cluster sums once
 * per cell.  acc[slot, j] = sum round(x_j * two_s) for the slots of the ACTIVE labellings (zeroed here; frozen ones keep
 * their values), bit-identical to vtf_wide_sums_full.  Active labellings are processed in groups of at most 8 with at
 * most 1024 label tuples and 48 cluster slots; k <= 48 per labelling.  *bad (nullable) = labels outside [0, k).
 * ws: 64 bytes (chunk counters).  Replaces SK/cluster/_k_means_lloyd.pyx:215-218 behind grouping.py:100.          */
int vtf_cell_sums(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                  const int32_t* k_off_h, uint32_t active, double two_s, int64_t* acc, uint32_t* bad, void* ws,
                  void* stream);

/* Dispersion for narrow k-ranges, organised by cells like vtf_cell_sums: out[slot] = sum_{i in slot} |x_i - C64[slot]|^2 *
 * sq_unit and out[ktot + slot] = sum |x_i - C64[slot]| * rt_unit as int64 (zeroed here).  Labelling q measures against the
 * centres of slots [k_off_h[q], k_off_h[q+1]) with the labels of row row_of_h[q] of the label matrix (NULL: row q): several
 * labellings may share a label row (a fit measured against its centres AND against the means of its final labels).  Per
 * row: e = (x - fp32(c)) - (c - fp32(c)) and the sum of e^2 in fp32 (the reference's inertia is fp32 too; ~1e-6 relative),
 * then exact integer accumulation: the result does not depend on the order of rows, CTAs or ranks.  One launch: at most 12
 * labellings over at most 8 distinct label rows / 1024 joint label tuples, k <= 48, centres within 160 KB; d % 128 == 0,
 * d <= 1024.  The units must keep the totals below 2^62.  ws: 64 bytes.  Replaces vtf_cluster_dispersion_multi on the
 * default path (SK/cluster/_k_means_common.pyx:94-124, SK/metrics/cluster/_unsupervised.py:394-399, 467-474).            */
int vtf_cell_dispersion(const float* X, int64_t n, int d, const int32_t* labels, int64_t label_stride, int n_lab,
                        const int32_t* k_off_h, const int32_t* row_of_h, const double* C64, double sq_unit, double rt_unit,
                        int64_t* out, uint32_t* bad, void* ws, void* stream);

/* Cell layout of the silhouette sweep on the device (replaces a torch.unique / argsort chain): labels[n_lab][label_stride]
 * are ALL labellings of the k-range (k_h[q] clusters each, at most 4096 label tuples in total).  Occupied tuples = cells,
 * numbered g = 0..G-1 in key order (key = sum_q label_q * prod_{q' < q} k_q'); cell g takes ceil32(size) positions, the
 * total is padded to a multiple of 256.  perm[p] = row at position p (-1: padding), chunk_cell[p/32] = cell of the 32
 * positions, c2c[q * c2c_stride + g] = cluster of cell g in labelling q, meta = {G, P, positions in use}.  perm_cap >=
 * ceil256(N + 31 * min(tuples, N)), chunk_cell holds perm_cap/32 entries, c2c_stride >= min(tuples, N).
 * ws: vtf_build_cells_ws_bytes(N).                                                                           */
size_t vtf_build_cells_ws_bytes(int64_t N);
int vtf_build_cells(const int32_t* labels, int64_t label_stride, int64_t N, int n_lab, const int32_t* k_h, int64_t* perm,
                    int64_t perm_cap, int32_t* chunk_cell, int32_t* c2c, int c2c_stride, int64_t* meta, void* ws,
                    void* stream);

/* ---- a3: nearest-reference classification (grouping.py:51-56) -----------------------------
 * Xn, Rn are cosine-normalised.  dist[N,R] = clip(1 - Xn.Rn^T, 0, 2) (fp32, kept for the
 * '%.4f' log, grouping.py:63); label[i] = argmin (first index), or R when
 * min >= thr (fp32 compare, grouping.py:55); thr < 0 or NaN disables the 'other' class.    */
int vtf_cosine_classify(const float* Xn, int64_t N, int d, const float* Rn, int R, float thr,
                        float* dist, int64_t* label, void* stream);

/* ---- a1: all-pairs duplicate marking (dupes.py:60-65) -------------------------------------
 * Xn cosine-normalised [N,d].  For the row tiles this rank owns (tile t of 64 rows belongs to
 * rank t % world) computes key[i] = (bits(min_{j<i} D_ij) << 32) | argmin_j with
 * D = clip(1 - xn_i.xn_j, 0, 2) in fp32 FMA arithmetic; row 0 keeps (bits(10000.f)<<32)|0,
 * which is what the reference's +10000 mask (dupes.py:62) yields.  key must be pre-filled
 * with vtf_dedup_init.  min/argmin decode: dist = float(key >> 32), j = key & 0xffffffff.   */
int vtf_dedup_init(uint64_t* key, int64_t N, void* stream);
int vtf_cosine_dedup(const float* Xn, int64_t N, int d, int rank, int world,
                     uint64_t* key, void* stream);
/* Tensor-core form of the same (tcgen05, CTA pairs; used for large N).  Decisions must match an fp32 SGEMM,
 * so pack splits scale*Xn into two fp16 pieces hi + lo (22 significant bits; rows >= N zero, P % 256 == 0)
 * and the sweep accumulates hi.hi + hi.lo + lo.hi in fp32 in TMEM, then applies the same
 * clip(1 - S, 0, 2) / j < i / min / first-argmin epilogue.  Units of 256 rows, unit u owned by rank u % world
 * (counted from the bottom so the triangular work balances).  key as for vtf_cosine_dedup.   */
int vtf_dedup_pack(const float* Xn, int64_t N, int d, int64_t P, float scale, void* Hi, void* Lo, void* stream);
int vtf_cosine_dedup_tc(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale,
                        int rank, int world, uint64_t* key, void* stream);
/* The tensor core accumulates fp32 with truncation (~3e-6 relative after 96 steps): refine re-evaluates
 * every owned row's winner, key[i] = (bits(clip(1 - xn_i.xn_j, 0, 2)) << 32) | j with an exact fp32 dot.  */
int vtf_dedup_refine(const float* Xn, int64_t N, int d, int rank, int world, uint64_t* key, void* stream);
/* Two-stage form (the default for large N): a SCREEN with one MMA per K step on the rounded operand Hi alone -- a third of
 * the tensor work of vtf_cosine_dedup_tc -- that keeps, per owned row and column half, the four largest Gram values among the
 * earlier rows: cand[N][8] = {float bits of g, column} (half 0 in descending order, then half 1; column 0xffffffff =
 * none).  refine4 evaluates those candidates with exact fp32 dot products, writes key[i] = (bits(min dist) << 32) | first
 * argmin, and flags unit_flag[i / 256] (*n_flagged += newly flagged units) for rows whose decision at `thr` the screen
 * cannot guarantee: exact minimum above thr although an unseen column might be below (1 - g_unseen/scale^2 - eps <= thr;
 * eps = bound of the screen's error, 1e-3 for unit rows; thr = NaN: never).  The flagged units are re-swept with the split
 * operands by vtf_cosine_dedup_tc_units (unit_list: device array of unit indices) followed by vtf_dedup_refine.
 * The screen cuts the triangle into column chunks of <= 48 MB of operand (one launch each, so a chunk comes from DRAM once)
 * and claims the (unit x chunk) items of a launch from work_counters[launch], longest first: the caller passes
 * vtf_dedup_screen_launches(N, d) ZEROED counters.                                                                      */
int vtf_dedup_screen_launches(int64_t N, int d);
int vtf_dedup_screen(const void* Hi, int64_t N, int64_t P, int d, int rank, int world, void* cand, uint32_t* work_counters,
                     int n_counters, void* stream);
int vtf_dedup_refine4(const float* Xn, int64_t N, int d, int rank, int world, const void* cand, float scale, float thr,
                      float eps, uint64_t* key, int32_t* unit_flag, uint32_t* n_flagged, void* stream);
int vtf_cosine_dedup_tc_units(const void* Hi, const void* Lo, int64_t N, int64_t P, int d, float scale,
                              const int32_t* unit_list, int64_t n_list, uint64_t* key, void* stream);
/* 'hash' branch (dupes.py:55-58; detection.py:57-59): bits[N,64] int64 0/1 average hashes (dupes.py:11-15)
 * -> packed[N] uint64; key[i] = (min_{j<i} popc(h_i ^ h_j) << 32) | argmin_j, row 0 = 10000 << 32 (the
 * reference's uint16 distance + 10000 mask).  Rows are owned in 256-row tiles, tile t by rank t % world.  */
int vtf_hash_pack(const int64_t* bits, int64_t N, uint64_t* packed, void* stream);
int vtf_hash_dedup(const uint64_t* packed, int64_t N, int rank, int world, uint64_t* key, void* stream);
/* X_out[r,:] = X[idx[r],:]  -- np.delete(X, idx, axis=0) with idx = kept rows (dupes.py:68) */
int vtf_gather_rows(const float* X, int d, const int64_t* idx, int64_t n_out, float* X_out,
                    void* stream);

/* ---- a8: silhouette (SK/metrics/cluster/_unsupervised.py:149-200, 211-323) ----------------
 * Exact path (fp32 FMA Gram, any N; used for small N and as the on-device cross-check):
 * rowsum[i, c] += sum_{j: label_j == c} sqrt(max(n_i + n_j - 2 x_i.x_j, 0)), D_ii = 0, for
 * rows i in [row_begin, row_end).  rowsum is fp64 [N, k], zeroed by the caller.            */
int vtf_silhouette_rowsums_exact(const float* X, const float* sqnorm, int64_t N, int d,
                                 const int32_t* labels, int k, int64_t row_begin,
                                 int64_t row_end, double* rowsum, void* stream);
int vtf_row_sqnorms(const float* X, int64_t N, int d, float* sqnorm, void* stream);

/* Tensor-core path (tcgen05 + TMA, the dominant kernel at BASELINE sizes).
 * Distances do not depend on k, so ONE sweep over the Gram matrix serves every labelling of
 * the k-range: the caller groups columns by their joint label tuple ("cell"), pads every
 * cell to a multiple of 32 columns, and the kernel accumulates, for each row, one distance
 * sum per cell.  See DESIGN.md for the layout.
 *
 * pack: Xh[p,:] = fp16(scale * X[perm[p],:]) for p < P (rows with perm[p] < 0 are zero padding) and the
 * augmented operand aug = [A flavour | B flavour], each P*16 fp16 in UMMA no-swizzle K-major tile order:
 * three fp16 pieces of m_p * 2^-shift (m_p = |Xh[p,:]|^2 / 2 from the ROUNDED values, so the distance is
 * the exact distance between the rounded points) against the constant 2^shift, so that ONE extra K=16 MMA
 * adds -m_p - m_q to the Gram accumulator.  Padding rows get an all-zero augmented block, so every accumulator
 * entry that involves one is exactly 0 (distance 0); pad_m (>= 4 * max m) only bounds the shift.  half_sqn[P]
 * (nullable) receives m (-pad_m for padding rows).  P % 256 == 0, d % 64 == 0.                       */
int vtf_sil_pack(const float* X, int64_t N, int d, const int64_t* perm, int64_t P, float scale,
                 int shift, float pad_m, void* Xh, void* aug, float* half_sqn, void* stream);
/* diag: out[p] = out_scale * T_pp for every row, T exactly as the sweeps accumulate it (same MMA shape and order, the
 * augmented block last).  With an all-zero `aug` this is the squared norm of row p of Xh IN THE TENSOR CORE'S ARITHMETIC
 * (truncating fp32 accumulation); out_scale = 0.5 gives the half squared norms m.  aug (below) builds the augmented
 * operand from such m[P] (perm[p] < 0, or p >= N without perm = padding row: all-zero block).  Norms from here make T_pq = g_pq - m_p - m_q
 * cancel when rows p and q coincide and take the accumulation bias out of every small distance -- with the fp32-FMA norms
 * of vtf_sil_pack that bias is a floor of ~2e-3 |x| (duplicate frames, tight clusters).  With the final aug,
 * vtf_sil_diag(out_scale = 1) returns the residual T_pp the sweep would see for a duplicate pair (0 when exact).       */
int vtf_sil_diag(const void* Xh, const void* aug, int64_t P, int d, float out_scale, float* out, void* stream);
int vtf_sil_aug(const float* m, const int64_t* perm, int64_t N, int64_t P, int shift, void* aug, void* stream);
/* out[0] = max |x| (picks the power-of-two fp16 scale).                                     */
int vtf_absmax(const float* X, int64_t N, int d, float* out, void* stream);
/* sweep: cellsum[g, p] += sum_{q in cell g} sqrt(max(2 (m_p + m_q - xh_p.xh_q), 0)) / scale
 * for row blocks [rb_begin, rb_end) of 128 rows, D_pp = 0.  chunk_cell[P/32] gives the cell
 * of every 32-column chunk.  cellsum is fp32 [n_cells, P], zeroed by the caller.
 * Requires sm_100a; d in {64..1024}, d % 64 == 0.  Even row-block ranges run as CTA pairs
 * (cta_group::2); VTF_SWEEP_VARIANT=1 in the environment forces the single-CTA form.        */
int vtf_sil_sweep(const void* Xh, const void* aug, int64_t P, int d, float scale,
                  const int32_t* chunk_cell, int n_cells, int64_t rb_begin, int64_t rb_end,
                  float* cellsum, void* stream);
/* Symmetric form (D_pq = D_qp): only tiles on or above the diagonal are computed -- half the MMAs -- and every
 * off-diagonal block also feeds the row sums of its columns (warp transpose-reduce in the epilogue).  Rank r of
 * `world` sweeps unit pairs (q, U-1-q), q = r, r + world, ... (U = P/256 units of 256 rows); each rank's cellsum
 * then holds partial sums for ALL rows and the caller adds them (all-reduce) before vtf_sil_finish_cells.
 * work_counter != NULL switches to the DYNAMIC schedule: the pairs are cut into sub-ranges and every CTA pair of every
 * rank claims the next one with an atomicAdd (system scope) on *work_counter, which must be 0 before the first rank
 * starts and is shared by all ranks (vtf_shared_alloc on one rank, vtf_shared_open on the others: the claims travel
 * over NVLink).  A slower GPU then simply takes fewer items.  tiles_done (nullable) += 256 x 256 tiles computed here. */
int vtf_sil_sweep_sym(const void* Xh, const void* aug, int64_t P, int d, float scale,
                      const int32_t* chunk_cell, int n_cells, int rank, int world, float* cellsum,
                      uint32_t* work_counter, uint64_t* tiles_done, void* stream);
/* Debugging aid: with VTF_SWEEP_TIMING=1 in the environment vtf_sil_sweep_sym records the %globaltimer value of its first CTA
 * entry and last CTA exit; this returns the pair of the last launch (ns) after synchronising the device.            */
int vtf_debug_sweep_timing(uint64_t* out2);
/* finish: for labelling l (k_l clusters), per real row: a = S_own/(n_own-1),
 * b = min_{c != own} S_c/n_c, s = (b-a)/max(a,b) (0 for singleton clusters, :323), and
 * out[l] += sum_i s_i in fp64.  From the exact path: vtf_sil_finish_rows on rowsum[N,k].    */
int vtf_sil_finish_rows(const double* rowsum, int64_t N, int k, const int32_t* labels,
                        const int64_t* counts, int64_t row_begin, int64_t row_end,
                        double* out_sum, void* stream);
/* From the tensor-core path: cellsum[n_cells,P] plus cell_to_cluster[n_cells] for this
 * labelling; perm[p] < 0 marks padding rows.                                               */
int vtf_sil_finish_cells(const float* cellsum, int64_t P, int n_cells,
                         const int32_t* cell_to_cluster, int k, const int64_t* perm,
                         const int32_t* labels, const int64_t* counts,
                         int64_t p_begin, int64_t p_end, double* out_sum, void* stream);

/* ---- f3: test_grouping (grouping.py:140-155) -- rand_score's contingency table ------------------
 * table[a[i]*nb + b[i]] += 1 over the rows given (sklearn.metrics.cluster.contingency_matrix, the input of
 * pair_confusion_matrix / rand_score, SK/metrics/cluster/_supervised.py:106-192, 195-271, 280-357); a in [0,na), b in [0,nb),
 * anything else is counted in *bad.  Shards add across ranks.                                          */
int vtf_contingency(const int32_t* a, const int32_t* b, int64_t N, int na, int nb, int64_t* table, int64_t* bad,
                    void* stream);

/* ---- device memory shared between the ranks of one node (CUDA IPC; the sweep's work counter) ------------------
 * alloc: cudaMalloc + zero + export a 64-byte handle (owner rank).  open: map another process's allocation (peer
 * access over NVLink is enabled on first use).  zero: stream-ordered memset, valid on owner and opened pointers.   */
int vtf_shared_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int vtf_shared_open(const unsigned char handle[64], void** ptr);
int vtf_shared_zero(void* ptr, size_t bytes, void* stream);
int vtf_shared_close(void* ptr);
int vtf_shared_free(void* ptr);

/* Host matrix in PAGEABLE memory -> device (main.py:71 hands the grouping stage an ordinary numpy array): n_threads native
 * threads copy 8 MB pieces into page-locked double buffers and queue each piece's DMA on their own streams, so the host copies
 * run in parallel and overlap the transfers; `stream` waits for all pieces (nothing is ordered before that).  Returns after the
 * host side is done: the source may be reused at once.  Pinned sources: use one plain cudaMemcpyAsync instead.            */
int vtf_upload_pageable(const void* src_host, void* dst_dev, size_t bytes, int n_threads, void* stream);
/* ---- f4: output stage (grouping.py:78-84, 122-137; dupes.py:70-79) -- HOST code, no kernel -----------------------
 * A batch of independent file operations on native threads: op 0 = copy src -> dst (shutil.copyfile: contents only,
 * copy_file_range with a read/write fallback), 1 = move (os.replace), 2 = remove (os.remove; dst ignored).  Paths are
 * NUL-terminated strings inside a host blob, path i at blob + off[i].  Every operation is attempted; returns the number
 * of failures (-1: bad arguments) and reports the lowest failing index with its errno.                       */
int64_t vtf_file_ops(int op, const char* src_blob_h, const int64_t* src_off_h, const char* dst_blob_h,
                     const int64_t* dst_off_h, int64_t n, int threads, int64_t* first_index_h, int* first_errno_h);

#ifdef __cplusplus
}
#endif
#endif /* VTF_B200_H */
