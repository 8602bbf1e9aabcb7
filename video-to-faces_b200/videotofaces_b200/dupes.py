"""Duplicate removal over embeddings -- drop-in for videotofaces.dupes.remove_dupes_overall
(/root/reference/src/videotofaces/dupes.py:51-92).

Both branches run on the GPU and never materialise the N x N matrix: 'enc' (cosine distance to every
EARLIER row, mark when min <= thr; called from main.py:74) and 'hash' (Hamming distance between 64-bit
average hashes, dupes.py:55-58, called from detection.py:57-59 -- the first "next" row of SURVEY.md
section 8f).  File moves and log_dupes{2,3}.csv keep the reference's format.
"""
import os
import os.path as osp

import numpy as np

from .fileops import move_files, remove_files
from .runtime import get_engine, is_main, sync

TIE_EPS = 1e-6        # |distance - threshold| at or below this is reported as a tie (BASELINE.json north_star)


def _write_dupes_log(path, filenames, mins, inds, is_dup, measure_name):
    # dupes.py:85-88: full paths in the first two columns, rows start at i = 1
    with open(path, 'w') as f:
        f.write('file_name,nearest_in_prev,' + measure_name + ',marked_as_duplicate\n')
        for i in range(1, len(filenames)):
            f.write('%s,%s,%s,%s\n' % (filenames[i], filenames[inds[i]], str(mins[i]), '1' if is_dup[i] else '0'))


def dedup_rows(X, threshold, keep_on_device=False):
    """The arithmetic of remove_dupes_overall's 'enc' branch (dupes.py:60-68) without file system effects: returns
    dict(kept [N - n_dup, d] (numpy for a host X, a CUDA tensor for a device X or with keep_on_device), idx int64 [n_dup],
    mins fp32 [N], inds int64 [N]).  keep_on_device: the surviving rows stay in HBM for the grouping call that follows
    (main.py:74-80 only passes them on) -- one upload of X instead of upload + download + upload."""
    eng = get_engine()
    on_device = hasattr(X, 'is_cuda') and X.is_cuda                            # encoder output still on the GPU (section 8f item 2)
    if keep_on_device and not on_device:
        X = eng.to_device(X)                                                   # the ONE host-to-device copy of the flow
        on_device = True
    mins, inds, idx = eng.dedup(X, threshold)                                  # dupes.py:60-65
    if on_device:
        kept = eng.delete_rows(X, idx) if idx.size else X.clone()              # stays in HBM for the grouping that follows
    else:
        kept = eng.to_host([eng.delete_rows(X, idx)])[0] if idx.size else np.array(X, dtype=np.float32, copy=True)
    return dict(kept=kept, idx=idx, mins=mins, inds=inds)


def remove_dupes_overall(X, filenames, dup_params, keep_on_device=False):
    """Returns (X_kept float32 [N - n_dup, d], goods) exactly like dupes.py:51-92.  keep_on_device (used by video_to_faces,
    which only hands X_kept to the grouping call): X_kept is a CUDA tensor instead of a numpy array ('enc' branch)."""
    measure_type, threshold, save_dupes, out_dir = dup_params
    eng = get_engine()
    n = X.shape[0]
    is_hash = measure_type == 'hash'
    if is_hash:
        mins, inds, idx = eng.dedup_hash(X, threshold)                         # dupes.py:55-58, 62-65
    else:
        r = dedup_rows(X, threshold, keep_on_device)                           # dupes.py:60-65, 68
        mins, inds, idx = r['mins'], r['inds'], r['idx']
    is_dup = np.zeros(n, dtype=bool)
    is_dup[idx] = True
    dupes = [fn for i, fn in enumerate(filenames) if is_dup[i]]                # dupes.py:66-67
    goods = [fn for i, fn in enumerate(filenames) if not is_dup[i]]
    if is_hash:
        X_kept = np.delete(np.asarray(X), idx, axis=0)                         # 64 ints per row: host bookkeeping
    else:
        X_kept = r['kept']

    face_dir = osp.join(out_dir, 'faces')
    if not is_main():                                                          # several ranks: rank 0 owns files, logs and console
        pass
    elif not save_dupes:                                                       # dupes.py:70-74
        remove_files(osp.join(face_dir, osp.basename(fn)) for fn in dupes)
    else:                                                                      # dupes.py:75-88
        mdigit, mname = ('2', 'hash_diff') if is_hash else ('3', 'distance')
        dup_dir = osp.join(out_dir, 'intermediate', 'dupes' + mdigit)
        os.makedirs(dup_dir, exist_ok=True)
        move_files((osp.join(face_dir, osp.basename(fn)), osp.join(dup_dir, osp.basename(fn))) for fn in dupes)
        _write_dupes_log(osp.join(out_dir, 'intermediate', 'log_dupes' + mdigit + '.csv'), filenames, mins, inds, is_dup, mname)

    if is_main():
        if not is_hash and idx.any():                                          # dupes.py:90-91
            print('Removed %u near-duplicates' % idx.shape[0])
        if not is_hash and n > 1:
            ties = int(np.count_nonzero(np.abs(mins[1:] - np.float32(threshold)) <= TIE_EPS))
            if ties:                                                           # silent in the reference; north_star asks for it
                print('NOTE: %u images have their nearest earlier image within %g of enc_dup_thr=%s' % (ties, TIE_EPS, str(threshold)))
    sync()
    return X_kept, goods
