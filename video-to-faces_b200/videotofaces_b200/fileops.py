"""The output stage of the grouping path: the reference moves, copies and deletes one
file per Python loop iteration (/root/reference/src/videotofaces/grouping.py:78-84, 122-137, dupes.py:70-79).  Once the
arithmetic takes seconds these loops are the wall-clock floor (1M faces = 1M shutil.copyfile + 1M os.remove calls).

All file operations of the drop-in go through this module.  Batches of VTF_IO_MIN_BATCH (256) files or more are handed to
the native library in ONE call (include/vtf_b200.h, vtf_file_ops: native threads, no GIL, copy_file_range); the paths
travel as one NUL-separated blob plus an offset array.  Smaller batches, and VTF_IO_THREADS=0, use the reference's plain
loop.  The resulting directory tree is the same either way (tests/test_output_stage.py); errors surface like in the loop
-- the exception of the first failing file -- but only after the whole batch has been attempted.

Measured in the build container (8 cores; 20 000 files of 20 KB copied into 6 folders, then removed; tools/fileops_bench.py,
profiles/r02_fileops_bench.json): see DESIGN.md section 8.
"""
import ctypes
import os
import shutil

import numpy as np

MIN_BATCH = int(os.environ.get('VTF_IO_MIN_BATCH', '256'))


def io_threads():
    """Native worker threads: VTF_IO_THREADS, default min(8, cores); 0 = the reference's plain Python loop."""
    v = os.environ.get('VTF_IO_THREADS')
    if v is not None and v != '':
        return max(0, int(v))
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    return max(1, min(8, n))


def _blob(paths):
    enc = [os.fsencode(p) for p in paths]
    lens = np.fromiter((len(b) + 1 for b in enc), dtype=np.int64, count=len(enc))
    off = np.zeros(len(enc), dtype=np.int64)
    np.cumsum(lens[:-1], out=off[1:])
    return b'\0'.join(enc) + b'\0', off


_PLAIN = {0: shutil.copyfile, 1: os.replace, 2: os.remove}


def _run(op, srcs, dsts=None):
    srcs = list(srcs)
    threads = io_threads()
    if threads == 0 or len(srcs) < MIN_BATCH:
        fn = _PLAIN[op]
        for i, s in enumerate(srcs):
            fn(s) if dsts is None else fn(s, dsts[i])
        return
    from . import _native
    lib = _native.load()                        # raises when the library is missing: no silent fallback for big batches
    sb, so = _blob(srcs)
    if dsts is not None:
        db, do = _blob(dsts)
        dptr, doff = db, do.ctypes.data_as(ctypes.c_void_p)
    else:
        dptr, doff = None, None
    first, err = ctypes.c_int64(-1), ctypes.c_int(0)
    failed = lib.vtf_file_ops(op, sb, so.ctypes.data_as(ctypes.c_void_p), dptr, doff, len(srcs), threads,
                              ctypes.byref(first), ctypes.byref(err))
    if failed < 0:
        raise ValueError('vtf_file_ops: bad arguments')
    if failed:
        i = first.value
        raise OSError(err.value, os.strerror(err.value), srcs[i], None, None if dsts is None else dsts[i])


def copy_files(pairs):
    """shutil.copyfile(src, dst) for every (src, dst)  -- grouping.py:130."""
    pairs = list(pairs)
    _run(0, [p[0] for p in pairs], [p[1] for p in pairs])


def move_files(pairs):
    """os.replace(src, dst) for every (src, dst)  -- grouping.py:80, dupes.py:79."""
    pairs = list(pairs)
    _run(1, [p[0] for p in pairs], [p[1] for p in pairs])


def remove_files(paths):
    """os.remove(p) for every p  -- grouping.py:136-137, dupes.py:72-73."""
    _run(2, list(paths))
