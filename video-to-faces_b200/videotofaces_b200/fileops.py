"""The output stage of the grouping path: the reference moves, copies and deletes one
file per Python loop iteration (/root/reference/src/videotofaces/grouping.py:78-84, 122-137, dupes.py:70-79).  Once the
arithmetic takes seconds these loops are the wall-clock floor (1M faces = 1M copyfile + 1M remove calls).

All file operations of the drop-in go through this module so that they can be batched onto a thread pool
(VTF_IO_THREADS=n): the calls are independent and the resulting directory tree is the same in any order; errors
surface like in the loop (the first one is raised) but only after the batch has drained.  MEASURED on the build
container (8 cores, overlay file system, 20 000 files of 20 KB, copy into 6 folders): plain loop 14 200 files/s,
4 threads 1 700 files/s, 4 processes 4 100 files/s -- metadata operations serialise in that file system and the GIL
hand-offs between threads cost more than the copies, so the DEFAULT IS THE PLAIN LOOP (1 thread).  The pool is for
file systems where a call waits on the network (NFS, FUSE object stores), which this image cannot reproduce.
"""
import os
import shutil
from concurrent.futures import ThreadPoolExecutor

_CHUNK = 256            # files per task: amortises the executor overhead, keeps the pool busy


def io_threads():
    v = os.environ.get('VTF_IO_THREADS')
    return max(1, int(v)) if v else 1


def _run(fn, pairs):
    """fn(*pair) for every pair; threaded when there is enough work."""
    pairs = list(pairs)
    workers = io_threads()
    if workers == 1 or len(pairs) < 2 * _CHUNK:
        for p in pairs:
            fn(*p)
        return

    def task(lo):
        first = None
        for p in pairs[lo:lo + _CHUNK]:
            try:
                fn(*p)
            except Exception as e:  # noqa: BLE001  (kept, raised once the batch is through)
                first = first or e
        if first is not None:
            raise first

    with ThreadPoolExecutor(max_workers=workers) as ex:
        futures = [ex.submit(task, lo) for lo in range(0, len(pairs), _CHUNK)]
        first = None
        for f in futures:
            e = f.exception()
            if e is not None and first is None:
                first = e
        if first is not None:
            raise first


def copy_files(pairs):
    """shutil.copyfile(src, dst) for every (src, dst)  -- grouping.py:130."""
    _run(shutil.copyfile, pairs)


def move_files(pairs):
    """os.replace(src, dst) for every (src, dst)  -- grouping.py:80, dupes.py:79."""
    _run(os.replace, pairs)


def remove_files(paths):
    """os.remove(p) for every p  -- grouping.py:136-137, dupes.py:72-73."""
    _run(os.remove, ((p,) for p in paths))
