"""Output stage: files per second of the reference's per-file Python loops against the native batch
(vtf_file_ops) on the file systems of this machine.

    python tools/fileops_bench.py [--files 20000] [--bytes 20000] [--dirs /tmp,/dev/shm] > profiles/r02_fileops_bench.json
"""
import argparse
import json
import os
import os.path as osp
import shutil
import sys
import tempfile
import time

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, osp.join(ROOT, 'video-to-faces_b200'))

import numpy as np  # noqa: E402


def make(root, n, size):
    os.makedirs(root)
    payload = os.urandom(size)
    paths = []
    for i in range(n):
        p = osp.join(root, 'f%06d.jpg' % i)
        with open(p, 'wb') as f:
            f.write(payload)
        paths.append(p)
    return paths


def run(base, n, size, threads):
    from videotofaces_b200 import fileops
    os.environ['VTF_IO_THREADS'] = str(threads)
    root = tempfile.mkdtemp(dir=base)
    try:
        paths = make(osp.join(root, 'faces'), n, size)
        labels = np.random.default_rng(0).integers(0, 6, n)
        for j in range(6):
            os.makedirs(osp.join(root, 'faces', str(j)))
        os.makedirs(osp.join(root, 'dupes'))
        os.sync()
        time.sleep(0.5)
        t0 = time.perf_counter()
        fileops.copy_files((p, osp.join(root, 'faces', str(l), osp.basename(p))) for p, l in zip(paths, labels))   # grouping.py:129-130
        t1 = time.perf_counter()
        fileops.remove_files(paths)                                                                               # grouping.py:136-137
        t2 = time.perf_counter()
        moved = [osp.join(root, 'faces', str(l), osp.basename(p)) for p, l in zip(paths, labels)]
        fileops.move_files((p, osp.join(root, 'dupes', osp.basename(p))) for p in moved)                           # dupes.py:79
        t3 = time.perf_counter()
        return dict(threads=threads, copy_files_per_s=n / (t1 - t0), remove_files_per_s=n / (t2 - t1), move_files_per_s=n / (t3 - t2),
                    cluster_faces_output_s=t2 - t0)
    finally:
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--files', type=int, default=20000)
    ap.add_argument('--bytes', type=int, default=20000)
    ap.add_argument('--dirs', default='/tmp,/dev/shm')
    args = ap.parse_args()
    out = dict(files=args.files, bytes=args.bytes, cores=os.cpu_count(), results={})
    for base in args.dirs.split(','):
        if not osp.isdir(base):
            continue
        fs = os.popen("df -T %s | tail -1 | awk '{print $2}'" % base).read().strip()
        rows = []
        # a disk-backed file system throttles writers once earlier runs' dirty pages are being written back, so whoever
        # runs later looks slower: flush before every run, visit the modes in both orders, keep each mode's best
        best = {}
        for threads in (0, 1, 2, 4, 8, 8, 4, 2, 1, 0):
            os.sync()
            time.sleep(0.5)
            r = run(base, args.files, args.bytes, threads)
            if threads not in best or r['cluster_faces_output_s'] < best[threads]['cluster_faces_output_s']:
                best[threads] = r
        for threads in (0, 1, 2, 4, 8):
            best[threads]['mode'] = 'python loop (reference)' if threads == 0 else 'native batch'
            rows.append(best[threads])
        out['results']['%s (%s)' % (base, fs)] = rows
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main()
