"""Build libvtf_b200.so (sm_100a only) in-tree with nvcc; no JIT cache, no torch extension glue.

    python video-to-faces_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import os.path as osp
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = osp.dirname(osp.abspath(__file__))
CSRC = osp.join(HERE, 'csrc')
OUT_DIR = osp.join(HERE, 'videotofaces_b200', '_lib')
LIB = osp.join(OUT_DIR, 'libvtf_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def sources():
    return sorted(osp.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def newest_input():
    files = sources() + [osp.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    files.append(osp.join(HERE, '..', 'include', 'vtf_b200.h'))
    return max(osp.getmtime(f) for f in files)


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and osp.exists(LIB) and osp.getmtime(LIB) >= newest_input():
        return LIB
    if not osp.exists(NVCC):
        raise RuntimeError('nvcc not found at %s and no prebuilt %s' % (NVCC, LIB))
    obj_dir = osp.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src):
        obj = osp.join(obj_dir, osp.basename(src)[:-3] + '.o')
        cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources()))
    log = osp.join(obj_dir, 'ptxas.log')
    with open(log, 'w') as f:
        for obj, err in results:
            f.write('==== %s\n%s\n' % (obj, err))
    cmd = [NVCC, '-shared', '-o', LIB] + [o for o, _ in results] + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    if verbose:
        print(open(log).read())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
