"""`python -m videotofaces_b200 ...` -- the reference CLI (videotofaces/__main__.py:21-58) for the
flags the grouping stage reads, with the same names and defaults (note the reference's own
mismatch: --clusters defaults to '2-9' here but None -> 2..8 in Python, __main__.py:50 vs
prep.py:49-50).  Flags of the detection stage are accepted and ignored so existing command lines
keep working; --features / --ref-features feed precomputed embeddings.
"""
import argparse

from .main import video_to_faces


def build_parser():
    p = argparse.ArgumentParser(prog='python -m videotofaces_b200',
                                description='B200-native grouping stage of video-to-faces')
    p.add_argument('-i', '--input-path', metavar='PATH')
    p.add_argument('-e', '--input-ext', metavar='EXTENSIONS')
    p.add_argument('-o', '--out-dir', metavar='PATH')
    p.add_argument('-op', '--out-prefix', metavar='TEXT', default='')
    p.add_argument('-s', '--style', metavar='TEXT', required=True)
    p.add_argument('-m', '--mode', metavar='TEXT', default='grouping')
    p.add_argument('-d', '--device', metavar='TEXT')
    p.add_argument('--save-dupes', action='store_true')
    p.add_argument('--enc-model', metavar='TEXT', default='default')
    p.add_argument('--enc-batch-size', metavar='INT', type=int, default=16)
    p.add_argument('--enc-area', metavar='N', type=float, nargs=4)
    p.add_argument('--enc-dup-thr', metavar='FLOAT', type=float, default=0.25)
    p.add_argument('--group-mode', metavar='TEXT', default='clustering')
    p.add_argument('--clusters', metavar='TEXT', default='2-9')
    p.add_argument('--clusters-save-all', action='store_true')
    p.add_argument('--random-state', metavar='INT', type=int, default=0)
    p.add_argument('--ref-dir', metavar='PATH')
    p.add_argument('--enc-oth-thr', metavar='FLOAT', type=float, default=0.9)
    p.add_argument('--group-log', action='store_true')
    p.add_argument('--features', metavar='NPY', help='precomputed embeddings [N,d] float32 (.npy), rows in sorted file order')
    p.add_argument('--normalize-features', action='store_true',
                   help='--features are raw encoder outputs: apply the L2 normalisation FaceNet ends with, fused into the first pass')
    p.add_argument('--ref-features', metavar='NPY', help='precomputed reference embeddings [R,d] (.npy), rows in sorted class order')
    # detection-stage flags of the reference CLI: parsed for compatibility, unused here
    for flag, kw in (('--save-frames', dict(action='store_true')), ('--save-rejects', dict(action='store_true')),
                     ('--video-step', dict(type=float, default=1)), ('--video-fragment', dict(type=float, nargs=2)),
                     ('--video-area', dict(type=int, nargs=4)), ('--video-reader', dict(default='opencv')),
                     ('--det-model', dict(default='default')), ('--det-batch-size', dict(type=int, default=4)),
                     ('--det-min-score', dict(type=float, default=0.4)), ('--det-min-size', dict(type=int, default=50)),
                     ('--det-min-border', dict(type=int, default=5)),
                     ('--det-scale', dict(type=float, nargs=4, default=[1.5, 1.5, 2.2, 1.2])),
                     ('--det-square', dict(action='store_true')), ('--hash-thr', dict(type=int, default=8))):
        p.add_argument(flag, help=argparse.SUPPRESS, **kw)
    return p


def _init_ranks():
    """`torchrun --nproc-per-node G -m videotofaces_b200 ...` runs one process per GPU: join the NCCL group before the
    engine is built so that every phase shards its rows over the ranks (rank 0 owns files, logs and console).  A plain
    `python -m videotofaces_b200` stays single-process like the reference."""
    import os
    if int(os.environ.get('WORLD_SIZE', '1')) <= 1:
        return False
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return True


def main(argv=None):
    args = build_parser().parse_args(argv)
    joined = _init_ranks()
    try:
        video_to_faces(**vars(args))
    finally:
        if joined:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
