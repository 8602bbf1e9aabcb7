"""Process-wide engine: one GPU per process (LOCAL_RANK picks it), NCCL group if initialised."""
import os

_engine = None


def get_engine():
    """The lazily built GroupingEngine on this process's B200.  Raises without a GPU or without the
    compiled extension: the product has no CPU path."""
    global _engine
    if _engine is None:
        import torch
        from .kernels import CudaKernels
        from .engine import GroupingEngine
        from .dist import Comm
        if not torch.cuda.is_available():
            raise RuntimeError('videotofaces_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        dev = int(os.environ.get('LOCAL_RANK', '0')) % max(torch.cuda.device_count(), 1)
        torch.cuda.set_device(dev)
        _engine = GroupingEngine(CudaKernels('cuda:%d' % dev), Comm())
    return _engine


def reset_engine(engine=None):
    """Install (tests: a stand-in) or drop the process-wide engine."""
    global _engine
    _engine = engine


def is_main():
    """True on the rank that owns the side effects of the drop-in entry points (files, logs, console).  With one
    process (the reference's only mode) that is always the caller; under torchrun every rank computes its shard and
    rank 0 alone touches the file system ('several GPUs')."""
    return get_engine().comm.rank == 0


def sync():
    """Barrier over the ranks of the engine's group (no-op for one process): the entry points return only after
    rank 0 has finished moving files, so a caller on any rank sees the final tree."""
    get_engine().comm.barrier()
