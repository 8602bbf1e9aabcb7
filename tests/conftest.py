import os
import os.path as osp
import sys

import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (ROOT, osp.join(ROOT, 'video-to-faces_b200'), osp.dirname(osp.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = osp.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def engine():
    """The real engine on cuda:0 (gpu tests only)."""
    from videotofaces_b200.runtime import get_engine
    return get_engine()
