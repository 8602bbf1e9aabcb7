"""The C-ABI shared library loads without a GPU and exports every symbol include/vtf_b200.h declares;
compute entry points refuse to run without a device (no CPU fallback).  CPU only."""
import ctypes
import os.path as osp
import re

import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


def _header_functions():
    txt = open(osp.join(ROOT, 'include', 'vtf_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(vtf_[a-z0-9_]+)\s*\(', txt)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from videotofaces_b200 import _native
    return _native.load()


def test_header_declares_functions():
    fns = _header_functions()
    assert len(fns) >= 25 and 'vtf_sil_sweep' in fns and 'vtf_cosine_dedup' in fns


def test_every_declared_symbol_is_exported(lib):
    for name in _header_functions():
        assert hasattr(lib, name), 'libvtf_b200.so does not export %s' % name


def test_python_binding_matches_header(lib):
    from videotofaces_b200 import _native
    assert sorted(_native.SIGNATURES) == _header_functions()


def test_version_and_launch_counter(lib):
    assert lib.vtf_version() >= 100
    assert lib.vtf_launch_count() >= 0


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from videotofaces_b200 import _native
    from videotofaces_b200.runtime import get_engine, reset_engine
    reset_engine(None)
    with pytest.raises(RuntimeError):
        get_engine()
    with pytest.raises(_native.NativeError):
        _native.require_gpu()
    assert lib.vtf_device_check() != 0          # reports an error instead of computing on the host


def test_sass_contains_blackwell_tensor_path():
    """cuobjdump evidence that the sweep is a tcgen05/TMA kernel (B200_PROFILING.md mnemonics)."""
    import shutil
    import subprocess
    from videotofaces_b200 import _native
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    sass = subprocess.run(['cuobjdump', '-sass', _native.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):
        assert mnemonic in sass, mnemonic


def test_dedup_screen_launch_plan_is_host_arithmetic(monkeypatch):
    """vtf_dedup_screen_launches(N, d): column chunks (= launches = claim counters) of the dedup screen.  One launch while the
    fp16 operand fits in L2 with room to spare (64 MB), else chunks of at most 48 MB; VTF_SCREEN_CHUNK overrides (tests)."""
    from videotofaces_b200 import _native as nv
    lib = nv.load()
    monkeypatch.delenv('VTF_SCREEN_CHUNK', raising=False)
    assert lib.vtf_dedup_screen_launches(0, 512) == 0 and lib.vtf_dedup_screen_launches(1, 512) == 0
    assert lib.vtf_dedup_screen_launches(2, 512) == 1
    assert lib.vtf_dedup_screen_launches(50_000, 512) == 1                 # config 2: 51 MB of operand
    units = (200_000 + 255) // 256                                         # 205 MB: chunks of 48 MB / (256 * 512 * 2) = 192 tiles
    assert lib.vtf_dedup_screen_launches(200_000, 512) == (units + 191) // 192
    units = (4_000_000 + 255) // 256                                       # config 5: 128 tiles per chunk at d = 768
    assert lib.vtf_dedup_screen_launches(4_000_000, 768) == (units + 127) // 128
    monkeypatch.setenv('VTF_SCREEN_CHUNK', '2')
    assert lib.vtf_dedup_screen_launches(3000, 512) == 6
