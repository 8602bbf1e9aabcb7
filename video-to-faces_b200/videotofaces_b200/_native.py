"""ctypes binding of libvtf_b200.so (include/vtf_b200.h).

The library is the only compute back end: if it is missing, or no sm_100 GPU is visible, every
entry point raises -- there is no CPU fallback (BASELINE.json north_star).
"""
import ctypes
import os
import os.path as osp

c_i64, c_int, c_f32, c_vp, c_sz = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

LIB_PATH = os.environ.get('VTF_LIB') or osp.join(osp.dirname(osp.abspath(__file__)), '_lib', 'libvtf_b200.so')   # VTF_LIB: A/B builds

# name -> (restype, argtypes); mirrors include/vtf_b200.h one to one (tests/test_cabi.py checks it)
SIGNATURES = {
    'vtf_version': (c_int, []),
    'vtf_last_error': (ctypes.c_char_p, []),
    'vtf_launch_count': (c_i64, []),
    'vtf_device_check': (c_int, []),
    'vtf_normalize_rows': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'vtf_col_stats_ws_bytes': (c_sz, [c_int]),
    'vtf_col_stats': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_center_rows': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    'vtf_normalize_col_stats_ws_bytes': (c_sz, [c_int]),
    'vtf_normalize_col_stats': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_center_rows_stats': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_build_cells_ws_bytes': (c_sz, [c_i64]),
    'vtf_build_cells': (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    'vtf_kpp_ws_bytes': (c_sz, [c_int]),
    'vtf_kpp_dist': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_kpp_rownorm64': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'vtf_kpp_pick': (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_kpp_pick_ws_bytes': (c_sz, [c_i64]),
    'vtf_kpp_pick_fast': (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_contingency': (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    'vtf_kpp_round_prep': (c_int, [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_kpp_unsort': (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    'vtf_kpp_pick_serial': (c_int, [c_vp, c_i64, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_lloyd_assign': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_cluster_sums_ws_bytes': (c_sz, [c_int, c_int]),
    'vtf_cluster_sums': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'vtf_cluster_dispersion': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_lloyd_update': (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_lloyd_assign_multi': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, ctypes.c_uint32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    'vtf_cluster_sums_multi': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    'vtf_cluster_dispersion_multi': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_lloyd_update_multi': (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, ctypes.c_uint32, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    'vtf_wide_pack_centres': (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vtf_wide_assign': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_f32, c_f32, c_vp, c_i64,
                                c_vp]),
    'vtf_wide_refine': (c_int, [c_vp, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    'vtf_wide_count': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_vp, ctypes.c_uint32, c_vp, c_vp, c_vp]),
    'vtf_wide_commit': (c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_vp, ctypes.c_uint32, c_vp, c_i64, c_vp, c_vp,
                                ctypes.c_uint32, c_vp]),
    'vtf_wide_sums_delta': (c_int, [c_vp, c_int, c_vp, c_vp, ctypes.c_uint32, c_int, c_vp, ctypes.c_double, c_vp, c_vp]),
    'vtf_wide_sums_ws_bytes': (c_sz, [c_int, c_int]),
    'vtf_wide_sums_full': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, ctypes.c_uint32, c_int, c_int, c_vp, c_i64,
                                   ctypes.c_double, c_vp, c_vp, c_vp]),
    'vtf_wide_update': (c_int, [c_vp, ctypes.c_double, c_vp, c_vp, c_int, c_vp, ctypes.c_uint32, c_int, c_vp, c_vp, c_vp, c_int,
                                c_vp]),
    'vtf_cell_sums': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, ctypes.c_uint32, ctypes.c_double, c_vp, c_vp, c_vp, c_vp]),
    'vtf_cell_dispersion': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, ctypes.c_double, ctypes.c_double, c_vp,
                                    c_vp, c_vp, c_vp]),
    'vtf_cosine_classify': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_f32, c_vp, c_vp, c_vp]),
    'vtf_dedup_init': (c_int, [c_vp, c_i64, c_vp]),
    'vtf_cosine_dedup': (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp]),
    'vtf_dedup_pack': (c_int, [c_vp, c_i64, c_int, c_i64, c_f32, c_vp, c_vp, c_vp]),
    'vtf_cosine_dedup_tc': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_f32, c_int, c_int, c_vp, c_vp]),
    'vtf_debug_sweep_timing': (c_int, [c_vp]),
    'vtf_upload_pageable': (c_int, [c_vp, c_vp, c_sz, c_int, c_vp]),
    'vtf_dedup_screen_launches': (c_int, [c_i64, c_int]),
    'vtf_dedup_screen': (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp]),
    'vtf_dedup_refine4': (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'vtf_cosine_dedup_tc_units': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_f32, c_vp, c_i64, c_vp, c_vp]),
    'vtf_dedup_refine': (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp]),
    'vtf_hash_pack': (c_int, [c_vp, c_i64, c_vp, c_vp]),
    'vtf_hash_dedup': (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    'vtf_gather_rows': (c_int, [c_vp, c_int, c_vp, c_i64, c_vp, c_vp]),
    'vtf_silhouette_rowsums_exact': (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_i64, c_i64, c_vp, c_vp]),
    'vtf_row_sqnorms': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'vtf_sil_pack': (c_int, [c_vp, c_i64, c_int, c_vp, c_i64, c_f32, c_int, c_f32, c_vp, c_vp, c_vp, c_vp]),
    'vtf_sil_diag': (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp]),
    'vtf_sil_aug': (c_int, [c_vp, c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    'vtf_absmax': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'vtf_sil_sweep': (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_int, c_i64, c_i64, c_vp, c_vp]),
    'vtf_sil_sweep_sym': (c_int, [c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vtf_shared_alloc': (c_int, [c_sz, c_vp, c_vp]),
    'vtf_shared_open': (c_int, [c_vp, c_vp]),
    'vtf_shared_zero': (c_int, [c_vp, c_sz, c_vp]),
    'vtf_shared_close': (c_int, [c_vp]),
    'vtf_shared_free': (c_int, [c_vp]),
    'vtf_sil_finish_rows': (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
    'vtf_file_ops': (c_i64, [c_int, ctypes.c_char_p, c_vp, ctypes.c_char_p, c_vp, c_i64, c_int, c_vp, c_vp]),
    'vtf_sil_finish_cells': (c_int, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp]),
}


class NativeError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (no GPU needed for loading / symbol checks)."""
    global _lib
    if _lib is None:
        if not osp.exists(LIB_PATH):
            raise NativeError('CUDA extension %s is missing: run `python video-to-faces_b200/build.py` '
                              '(there is no CPU fallback)' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().vtf_last_error().decode('utf-8', 'replace')


def call(name, *args):
    """Invoke an int-returning entry point and raise on a negative status."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise NativeError('%s failed (%d): %s' % (name, rc, last_error()))


def launch_count():
    return int(load().vtf_launch_count())


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        raise NativeError('videotofaces_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    call('vtf_device_check')


_empty_stub = {}


def ptr(t):
    """Device (or host) address of a torch tensor, None -> NULL.  A tensor without elements (a rank whose row shard is
    empty) has no storage and data_ptr() == 0; the entry points treat NULL as a missing argument, so such a tensor is
    passed as the address of a small per-device stub that no kernel dereferences (every entry point returns early or
    launches nothing for zero rows)."""
    if t is None:
        return None
    if t.numel() == 0:
        import torch
        stub = _empty_stub.get(t.device)
        if stub is None:
            stub = _empty_stub[t.device] = torch.zeros(64, dtype=torch.uint8, device=t.device)
        return c_vp(stub.data_ptr())
    return c_vp(t.data_ptr())


def stream():
    import torch
    return c_vp(torch.cuda.current_stream().cuda_stream)
