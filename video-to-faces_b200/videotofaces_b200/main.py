"""Orchestrator with the reference's signature (videotofaces.main.video_to_faces,
/root/reference/src/videotofaces/main.py:13-81), restricted to what this package owns: the grouping
stage.  Detection and the encoder networks are out of scope; embeddings arrive
either from the upstream package, if it is installed next to this one, or as arrays via the extra
keyword arguments `features=` / `ref_features=` (an ndarray or a path to an .npy file).
"""
import os.path as osp

import numpy as np

from .prep import validate_args, get_clusters, get_class_ref, get_paths_for_grouping
from .grouping import cluster_faces, classify_faces, test_grouping
from .dupes import remove_dupes_overall
from .runtime import is_main


def _load_features(f):
    if isinstance(f, str):
        f = np.load(f)
    if hasattr(f, 'is_cuda') and f.is_cuda:        # encoder output still in HBM: no trip through the host (section 8f item 2)
        import torch
        return f.detach().to(torch.float32).contiguous()
    return np.ascontiguousarray(f, dtype=np.float32)


def _upstream_encoder(style, enc_model, device, paths, bs, area):
    try:
        from videotofaces.grouping import get_encoder_model, encode_faces   # the unmodified reference package
    except Exception as e:  # noqa: BLE001
        print('ERROR: no embeddings given (features=...) and the upstream videotofaces encoders are not importable: %s' % e)
        return None, None
    model = get_encoder_model(style, enc_model, device)
    return model, encode_faces(paths, model, bs, area)


def video_to_faces(input_path=None, input_ext=None,
                   mode='full', style='anime', device=None,
                   out_dir=None, out_prefix='', resize_to=None,
                   save_frames=False, save_rejects=False, save_dupes=False,
                   video_step=1, video_fragment=None, video_area=None, video_reader='opencv',
                   det_model='default', det_batch_size=4, det_min_score=0.4, det_min_size=50,
                   det_min_border=5, det_scale=(1.5, 1.5, 2.2, 1.2), det_square=True,
                   hash_thr=8,
                   enc_model='default', enc_batch_size=16, enc_area=None,
                   group_mode='clustering', clusters=None, clusters_save_all=False,
                   ref_dir=None, random_state=0, group_log=True,
                   enc_dup_thr=0.25, enc_oth_thr=0.9,
                   _test_enc=False, _test_exclude_other=False,
                   features=None, ref_features=None, normalize_features=False):
    if not validate_args(mode, input_path, out_dir, style, group_mode, video_reader, det_model, enc_model):
        return
    if mode != 'grouping':
        print('ERROR: videotofaces_b200 implements mode="grouping" only; run detection with the upstream package '
              '(python -m videotofaces -m detection ...) and then group its output here')
        return
    if enc_model == 'default':
        enc_model = 'vit_b' if style == 'anime' else 'facenet_vgg'              # main.py:33-34
    if not out_dir:
        out_dir = input_path if osp.isdir(input_path) else osp.dirname(osp.abspath(input_path))

    refs = None
    if group_mode == 'clustering' or _test_enc:                                # main.py:41-44
        clusters = get_clusters(clusters)
        if not clusters:
            return
    if group_mode == 'classification' or _test_enc:                            # main.py:45-48
        refs = get_class_ref(ref_dir, out_dir)
        if not refs:
            return
    imgpaths = get_paths_for_grouping(out_dir)                                 # main.py:49-52
    if not imgpaths:
        return

    if _test_enc:                                                              # main.py:65-69
        test_params = (style, enc_model, device, out_dir, _test_exclude_other, enc_batch_size, enc_area, enc_oth_thr,
                       random_state)
        X = _load_features(features) if features is not None else None
        R = _load_features(ref_features) if ref_features is not None else None
        test_grouping(imgpaths, refs, test_params, X=X, R=R)
        return

    encoder = None
    if features is None:
        encoder, feats = _upstream_encoder(style, enc_model, device, imgpaths, enc_batch_size, enc_area)
        if feats is None:
            return
    else:
        feats = _load_features(features)
        if feats.shape[0] != len(imgpaths):
            print('ERROR: features has %u rows but %u images were found' % (feats.shape[0], len(imgpaths)))
            return

    if enc_dup_thr and enc_dup_thr != -1:                                      # main.py:72-74
        # the surviving rows stay on the device: the grouping call below is their only consumer
        feats, imgpaths = remove_dupes_overall(feats, imgpaths, ('enc', enc_dup_thr, save_dupes, out_dir), keep_on_device=True)
    if group_mode == 'clustering':                                             # main.py:75-77
        # normalize_features: `features` are raw encoder outputs (FaceNet without its closing F.normalize, facenet.py:153);
        # the normalisation is fused into the first pass of the clustering (dedup / classify normalise on their own)
        cluster_faces(imgpaths, feats, (clusters, clusters_save_all, random_state, group_log, out_dir), normalize=normalize_features)
    if group_mode == 'classification':                                         # main.py:78-80
        R = _load_features(ref_features) if ref_features is not None else None
        classify_faces(imgpaths, feats, encoder, (refs, enc_oth_thr, group_log, out_dir), R=R)
    if is_main():
        print('Done')                                                          # main.py:82
