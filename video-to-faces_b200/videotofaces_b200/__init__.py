"""videotofaces_b200 -- B200-native grouping stage of video-to-faces.

Drop-in for the post-encoder path of the reference package `videotofaces`
(/root/reference/src/videotofaces): same entry points, arguments, logs and CLI flags,

    from videotofaces_b200 import video_to_faces
    from videotofaces_b200.dupes import remove_dupes_overall          # dupes.py:51
    from videotofaces_b200.grouping import classify, classify_faces, cluster_faces   # grouping.py:50,69,92

with the scikit-learn arithmetic replaced by hand-written sm_100a CUDA kernels behind the C ABI
of include/vtf_b200.h.  Video decoding, detection and the encoder networks are out of scope.
"""
from .main import video_to_faces  # noqa: F401
from .runtime import get_engine, reset_engine  # noqa: F401

__all__ = ['video_to_faces', 'get_engine', 'reset_engine']
