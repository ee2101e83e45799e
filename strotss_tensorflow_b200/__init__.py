"""B200-native implementation of the STROTSS per-iteration loss hot path
(interaction-lab-uh/STROTSS-tensorflow: nn/losses.py, run_strotss.py:21-40,140).

Hand-written sm_100a CUDA (tcgen05 + TMA) behind a C ABI (include/strotss_b200.h); this package is
the host-side mirror of the reference's Python interface.  Importing it needs only torch; calling
anything needs the built library and a B200 -- there is no fallback.
"""
from .losses import dist_metrics, moment_matching, relaxed_emd, reshape_2d, self_similarity
from .modules import ContentLoss, MaskedStrotssLoss, StrotssLoss, StyleLoss
from .runtime import Handle, shared_handle
from .sampling import Sampling
from .strotss_utils import (RMSprop, convert_rgb_to_yuv, fold_laplacian_pyramid, make_laplacian, make_laplacian_pyramid,
                            resize, resize_like)

__all__ = ["relaxed_emd", "moment_matching", "self_similarity", "dist_metrics", "reshape_2d", "convert_rgb_to_yuv",
           "ContentLoss", "StyleLoss", "StrotssLoss", "MaskedStrotssLoss", "Handle", "shared_handle", "Sampling", "fold_laplacian_pyramid", "make_laplacian",
           "make_laplacian_pyramid", "resize", "resize_like", "RMSprop"]
