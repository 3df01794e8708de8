"""cniic_b200 -- B200-native (sm_100a CUDA) implementation of cniic's K-means / voronoi / pre-Huffman hot path.

Host layer above the C ABI (include/cniic_b200.h).  The product never imports ``oracle`` and has no CPU fallback.
"""
from . import _lib
from ._lib import (ERR_BAD_ARG, ERR_BUFFER_TOO_SMALL, ERR_CUDA, ERR_DECODE, ERR_NCCL, ERR_TOO_FEW_ACTIVE,
                   ERR_TOO_FEW_POINTS, ERR_UNSUPPORTED, MAX_DIM, MAX_K, OK, POINTS_RGB, POINTS_XYRGB, TIE_KEEP_CURRENT,
                   TIE_LOWEST_INDEX, build)
from .api import (CniicError, Context, KMeansResult, KMeansSession, kmeans_cluster, kmeans_reset_batch, kmeans_run_batch, synth_image_device,
                  synth_image_host)

__all__ = ["Context", "KMeansSession", "KMeansResult", "CniicError", "synth_image_host", "synth_image_device", "kmeans_cluster", "kmeans_reset_batch", "kmeans_run_batch", "build",
           "OK", "TIE_KEEP_CURRENT", "TIE_LOWEST_INDEX", "POINTS_RGB", "POINTS_XYRGB"]
