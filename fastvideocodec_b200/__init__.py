"""fastvideocodec_b200 — B200-native (sm_100a) DVC P-frame coding hot path.

Drop-in for the reference's ``DVC.net.VideoCompressor`` path (see DESIGN.md / INTEGRATION.md).
Importing the package does not load the CUDA library; the first compute call does and fails
loudly if it is not built or no GPU is present.
"""
from .net import VideoCompressor, load_model, save_model  # noqa: F401
from .gop import (AverageMeter, PSNR, get_codec_model, get_DVC_pretrained, parallel_compression,  # noqa: F401
                  reduce_stats, shard_gops, stats_vector, summarize)

from .lsvc import LSVC, generate_graph, graph_from_batch, refidx_from_graph  # noqa: F401

__version__ = "0.1.0"
