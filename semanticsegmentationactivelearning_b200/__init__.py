"""B200-native pool acquisition scoring for semantic-segmentation active learning.

Drop-in for the pool-scoring path of alfrunesiq/SemanticSegmentationActiveLearning
(active_learning.py:229-275 and :682-715): logits -> softmax -> entropy / margin / max-prob /
MC-variance confidence -> per-image score -> k lowest-confidence images.  The arithmetic runs in
hand-written sm_100a CUDA kernels behind a C ABI (include/alscore.h); there is no CPU fallback.
"""
from ._lib import AlscoreUnavailable, LIB_PATH
from .acquisition import MEASURES, Scorer, default_scorer, measure_id, rank_confidence
from . import al_loop
from .distributed import comm_init_torch, rank_confidence_sharded, rank_confidence_sharded_device, shard_bounds

__all__ = ["al_loop", "AlscoreUnavailable", "LIB_PATH", "MEASURES", "Scorer", "default_scorer", "measure_id",
           "comm_init_torch", "rank_confidence", "rank_confidence_sharded", "rank_confidence_sharded_device", "shard_bounds"]
__version__ = "0.1.1"
