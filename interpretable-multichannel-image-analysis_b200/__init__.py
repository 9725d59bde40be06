"""B200-native per-object, per-channel hand-crafted feature extraction.

Drop-in for the extraction step of
aliechoes/interpretable-multichannel-image-analysis' ``channel_importance_hand_crafted_features``
notebook (cell 13 functions + the cell 17 loop).  All arithmetic runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/imfeat.h``; there is no CPU fallback.
"""
from . import ablation, batcher, distributed, postprocess, schema, synth
from .batcher import PinnedBatcher
from .postprocess import MinMaxScaler
from ._lib import ImfeatError
from .extractor import (FeatureExtractor, basic_statistical_features, extract_features, from_unit_float,
                        get_extractor, glcm_features, pack_mask_bits, plane_stride_for)
from .schema import feature_columns

__all__ = [
    "FeatureExtractor", "ImfeatError", "basic_statistical_features", "extract_features",
    "feature_columns", "from_unit_float", "get_extractor", "glcm_features", "pack_mask_bits", "plane_stride_for", "schema",
    "ablation", "batcher", "distributed", "postprocess", "synth", "PinnedBatcher", "MinMaxScaler",
]
