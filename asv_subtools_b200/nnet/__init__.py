"""Host-side mirror of the reference's `libs.nnet` surface for the extraction path: parameter
containers with the reference's names/shapes (so reference checkpoints load unchanged) and the
`TopVirtualNnet` plugin base.  All arithmetic is delegated to the native library."""
from .components import FTdnnBlock, ReluBatchNormTdnnLayer, TdnnAffine  # noqa: F401
from .pooling import (AttentionAlphaComponent, AttentiveStatisticsPooling, GlobalMultiHeadAttentionPooling,  # noqa: F401
                      LDEPooling, MultiHeadAttentionPooling, MultiResolutionMultiHeadAttentionPooling, StatisticsPooling,
                      xivec_stdinit_softplus2_prec_pooling)
from .framework import (AttentionPoolingExtractor, TopVirtualNnet, build_tdnn_extractor,  # noqa: F401
                        for_extract_embedding)
