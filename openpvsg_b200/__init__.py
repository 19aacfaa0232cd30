"""openpvsg_b200 -- B200 (sm_100a) backend of the OpenPVSG inference hot path.

Importing the package registers the reference's module names
(``Mask2FormerVideoCustom``, ``Mask2FormerVideoHead``, ``MaskFormerFusionHeadCustom``,
``SinePositionalEncoding3D``, ...; reference models/__init__.py:1-12) in the mmcv-style
registries of ``openpvsg_b200.registry``.  The arithmetic lives in libpvsg_sm100.so
(include/pvsg.h); there is no PyTorch / CPU fallback.
"""
from .registry import (build_backbone, build_detector, build_head, load_config,  # noqa: F401
                       DETECTORS, HEADS, BACKBONES, POSITIONAL_ENCODING)
from . import mask2former  # noqa: F401  (registers the modules)
from . import swin  # noqa: F401  (registers SwinTransformer)
from .swin import SwinTransformer  # noqa: F401
from .mask2former import (Mask2FormerCustom, Mask2FormerHeadCustom, Mask2FormerVideoCustom,  # noqa: F401
                          Mask2FormerVideoCustomMinVIS, Mask2FormerVideoHead, MaskFormerFusionHeadCustom,
                          SinePositionalEncoding3D)
from .relation_head import (ObjectEncoder, PairProposalNetwork, TemporalTransformer, VanillaModel,  # noqa: F401
                            HandcraftedFilter, Learnable1DConv, pick_top_pairs_eval, concatenate_sub_obj,
                            generate_results, generate_pairwise_results)

__version__ = '0.1.0'
