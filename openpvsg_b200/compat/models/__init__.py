"""`import models` of the reference's tools (tools/test.py:27): registers the detectors / heads / positional
encodings of reference models/__init__.py:1-12 -- here the B200 implementations, under the same names."""
import openpvsg_b200  # noqa: F401  (registration side effect)
from openpvsg_b200.mask2former import (Mask2FormerCustom, Mask2FormerHeadCustom, Mask2FormerVideoCustom,  # noqa: F401
                                       Mask2FormerVideoCustomMinVIS, Mask2FormerVideoHead, MaskFormerFusionHeadCustom,
                                       SinePositionalEncoding3D)
