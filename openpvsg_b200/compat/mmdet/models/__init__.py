from openpvsg_b200.registry import (BACKBONES, DETECTORS, HEADS, build_backbone, build_detector,  # noqa: F401
                                    build_head)
