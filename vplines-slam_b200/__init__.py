"""vplines-slam_b200: B200-native line front end (LSD -> LBD -> Hamming kNN).

The product is libvplines_b200.so (CUDA, sm_100a) behind the C ABI in
include/vpl_capi.h; this package is the Python host-side mirror of the
OpenCV-3.4 line_descriptor surface the reference's line tracker sits on
(LSDDetector / BinaryDescriptor / BinaryDescriptorMatcher) plus the batch driver.
Nothing here computes on the CPU.
"""
from . import capi  # noqa: F401
from .capi import Context, VplError  # noqa: F401
from .line_descriptor import (BinaryDescriptor, BinaryDescriptorMatcher, DMatch, KeyLine,  # noqa: F401
                              LSDDetector)
from .driver import FrontEnd, LineFrontEnd, LinePipeline, VanishingPoints, shard_range  # noqa: F401
