"""B200-native ORB front-end and bundle-adjustment engine behind the class signatures of
b51/ceres_mono_orb_slam2 (ORBextractor, ORBmatcher, CeresOptimizer).  See DESIGN.md.

All compute lives in libcmos_b200.so (hand-written sm_100a CUDA behind the C ABI in include/cmos_b200.h).
Nothing in this package falls back to the CPU.
"""
from ._lib import KP_DTYPE, CmosError, LIB_PATH  # noqa: F401
from .orb_extractor import ORBextractor  # noqa: F401
from .orb_matcher import ORBmatcher, Camera  # noqa: F401
from .ceres_optimizer import CeresOptimizer  # noqa: F401
from .tracking import TrackingFrontEnd  # noqa: F401
from .kf_matcher import KeyFrameMatcher, FeatureVector  # noqa: F401
from .map_points import MapPointOps  # noqa: F401
from .vocabulary import ORBVocabulary  # noqa: F401

__all__ = ["ORBextractor", "ORBmatcher", "Camera", "CeresOptimizer", "TrackingFrontEnd", "KeyFrameMatcher", "FeatureVector", "MapPointOps", "ORBVocabulary", "KP_DTYPE", "CmosError", "LIB_PATH"]
