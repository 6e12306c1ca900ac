"""clskd_b200: B200-native (sm_100a) implementation of the DCCRN teacher->student distillation
step of KhanhNguyen4999/Speech-Enhancement-CLSKD behind the reference's Python module / loss API.

    from clskd_b200 import DCCRN, framework, tools_for_loss, feature_extraction, set_precision

All compute goes through libclskd_sm100.so (include/clskd.h); there is no CPU fallback.
"""
from . import _lib
from .ops import get_precision, set_precision
from . import config, tools_for_loss, tools_for_model, feature_extraction, framework, distill, metrics, lightning
from .DCCRN import DCCRN

__all__ = ["DCCRN", "config", "tools_for_loss", "tools_for_model", "feature_extraction", "framework",
           "distill", "metrics", "lightning", "set_precision", "get_precision"]
