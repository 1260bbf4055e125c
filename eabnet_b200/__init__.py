"""eabnet_b200 - B200-native (sm_100a) implementation of EaBNet's inference hot path.

    from eabnet_b200 import EaBNet              # same constructor / state_dict / forward as the reference
    from eabnet_b200 import stft_compress, istft
"""
from .model import EaBNet, numParams  # noqa: F401
from .signal import istft, stft_compress  # noqa: F401

__all__ = ["EaBNet", "numParams", "stft_compress", "istft"]
