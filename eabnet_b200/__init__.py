"""eabnet_b200 - B200-native (sm_100a) implementation of EaBNet's inference hot path.

    from eabnet_b200 import EaBNet              # same constructor / state_dict / forward as the reference
    from eabnet_b200 import stft_compress, istft
    from eabnet_b200 import wav_read, resample, wav_write, enhance_file            # enhance.py's file edges
    from eabnet_b200 import GaGNet, EaBNetWithPostNet, make_eabnet_with_postnet     # the post-filter enhance.py runs
"""
from .model import EaBNet, numParams  # noqa: F401
from .postnet import EaBNetWithPostNet, GaGNet, make_eabnet_with_postnet, make_gag_net  # noqa: F401
from .signal import enhance_file, istft, resample, stft_compress, wav_bytes, wav_read, wav_write  # noqa: F401

__all__ = ["EaBNet", "GaGNet", "EaBNetWithPostNet", "make_gag_net", "make_eabnet_with_postnet", "numParams",
           "stft_compress", "istft", "wav_read", "wav_write", "wav_bytes", "resample", "enhance_file"]
