"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product): `torchaudio.transforms.Resample(orig, new)` with its
default arguments, which enhance.py:36-37 applies when the file is not at 16 kHz.  The algorithm lives in a third-party
dependency of the reference (torchaudio, not vendored; the image has 2.11): `torchaudio.functional.functional.
_get_sinc_resample_kernel` + `_apply_sinc_resample_kernel` (resampling_method "sinc_interp_hann", lowpass_filter_width 6,
rolloff 0.99).  Restated here in numpy; pinned by tests/golden/resample_*.npz = outputs of torchaudio itself
(tools/make_golden_resample.py)."""
import math

import numpy as np


def sinc_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t = t * base_freq
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * scale
    return k.astype(np.float32), width, orig, new          # [new, 2 width + orig]


def resample(wave: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """wave [..., length] float32 -> [..., ceil(length * new / orig)]"""
    if orig_freq == new_freq:
        return wave
    k, width, orig, new = sinc_kernel(orig_freq, new_freq)
    shape = wave.shape
    x = np.asarray(wave, dtype=np.float32).reshape(int(np.prod(shape[:-1])), shape[-1])
    L = x.shape[1]
    xp = np.pad(x, ((0, 0), (width, width + orig)))
    K = k.shape[1]
    nblk = (xp.shape[1] - K) // orig + 1
    # conv1d(stride = orig): windows [rows, nblk, K] against kernels [new, K]
    win = np.lib.stride_tricks.sliding_window_view(xp, K, axis=1)[:, ::orig][:, :nblk]
    out = np.einsum("rnk,jk->rnj", win.astype(np.float64), k.astype(np.float64)).astype(np.float32)
    out = out.reshape(x.shape[0], -1)
    target = int(math.ceil(new * L / orig))
    return out[:, :target].reshape(*shape[:-1], target)
