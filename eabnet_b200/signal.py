"""STFT front end / iSTFT back end of the path as standalone calls (test.py:20-47, enhance.py:59-61)."""
from __future__ import annotations

import torch

from . import _lib


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def stft_compress(wave: torch.Tensor) -> torch.Tensor:
    """wave [B,M,L] (CUDA fp32) -> square-root-compressed spectrum [B,T,161,M,2], T = 1 + L//160.
    Mirrors the `noisy_stft` branch of `prepare_data(x, target, device, args)` with the reference's fixed
    signal constants (sr*win_size = 320, sr*win_shift = 160, fft_num = 320)."""
    if not wave.is_cuda or wave.dtype != torch.float32 or wave.ndim != 3:
        raise TypeError("stft_compress: expected a CUDA float32 tensor [B,M,L] (no CPU fallback)")
    B, M, L = wave.shape
    x = wave.contiguous()
    out = torch.empty((B, 1 + L // 160, 161, M, 2), dtype=torch.float32, device=wave.device)
    with torch.cuda.device(wave.device):
        _lib.check(_lib.load().eab_stft(x.data_ptr(), out.data_ptr(), B, M, L, _stream(wave.device)), "eab_stft")
    return out


def istft(spec: torch.Tensor) -> torch.Tensor:
    """spec [B,2,T,161] (CUDA fp32) -> wave [B,160*(T-1)]: the torch.istft call of enhance.py:59-61."""
    if not spec.is_cuda or spec.dtype != torch.float32 or spec.ndim != 4 or spec.shape[1] != 2 or spec.shape[3] != 161:
        raise TypeError("istft: expected a CUDA float32 tensor [B,2,T,161] (no CPU fallback)")
    B, _, T, _ = spec.shape
    x = spec.contiguous()
    out = torch.empty((B, 160 * (T - 1)), dtype=torch.float32, device=spec.device)
    with torch.cuda.device(spec.device):
        _lib.check(_lib.load().eab_istft(x.data_ptr(), out.data_ptr(), B, T, _stream(spec.device)), "eab_istft")
    return out
