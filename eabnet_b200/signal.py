"""STFT front end / iSTFT back end of the path as standalone calls (test.py:20-47, enhance.py:59-61), and the I/O edges of
enhance.py: wav container (torchaudio.load / scipy wavfile.write) and sample-rate conversion (torchaudio Resample)."""
from __future__ import annotations

import ctypes as C

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


def wav_read(path_or_bytes, pcm16: bool = False):
    """`torchaudio.load(path)` (enhance.py:35) for RIFF/WAVE files: -> (float32 CPU tensor [channels, frames], sample_rate).
    pcm16=True returns the raw int16 samples of a 16-bit PCM file instead (input of the int16 front doors)."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    buf = (C.c_char * len(data)).from_buffer_copy(data)
    lib = _lib.load()
    ch, sr, bits, isf = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    frames = C.c_int64()
    _lib.check(lib.eab_wav_info(buf, len(data), C.byref(ch), C.byref(sr), C.byref(frames), C.byref(bits), C.byref(isf)), "eab_wav_info")
    out = torch.empty((ch.value, frames.value), dtype=torch.int16 if pcm16 else torch.float32)
    if out.numel() == 0:
        if pcm16 and (bits.value != 16 or isf.value):
            raise RuntimeError("eabnet_b200 (eab_wav_decode): the int16 output needs a 16-bit PCM file")
        return out, sr.value
    _lib.check(lib.eab_wav_decode(buf, len(data), None if pcm16 else out.data_ptr(), out.data_ptr() if pcm16 else None), "eab_wav_decode")
    return out, sr.value


def wav_bytes(sample_rate: int, data: torch.Tensor) -> bytes:
    """The file `scipy.io.wavfile.write(path, sample_rate, data.numpy())` writes (enhance.py:63), as bytes: data [frames] or
    [frames, channels], float32 (IEEE-float WAV) or int16 (PCM WAV)."""
    if data.is_cuda or data.dtype not in (torch.float32, torch.int16) or data.ndim not in (1, 2):
        raise TypeError("wav_bytes: expected a float32 or int16 CPU tensor [frames] or [frames, channels]")
    x = data.contiguous()
    frames, ch = x.shape[0], (1 if x.ndim == 1 else x.shape[1])
    pcm = x.dtype == torch.int16
    lib = _lib.load()
    n = lib.eab_wav_encode_bytes(frames, ch, int(pcm))
    buf = (C.c_char * n)()
    ptr = x.data_ptr() if x.numel() else torch.zeros(1, dtype=x.dtype).data_ptr()       # (an empty tensor has a null pointer)
    _lib.check(lib.eab_wav_encode(None if pcm else ptr, ptr if pcm else None, frames, ch, int(sample_rate), buf, n),
               "eab_wav_encode")
    return bytes(buf)


def wav_write(path, sample_rate: int, data: torch.Tensor) -> None:
    with open(path, "wb") as f:
        f.write(wav_bytes(sample_rate, data))


def resample(wave: torch.Tensor, orig_freq: int, new_freq: int) -> torch.Tensor:
    """`torchaudio.transforms.Resample(orig_freq, new_freq)(wave)` with its default kernel (enhance.py:36-37) for a CUDA
    float32 tensor [..., length] -> [..., ceil(length * new / orig)]."""
    if not wave.is_cuda or wave.dtype != torch.float32 or wave.ndim < 1:
        raise TypeError("resample: expected a CUDA float32 tensor [..., length] (no CPU fallback)")
    L = wave.shape[-1]
    rows = 1
    for d in wave.shape[:-1]:
        rows *= d
    x = wave.contiguous().view(rows, L)
    lib = _lib.load()
    Lo = lib.eab_resample_length(L, int(orig_freq), int(new_freq))
    if Lo < 0:
        raise ValueError("resample: bad length / rates")
    out = torch.empty((x.shape[0], Lo), dtype=torch.float32, device=wave.device)
    with torch.cuda.device(wave.device):
        _lib.check(lib.eab_resample(x.data_ptr(), out.data_ptr(), x.shape[0], L, int(orig_freq), int(new_freq), _stream(wave.device)),
                   "eab_resample")
    return out.view(*wave.shape[:-1], Lo)


def enhance_file(model, in_path, out_path, mic_order=None, device: torch.device | str = "cuda") -> torch.Tensor:
    """enhance.py:35-63 for one file: read the wav, convert to 16 kHz if needed, permute the microphones
    (`noisy.index_select(0, indices)`), run `model.enhance` (an EaBNet, or an EaBNetWithPostNet as enhance.py uses) and
    write the float32 result with scipy's wav layout.  Returns the enhanced wave [1, 160*(L//160)] (CUDA)."""
    noisy, sr = wav_read(in_path)
    dev = torch.device(device)
    x = noisy.to(dev)
    if sr != 16000:
        x = resample(x, sr, 16000)
    if mic_order is not None:
        x = x.index_select(0, torch.as_tensor(list(mic_order), device=dev))
    with torch.no_grad():
        y = model.enhance(x.unsqueeze(0).contiguous())
    wav_write(out_path, 16000, y[0].cpu())
    return y
