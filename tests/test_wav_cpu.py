"""CPU: the wav container of enhance.py's file edges (host C++ behind the C ABI, no GPU work) against scipy.io.wavfile, and the
resample oracle against torchaudio's own outputs (tests/golden/resample_*.npz, tools/make_golden_resample.py)."""
import io
import os
import struct

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from conftest import GOLDEN
from oracle import resample_oracle as R


@pytest.fixture(scope="module", autouse=True)
def _built():
    from eabnet_b200 import build
    build.build()


def _scipy_bytes(sr, x):
    b = io.BytesIO()
    wavfile.write(b, sr, x)
    return b.getvalue()


@pytest.mark.parametrize("shape", [(1000,), (1001,), (333, 3), (0,), (1, 9)])
def test_encode_is_scipy_byte_for_byte(shape):
    """enhance.py:63 `wavfile.write(path, 16000, esti_wav[0])`: float32 -> IEEE-float WAV with fact chunk; int16 -> PCM WAV"""
    from eabnet_b200 import wav_bytes, wav_read
    rng = np.random.default_rng(len(shape) * 100 + shape[0])
    x = (0.1 * rng.standard_normal(shape)).astype(np.float32)
    assert wav_bytes(16000, torch.from_numpy(x)) == _scipy_bytes(16000, x)
    xi = (x * 32767).astype(np.int16)
    assert wav_bytes(44100, torch.from_numpy(xi)) == _scipy_bytes(44100, xi)
    # and back: planar [channels, frames] like torchaudio.load, int16 / 32768
    y, sr = wav_read(_scipy_bytes(44100, xi))
    ch = 1 if len(shape) == 1 else shape[1]
    planar = xi.reshape(shape[0], ch).T
    assert sr == 44100 and torch.equal(y, torch.from_numpy(planar.astype(np.float32) / 32768.0))
    y16, _ = wav_read(_scipy_bytes(44100, xi), pcm16=True)
    assert torch.equal(y16, torch.from_numpy(planar.copy()))
    yf, sr = wav_read(_scipy_bytes(16000, x))
    assert sr == 16000 and torch.equal(yf, torch.from_numpy(x.reshape(shape[0], ch).T.copy()))


def test_decode_other_encodings_like_scipy():
    from eabnet_b200 import wav_read
    rng = np.random.default_rng(5)
    x = rng.standard_normal((257, 2))
    for arr, scale in ((np.clip(x * 60 + 128, 0, 255).astype(np.uint8), None), ((x * 2e8).astype(np.int32), 2.0 ** 31),
                       (x.astype(np.float64), 1.0)):
        data = _scipy_bytes(8000, arr)
        y, sr = wav_read(data)
        if arr.dtype == np.uint8:
            ref = (arr.astype(np.float32) - 128.0) / 128.0
        else:
            ref = (arr / scale).astype(np.float32)
        assert sr == 8000 and np.array_equal(y.numpy(), ref.T), arr.dtype
    # 24-bit PCM + a LIST chunk with an odd size before the data chunk + WAVE_FORMAT_EXTENSIBLE
    s24 = (x[:, 0] * 2e6).astype(np.int32)
    raw = b"".join(struct.pack("<i", int(v))[:3] for v in s24)
    fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 22050, 22050 * 3, 3, 24) + struct.pack("<HHI", 22, 24, 4) + struct.pack("<H", 1) + b"\x00" * 14
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + b"data" + struct.pack("<I", len(raw)) + raw
    y, sr = wav_read(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert sr == 22050 and np.array_equal(y.numpy()[0], (s24.astype(np.float64) / 2.0 ** 23).astype(np.float32))


def test_errors_are_loud():
    from eabnet_b200 import wav_bytes, wav_read
    with pytest.raises(RuntimeError, match="RIFF"):
        wav_read(b"not a wav file at all")
    good = _scipy_bytes(16000, np.zeros(10, np.int16))
    with pytest.raises(RuntimeError, match="unsupported encoding"):
        wav_read(good[:20] + struct.pack("<H", 2) + good[22:])          # ADPCM tag
    with pytest.raises(RuntimeError, match="16-bit PCM"):
        wav_read(_scipy_bytes(16000, np.zeros(10, np.float32)), pcm16=True)
    with pytest.raises(TypeError):
        wav_bytes(16000, torch.zeros(4, dtype=torch.float64))


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.startswith("resample_")))
def test_resample_oracle_matches_torchaudio_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    y = R.resample(z["x"], int(z["orig"]), int(z["new"]))
    assert y.shape == z["y"].shape
    assert np.abs(y - z["y"]).max() <= 3e-5            # torchaudio accumulates ~460 taps in fp32, the oracle in fp64


def test_resample_length_matches_torchaudio_rule():
    from eabnet_b200 import _lib
    lib = _lib.load()
    for L, o, n in ((96000, 44100, 16000), (1, 48000, 16000), (0, 8000, 16000), (12345, 22050, 16000), (7, 16000, 16000)):
        g = np.gcd(o, n)
        assert lib.eab_resample_length(L, o, n) == -(-(n // g) * L // (o // g))
