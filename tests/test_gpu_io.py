"""GPU (-m gpu): the file edges of enhance.py:35-63 - sample-rate conversion against torchaudio's own outputs (golden vectors)
and the numpy oracle, and one file through the whole script sequence (wav -> 16 kHz -> microphone permutation -> network ->
iSTFT -> wav) against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import eabnet_oracle as O
from oracle import resample_oracle as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(f for f in os.listdir(GOLDEN) if f.startswith("resample_")))
def test_resample_matches_torchaudio_golden(name):
    from eabnet_b200 import resample
    z = np.load(os.path.join(GOLDEN, name))
    orig, new = int(z["orig"]), int(z["new"])
    y = resample(torch.from_numpy(z["x"]).cuda(), orig, new).cpu().numpy()
    assert y.shape == z["y"].shape
    assert np.abs(y - z["y"]).max() <= 3e-5                       # torchaudio's own fp32 accumulation over ~460 taps
    assert np.abs(y - R.resample(z["x"], orig, new)).max() <= 5e-6  # the fp64-accumulating oracle


@pytest.mark.parametrize("orig,new,shape", [(44100, 16000, (2, 9, 44100)), (48000, 16000, (1, 30001)), (8000, 16000, (5,)),
                                            (16000, 16000, (3, 100)), (96000, 16000, (1, 1)), (11025, 16000, (2, 0))])
def test_resample_shapes_and_edges(orig, new, shape):
    from eabnet_b200 import resample
    g = torch.Generator().manual_seed(orig // 100 + len(shape))
    x = 0.2 * torch.randn(*shape, generator=g)
    y = resample(x.cuda(), orig, new).cpu()
    ref = R.resample(x.numpy(), orig, new)
    assert tuple(y.shape) == ref.shape
    if y.numel():
        assert np.abs(y.numpy() - ref).max() <= 5e-6
    if orig == new:
        assert torch.equal(y, x)


def test_resample_rejects_cpu_tensors():
    from eabnet_b200 import resample
    with pytest.raises(TypeError):
        resample(torch.zeros(4), 44100, 16000)


def test_enhance_file_is_the_enhance_py_sequence(tmp_path):
    """44.1 kHz 16-bit 8-channel file -> enhance.py's steps; compared with torchaudio-convention decode + resample oracle +
    index_select + the network oracle + iSTFT"""
    from eabnet_b200 import EaBNet, enhance_file, wav_read, wav_write
    cfg = O.make_cfg(M=8)
    sd = O.make_weights(cfg, 3, "B")
    net = EaBNet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    g = torch.Generator().manual_seed(9)
    pcm = (0.1 * torch.randn(22050, 8, generator=g) * 32768).round().clamp(-32768, 32767).to(torch.int16)     # 0.5 s
    src, dst = str(tmp_path / "in.wav"), str(tmp_path / "out.wav")
    wav_write(src, 44100, pcm)
    indices = [7, 0, 1, 2, 3, 4, 5, 6]                           # enhance.py:41
    y = enhance_file(net, src, dst, mic_order=indices)
    # oracle chain
    noisy = pcm.T.float() / 32768.0                              # torchaudio.load
    noisy = torch.from_numpy(R.resample(noisy.numpy(), 44100, 16000))
    noisy = noisy.index_select(0, torch.tensor(indices)).unsqueeze(0)
    ref = O.istft(O.forward(sd, O.stft_compress(noisy), cfg))
    assert tuple(y.shape) == tuple(ref.shape) == (1, 160 * (noisy.shape[-1] // 160))
    assert (y.cpu() - ref).abs().max() <= 2e-4 * max(1.0, float(ref.abs().max()))
    back, sr = wav_read(dst)                                     # the written file holds exactly the returned wave
    assert sr == 16000 and torch.equal(back, y.cpu())
    # and it is the file scipy would have written
    import io
    from scipy.io import wavfile
    b = io.BytesIO()
    wavfile.write(b, 16000, y[0].cpu().numpy())
    assert open(dst, "rb").read() == b.getvalue()
