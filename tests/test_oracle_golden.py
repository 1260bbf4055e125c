"""CPU: the oracle (oracle/eabnet_oracle.py) against the reference's own outputs.

The reference has no tests or golden vectors; the pin is tests/golden/*.npz, written by
tools/make_golden.py from the unmodified /root/reference module.  Tolerances are fp32 round-off only: the
oracle uses the same torch primitives in a different composition (manual LSTM, functional norms)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import eabnet_oracle as O


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    cfg = O.make_cfg(**g["cfg"])
    sd = O.make_weights(cfg, seed=0, variant=g["variant"])
    wave, _ = O.make_wave(g["B"], cfg["M"], g["L"], seed=1234)
    spec = O.stft_compress(wave)
    assert spec.shape == g["spec"].shape
    assert np.abs(spec.numpy() - g["spec"]).max() <= 2e-6
    out = O.forward(sd, torch.from_numpy(g["spec"]), cfg)
    assert out.shape == g["out"].shape
    scale = max(1.0, float(np.abs(g["out"]).max()))
    assert np.abs(out.numpy() - g["out"]).max() <= 2e-5 * scale
    if "wav" in g:
        wav = O.istft(torch.from_numpy(g["out"]))
        assert wav.shape == g["wav"].shape == (g["B"], 160 * (g["L"] // 160))
        assert np.abs(wav.numpy() - g["wav"]).max() <= 1e-6 * scale


def test_param_table_sanity():
    shapes = O.param_shapes()
    assert len(shapes) == 498
    assert sum(int(np.prod(s)) for s in shapes.values()) == 2838610
    assert shapes["en.meta_unet_list.0.in_conv.0.conv.1.weight"] == (128, 18, 2, 5)
    assert shapes["de.last_conv.0.conv.0.weight"] == (128, 128, 2, 5)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present on this box")
def test_oracle_matches_live_reference():
    sys.path.insert(0, "/root/reference")
    from EaBNet import EaBNet
    cfg = O.make_cfg()
    sd = O.make_weights(cfg, seed=3, variant="B")
    net = EaBNet(**cfg).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(O.param_shapes(cfg).items())
    net.load_state_dict(sd, strict=True)
    wave, _ = O.make_wave(1, 9, 4800, seed=7)
    spec = O.stft_compress(wave)
    with torch.no_grad():
        ref = net(spec)
    assert (O.forward(sd, spec, cfg) - ref).abs().max() <= 2e-5


def test_si_sdr_definition():
    rng = np.random.RandomState(0)
    s = rng.randn(1000)
    assert O.si_sdr(s, 3.0 * s + 1e-3 * rng.randn(1000)) > 50
    assert abs(O.si_sdr(s, s + rng.randn(1000))) < 1.5


def test_lstm_fused_equals_stepwise_restatement():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 33, 64, generator=g)
    w = [torch.randn(256, 64, generator=g) * 0.125, torch.randn(256, 64, generator=g) * 0.125,
         torch.randn(256, generator=g) * 0.1, torch.randn(256, generator=g) * 0.1]
    assert (O.lstm_layer(x, *w) - O.lstm_layer_stepwise(x, *w)).abs().max() <= 2e-6
    xd = x.double()
    wd = [t.double() for t in w]
    assert (O.lstm_layer(xd, *wd) - O.lstm_layer_stepwise(xd, *wd)).abs().max() <= 1e-12
