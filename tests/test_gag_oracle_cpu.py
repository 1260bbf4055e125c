"""CPU: the GaGNet / EaBNetWithPostNet oracle (oracle/gagnet_oracle.py) against the reference's own outputs
(tests/golden/gag_*.npz, written by tools/make_golden_gag.py from the unmodified /root/reference modules), and the
product's GaGNet parameter table against the reference's state_dict contract."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import gag_golden_cases, load_golden
from oracle import eabnet_oracle as O
from oracle import gagnet_oracle as G


@pytest.mark.parametrize("name", gag_golden_cases())
def test_gag_oracle_matches_reference_golden(name):
    g = load_golden(name)
    cfg = G.make_gag_cfg(**g["cfg"])
    sd = G.make_gag_weights(cfg, 0, g["variant"])
    outs = torch.stack(G.gag_forward(sd, torch.from_numpy(g["inpt"]), torch.from_numpy(g["pre"]), cfg)).numpy()
    assert outs.shape == g["outs"].shape == (cfg["q"], g["B"], 2, 161, 1 + g["L"] // 160)
    assert np.abs(outs - g["outs"]).max() <= 2e-5 * max(1.0, float(np.abs(g["outs"]).max()))


def test_postnet_wrapper_oracle_matches_reference_golden():
    g = load_golden("gag_wrapper_default_b1_t21")
    sd = G.make_postnet_weights(None, None, 0, "B")
    r = G.postnet_forward(sd, torch.from_numpy(g["spec"]), O.make_cfg(), G.make_gag_cfg(), ref_mic=0)
    scale = max(1.0, float(np.abs(g["esti"]).max()))
    assert np.abs(r["esti0_stft"].numpy() - g["esti0"]).max() <= 2e-5 * scale
    assert np.abs(r["esti_stft"].numpy() - g["esti"]).max() <= 5e-5 * scale
    assert np.abs(torch.stack(r["esti1_stft_list"]).numpy() - g["stages"]).max() <= 5e-5 * scale


def test_gag_param_table_sanity():
    shapes = G.gag_param_shapes()
    assert len(shapes) == 815
    assert sum(int(np.prod(s)) for s in shapes.values()) == 5950697          # GaGNet.py default configuration
    assert shapes["en.meta_unet_list.0.in_conv.0.conv.1.weight"] == (128, 4, 2, 5)
    assert shapes["gags.0.glance_block.in_conv_main.weight"] == (256, 578, 1)


@pytest.mark.parametrize("over", [{}, {"is_u2": False, "norm_type": "BN", "is_squeezed": True},
                                  {"intra_connect": "add", "dilas": (1, 2, 4), "q": 2, "p": 1}])
def test_product_gag_state_dict_contract(over):
    """eab_gag_create's parameter table (names, order, shapes) == the oracle's independent table (== the reference's)."""
    from eabnet_b200 import GaGNet
    cfg = G.make_gag_cfg(**over)
    net = GaGNet(**cfg)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(G.gag_param_shapes(cfg).items())
    assert net.cin == 2 and net.dilas == list(cfg["dilas"]) and net.acti_type == "sigmoid"
    net.load_state_dict(G.make_gag_weights(cfg, 0, "B"), strict=True)


def test_product_wrapper_contract_and_errors():
    from eabnet_b200 import GaGNet, make_eabnet_with_postnet
    from eabnet_b200.postnet import default_postnet_args
    w = make_eabnet_with_postnet(default_postnet_args(freeze_eabnet=True))
    sd = G.make_postnet_weights(None, None, 0, "A")
    assert list(w.state_dict().keys()) == list(sd.keys()) and len(sd) == 498 + 815
    w.load_state_dict(sd, strict=True)
    assert not any(p.requires_grad for p in w.eabnet.parameters()) and all(p.requires_grad for p in w.postnet.parameters())
    w.unfreeze_eabnet()
    assert all(p.requires_grad for p in w.eabnet.parameters())
    with pytest.raises(RuntimeError):
        GaGNet(acti_type="gelu")
    with pytest.raises(RuntimeError):
        GaGNet(cin=1)
    with pytest.raises(RuntimeError):                                 # no CPU fallback
        with torch.no_grad():
            GaGNet()(torch.zeros(1, 2, 4, 161), torch.zeros(1, 2, 4, 161))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present on this box")
def test_gag_oracle_matches_live_reference():
    sys.path.insert(0, "/root/reference")
    from GaGNet import GaGNet
    cfg = G.make_gag_cfg(p=1, q=2)
    sd = G.make_gag_weights(cfg, 5, "B")
    net = GaGNet(**{**cfg, "dilas": list(cfg["dilas"])}).eval()
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(G.gag_param_shapes(cfg).items())
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(3)
    x, pre = torch.randn(2, 2, 19, 161, generator=g), 0.5 * torch.randn(2, 2, 19, 161, generator=g)
    with torch.no_grad():
        ref = net(x, pre)
    for a, b in zip(ref, G.gag_forward(sd, x, pre, cfg)):
        assert (a - b).abs().max() <= 2e-5 * max(1.0, float(a.abs().max()))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not present on this box")
def test_default_postnet_args_equal_reference_argparse_defaults():
    """eabnet_b200.postnet.default_postnet_args() mirrors the argparse defaults of train_distributed.py:277-318"""
    import ast
    from eabnet_b200.postnet import default_postnet_args
    src = open("/root/reference/train_distributed.py").read()
    ref = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            name = getattr(node.args[0], "value", "")
            if not isinstance(name, str) or not name.startswith("--"):
                continue
            for kw in node.keywords:
                if kw.arg == "default":
                    try:
                        ref[name[2:]] = ast.literal_eval(kw.value)
                    except Exception:
                        pass
    mine = vars(default_postnet_args())
    checked = 0
    assert ref["M"] == 8 and mine["M"] == 9          # the script's default is its 8-mic checkpoint; ours is EaBNet.py's / BASELINE's 9
    for k, v in mine.items():
        if k in ref and k != "M":
            a, b = (list(v) if isinstance(v, (tuple, list)) else v), (list(ref[k]) if isinstance(ref[k], (tuple, list)) else ref[k])
            assert a == b, (k, v, ref[k])
            checked += 1
    assert checked >= 24, checked
