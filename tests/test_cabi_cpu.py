"""CPU: the C-ABI library loads, exports every symbol include/eabnet_b200.h declares, and its parameter table
is the reference's state_dict (no compute call is made - there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import eabnet_oracle as O

CASES = [{}, {"norm_type": "BN"}, {"is_u2": False, "bf_type": "cnn", "M": 8},
         {"topo_type": "miso", "intra_connect": "add", "M": 1}, {"is_causal": False, "p": 4, "q": 2}]


@pytest.fixture(scope="module")
def lib():
    from eabnet_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "eabnet_b200.h")).read()
    declared = set(re.findall(r"EAB_API[^;(]*?\b(eab_\w+)\s*\(", hdr))
    from eabnet_b200 import _lib
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.eab_build_info().decode().startswith("sm_100a;")


@pytest.mark.parametrize("over", CASES)
def test_state_dict_contract(over):
    from eabnet_b200 import EaBNet
    cfg = O.make_cfg(**over)
    net = EaBNet(**cfg)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(O.param_shapes(cfg).items())
    for k in ("k1", "k2", "c", "M", "embed_dim", "kd1", "cd1", "d_feat", "p", "q", "is_causal", "is_u2", "bf_type",
              "topo_type", "intra_connect", "norm_type"):
        assert getattr(net, k) == cfg[k]
    sd = O.make_weights(cfg)
    net.load_state_dict(sd, strict=True)
    back = net.state_dict()
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    wrapped = {"eabnet." + k: v for k, v in sd.items()}         # a wrapper checkpoint loads nothing (test.py:165)
    res = net.load_state_dict(wrapped, strict=False)
    assert len(res.unexpected_keys) == len(sd)


def test_default_param_count_and_init():
    from eabnet_b200 import EaBNet, numParams
    torch.manual_seed(0)
    net = EaBNet()
    assert numParams(net) == 2838610
    sd = net.state_dict()
    assert float(sd["en.last_conv.2.weight"].mean()) == 0.25
    assert float(sd["en.last_conv.1.norm.weight"].min()) == 1.0
    w = sd["de.last_conv.0.conv.0.weight"]
    assert float(w.abs().max()) <= 1.0 / (128 * 2 * 5) ** 0.5 + 1e-9          # torch fan_in of a ConvTranspose2d


def test_errors_are_loud(lib):
    from eabnet_b200 import EaBNet, _lib, stft_compress
    with pytest.raises(TypeError):
        EaBNet(norm_type="cLN")
    with pytest.raises(RuntimeError, match="d_feat"):
        EaBNet(d_feat=128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EaBNet()(torch.zeros(1, 4, 161, 9, 2))
    with pytest.raises(TypeError, match="no CPU fallback"):
        stft_compress(torch.zeros(1, 9, 1600))
    h = C.c_void_p()
    cfg = _lib.EabConfig(2, 3, 1, 3, 64, 9, 64, 5, 64, 256, 6, 3, 1, 1, 0, 0, 0, 0, 161)
    assert lib.eab_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.eab_set_param(h, b"no.such.key", None, 0) != 0
    assert b"unexpected key" in lib.eab_last_error()
    buf = (C.c_float * 4)()
    assert lib.eab_set_param(h, b"en.last_conv.2.weight", C.cast(buf, C.c_void_p), 4) != 0
    assert b"size mismatch" in lib.eab_last_error()
    assert lib.eab_workspace_bytes(h, 2, 21) > 0            # planning is pure host arithmetic
    lib.eab_destroy(h)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eabnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_stream_state_planning_is_host_only(lib):
    """eab_stream_state_bytes is a pure planning call (no GPU): BN + causal gives a size, InstanceNorm / non-causal
    configurations are refused with a message (streaming needs static normalisation)."""
    import ctypes as C
    from eabnet_b200 import _lib

    def make(norm, causal):
        cfg = _lib.EabConfig(2, 3, 1, 3, 64, 9, 64, 5, 64, 256, 6, 3, causal, 1, 0, 0, 0, norm, 161)
        h = C.c_void_p()
        assert lib.eab_create(C.byref(cfg), C.byref(h)) == 0
        return h

    h = make(1, 1)
    n1, n256 = lib.eab_stream_state_bytes(h, 1), lib.eab_stream_state_bytes(h, 256)
    assert n1 > 0 and n256 > 200 * n1 and n256 < 4 << 30
    lib.eab_destroy(h)
    for norm, causal in ((0, 1), (1, 0)):
        h = make(norm, causal)
        assert lib.eab_stream_state_bytes(h, 4) == 0
        assert b"streaming needs" in lib.eab_last_error()
        lib.eab_destroy(h)
