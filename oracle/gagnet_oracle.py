"""CPU oracle for the GaGNet post-filter and the EaBNetWithPostNet wrapper  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Functional torch-fp32 restatement (flat ``state_dict`` in, tensors out) of SURVEY.md section 8(f) rank 1:

    GaGNet.forward            GaGNet.py:75-89     (U2 / U-Net encoder on cat(inpt, pre_x), q glance-gaze modules)
    GlanceGazeModule.forward  GaGNet.py:120-134   (gain * |pre| * e^{j phase(pre)} + complex residual)
    GlanceBlock.forward       GaGNet.py:181-193   GazeBlock.forward   GaGNet.py:241-259
    SqueezedTCM (one branch)  GaGNet.py:285-326   encoders            GaGNet.py:329-413
    EaBNetWithPostNet.forward EaBNet.py:139-148   (reference microphone + detached EaBNet estimate -> post-filter)

The encoder blocks (gated conv, En_unet_module, NormSwitch, PReLU) are the ones of eabnet_oracle: GaGNet.py restates
the same classes.  Same rules as eabnet_oracle.py: only tests/, smoke() and bench.py's CPU baseline may import this;
the pin is the live reference module (tools/make_golden.py -> tests/golden/gag_*.npz), since the reference ships no
tests or golden vectors of its own.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import eabnet_oracle as E

Tensor = torch.Tensor
SD = Dict[str, Tensor]

DEFAULT_GAG_CFG = dict(cin=2, k1=(2, 3), k2=(1, 3), c=64, kd1=3, cd1=64, d_feat=256, p=2, q=3, dilas=(1, 2, 5, 9),
                       fft_num=320, is_u2=True, is_causal=True, is_squeezed=False, acti_type="sigmoid",
                       intra_connect="cat", norm_type="IN")


def make_gag_cfg(**kw) -> dict:
    cfg = dict(DEFAULT_GAG_CFG)
    cfg.update(kw)
    cfg["k1"], cfg["k2"], cfg["dilas"] = tuple(cfg["k1"]), tuple(cfg["k2"]), tuple(cfg["dilas"])
    return cfg


def gag_encoder(sd: SD, x: Tensor, cfg: dict) -> Tensor:
    """U2Net_Encoder / UNet_Encoder of GaGNet.py:329-413: only the bottleneck is returned; unlike EaBNet's plain
    U-Net encoder every layer of GaGNet's carries a norm (GaGNet.py:386-405)."""
    if cfg["is_u2"]:
        for i, (k, scale) in enumerate([((2, 5), 4), (cfg["k1"], 3), (cfg["k1"], 2), (cfg["k1"], 1)]):
            x = E.unet_module(sd, "en.meta_unet_list.%d" % i, x, k, scale, False, cfg)
        return E._gated_block(sd, "en.last_conv", x, cfg["k1"], cfg, False)
    for i in range(5):
        x = E._gated_block(sd, "en.unet_list.%d" % i, x, (2, 5) if i == 0 else cfg["k1"], cfg, False)
    return x


def gag_tcm(sd: SD, pfx: str, x: Tensor, dilation: int, cfg: dict) -> Tensor:
    """SqueezedTCM of GaGNet.py:285-326: 1x1 squeeze, ONE PReLU->norm->pad->dilated conv branch, PReLU->norm->1x1
    expand, residual (no bias anywhere)."""
    span = (cfg["kd1"] - 1) * dilation
    pad = (span, 0) if cfg["is_causal"] else (span // 2, span // 2)
    y = F.conv1d(x, sd[pfx + ".in_conv.weight"])
    y = E._norm(sd, pfx + ".d_conv.1", E._prelu(sd, pfx + ".d_conv.0.weight", y), cfg)
    y = F.conv1d(F.pad(y, pad), sd[pfx + ".d_conv.3.weight"], dilation=dilation)
    y = E._norm(sd, pfx + ".out_conv.1", E._prelu(sd, pfx + ".out_conv.0.weight", y), cfg)
    return F.conv1d(y, sd[pfx + ".out_conv.2.weight"]) + x


def _tcm_groups(sd: SD, pfx: str, x: Tensor, cfg: dict) -> Tensor:
    for g in range(cfg["p"]):
        for i, d in enumerate(cfg["dilas"]):
            x = gag_tcm(sd, "%s.%d.tcns.%d" % (pfx, g, i), x, d, cfg)
    return x


def _gated_in(sd: SD, pfx: str, feat: Tensor, pre: Tensor) -> Tensor:
    """in_conv_main(cat) * sigmoid(in_conv_gate(cat)) on cat(feat_x [B,256,T], pre_x viewed [B,2F,T])."""
    inpt = torch.cat((feat, pre.reshape(pre.shape[0], -1, pre.shape[-1])), dim=1)
    main = F.conv1d(inpt, sd[pfx + ".in_conv_main.weight"], sd[pfx + ".in_conv_main.bias"])
    gate = F.conv1d(inpt, sd[pfx + ".in_conv_gate.0.weight"], sd[pfx + ".in_conv_gate.0.bias"])
    return main * torch.sigmoid(gate)


def glance(sd: SD, pfx: str, feat: Tensor, pre: Tensor, cfg: dict) -> Tensor:
    x = _tcm_groups(sd, pfx + ".tcn_g", _gated_in(sd, pfx, feat, pre), cfg)
    g = F.conv1d(x, sd[pfx + ".linear_g.0.weight"], sd[pfx + ".linear_g.0.bias"])
    return {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu}[cfg["acti_type"]](g)


def gaze(sd: SD, pfx: str, feat: Tensor, pre: Tensor, cfg: dict) -> Tensor:
    x = _gated_in(sd, pfx, feat, pre)
    if cfg["is_squeezed"]:
        xr = xi = _tcm_groups(sd, pfx + ".tcm_ri", x, cfg)
    else:
        xr, xi = _tcm_groups(sd, pfx + ".tcm_r", x, cfg), _tcm_groups(sd, pfx + ".tcm_i", x, cfg)
    return torch.stack((F.conv1d(xr, sd[pfx + ".linear_r.weight"], sd[pfx + ".linear_r.bias"]),
                        F.conv1d(xi, sd[pfx + ".linear_i.weight"], sd[pfx + ".linear_i.bias"])), dim=1)


@torch.no_grad()
def gag_forward(sd: SD, inpt: Tensor, pre_x: Tensor, cfg: dict | None = None) -> List[Tensor]:
    """GaGNet.forward (GaGNet.py:75-89): inpt, pre_x [B,2,T,F] -> list of q estimates, each [B,2,F,T]."""
    cfg = make_gag_cfg() if cfg is None else cfg
    B, _, T, _ = inpt.shape
    feat = gag_encoder(sd, torch.cat((inpt, pre_x), dim=1), cfg)
    feat = feat.transpose(-2, -1).reshape(B, -1, T)                    # channel = c*Fb + f
    pre = pre_x.transpose(-2, -1).contiguous()                         # [B,2,F,T]
    outs = []
    for i in range(cfg["q"]):
        gain = glance(sd, "gags.%d.glance_block" % i, feat, pre, cfg)
        resi = gaze(sd, "gags.%d.gaze_block" % i, feat, pre, cfg)
        mag, pha = torch.norm(pre, dim=1), torch.atan2(pre[:, -1], pre[:, 0])
        filt = mag * gain
        pre = torch.stack((filt * torch.cos(pha), filt * torch.sin(pha)), dim=1) + resi
        outs.append(pre)
    return outs


@torch.no_grad()
def postnet_forward(sd: SD, noisy: Tensor, cfg_e: dict | None = None, cfg_g: dict | None = None,
                    ref_mic: int = 0) -> dict:
    """EaBNetWithPostNet.forward (EaBNet.py:139-148) on a wrapper state_dict (keys `eabnet.*`, `postnet.*`)."""
    sd_e = {k[len("eabnet."):]: v for k, v in sd.items() if k.startswith("eabnet.")}
    sd_g = {k[len("postnet."):]: v for k, v in sd.items() if k.startswith("postnet.")}
    est0 = E.forward(sd_e, noisy, cfg_e)
    ref = noisy[..., ref_mic, :].permute(0, 3, 1, 2)
    lst = gag_forward(sd_g, ref, est0, cfg_g)
    return {"esti0_stft": est0, "esti1_stft_list": lst, "esti_stft": lst[-1].permute(0, 1, 3, 2)}


# --------------------------------------------------------------------------------------------------------
def gag_param_shapes(cfg: dict | None = None) -> Dict[str, Tuple[int, ...]]:
    """Names / shapes of ``GaGNet(**cfg).state_dict()`` in registration order (en, gags; GaGNet.py:69-73)."""
    cfg = make_gag_cfg() if cfg is None else cfg
    c, k1, k2 = cfg["c"], cfg["k1"], cfg["k2"]
    bn = cfg["norm_type"] == "BN"
    out: Dict[str, Tuple[int, ...]] = {}

    def norm(p, ch):
        out[p + ".norm.weight"] = (ch,)
        out[p + ".norm.bias"] = (ch,)
        if bn:
            out[p + ".norm.running_mean"] = (ch,)
            out[p + ".norm.running_var"] = (ch,)
            out[p + ".norm.num_batches_tracked"] = ()

    def gated(p, cin, cout, k):
        sub = ".conv.1" if k[0] > 1 else ".conv"
        out[p + ".0" + sub + ".weight"] = (2 * cout, cin, *k)
        out[p + ".0" + sub + ".bias"] = (2 * cout,)
        norm(p + ".1", cout)
        out[p + ".2.weight"] = (cout,)

    def module(p, cin, k_in, scale):
        gated(p + ".in_conv", cin, c, k_in)
        for i in range(scale):
            q = "%s.enco.%d.conv" % (p, i)
            out[q + ".0.weight"] = (c, c, *k2)
            out[q + ".0.bias"] = (c,)
            norm(q + ".1", c)
            out[q + ".2.weight"] = (c,)
        for i in range(scale):
            q = "%s.deco.%d.deconv" % (p, i)
            cin_d = c if (i == 0 or cfg["intra_connect"] == "add") else 2 * c
            out[q + ".0.weight"] = (cin_d, c, *k2)
            out[q + ".0.bias"] = (c,)
            norm(q + ".1", c)
            out[q + ".2.weight"] = (c,)

    cin = 2 * cfg["cin"]
    if cfg["is_u2"]:
        module("en.meta_unet_list.0", cin, (2, 5), 4)
        for i in (1, 2, 3):
            module("en.meta_unet_list.%d" % i, c, k1, 4 - i)
        gated("en.last_conv", c, 64, k1)
    else:
        gated("en.unet_list.0", cin, c, (2, 5))
        for i in (1, 2, 3):
            gated("en.unet_list.%d" % i, c, c, k1)
        gated("en.unet_list.4", c, 64, k1)

    cd, df, kd, Fq = cfg["cd1"], cfg["d_feat"], cfg["kd1"], cfg["fft_num"] // 2 + 1
    ci = 2 * Fq + df

    def tcm_groups(p):
        for g in range(cfg["p"]):
            for i in range(len(cfg["dilas"])):
                t = "%s.%d.tcns.%d" % (p, g, i)
                out[t + ".in_conv.weight"] = (cd, df, 1)
                out[t + ".d_conv.0.weight"] = (cd,)
                norm(t + ".d_conv.1", cd)
                out[t + ".d_conv.3.weight"] = (cd, cd, kd)
                out[t + ".out_conv.0.weight"] = (cd,)
                norm(t + ".out_conv.1", cd)
                out[t + ".out_conv.2.weight"] = (df, cd, 1)

    def in_convs(p):
        out[p + ".in_conv_main.weight"] = (df, ci, 1)
        out[p + ".in_conv_main.bias"] = (df,)
        out[p + ".in_conv_gate.0.weight"] = (df, ci, 1)
        out[p + ".in_conv_gate.0.bias"] = (df,)

    for i in range(cfg["q"]):
        p = "gags.%d.glance_block" % i
        in_convs(p)
        tcm_groups(p + ".tcn_g")
        out[p + ".linear_g.0.weight"] = (Fq, df, 1)
        out[p + ".linear_g.0.bias"] = (Fq,)
        p = "gags.%d.gaze_block" % i
        in_convs(p)
        if cfg["is_squeezed"]:
            tcm_groups(p + ".tcm_ri")
        else:
            tcm_groups(p + ".tcm_r")
            tcm_groups(p + ".tcm_i")
        for n in ("linear_r", "linear_i"):
            out["%s.%s.weight" % (p, n)] = (Fq, df, 1)
            out["%s.%s.bias" % (p, n)] = (Fq,)
    return out


def make_gag_weights(cfg: dict | None = None, seed: int = 0, variant: str = "B") -> SD:
    """Name-seeded synthetic weights with torch's default scales (same scheme as eabnet_oracle.make_weights)."""
    import zlib
    import numpy as np

    shapes = gag_param_shapes(cfg)
    sd: SD = {}
    for name, shape in shapes.items():
        rng = np.random.RandomState((zlib.crc32(("gag." + name).encode()) + 7919 * seed) % (2 ** 31))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(0, dtype=torch.long)
            continue
        if leaf == "running_mean":
            v = rng.normal(0, 0.1, shape) if variant == "B" else np.zeros(shape)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, shape) if variant == "B" else np.ones(shape)
        elif ".norm." in name and leaf == "weight":
            v = rng.uniform(0.5, 1.5, shape) if variant == "B" else np.ones(shape)
        elif ".norm." in name and leaf == "bias":
            v = rng.normal(0, 0.1, shape) if variant == "B" else np.zeros(shape)
        elif len(shape) == 1 and leaf == "weight":
            v = rng.uniform(0.05, 0.5, shape) if variant == "B" else np.full(shape, 0.25)
        else:
            wshape = shapes[name[:-4] + "weight"] if leaf == "bias" else shape
            fan_in = wshape[1] * int(np.prod(wshape[2:]))
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, shape)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape).clone()
    return sd


def make_postnet_weights(cfg_e: dict | None = None, cfg_g: dict | None = None, seed: int = 0, variant: str = "B") -> SD:
    sd = {"eabnet." + k: v for k, v in E.make_weights(cfg_e, seed, variant).items()}
    sd.update({"postnet." + k: v for k, v in make_gag_weights(cfg_g, seed, variant).items()})
    return sd
