"""CPU oracle for the EaBNet inference hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a from-scratch, *functional* restatement (plain torch fp32/fp64 ops on the CPU, driven by a
flat ``state_dict``) of what the reference computes on the path

    waveform -> STFT + sqrt compression -> U2/U-Net encoder -> squeezed TCM stack -> decoder
             -> beam-weight head (LSTM | 1x1 conv) -> complex filter-and-sum -> iSTFT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it, and only as the checker or the timed CPU baseline.  ``eabnet_b200`` never imports it.

Parity pin: the reference repo ships no tests / golden vectors, so the oracle is pinned
against outputs of the *reference module itself*, imported from /root/reference in the build container by
``tools/make_golden.py``; the resulting vectors live in ``tests/golden/`` and are checked by
``tests/test_oracle_golden.py`` (CPU).  Where /root/reference is importable the same test also compares
live.

Reference lines each function follows (all in /root/reference/):
    stft_compress      test.py:20-47 (== train_distributed.py:68-95)
    istft              enhance.py:59-61, test.py:189-190
    gated_conv2d       EaBNet.py:434-460          gated_deconv2d   EaBNet.py:463-490, 617-624
    norm               EaBNet.py:662-694          unet_module      EaBNet.py:331-431, 493-503
    encoder/decoder    EaBNet.py:157-328          tcm / tcm stack  EaBNet.py:100-106, 506-578
    lstm_head          EaBNet.py:581-614          forward          EaBNet.py:88-125
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

DEFAULT_CFG = dict(k1=(2, 3), k2=(1, 3), c=64, M=9, embed_dim=64, kd1=5, cd1=64, d_feat=256, p=6, q=3,
                   is_causal=True, is_u2=True, bf_type="lstm", topo_type="mimo", intra_connect="cat",
                   norm_type="IN")

N_FFT, HOP = 320, 160


def make_cfg(**kw) -> dict:
    cfg = dict(DEFAULT_CFG)
    cfg.update(kw)
    cfg["k1"], cfg["k2"] = tuple(cfg["k1"]), tuple(cfg["k2"])
    return cfg


# --------------------------------------------------------------------------------------------------------
# signal front / back end
# --------------------------------------------------------------------------------------------------------
def stft_compress(wave: Tensor) -> Tensor:
    """wave [B, M, L] -> compressed spectrum [B, T, F, M, 2]  (test.py:32-43).

    torch.stft(320, 160, 320, periodic hann, center=True/reflect, onesided), then |z|^0.5 with the phase
    kept, written the way the reference does it (norm ** 0.5, atan2, cos/sin) so that the rounding matches.
    """
    B, M, L = wave.shape
    win = torch.hann_window(N_FFT, dtype=wave.dtype, device=wave.device)
    z = torch.stft(wave.reshape(B * M, L), N_FFT, HOP, N_FFT, win, return_complex=True)   # [BM, F, T]
    z = torch.view_as_real(z)                                                            # [BM, F, T, 2]
    Fq, T = z.shape[1], z.shape[2]
    z = z.view(B, M, Fq, T, 2).permute(0, 3, 2, 1, 4)                                      # [B, T, F, M, 2]
    mag = torch.norm(z, dim=-1) ** 0.5
    pha = torch.atan2(z[..., 1], z[..., 0])
    return torch.stack((mag * torch.cos(pha), mag * torch.sin(pha)), dim=-1).contiguous()


def istft(spec: Tensor) -> Tensor:
    """spec [B, 2, T, F] -> wave [B, 160 (T-1)]  (enhance.py:59-61); no de-compression, as the reference."""
    z = torch.view_as_complex(spec.permute(0, 3, 2, 1).contiguous())                      # [B, F, T]
    win = torch.hann_window(N_FFT, dtype=spec.dtype, device=spec.device)
    return torch.istft(z, N_FFT, HOP, N_FFT, win)


# --------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------
def _norm(sd: SD, pfx: str, x: Tensor, cfg: dict) -> Tensor:
    """NormSwitch (EaBNet.py:662-694).  IN: per-(b,c) batch statistics always; BN: running stats (eval)."""
    g, b = sd[pfx + ".norm.weight"], sd[pfx + ".norm.bias"]
    if cfg["norm_type"] == "IN":
        return F.instance_norm(x, None, None, g, b, True, 0.1, 1e-5)
    if cfg["norm_type"] == "BN":
        return F.batch_norm(x, sd[pfx + ".norm.running_mean"], sd[pfx + ".norm.running_var"], g, b,
                            False, 0.1, 1e-5)
    raise ValueError("norm_type %r cannot be constructed in the reference either" % cfg["norm_type"])


def _prelu(sd: SD, name: str, x: Tensor) -> Tensor:
    return F.prelu(x, sd[name])


def _gate(x: Tensor) -> Tensor:
    a, g = x.chunk(2, dim=1)
    return a * torch.sigmoid(g)


def gated_conv2d(sd: SD, pfx: str, x: Tensor, k: Tuple[int, int]) -> Tensor:
    """GateConv2d (EaBNet.py:434-460): causal top pad kt-1, Conv2d stride (1,2), value*sigmoid(gate)."""
    kt = k[0]
    if kt > 1:
        x = F.pad(x, (0, 0, kt - 1, 0))
        w, b = sd[pfx + ".conv.1.weight"], sd[pfx + ".conv.1.bias"]
    else:
        w, b = sd[pfx + ".conv.weight"], sd[pfx + ".conv.bias"]
    return _gate(F.conv2d(x, w, b, stride=(1, 2)))


def gated_deconv2d(sd: SD, pfx: str, x: Tensor, k: Tuple[int, int]) -> Tensor:
    """GateConvTranspose2d + Chomp_T (EaBNet.py:463-490, 617-624)."""
    kt = k[0]
    if kt > 1:
        y = F.conv_transpose2d(x, sd[pfx + ".conv.0.weight"], sd[pfx + ".conv.0.bias"], stride=(1, 2))
        y = y[:, :, :-(kt - 1), :]
    else:
        y = F.conv_transpose2d(x, sd[pfx + ".conv.weight"], sd[pfx + ".conv.bias"], stride=(1, 2))
    return _gate(y)


def _gated_block(sd: SD, pfx: str, x: Tensor, k, cfg: dict, deconv: bool, with_norm: bool = True) -> Tensor:
    """Sequential(gated (de)conv, [NormSwitch], PReLU) with indices 0, 1, 2 (or 0, 1 without the norm)."""
    y = (gated_deconv2d if deconv else gated_conv2d)(sd, pfx + ".0", x, k)
    if with_norm:
        y = _norm(sd, pfx + ".1", y, cfg)
        return _prelu(sd, pfx + ".2.weight", y)
    return _prelu(sd, pfx + ".1.weight", y)


def unet_module(sd: SD, pfx: str, x: Tensor, k_in, scale: int, deconv: bool, cfg: dict) -> Tensor:
    """En_unet_module (EaBNet.py:331-388) with Conv2dunit / Deconv2dunit / Skip_connect (:391-431, :493-503)."""
    x0 = _gated_block(sd, pfx + ".in_conv", x, k_in, cfg, deconv)
    y, keep = x0, []
    for i in range(scale):
        p = "%s.enco.%d.conv" % (pfx, i)
        y = F.conv2d(y, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=(1, 2))
        y = _prelu(sd, p + ".2.weight", _norm(sd, p + ".1", y, cfg))
        keep.append(y)
    for i in range(scale):
        p = "%s.deco.%d.deconv" % (pfx, i)
        if i > 0:
            aux = keep[-(i + 1)]
            y = torch.cat((y, aux), dim=1) if cfg["intra_connect"] == "cat" else y + aux
        y = F.conv_transpose2d(y, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=(1, 2))
        y = _prelu(sd, p + ".2.weight", _norm(sd, p + ".1", y, cfg))
    return x0 + y


def _tap(taps, name: str, x: Tensor) -> None:
    """record a channels-last copy [B,T,F,C] of an intermediate (debug aid for the stage-wise GPU tests)"""
    if taps is not None:
        taps[name] = x.permute(0, 2, 3, 1).contiguous() if x.dim() == 4 else x


def encoder(sd: SD, x: Tensor, cfg: dict) -> Tuple[Tensor, List[Tensor]]:
    """U2Net_Encoder (EaBNet.py:157-197) or UNet_Encoder (:199-239)."""
    skips = []
    if cfg["is_u2"]:
        for i, (k, scale) in enumerate([((2, 5), 4), (cfg["k1"], 3), (cfg["k1"], 2), (cfg["k1"], 1)]):
            x = unet_module(sd, "en.meta_unet_list.%d" % i, x, k, scale, False, cfg)
            skips.append(x)
        x = _gated_block(sd, "en.last_conv", x, cfg["k1"], cfg, False)
        skips.append(x)
    else:
        for i in range(5):
            k = (2, 5) if i == 0 else cfg["k1"]
            x = _gated_block(sd, "en.unet_list.%d" % i, x, k, cfg, False, with_norm=i not in (1, 2))
            skips.append(x)
    return x, skips


def decoder(sd: SD, x: Tensor, skips: List[Tensor], cfg: dict, taps=None) -> Tensor:
    """U2Net_Decoder (EaBNet.py:241-279) or UNet_Decoder (:282-328)."""
    if cfg["is_u2"]:
        for i in range(4):
            x = unet_module(sd, "de.meta_unet_list.%d" % i, torch.cat((x, skips[-(i + 1)]), dim=1),
                            cfg["k1"], i + 1, True, cfg)
            _tap(taps, "de.%d" % i, x)
        return _gated_block(sd, "de.last_conv", torch.cat((x, skips[0]), dim=1), (2, 5), cfg, True)
    for i in range(5):
        k = (2, 5) if i == 4 else cfg["k1"]
        x = _gated_block(sd, "de.unet_list.%d" % i, torch.cat((x, skips[-(i + 1)]), dim=1), k, cfg, True)
        if i < 4:
            _tap(taps, "de.%d" % i, x)
    return x


def tcm(sd: SD, pfx: str, x: Tensor, dilation: int, cfg: dict) -> Tensor:
    """SqueezedTCM (EaBNet.py:532-578): 1x1 squeeze, two PReLU->norm->pad->dilated-conv branches (the second
    through a sigmoid), product, PReLU->norm->1x1 expand, residual."""
    span = (cfg["kd1"] - 1) * dilation
    pad = (span, 0) if cfg["is_causal"] else (span // 2, span // 2)
    y = F.conv1d(x, sd[pfx + ".in_conv.weight"])

    def branch(name: str) -> Tensor:
        z = _prelu(sd, "%s.%s.0.weight" % (pfx, name), y)
        z = _norm(sd, "%s.%s.1" % (pfx, name), z, cfg)
        return F.conv1d(F.pad(z, pad), sd["%s.%s.3.weight" % (pfx, name)], dilation=dilation)

    z = branch("left_conv") * torch.sigmoid(branch("right_conv"))
    z = _norm(sd, pfx + ".out_conv.1", _prelu(sd, pfx + ".out_conv.0.weight", z), cfg)
    return F.conv1d(z, sd[pfx + ".out_conv.2.weight"]) + x


def tcm_stack(sd: SD, x: Tensor, cfg: dict) -> Tensor:
    """x [B, 64, T, Fb] -> same shape (EaBNet.py:99-106): channel = c*Fb + f, chained groups, summed."""
    B, C, T, Fb = x.shape
    y = x.transpose(-2, -1).reshape(B, C * Fb, T)
    acc = torch.zeros_like(y)
    for g in range(cfg["q"]):
        for i in range(cfg["p"]):
            y = tcm(sd, "stcns.%d.tcm_list.%d" % (g, i), y, 2 ** i, cfg)
        acc = acc + y
    return acc.view(B, C, Fb, T).transpose(-2, -1).contiguous()


def lstm_layer(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """nn.LSTM(batch_first, 1 layer, zero initial state) through torch's own fused CPU primitive - the same call
    the reference makes (EaBNet.py:610-611), so the CPU baseline is not handicapped by a Python time loop.
    tests/test_oracle_golden.py checks it against the step-by-step restatement below."""
    H = w_hh.shape[1]
    rnn = torch.nn.LSTM(input_size=w_ih.shape[1], hidden_size=H, batch_first=True).to(device=x.device, dtype=x.dtype)
    with torch.no_grad():
        rnn.weight_ih_l0.copy_(w_ih)
        rnn.weight_hh_l0.copy_(w_hh)
        rnn.bias_ih_l0.copy_(b_ih)
        rnn.bias_hh_l0.copy_(b_hh)
        return rnn(x)[0]


def lstm_layer_stepwise(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """The recurrence written out: gate order i,f,g,o, both biases added, h0 = c0 = 0."""
    N, T, _ = x.shape
    H = w_hh.shape[1]
    gx = x @ w_ih.t() + (b_ih + b_hh)
    h = x.new_zeros(N, H)
    c = x.new_zeros(N, H)
    out = []
    for t in range(T):
        g = gx[:, t] + h @ w_hh.t()
        i, f, gg, o = g.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out.append(h)
    return torch.stack(out, dim=1)


def lstm_head(sd: SD, emb: Tensor, cfg: dict, taps=None) -> Tensor:
    """LSTM_BF (EaBNet.py:581-614): emb [B, C, T, F] -> beam weights [B, T, F, M, 2]."""
    B, C, T, Fq = emb.shape
    x = F.layer_norm(emb.permute(0, 3, 2, 1), (C,), sd["bf_map.norm.weight"], sd["bf_map.norm.bias"], 1e-5)
    x = x.reshape(B * Fq, T, C)
    for r in ("rnn1", "rnn2"):
        x = lstm_layer(x, *(sd["bf_map.%s.%s_l0" % (r, n)] for n in
                            ("weight_ih", "weight_hh", "bias_ih", "bias_hh")))
        if taps is not None:
            taps["h1" if r == "rnn1" else "h2"] = x.view(B, Fq, T, -1).transpose(1, 2).contiguous()
    x = x.view(B, Fq, T, -1).transpose(1, 2)
    x = torch.relu(F.linear(x, sd["bf_map.w_dnn.0.weight"], sd["bf_map.w_dnn.0.bias"]))
    x = F.linear(x, sd["bf_map.w_dnn.2.weight"], sd["bf_map.w_dnn.2.bias"])
    return x.reshape(B, T, Fq, cfg["M"], 2)


def filter_and_sum(w: Tensor, inpt: Tensor) -> Tensor:
    """EaBNet.py:114-117: complex multiply-accumulate over the mic axis -> [B, 2, T, F]."""
    wr, wi, xr, xi = w[..., 0], w[..., 1], inpt[..., 0], inpt[..., 1]
    return torch.stack(((wr * xr - wi * xi).sum(-1), (wr * xi + wi * xr).sum(-1)), dim=1)


def network_embedding(sd: SD, inpt: Tensor, cfg: dict, taps=None) -> Tensor:
    """inpt [B,T,F,M,2] -> decoder output [B, embed_dim, T, F]  (EaBNet.py:95-107)."""
    B, T, Fq, M, _ = inpt.shape
    x = inpt.transpose(-2, -1).reshape(B, T, Fq, 2 * M).permute(0, 3, 1, 2)           # channel = ri*M + m
    x, skips = encoder(sd, x, cfg)
    for i, s in enumerate(skips):
        _tap(taps, "en.%d" % i, s)
    x = tcm_stack(sd, x, cfg)
    _tap(taps, "tcm", x)
    emb = decoder(sd, x, skips, cfg, taps)
    _tap(taps, "embed", emb)
    return emb


@torch.no_grad()
def forward(sd: SD, inpt: Tensor, cfg: dict | None = None, taps=None) -> Tensor:
    """EaBNet.forward (EaBNet.py:88-125).  inpt [B,T,F,M,2] or [B,T,F,2]; returns [B,2,T,F] ([B,2,T] miso).
    `taps`, if a dict, receives channels-last copies of the main intermediates."""
    cfg = make_cfg() if cfg is None else cfg
    if inpt.dim() == 4:
        inpt = inpt.unsqueeze(-2)
    B, T, Fq, M, _ = inpt.shape
    emb = network_embedding(sd, inpt, cfg, taps)
    if cfg["topo_type"] == "mimo":
        if cfg["bf_type"] == "lstm":
            w = lstm_head(sd, emb, cfg, taps)
        else:
            w = F.conv2d(emb, sd["bf_map.weight"], sd["bf_map.bias"])
            w = w.view(B, M, -1, T, Fq).permute(0, 3, 4, 1, 2)
        return filter_and_sum(w, inpt)
    w = F.conv2d(emb, sd["bf_map.weight"], sd["bf_map.bias"]).permute(0, 2, 3, 1)       # [B,T,F,2]
    wr, wi, xr, xi = w[..., 0], w[..., 1], inpt[..., 0, 0], inpt[..., 0, 1]
    # the reference sums the (already mic-free) product over its last axis, i.e. over F (EaBNet.py:123-124)
    return torch.stack(((wr * xr - wi * xi).sum(-1), (wr * xi + wi * xr).sum(-1)), dim=1)


@torch.no_grad()
def enhance(sd: SD, wave: Tensor, cfg: dict | None = None) -> Tensor:
    """wave [B,M,L] -> enhanced wave [B, 160*(L//160)]: prepare_data -> forward -> istft (test.py:178-190)."""
    return istft(forward(sd, stft_compress(wave), cfg))


# --------------------------------------------------------------------------------------------------------
# parameter table (names / shapes of the reference state_dict) and seeded weights / inputs
# --------------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict | None = None) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of the reference ``EaBNet(**cfg).state_dict()`` in registration order
    (en, de, bf_map, stcns - EaBNet.py:68-86), derived independently of the product's C++ table."""
    cfg = make_cfg() if cfg is None else cfg
    c, M, E, k1, k2 = cfg["c"], cfg["M"], cfg["embed_dim"], cfg["k1"], cfg["k2"]
    bn = cfg["norm_type"] == "BN"
    out: Dict[str, Tuple[int, ...]] = {}

    def norm(p, ch):
        out[p + ".norm.weight"] = (ch,)
        out[p + ".norm.bias"] = (ch,)
        if bn:
            out[p + ".norm.running_mean"] = (ch,)
            out[p + ".norm.running_var"] = (ch,)
            out[p + ".norm.num_batches_tracked"] = ()

    def gated(p, cin, cout, k, deconv, with_norm=True):
        sub = (".conv.0" if deconv else ".conv.1") if k[0] > 1 else ".conv"
        out[p + ".0" + sub + ".weight"] = (cin, 2 * cout, *k) if deconv else (2 * cout, cin, *k)
        out[p + ".0" + sub + ".bias"] = (2 * cout,)
        if with_norm:
            norm(p + ".1", cout)
            out[p + ".2.weight"] = (cout,)
        else:
            out[p + ".1.weight"] = (cout,)

    def module(p, cin, k_in, scale, deconv):
        gated(p + ".in_conv", cin, c, k_in, deconv)
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

    if cfg["is_u2"]:
        module("en.meta_unet_list.0", 2 * M, (2, 5), 4, False)
        for i in (1, 2, 3):
            module("en.meta_unet_list.%d" % i, c, k1, 4 - i, False)
        gated("en.last_conv", c, 64, k1, False)
        module("de.meta_unet_list.0", 128, k1, 1, True)
        for i in (1, 2, 3):
            module("de.meta_unet_list.%d" % i, 2 * c, k1, i + 1, True)
        gated("de.last_conv", 2 * c, E, (2, 5), True)
    else:
        gated("en.unet_list.0", 2 * M, c, (2, 5), False)
        gated("en.unet_list.1", c, c, k1, False, with_norm=False)
        gated("en.unet_list.2", c, c, k1, False, with_norm=False)
        gated("en.unet_list.3", c, c, k1, False)
        gated("en.unet_list.4", c, 64, k1, False)
        gated("de.unet_list.0", 128, c, k1, True)
        for i in (1, 2, 3):
            gated("de.unet_list.%d" % i, 2 * c, c, k1, True)
        gated("de.unet_list.4", 2 * c, E, (2, 5), True)

    if cfg["topo_type"] == "mimo" and cfg["bf_type"] == "lstm":
        H = 64
        for r, cin in (("rnn1", E), ("rnn2", H)):
            out["bf_map.%s.weight_ih_l0" % r] = (4 * H, cin)
            out["bf_map.%s.weight_hh_l0" % r] = (4 * H, H)
            out["bf_map.%s.bias_ih_l0" % r] = (4 * H,)
            out["bf_map.%s.bias_hh_l0" % r] = (4 * H,)
        out["bf_map.w_dnn.0.weight"] = (H, H)
        out["bf_map.w_dnn.0.bias"] = (H,)
        out["bf_map.w_dnn.2.weight"] = (2 * M, H)
        out["bf_map.w_dnn.2.bias"] = (2 * M,)
        out["bf_map.norm.weight"] = (E,)
        out["bf_map.norm.bias"] = (E,)
    else:
        n = 2 * M if cfg["topo_type"] == "mimo" else 2
        out["bf_map.weight"] = (n, E, 1, 1)
        out["bf_map.bias"] = (n,)

    cd, df, kd = cfg["cd1"], cfg["d_feat"], cfg["kd1"]
    for g in range(cfg["q"]):
        for i in range(cfg["p"]):
            p = "stcns.%d.tcm_list.%d" % (g, i)
            out[p + ".in_conv.weight"] = (cd, df, 1)
            for br in ("left_conv", "right_conv"):
                out["%s.%s.0.weight" % (p, br)] = (cd,)
                norm("%s.%s.1" % (p, br), cd)
                out["%s.%s.3.weight" % (p, br)] = (cd, cd, kd)
            out[p + ".out_conv.0.weight"] = (cd,)
            norm(p + ".out_conv.1", cd)
            out[p + ".out_conv.2.weight"] = (df, cd, 1)
    return out


def make_weights(cfg: dict | None = None, seed: int = 0, variant: str = "B") -> SD:
    """Deterministic synthetic weights that do not depend on torch's RNG stream or module construction
    order (each tensor is seeded from its own name).  Scales follow torch's default initialisers so that
    activations stay O(1).  variant "A": norm gamma=1, beta=0, PReLU 0.25 (torch defaults);
    variant "B": gamma~U(.5,1.5), beta~N(0,.1), PReLU~U(.05,.5), BN running stats randomised - catches
    fusions that drop an affine term."""
    import zlib
    import numpy as np

    sd: SD = {}
    for name, shape in param_shapes(cfg).items():
        rng = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
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
        elif len(shape) == 1 and leaf == "weight":                      # PReLU slopes
            v = rng.uniform(0.05, 0.5, shape) if variant == "B" else np.full(shape, 0.25)
        elif "rnn" in name:
            v = rng.uniform(-0.125, 0.125, shape)                       # 1/sqrt(hidden)
        else:
            if leaf == "bias":
                wshape = param_shapes(cfg)[name[:-4] + "weight"]
            else:
                wshape = shape
            # fan_in as torch computes it: size(1) * receptive field (also for ConvTranspose weights)
            fan_in = wshape[1] * int(np.prod(wshape[2:])) if len(wshape) > 1 else wshape[0]
            bound = 1.0 / math.sqrt(fan_in)
            v = rng.uniform(-bound, bound, shape)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32)).reshape(shape).clone()
    return sd


def make_wave(B: int, M: int, L: int, seed: int = 1234) -> Tuple[Tensor, Tensor]:
    """Synthetic multichannel mixture: a 5-harmonic source with a 4 Hz envelope, delayed by
    m samples at mic m, plus white noise.  Returns (x [B,M,L], clean source [B,L])."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L + M, dtype=torch.float64) / 16000.0
    xs, ss = [], []
    for _ in range(B):
        f0 = 100.0 + 150.0 * torch.rand(1, generator=g, dtype=torch.float64).item()
        s = sum(torch.sin(2 * math.pi * f0 * (h + 1) * t) / (h + 1) for h in range(5))
        s = 0.1 * s / s.abs().max() * (0.5 - 0.5 * torch.cos(2 * math.pi * 4.0 * t))
        s = s.float()
        mics = torch.stack([s[M - m: M - m + L] for m in range(M)])          # delay of m samples
        xs.append(mics + 0.05 * torch.randn(M, L, generator=g))
        ss.append(s[M: M + L])
    return torch.stack(xs), torch.stack(ss)


def si_sdr(s, s_hat) -> float:
    """metrics.py:71-75 restated (numpy in, dB out)."""
    import numpy as np
    s, s_hat = np.asarray(s, dtype=np.float64), np.asarray(s_hat, dtype=np.float64)
    a = np.dot(s_hat, s) / np.dot(s, s)
    return float(10 * np.log10(np.sum((a * s) ** 2) / np.sum((a * s - s_hat) ** 2)))
