"""Drop-in `GaGNet`, `EaBNetWithPostNet`, `make_gag_net` and `make_eabnet_with_postnet`:
what `enhance.py:21-22,49-52` and `train_distributed.py:181` build and call, over the same sm_100a kernels.

    GaGNet(cin, k1, k2, c, kd1, cd1, d_feat, p, q, dilas, fft_num, is_u2, is_causal, is_squeezed, acti_type,
           intra_connect, norm_type).forward(inpt [B,2,T,F], pre_x [B,2,T,F]) -> list of q tensors [B,2,F,T]   (GaGNet.py:5-89)
    EaBNetWithPostNet(args).forward(noisy_stft [B,T,F,M,2]) -> {"esti0_stft", "esti1_stft_list", "esti_stft"}  (EaBNet.py:127-148)

State-dict keys are the reference's (`en.*`, `gags.*`; wrapper: `eabnet.*`, `postnet.*`), so the checkpoints enhance.py
loads with `load_state_dict` fit.  Inference only, CUDA only, no fallback - as eabnet_b200.EaBNet.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn

from . import _lib
from .model import EaBNet, _Native, _NativeModule, _ptr

_ACTI = {"sigmoid": 0, "tanh": 1, "relu": 2}
_INTRA = {"cat": 0, "add": 1}
_NORM = {"IN": 0, "BN": 1}


class GaGNet(_NativeModule):
    def __init__(self, cin: int = 2, k1: tuple = (2, 3), k2: tuple = (1, 3), c: int = 64, kd1: int = 3, cd1: int = 64,
                 d_feat: int = 256, p: int = 2, q: int = 3, dilas: list = (1, 2, 5, 9), fft_num: int = 320,
                 is_u2: bool = True, is_causal: bool = True, is_squeezed: bool = False, acti_type: str = "sigmoid",
                 intra_connect: str = "cat", norm_type: str = "IN"):
        super().__init__()
        self.cin, self.k1, self.k2, self.c, self.kd1, self.cd1 = cin, tuple(k1), tuple(k2), c, kd1, cd1
        self.d_feat, self.p, self.q, self.dilas, self.fft_num = d_feat, p, q, list(dilas), fft_num
        self.is_u2, self.is_causal, self.is_squeezed = is_u2, is_causal, is_squeezed
        self.acti_type, self.intra_connect, self.norm_type = acti_type, intra_connect, norm_type
        if acti_type not in _ACTI:
            raise RuntimeError("a activation function must be assigned!")          # GaGNet.py:171-172
        if intra_connect not in _INTRA or norm_type not in _NORM:
            raise ValueError("unknown intra_connect / norm_type")
        if len(self.dilas) > 8:
            raise ValueError("at most 8 dilation rates")
        cfg = _lib.EabGagConfig()
        cfg.cin, cfg.k1_t, cfg.k1_f, cfg.k2_t, cfg.k2_f = cin, self.k1[0], self.k1[1], self.k2[0], self.k2[1]
        cfg.c, cfg.kd1, cfg.cd1, cfg.d_feat, cfg.p, cfg.q = c, kd1, cd1, d_feat, p, q
        cfg.n_dilas = len(self.dilas)
        for i, d in enumerate(self.dilas):
            cfg.dilas[i] = int(d)
        cfg.fft_num, cfg.is_u2, cfg.is_causal, cfg.is_squeezed = fft_num, int(bool(is_u2)), int(bool(is_causal)), int(bool(is_squeezed))
        cfg.acti_type, cfg.intra_connect, cfg.norm_type = _ACTI[acti_type], _INTRA[intra_connect], _NORM[norm_type]
        self._setup_native(_Native(cfg, "eab_gag_create"))

    def forward_time_major(self, inpt: torch.Tensor, pre_x: torch.Tensor) -> torch.Tensor:
        """Same computation, all q estimates as one [q,B,2,T,F] tensor (the layout the kernels write)."""
        if inpt.ndim != 4 or inpt.shape[1] != 2 or tuple(pre_x.shape) != tuple(inpt.shape):
            raise ValueError("expected inpt and pre_x of shape [B,2,T,F], got %s and %s" % (tuple(inpt.shape), tuple(pre_x.shape)))
        self._check_input(inpt, "inpt")
        self._check_input(pre_x, "pre_x")
        B, _, T, Fq = inpt.shape
        if Fq != self.fft_num // 2 + 1:
            raise RuntimeError("expected F=%d, got %d" % (self.fft_num // 2 + 1, Fq))
        dev = inpt.device
        pre = pre_x.contiguous()
        strides = (C.c_int64 * 4)(*[int(v) for v in inpt.stride()])          # any view is read in place (e.g. the ref-mic slice)
        with torch.cuda.device(dev):
            self._sync_params(dev)
            lib, h = self._native.lib, self._native.h
            nbytes = lib.eab_gag_workspace_bytes(h, B, T)
            if nbytes == 0:
                _lib.check(1, "eab_gag_workspace_bytes")
            ws = self._workspace(nbytes, dev)
            out = torch.empty((self.q, B, 2, T, Fq), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.eab_gag_forward(h, _ptr(inpt), C.byref(strides), _ptr(pre), _ptr(out), B, T, _ptr(ws), ws.numel(),
                                           stream), "eab_gag_forward")
        return out

    def forward(self, inpt: torch.Tensor, pre_x: torch.Tensor) -> List[torch.Tensor]:
        """inpt, pre_x [B,2,T,F] -> list of q estimates, each [B,2,F,T]   (GaGNet.py:75-89)."""
        if inpt.ndim == 3:
            raise RuntimeError("GaGNet with single-plane (cin=1) inputs fails in the reference's glance block as well "
                               "(GaGNet.py:160,189)")
        out = self.forward_time_major(inpt, pre_x)
        return [out[i].transpose(-2, -1) for i in range(self.q)]

    def stream(self, n_streams: int, device=None) -> "GaGNetStream":
        return GaGNetStream(self, n_streams, device)


class GaGNetStream:
    """GaGNet.forward one frame at a time for S concurrent streams (is_causal, norm_type "BN"): eab_gag_stream_step_spec."""

    def __init__(self, net: GaGNet, n_streams: int, device=None):
        dev = torch.device(device) if device is not None else next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("eabnet_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.net, self.S, self.dev = net, int(n_streams), dev
        lib, h = net._native.lib, net._native.h
        with torch.cuda.device(dev):
            net._sync_params(dev)
        nbytes = lib.eab_stream_state_bytes(h, self.S)
        if nbytes == 0:
            _lib.check(1, "eab_stream_state_bytes")
        self.state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.reset()

    def reset(self) -> None:
        with torch.cuda.device(self.dev):
            self.net._sync_params(self.dev)
            st = torch.cuda.current_stream(self.dev).cuda_stream
            _lib.check(self.net._native.lib.eab_stream_reset(self.net._native.h, _ptr(self.state), self.state.numel(), self.S, st),
                       "eab_stream_reset")

    def step_spec(self, inpt_frame: torch.Tensor, pre_frame: torch.Tensor) -> torch.Tensor:
        """inpt_frame, pre_frame [S,2,F] -> the q modules' estimates of this frame [q,S,2,F]"""
        Fq = self.net.fft_num // 2 + 1
        for t in (inpt_frame, pre_frame):
            if tuple(t.shape) != (self.S, 2, Fq) or t.dtype != torch.float32 or t.device != self.dev:
                raise ValueError("expected float32 [%d,2,%d] tensors on %s" % (self.S, Fq, self.dev))
        a, b = inpt_frame.contiguous(), pre_frame.contiguous()
        out = torch.empty((self.net.q, self.S, 2, Fq), dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            _lib.check(self.net._native.lib.eab_gag_stream_step_spec(self.net._native.h, _ptr(self.state), self.state.numel(), _ptr(a),
                                                                     _ptr(b), _ptr(out), self.S, st), "eab_gag_stream_step_spec")
        return out


def make_gag_net(args) -> GaGNet:
    """GaGNet.py:645-665 (minus the `.cuda()`: move the wrapper with `.to(device)` like enhance.py:22 does)."""
    return GaGNet(cin=2, k1=args.gagnet_k1, k2=args.gagnet_k2, c=args.gagnet_c, kd1=args.gagnet_kd1, cd1=args.gagnet_cd1,
                  d_feat=args.gagnet_d_feat, p=args.gagnet_p, q=args.gagnet_q, dilas=args.gagnet_dilas,
                  fft_num=args.gagnet_fft_num, is_u2=args.gagnet_is_u2, is_causal=args.gagnet_is_causal,
                  is_squeezed=args.gagnet_is_squeezed, acti_type=args.gagnet_acti_type,
                  intra_connect=args.gagnet_intra_connect, norm_type=args.gagnet_norm_type)


class EaBNetWithPostNet(nn.Module):
    """EaBNet.py:127-154: beamformer, then the GaGNet post-filter on (reference microphone, beamformer estimate)."""

    def __init__(self, args):
        super().__init__()
        self.eabnet = EaBNet(k1=args.k1, k2=args.k2, c=args.c, M=args.M, embed_dim=args.embed_dim, kd1=args.kd1,
                             cd1=args.cd1, d_feat=args.d_feat, p=args.p, q=args.q, is_causal=args.is_causal,
                             is_u2=args.is_u2, bf_type=args.bf_type, topo_type=args.topo_type,
                             intra_connect=args.intra_connect, norm_type=args.norm_type)
        self.ref_mic = args.ref_mic
        self.postnet = make_gag_net(args)
        if getattr(args, "freeze_eabnet", False):
            self.freeze_eabnet()

    def forward(self, noisy_stft: torch.Tensor) -> dict:
        esti0_stft = self.eabnet(noisy_stft)
        inpt = noisy_stft[..., self.ref_mic, :].permute(0, 3, 1, 2)         # 'b t f c -> b c t f' as a view, no copy
        tm = self.postnet.forward_time_major(inpt, esti0_stft)
        return {"esti0_stft": esti0_stft,
                "esti1_stft_list": [tm[i].transpose(-2, -1) for i in range(self.postnet.q)],
                "esti_stft": tm[-1]}

    def enhance(self, wave: torch.Tensor) -> torch.Tensor:
        """wave [B,M,L] on the GPU -> enhanced [B,160*(L//160)]: the enhance.py:35-62 sequence (STFT + compression,
        beamformer, post-filter, iSTFT of `esti_stft`) as one native call (eab_enhance_postnet)."""
        e, g = self.eabnet, self.postnet
        e._check_input(wave, "wave")
        B, M, L = wave.shape
        if M != e.M:
            raise RuntimeError("expected %d microphones, got %d" % (e.M, M))
        dev = wave.device
        x = wave.contiguous()
        with torch.cuda.device(dev):
            e._sync_params(dev)
            g._sync_params(dev)
            lib = e._native.lib
            nbytes = lib.eab_enhance_postnet_workspace_bytes(e._native.h, g._native.h, B, L)
            if nbytes == 0:
                _lib.check(1, "eab_enhance_postnet_workspace_bytes")
            ws = e._workspace(nbytes, dev)
            out = torch.empty((B, 160 * (L // 160)), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.eab_enhance_postnet(e._native.h, g._native.h, int(self.ref_mic), _ptr(x), _ptr(out), B, L, _ptr(ws),
                                               ws.numel(), stream), "eab_enhance_postnet")
        return out

    def last_launch_count(self) -> int:
        return self.eabnet.last_launch_count()

    def stream(self, n_streams: int, graph: bool = False):
        """Causal hop-by-hop enhancement with the post-filter for `n_streams` concurrent streams (both networks is_causal and
        norm_type "BN"): eab_stream_step_postnet."""
        from .model import EaBNetStream
        return EaBNetStream(self.eabnet, n_streams, graph=graph, postnet=self.postnet, ref_mic=self.ref_mic)

    def graphed_enhance(self, wave: torch.Tensor):
        """`enhance` on a fixed device buffer captured once into a CUDA graph (see EaBNet.graphed_enhance)."""
        from .model import GraphedEnhance
        return GraphedEnhance(self, wave)

    def freeze_eabnet(self):
        for param in self.eabnet.parameters():
            param.requires_grad = False

    def unfreeze_eabnet(self):
        for param in self.eabnet.parameters():
            param.requires_grad = True


def make_eabnet_with_postnet(args) -> EaBNetWithPostNet:
    """EaBNet.py:815-816."""
    return EaBNetWithPostNet(args)


def default_postnet_args(**over):
    """The argparse defaults of train_distributed.py:277-318 / enhance.py as a namespace (convenience for tests / bench),
    except M: the script defaults to its 8-microphone checkpoint, this follows EaBNet.py's own default and BASELINE.json (9)."""
    import argparse
    d = dict(k1=(2, 3), k2=(1, 3), c=64, M=9, embed_dim=64, kd1=5, cd1=64, d_feat=256, p=6, q=3, is_causal=True, is_u2=True,
             bf_type="lstm", topo_type="mimo", intra_connect="cat", norm_type="IN", ref_mic=0, freeze_eabnet=False,
             gagnet_fft_num=320, gagnet_k1=(2, 3), gagnet_k2=(1, 3), gagnet_c=64, gagnet_kd1=3, gagnet_cd1=64,
             gagnet_d_feat=256, gagnet_p=2, gagnet_q=3, gagnet_dilas=[1, 2, 5, 9], gagnet_is_u2=True, gagnet_is_causal=True,
             gagnet_is_squeezed=False, gagnet_acti_type="sigmoid", gagnet_intra_connect="cat", gagnet_norm_type="IN")
    d.update(over)
    return argparse.Namespace(**d)
