"""Drop-in `EaBNet(nn.Module)`: the reference's constructor, attributes, state_dict and forward contract
(EaBNet.py:9-125) over the hand-written sm_100a kernels in libeabnet_b200.so.

The module holds ordinary nn.Parameters under exactly the reference's names (so `load_state_dict`, `.to()`,
`.state_dict()`, DDP wrapping ... behave), and a native handle that owns the packed copies the kernels read.
Packed weights are refreshed whenever a parameter's version counter or storage changes.  Inference only:
the kernels have no backward, and `forward` raises if autograd is recording on any input/parameter.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from . import _lib

_BF = {"lstm": 0, "cnn": 1}
_TOPO = {"mimo": 0, "miso": 1}
_INTRA = {"cat": 0, "add": 1}
_NORM = {"IN": 0, "BN": 1}
N_FREQ = 161


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


class _Native:
    """Owns an eab_model* and frees it with the module."""

    def __init__(self, cfg, creator: str = "eab_create"):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        _lib.check(getattr(self.lib, creator)(C.byref(cfg), C.byref(self.h)), creator)

    def __del__(self):
        try:
            if self.h:
                self.lib.eab_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def param_table(self) -> List[Tuple[str, Tuple[int, ...], int, int]]:
        out = []
        name, ndim, kind, fan = C.c_char_p(), C.c_int(), C.c_int(), C.c_int()
        shape = (C.c_int64 * 4)()
        for i in range(self.lib.eab_param_count(self.h)):
            _lib.check(self.lib.eab_param_info(self.h, i, C.byref(name), C.byref(ndim), C.byref(shape),
                                               C.byref(kind), C.byref(fan)))
            out.append((name.value.decode(), tuple(int(shape[k]) for k in range(ndim.value)), kind.value, fan.value))
        return out


class _NativeModule(nn.Module):
    """nn.Module whose parameter tree is declared by a native handle's state_dict table (shared by EaBNet and GaGNet)."""

    def _setup_native(self, native: "_Native") -> None:
        object.__setattr__(self, "_native", native)
        object.__setattr__(self, "_table", native.param_table())
        object.__setattr__(self, "_packed_key", None)
        object.__setattr__(self, "_ws", None)
        object.__setattr__(self, "_pack_device", None)
        for name, shape, kind, fan in self._table:
            self._register(name, self._init_tensor(shape, kind, fan), is_buffer=kind in (8, 9, 10))

    @property
    def _is_bn(self) -> bool:
        return getattr(self, "norm_type", "IN") == "BN"

    # ---------------------------------------------------------------- parameter tree
    @staticmethod
    def _init_tensor(shape, kind, fan) -> torch.Tensor:
        """torch's default initialisers for the layer kinds the reference uses."""
        if kind == 10:
            return torch.tensor(0, dtype=torch.long)
        t = torch.empty(shape, dtype=torch.float32)
        bound = 1.0 / math.sqrt(max(fan, 1))
        if kind in (0, 1, 5, 6, 7):
            t.uniform_(-bound, bound)           # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +) ; LSTM 1/sqrt(H)
        elif kind in (2, 9):
            t.fill_(1.0)
        elif kind in (3, 8):
            t.zero_()
        elif kind == 4:
            t.fill_(0.25)
        return t

    def _register(self, dotted: str, value: torch.Tensor, is_buffer: bool) -> None:
        parts = dotted.split(".")
        mod: nn.Module = self
        for part in parts[:-1]:
            if part not in mod._modules:
                mod.add_module(part, nn.Module())
            mod = mod._modules[part]
        if is_buffer:
            mod.register_buffer(parts[-1], value)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(value))

    def _tensor_of(self, dotted: str) -> torch.Tensor:
        mod: nn.Module = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            mod = mod._modules[part]
        t = mod._parameters.get(parts[-1])
        return t if t is not None else mod._buffers[parts[-1]]

    # ---------------------------------------------------------------- packing
    def _sync_params(self, device: torch.device) -> None:
        tensors = [self._tensor_of(n) for n, _, _, _ in self._table]
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self._packed_key:
            return
        lib, h = self._native.lib, self._native.h
        keep = []
        for (name, shape, kind, _), t in zip(self._table, tensors):
            if kind == 10:
                _lib.check(lib.eab_set_param(h, name.encode(), None, 0))
                continue
            if t.dtype != torch.float32:
                raise TypeError("eabnet_b200 computes in fp32; parameter %s is %s" % (name, t.dtype))
            ht = t.detach().to("cpu").contiguous()
            keep.append(ht)
            _lib.check(lib.eab_set_param(h, name.encode(), _ptr(ht), ht.numel()), name)
        with torch.cuda.device(device):
            _lib.check(lib.eab_commit_params(h, torch.cuda.current_stream(device).cuda_stream), "commit")
        object.__setattr__(self, "_packed_key", key)
        object.__setattr__(self, "_pack_device", device)

    def _workspace(self, nbytes: int, device: torch.device) -> torch.Tensor:
        ws = self._ws
        if ws is None or ws.device != device or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            object.__setattr__(self, "_ws", ws)
        return ws

    def _check_input(self, x: torch.Tensor, what: str) -> None:
        if not x.is_cuda:
            raise RuntimeError("eabnet_b200 runs on CUDA (sm_100a) only; %s is on %s - there is no CPU fallback"
                               % (what, x.device))
        if x.dtype != torch.float32:
            raise TypeError("%s must be float32" % what)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("eabnet_b200 is an inference path (no backward kernels): call it under "
                               "torch.no_grad() / torch.inference_mode()")
        if self.norm_type == "BN" and self.training:
            raise RuntimeError("norm_type='BN' is supported in eval() mode only (running statistics)")

    # ---------------------------------------------------------------- introspection for tests / bench
    def last_launch_count(self) -> int:
        return int(self._native.lib.eab_last_launch_count(self._native.h))

    def set_option(self, name: str, value: int) -> None:
        """kernel-selection / precision knobs of the native path (see include/eabnet_b200.h: eab_set_option)"""
        _lib.check(self._native.lib.eab_set_option(self._native.h, name.encode(), int(value)), "eab_set_option")
        self._opt_epoch = getattr(self, "_opt_epoch", 0) + 1          # streaming sessions re-plan their state after an option change

    def profile(self, on) -> None:
        """switch per-launch CUDA-event timing on/off for this thread's launches (2 = one entry per launch)"""
        _lib.check(self._native.lib.eab_profile_enable(self._native.h, int(on)))

    def profile_summary(self) -> list:
        """[{kernel, launches, ms, flops, bytes}] since the last call (synchronises the device)"""
        import json
        buf = C.create_string_buffer(1 << 18)
        n = self._native.lib.eab_profile_summary(self._native.h, buf, len(buf))
        if n < 0:
            _lib.check(1, "eab_profile_summary")
        return json.loads(buf.value.decode())

    def debug_counters(self) -> list:
        """16 cycle counters written by the kernel launch selected with option dbg_launch (diagnostics)"""
        buf = (C.c_uint64 * 16)()
        _lib.check(self._native.lib.eab_debug_counters(self._native.h, C.byref(buf)), "eab_debug_counters")
        return [int(v) for v in buf]

    def debug_tap(self, name: str, shape: Tuple[int, ...]) -> torch.Tensor:
        """Named intermediate of the last forward (normalised + activated), channels-last [B,T,F',C']."""
        dev = self._pack_device
        dst = torch.empty(shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            n = self._native.lib.eab_debug_tap(self._native.h, name.encode(), _ptr(dst), dst.numel(),
                                               torch.cuda.current_stream(dev).cuda_stream)
        if n != dst.numel():
            raise RuntimeError("debug_tap(%s): got %d elements, expected %d (%s)" % (
                name, n, dst.numel(), (self._native.lib.eab_last_error() or b"").decode()))
        return dst


class EaBNet(_NativeModule):
    def __init__(self, k1: tuple = (2, 3), k2: tuple = (1, 3), c: int = 64, M: int = 9, embed_dim: int = 64,
                 kd1: int = 5, cd1: int = 64, d_feat: int = 256, p: int = 6, q: int = 3, is_causal: bool = True,
                 is_u2: bool = True, bf_type: str = "lstm", topo_type: str = "mimo", intra_connect: str = "cat",
                 norm_type: str = "IN"):
        super().__init__()
        self.k1, self.k2, self.c, self.M, self.embed_dim = tuple(k1), tuple(k2), c, M, embed_dim
        self.kd1, self.cd1, self.d_feat, self.p, self.q = kd1, cd1, d_feat, p, q
        self.is_causal, self.is_u2, self.bf_type = is_causal, is_u2, bf_type
        self.intra_connect, self.topo_type, self.norm_type = intra_connect, topo_type, norm_type
        if norm_type == "cLN":
            # the reference raises a TypeError here as well (NormSwitch passes a string as num_features)
            raise TypeError("norm_type 'cLN' cannot be constructed (EaBNet.py:689,691)")
        for val, table, what in ((bf_type, _BF, "bf_type"), (topo_type, _TOPO, "topo_type"),
                                 (intra_connect, _INTRA, "intra_connect"), (norm_type, _NORM, "norm_type")):
            if val not in table:
                raise ValueError("unknown %s %r" % (what, val))
        cfg = _lib.EabConfig(self.k1[0], self.k1[1], self.k2[0], self.k2[1], c, M, embed_dim, kd1, cd1, d_feat, p, q,
                             int(bool(is_causal)), int(bool(is_u2)), _BF[bf_type], _TOPO[topo_type],
                             _INTRA[intra_connect], _NORM[norm_type], N_FREQ)
        self._setup_native(_Native(cfg))

    # ---------------------------------------------------------------- the reference contract
    def forward(self, inpt: torch.Tensor) -> torch.Tensor:
        """inpt [B,T,F,M,2] (or [B,T,F,2] for M = 1) -> [B,2,T,F]   (EaBNet.py:88-125)."""
        if inpt.ndim == 4:
            inpt = inpt.unsqueeze(-2)
        if inpt.ndim != 5 or inpt.shape[-1] != 2:
            raise ValueError("expected [B,T,F,M,2], got %s" % (tuple(inpt.shape),))
        self._check_input(inpt, "inpt")
        B, T, Fq, M, _ = inpt.shape
        if Fq != N_FREQ or M != self.M:
            raise RuntimeError("expected F=%d and M=%d, got F=%d, M=%d" % (N_FREQ, self.M, Fq, M))
        dev = inpt.device
        x = inpt.contiguous()
        with torch.cuda.device(dev):
            self._sync_params(dev)
            lib, h = self._native.lib, self._native.h
            nbytes = lib.eab_workspace_bytes(h, B, T)
            if nbytes == 0:
                _lib.check(1, "eab_workspace_bytes")
            ws = self._workspace(nbytes, dev)
            out = torch.empty((B, 2, T) if self.topo_type == "miso" else (B, 2, T, Fq), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.eab_forward(h, _ptr(x), _ptr(out), B, T, _ptr(ws), ws.numel(), stream), "eab_forward")
        return out

    # ---------------------------------------------------------------- streaming an InstanceNorm-trained model
    def to_batchnorm(self, specs) -> "EaBNet":
        """An InstanceNorm model cannot be stepped causally (its statistics span the utterance, EaBNet.py:45-48, 684-686).
        This returns the same network with norm_type="BN" - the same weights, and as running_mean / running_var of every norm
        the statistics its InstanceNorm saw on the calibration spectra `specs` (an iterable of [B,T,F,M,2] CUDA tensors, all
        positions pooled, biased variance like InstanceNorm's own).  Its state_dict is what the reference's
        EaBNet(norm_type="BN") loads; it streams (`.stream(n)`).  On a single calibration utterance the two models agree."""
        if self.norm_type != "IN":
            raise RuntimeError("to_batchnorm: the model already has static normalisation (norm_type=%r)" % self.norm_type)
        lib, h = self._native.lib, self._native.h
        sums, counts = {}, {}
        self.set_option("tcm_chain", 0)         # the chain kernel keeps the TCM statistics to itself
        self.set_option("norm_log", 1)
        try:
            with torch.no_grad():
                for spec in specs:
                    self.forward(spec)
                    st = torch.cuda.current_stream(spec.device).cuda_stream
                    for i in range(lib.eab_norm_stats_count(h)):
                        name, Cn, cnt = C.c_char_p(), C.c_int(), C.c_int64()
                        _lib.check(lib.eab_norm_stats(h, i, C.byref(name), C.byref(Cn), C.byref(cnt), None, st), "eab_norm_stats")
                        buf = torch.empty((Cn.value, 2), dtype=torch.float64)
                        _lib.check(lib.eab_norm_stats(h, i, None, None, None, buf.data_ptr(), st), "eab_norm_stats")
                        key = name.value.decode()
                        sums[key] = sums.get(key, 0) + buf
                        counts[key] = counts.get(key, 0) + cnt.value
        finally:
            self.set_option("norm_log", 0)
            self.set_option("tcm_chain", 1)
        if not sums:
            raise RuntimeError("to_batchnorm: no calibration data")
        bn = EaBNet(k1=self.k1, k2=self.k2, c=self.c, M=self.M, embed_dim=self.embed_dim, kd1=self.kd1, cd1=self.cd1,
                    d_feat=self.d_feat, p=self.p, q=self.q, is_causal=self.is_causal, is_u2=self.is_u2, bf_type=self.bf_type,
                    topo_type=self.topo_type, intra_connect=self.intra_connect, norm_type="BN").eval()
        sd = bn.state_dict()
        for k, v in self.state_dict().items():
            sd[k] = v.detach().clone()
        for key, sm in sums.items():
            assert key.endswith(".weight"), key
            mean = sm[:, 0] / counts[key]
            var = (sm[:, 1] / counts[key] - mean * mean).clamp_min(0)
            sd[key[:-len("weight")] + "running_mean"] = mean.float()
            sd[key[:-len("weight")] + "running_var"] = var.float()
        missing = [k for k in sd if k.endswith("running_mean") and k[:-len("running_mean")] + "weight" not in sums]
        if missing:
            raise RuntimeError("to_batchnorm: no statistics recorded for %s" % missing[:3])
        bn.load_state_dict(sd, strict=True)
        dev = next(self.parameters()).device
        return bn.to(dev)

    # ---------------------------------------------------------------- wave-to-wave (test.py:178-190)
    def enhance(self, wave: torch.Tensor, workspace: torch.Tensor | None = None) -> torch.Tensor:
        """wave [B,M,L] on the GPU -> enhanced [B,160*(L//160)]: STFT+compression, forward, iSTFT in one call.
        `workspace`: a caller-owned uint8 CUDA tensor of at least eab_enhance_workspace_bytes (calls that may run
        concurrently on different streams must not share the module's own workspace)."""
        self._check_input(wave, "wave")
        B, M, L = wave.shape
        if M != self.M:
            raise RuntimeError("expected %d microphones, got %d" % (self.M, M))
        dev = wave.device
        x = wave.contiguous()
        with torch.cuda.device(dev):
            self._sync_params(dev)
            lib, h = self._native.lib, self._native.h
            nbytes = lib.eab_enhance_workspace_bytes(h, B, L)
            if nbytes == 0:
                _lib.check(1, "eab_enhance_workspace_bytes")
            if workspace is not None:
                if workspace.device != dev or workspace.dtype != torch.uint8 or workspace.numel() < nbytes:
                    raise ValueError("workspace must be a uint8 tensor of at least %d bytes on %s" % (nbytes, dev))
                ws = workspace
            else:
                ws = self._workspace(nbytes, dev)
            out = torch.empty((B, 160 * (L // 160)), dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.eab_enhance(h, _ptr(x), _ptr(out), B, L, _ptr(ws), ws.numel(), stream), "eab_enhance")
        return out

    def enhance_host(self, wave: torch.Tensor, out: torch.Tensor | None = None,
                     device: torch.device | str = "cuda") -> torch.Tensor:
        """HOST wave [B,M,L] (ideally pinned) -> HOST enhanced wave; copies both ways inside the call."""
        if wave.is_cuda or wave.dtype != torch.float32:
            raise TypeError("enhance_host takes a float32 CPU tensor")
        B, M, L = wave.shape
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        x = wave.contiguous()
        if out is None:
            out = torch.empty((B, 160 * (L // 160)), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(dev):
            self._sync_params(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._native.lib.eab_enhance_host(self._native.h, _ptr(x), _ptr(out), B, L, stream),
                       "eab_enhance_host")
        return out

    def enhance_host_batches(self, waves, outs=None, device: torch.device | str = "cuda", mic_order=None):
        """A list of HOST batches [B,M,L] (same shape, ideally pinned) -> list of HOST enhanced batches.  Uploads,
        compute and downloads of consecutive batches overlap (eab_enhance_host_batches).  float32 batches give float32
        results; int16 batches (the PCM wire format: sample / 32768, microphone m = file channel mic_order[m]) give int16
        results through eab_enhance_host_batches_pcm16 at half the copy bytes."""
        if not waves:
            return []
        B, M, L = waves[0].shape
        dt = waves[0].dtype
        if dt not in (torch.float32, torch.int16):
            raise TypeError("enhance_host_batches takes float32 or int16 batches")
        for w in waves:
            if w.is_cuda or w.dtype != dt or tuple(w.shape) != (B, M, L) or not w.is_contiguous():
                raise TypeError("enhance_host_batches takes contiguous CPU tensors of one shape and dtype (float32 or int16)")
        if M != self.M:
            raise RuntimeError("expected %d microphones, got %d" % (self.M, M))
        if mic_order is not None and dt != torch.int16:
            raise TypeError("mic_order applies to int16 PCM batches")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if outs is None:
            outs = [torch.empty((B, 160 * (L // 160)), dtype=dt, pin_memory=True) for _ in waves]
        for o in outs:
            if o.is_cuda or o.dtype != dt or tuple(o.shape) != (B, 160 * (L // 160)) or not o.is_contiguous():
                raise TypeError("outs must be contiguous CPU tensors [B, 160 * (L // 160)] of the input dtype")
        n = len(waves)
        wp = (C.c_void_p * n)(*[_ptr(w) for w in waves])
        op = (C.c_void_p * n)(*[_ptr(o) for o in outs])
        with torch.cuda.device(dev):
            self._sync_params(dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            if dt == torch.int16:
                order = None
                if mic_order is not None:
                    if len(mic_order) != M:
                        raise ValueError("mic_order needs %d entries" % M)
                    order = (C.c_int * M)(*[int(v) for v in mic_order])
                _lib.check(self._native.lib.eab_enhance_host_batches_pcm16(self._native.h, wp, order, op, n, B, L, stream),
                           "eab_enhance_host_batches_pcm16")
            else:
                _lib.check(self._native.lib.eab_enhance_host_batches(self._native.h, wp, op, n, B, L, stream),
                           "eab_enhance_host_batches")
        return outs

    def enhance_pcm16(self, pcm: torch.Tensor, mic_order=None, out: torch.Tensor | None = None, postnet=None, ref_mic: int = 0,
                      device: torch.device | str = "cuda") -> torch.Tensor:
        """HOST int16 PCM [B,M,L] (file channel order) -> HOST int16 enhanced [B,160*(L//160)]: torchaudio.load's / 32768,
        enhance.py:41-42's microphone permutation, the network (+ `postnet`, a GaGNet), and the dataset tools' int16 writer, as one
        native call (eab_enhance_host_pcm16)."""
        if pcm.is_cuda or pcm.dtype != torch.int16 or pcm.ndim != 3:
            raise TypeError("enhance_pcm16 takes an int16 CPU tensor [B,M,L]")
        B, M, L = pcm.shape
        if M != self.M:
            raise RuntimeError("expected %d microphones, got %d" % (self.M, M))
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        x = pcm.contiguous()
        if out is None:
            out = torch.empty((B, 160 * (L // 160)), dtype=torch.int16, pin_memory=True)
        order = None
        if mic_order is not None:
            if len(mic_order) != M:
                raise ValueError("mic_order needs %d entries" % M)
            order = (C.c_int * M)(*[int(v) for v in mic_order])
        with torch.cuda.device(dev):
            self._sync_params(dev)
            gh = None
            if postnet is not None:
                postnet._sync_params(dev)
                gh = postnet._native.h
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(self._native.lib.eab_enhance_host_pcm16(self._native.h, gh, int(ref_mic), _ptr(x), order, _ptr(out), B, L, stream),
                       "eab_enhance_host_pcm16")
        return out

    def graphed_enhance(self, wave: torch.Tensor, private_workspace: bool = False) -> "GraphedEnhance":
        """The wave -> wave step on a fixed device buffer captured once into a CUDA graph (one replay = the kernel launches
        of eab_enhance without their host-side enqueue cost).  Re-capture after loading new weights.  With
        private_workspace=True the graph owns its workspace, so several graphs may replay concurrently on different streams."""
        return GraphedEnhance(self, wave, private_workspace)

    def stream(self, n_streams: int, device: torch.device | str | None = None) -> "EaBNetStream":
        """Carried-state, frame-by-frame inference for `n_streams` concurrent causal streams (eab_stream_*)."""
        return EaBNetStream(self, n_streams, device)


class GraphedEnhance:
    """eab_enhance on a fixed input buffer as a CUDA graph: `step()` replays it and returns the (static) output tensor.
    The input tensor is read in place, so new audio is enhanced by copying it into `wave` before the replay."""

    def __init__(self, net, wave: torch.Tensor, private_workspace: bool = False):
        # `net`: anything with enhance(wave) and last_launch_count() - EaBNet, or EaBNetWithPostNet (beamformer + post-filter)
        self.net, self.wave = net, wave
        dev = wave.device
        kw = {}
        if private_workspace:                              # EaBNet only: a workspace of its own (concurrent replays)
            B, _, L = wave.shape
            with torch.cuda.device(dev):
                net._sync_params(dev)                      # the plan (and its size) depends on the packed weights
            nbytes = net._native.lib.eab_enhance_workspace_bytes(net._native.h, B, L)
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            kw["workspace"] = self.workspace
        with torch.cuda.device(dev), torch.no_grad():
            net.enhance(wave, **kw)                        # packs weights, sizes the workspace, configures the kernels
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = net.enhance(wave, **kw)
        self.launches = net.last_launch_count()

    def step(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


class EaBNetStream:
    """`n_streams` concurrent causal streams stepped one 10 ms hop at a time (BASELINE configs[2]).

    step(hop [S,M,160])            -> enhanced hop [S,160], delayed by one hop (overlap-add needs the next frame)
    step_spec(frame [S,F,M,2])     -> [S,2,F], the column EaBNet.forward would produce for this frame
    Frame n of the result equals frame n of the offline forward on the whole signal (is_causal=True, norm_type='BN').
    With graph=True the whole step (stft frame, ~270 layer kernels, istft frame) is captured once into a CUDA graph
    and replayed: the frame counter lives in device memory, so the captured launches never change."""

    def __init__(self, net: EaBNet, n_streams: int, device=None, graph: bool = False, postnet=None, ref_mic: int = 0):
        """`postnet`: a GaGNet (is_causal, norm_type "BN") run behind the beamformer on every hop (enhance.py:49-62 as a stream,
        eab_stream_step_postnet); `ref_mic` is the microphone whose spectrum it filters (EaBNet.py:141)."""
        dev = torch.device(device) if device is not None else next(net.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("eabnet_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.net, self.S, self.dev, self.use_graph = net, int(n_streams), dev, graph
        lib, h = net._native.lib, net._native.h
        with torch.cuda.device(dev):
            net._sync_params(dev)                # the state layout depends on the committed weights (tensor-core images)
            if postnet is not None:
                postnet._sync_params(dev)
        nbytes = lib.eab_stream_state_bytes(h, self.S)
        if nbytes == 0:
            _lib.check(1, "eab_stream_state_bytes")
        self.state = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.postnet, self.ref_mic, self.gstate = postnet, int(ref_mic), None
        if postnet is not None:
            gbytes = lib.eab_stream_state_bytes(postnet._native.h, self.S)
            if gbytes == 0:
                _lib.check(1, "eab_stream_state_bytes (post-filter)")
            self.gstate = torch.empty(gbytes, dtype=torch.uint8, device=dev)
        self._graph = None
        self._hop_in = self._hop_out = None
        self.reset()

    def _states(self):
        yield self.net, self.state
        if self.postnet is not None:
            yield self.postnet, self.gstate

    def _epoch(self):
        return (getattr(self.net, "_opt_epoch", 0), getattr(self.postnet, "_opt_epoch", 0) if self.postnet is not None else 0)

    def reset(self) -> None:
        """All streams start over.  Also re-plans the state blobs: kernel-selection options (stream_umma, stream_lstm, ...) change
        their layout, so an option set after the session was created takes effect here (a step in between raises)."""
        with torch.cuda.device(self.dev):
            lib = self.net._native.lib
            need = lib.eab_stream_state_bytes(self.net._native.h, self.S)
            if need == 0:
                _lib.check(1, "eab_stream_state_bytes")
            if need != self.state.numel():
                self.state = torch.empty(need, dtype=torch.uint8, device=self.dev)
                self._graph = None
            if self.postnet is not None:
                gneed = lib.eab_stream_state_bytes(self.postnet._native.h, self.S)
                if gneed == 0:
                    _lib.check(1, "eab_stream_state_bytes (post-filter)")
                if gneed != self.gstate.numel():
                    self.gstate = torch.empty(gneed, dtype=torch.uint8, device=self.dev)
                    self._graph = None
            if getattr(self, "_planned", None) != self._epoch():
                self._graph = None                    # a captured step bakes the old kernel selection in
            self._planned = self._epoch()
            st = torch.cuda.current_stream(self.dev).cuda_stream
            for mod, state in self._states():
                mod._sync_params(self.dev)
                _lib.check(mod._native.lib.eab_stream_reset(mod._native.h, _ptr(state), state.numel(), self.S, st), "eab_stream_reset")

    def reset_stream(self, idx: int) -> None:
        """Stream `idx` leaves and a new one joins in its slot: it starts over at its frame 0 from the next step on, the other
        streams carry on bit-identically (eab_stream_reset_one; stream-ordered, valid between graph replays too)."""
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            for mod, state in self._states():
                _lib.check(mod._native.lib.eab_stream_reset_one(mod._native.h, _ptr(state), state.numel(), self.S, int(idx), st),
                           "eab_stream_reset_one")

    def _launch(self, hop: torch.Tensor, out: torch.Tensor) -> None:
        st = torch.cuda.current_stream(self.dev).cuda_stream
        lib, h = self.net._native.lib, self.net._native.h
        if self.postnet is not None:
            fn = lib.eab_stream_step_postnet_pcm16 if hop.dtype == torch.int16 else lib.eab_stream_step_postnet
            _lib.check(fn(h, _ptr(self.state), self.state.numel(), self.postnet._native.h, _ptr(self.gstate), self.gstate.numel(),
                          self.ref_mic, _ptr(hop), _ptr(out), self.S, st), "eab_stream_step_postnet")
        elif hop.dtype == torch.int16:
            _lib.check(lib.eab_stream_step_pcm16(h, _ptr(self.state), self.state.numel(), _ptr(hop), _ptr(out), self.S, st),
                       "eab_stream_step_pcm16")
        else:
            _lib.check(lib.eab_stream_step(h, _ptr(self.state), self.state.numel(), _ptr(hop), _ptr(out), self.S, st),
                       "eab_stream_step")

    def step(self, hop: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """One 10 ms hop for every stream: float32 [S,M,160] -> float32 [S,160], or the 16-bit PCM wire format
        int16 [S,M,160] -> int16 [S,160] (eab_stream_step_pcm16)."""
        if tuple(hop.shape) != (self.S, self.net.M, 160) or hop.dtype not in (torch.float32, torch.int16) or hop.device != self.dev:
            raise ValueError("expected a float32 or int16 [%d,%d,160] tensor on %s" % (self.S, self.net.M, self.dev))
        if self._planned != self._epoch():
            raise RuntimeError("kernel options changed since this streaming session was planned: call reset() first")
        hop = hop.contiguous()
        if out is None:
            out = torch.empty((self.S, 160), dtype=hop.dtype, device=self.dev)
        elif out.dtype != hop.dtype:
            raise ValueError("out must have the hop's dtype")
        with torch.cuda.device(self.dev):
            if not self.use_graph:
                self._launch(hop, out)
                return out
            if self._graph is None or self._hop_in.dtype != hop.dtype:
                self._hop_in, self._hop_out = torch.empty_like(hop), torch.empty_like(out)
                self.net._sync_params(self.dev)
                # one eager step first: first-use work of the library (constant tables, shared-memory attributes) is not
                # capturable; the carried state it advanced is put back before the capture
                snapshot = [state.clone() for _, state in self._states()]
                self._hop_in.zero_()
                self._launch(self._hop_in, self._hop_out)
                for (_, state), snap in zip(self._states(), snapshot):
                    state.copy_(snap)
                del snapshot
                if self.postnet is not None:
                    self.postnet._sync_params(self.dev)
                torch.cuda.current_stream(self.dev).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch(self._hop_in, self._hop_out)
                self._graph = g
            self._hop_in.copy_(hop, non_blocking=True)
            self._graph.replay()
            out.copy_(self._hop_out, non_blocking=True)
        return out

    def step_spec(self, frame: torch.Tensor) -> torch.Tensor:
        net = self.net
        if frame.ndim == 3:
            frame = frame.unsqueeze(-2)
        if tuple(frame.shape) != (self.S, N_FREQ, net.M, 2) or frame.dtype != torch.float32 or frame.device != self.dev:
            raise ValueError("expected a float32 [%d,%d,%d,2] tensor on %s" % (self.S, N_FREQ, net.M, self.dev))
        if self._planned != self._epoch():
            raise RuntimeError("kernel options changed since this streaming session was planned: call reset() first")
        frame = frame.contiguous()
        out = torch.empty((self.S, 2) if net.topo_type == "miso" else (self.S, 2, N_FREQ), dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            _lib.check(net._native.lib.eab_stream_step_spec(net._native.h, _ptr(self.state), self.state.numel(),
                                                            _ptr(frame), _ptr(out), self.S, st), "eab_stream_step_spec")
        return out


def numParams(net: nn.Module) -> int:
    """EaBNet.py:653-659."""
    return sum(int(p.numel()) for p in net.parameters() if p.requires_grad)
