"""First slice of the training step (train_distributed.py:214-230): the tail of the beamforming head and the loss as
`torch.autograd.Function`s whose forward AND backward are this library's hand-written kernels (csrc/head_bwd.cu) - no
autograd graph of torch ops, no library call.  The rest of the backward pass is not built; `EaBNet.forward` still refuses
autograd."""
from __future__ import annotations

import torch

from . import _lib


def _st(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _check(t: torch.Tensor, what: str):
    if not t.is_cuda or t.dtype != torch.float32:
        raise TypeError("%s: expected a CUDA float32 tensor (no CPU fallback)" % what)


class _HeadFilterSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h2, spec, W1, b1, W2, b2):
        for t, n in ((h2, "h2"), (spec, "spec"), (W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2")):
            _check(t, n)
        B, T, Fq, H = h2.shape
        M = spec.shape[-2]
        if H != 64 or tuple(spec.shape) != (B, T, Fq, M, 2) or tuple(W1.shape) != (64, 64) or tuple(W2.shape) != (2 * M, 64):
            raise ValueError("head_filter_sum: h2 [B,T,F,64], spec [B,T,F,M,2], W1 [64,64], W2 [2M,64]")
        h2c, sc = h2.contiguous(), spec.contiguous()
        W1c, b1c, W2c, b2c = W1.contiguous(), b1.contiguous(), W2.contiguous(), b2.contiguous()
        out = torch.empty((B, 2, T, Fq), dtype=torch.float32, device=h2.device)
        with torch.cuda.device(h2.device):
            _lib.check(_lib.load().eab_head_forward(W1c.data_ptr(), b1c.data_ptr(), W2c.data_ptr(), b2c.data_ptr(), h2c.data_ptr(),
                                                    sc.data_ptr(), out.data_ptr(), B, T, Fq, M, _st(h2.device)), "eab_head_forward")
        ctx.save_for_backward(h2c, sc, W1c, b1c, W2c, b2c)
        return out

    @staticmethod
    def backward(ctx, d_out):
        h2, spec, W1, b1, W2, b2 = ctx.saved_tensors
        B, T, Fq, _ = h2.shape
        M = spec.shape[-2]
        lib = _lib.load()
        g = d_out.contiguous()
        d_h2 = torch.empty_like(h2)
        grads = torch.empty(64 * 64 + 64 + 2 * M * 64 + 2 * M, dtype=torch.float32, device=h2.device)
        with torch.cuda.device(h2.device):
            ws = torch.empty(lib.eab_head_backward_workspace_bytes(M), dtype=torch.uint8, device=h2.device)
            _lib.check(lib.eab_head_backward(W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(), h2.data_ptr(), spec.data_ptr(),
                                             g.data_ptr(), d_h2.data_ptr(), grads.data_ptr(), B, T, Fq, M, ws.data_ptr(), ws.numel(),
                                             _st(h2.device)), "eab_head_backward")
        o = 0
        dW1 = grads[o:o + 4096].view(64, 64); o += 4096
        db1 = grads[o:o + 64]; o += 64
        dW2 = grads[o:o + 2 * M * 64].view(2 * M, 64); o += 2 * M * 64
        db2 = grads[o:o + 2 * M]
        return d_h2, None, dW1, db1, dW2, db2


def head_filter_sum(h2, spec, W1, b1, W2, b2):
    """out [B,2,T,F] = filter-and-sum of spec [B,T,F,M,2] with the beam weights w_dnn(h2) (EaBNet.py:593-597, 612-613, 114-117);
    differentiable with respect to h2 and the four head parameters."""
    return _HeadFilterSum.apply(h2, spec, W1, b1, W2, b2)


class _ComMagMse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, esti, label, frames, freq_major=False):
        _check(esti, "esti")
        _check(label, "label")
        if esti.ndim != 4 or esti.shape[1] != 2 or tuple(label.shape) != tuple(esti.shape):
            raise ValueError("com_mag_mse_loss: esti and label [B,2,T,F]")
        if freq_major:
            B, _, Fq, T = esti.shape
        else:
            B, _, T, Fq = esti.shape
        e, l = esti.contiguous(), label.contiguous()
        total = 0
        if frames is not None:
            total = int(frames.sum())
            frames = frames.to(device=esti.device, dtype=torch.int32).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=esti.device)
        scratch = torch.empty(2, dtype=torch.float64, device=esti.device)
        with torch.cuda.device(esti.device):
            fn = _lib.load().eab_loss_com_mag_mse_fm if freq_major else _lib.load().eab_loss_com_mag_mse
            _lib.check(fn(e.data_ptr(), l.data_ptr(), frames.data_ptr() if frames is not None else None,
                                                        total, B, T, Fq, loss.data_ptr(), scratch.data_ptr(), _st(esti.device)),
                       "eab_loss_com_mag_mse")
        ctx.save_for_backward(e, l)
        ctx.frames, ctx.total, ctx.freq_major, ctx.dims = frames, total, bool(freq_major), (B, T, Fq)
        return loss

    @staticmethod
    def backward(ctx, g):
        e, l = ctx.saved_tensors
        B, T, Fq = ctx.dims
        d = torch.empty_like(e)
        gc = g.contiguous().float()
        with torch.cuda.device(e.device):
            fn = _lib.load().eab_loss_com_mag_mse_fm_backward if ctx.freq_major else _lib.load().eab_loss_com_mag_mse_backward
            _lib.check(fn(e.data_ptr(), l.data_ptr(),
                                                                 ctx.frames.data_ptr() if ctx.frames is not None else None, ctx.total,
                                                                 B, T, Fq, gc.data_ptr(), d.data_ptr(), _st(e.device)),
                       "eab_loss_com_mag_mse_backward")
        return d, None, None, None


def _frames(frame_list, B, T):
    if frame_list is None:
        return None
    frames = torch.as_tensor(frame_list, dtype=torch.int64)
    if frames.numel() != B or int(frames.max()) != T or int(frames.min()) < 1:
        raise ValueError("frame_list: one entry per utterance, the longest equal to T (the reference pads to it)")
    return frames


def com_mag_mse_loss(esti, label, frame_list=None):
    """com_mag_mse_loss(esti, label, frame_list) of EaBNet.py:627-640 for esti / label [B,2,T,F]; frame_list: valid frames per
    utterance (a list or int tensor; None = all T frames, what train_distributed.py:221 passes).  Differentiable in esti."""
    return _ComMagMse.apply(esti, label, _frames(frame_list, esti.shape[0], esti.shape[2]), False)


def stagewise_com_mag_mse_loss(esti_list, label, frame_list=None):
    """stagewise_com_mag_mse_loss (GaGNet.py:601-619): estimates and label [B,2,F,T]; every stage weighted 0.1, the last one 1."""
    frames = _frames(frame_list, label.shape[0], label.shape[3])
    alpha = [0.1] * len(esti_list)
    alpha[-1] = 1.0
    total = None
    for a, e in zip(alpha, esti_list):
        term = a * _ComMagMse.apply(e, label, frames, True)
        total = term if total is None else total + term
    return total


def eabnet_with_postnet_loss(output, label, frame_list=None):
    """eabnet_with_postnet_loss (EaBNet.py:642-650) on the wrapper's output dict; label [B,2,T,F]."""
    loss0 = com_mag_mse_loss(output["esti0_stft"], label, frame_list)
    loss1 = stagewise_com_mag_mse_loss(output["esti1_stft_list"], label.permute(0, 1, 3, 2), frame_list)
    return {"eabnet": loss0, "postnet": loss1, "final": loss0 + loss1}
