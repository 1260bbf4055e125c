"""CPU / torch-eager restatement (TEST + BASELINE INFRASTRUCTURE, never imported by the product) of the reference's TRAINING
step - train_distributed.py:214-230: prepare_data -> EaBNetWithPostNet -> eabnet_with_postnet_loss -> backward ->
clip_grad_norm_(1.0) -> Adam - as a differentiable module over the oracle's functional forward.  bench.py's `config5`
object times it (under DistributedDataParallel when launched with torchrun) as the stated baseline of BASELINE configs[4];
parity unpinned beyond what the forward oracle's golden vectors pin (the reference has no training fixtures)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import eabnet_oracle as E
from . import gagnet_oracle as G


def _lstm_layer_autograd(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """nn.LSTM(batch_first, one layer, zero state) as the fused primitive, differentiable in the four parameters
    (E.lstm_layer copies them into a fresh module under no_grad, which is right for inference only)."""
    H = w_hh.shape[1]
    z = x.new_zeros(1, x.shape[0], H)
    return torch._VF.lstm(x, (z, z), [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, torch.is_grad_enabled(), False, True)[0]


def com_mag_mse_loss(esti: Tensor, label: Tensor, frame_list) -> Tensor:
    """EaBNet.py:627-640"""
    with torch.no_grad():
        mask = nn.utils.rnn.pad_sequence([torch.ones((n, esti.shape[-1]), dtype=esti.dtype) for n in frame_list],
                                         batch_first=True).to(esti.device)
        cmask = torch.stack((mask, mask), dim=1)
    me, ml = torch.norm(esti, dim=1), torch.norm(label, dim=1)
    loss1 = (((me - ml) ** 2.0) * mask).sum() / mask.sum()
    loss2 = (((esti - label) ** 2.0) * cmask).sum() / cmask.sum()
    return 0.5 * (loss1 + loss2)


def stagewise_com_mag_mse_loss(esti_list, label: Tensor, frame_list) -> Tensor:
    """GaGNet.py:601-619 (label and estimates [B,2,F,T]; every stage weighted 0.1, the last one 1)"""
    alpha = [0.1] * len(esti_list)
    alpha[-1] = 1.0
    with torch.no_grad():
        mask = nn.utils.rnn.pad_sequence([torch.ones((n, label.shape[-2]), dtype=label.dtype) for n in frame_list],
                                         batch_first=True).to(label.device).transpose(-2, -1).contiguous()
        cmask = torch.stack((mask, mask), dim=1)
    loss1, loss2 = 0.0, 0.0
    ml = torch.norm(label, dim=1)
    for a, e in zip(alpha, esti_list):
        loss1 = loss1 + a * (((e - label) ** 2.0) * cmask).sum() / cmask.sum()
        loss2 = loss2 + a * (((torch.norm(e, dim=1) - ml) ** 2.0) * mask).sum() / mask.sum()
    return 0.5 * (loss1 + loss2)


class TrainableEaBNetWithPostNet(nn.Module):
    """The wrapper's parameters as nn.Parameters (so DDP / Adam / clip_grad_norm_ see what they see in the reference: 8.79 M
    values, 35.16 MB of fp32 gradients), forward = the oracle's functional EaBNetWithPostNet.forward with autograd on."""

    def __init__(self, cfg_e=None, cfg_g=None, seed: int = 0, ref_mic: int = 0):
        super().__init__()
        self.cfg_e = E.make_cfg() if cfg_e is None else cfg_e
        self.cfg_g = G.make_gag_cfg() if cfg_g is None else cfg_g
        self.ref_mic = ref_mic
        sd = G.make_postnet_weights(self.cfg_e, self.cfg_g, seed, "B")
        self.names = list(sd)
        self.values = nn.ParameterList([nn.Parameter(v.clone(), requires_grad=v.is_floating_point()) for v in sd.values()])

    def forward(self, noisy_stft: Tensor) -> dict:
        sd = dict(zip(self.names, self.values))
        sd_e = {k[len("eabnet."):]: v for k, v in sd.items() if k.startswith("eabnet.")}
        sd_g = {k[len("postnet."):]: v for k, v in sd.items() if k.startswith("postnet.")}
        saved = E.lstm_layer
        E.lstm_layer = _lstm_layer_autograd
        try:
            est0 = E.forward.__wrapped__(sd_e, noisy_stft, self.cfg_e)
            ref = noisy_stft[..., self.ref_mic, :].permute(0, 3, 1, 2)
            lst = G.gag_forward.__wrapped__(sd_g, ref, est0, self.cfg_g)
        finally:
            E.lstm_layer = saved
        return {"esti0_stft": est0, "esti1_stft_list": lst, "esti_stft": lst[-1].permute(0, 1, 3, 2)}


def loss_fn(output: dict, label: Tensor, frame_list) -> dict:
    """eabnet_with_postnet_loss, EaBNet.py:642-650"""
    l0 = com_mag_mse_loss(output["esti0_stft"], label, frame_list)
    l1 = stagewise_com_mag_mse_loss(output["esti1_stft_list"], label.permute(0, 1, 3, 2), frame_list)
    return {"eabnet": l0, "postnet": l1, "final": l0 + l1}
