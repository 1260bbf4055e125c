"""CPU: the training-step oracle (oracle/train_oracle.py, the baseline bench.py's config5 object times): the parameter set is
the reference wrapper's (8 789 307 trainable values = 35.16 MB of fp32 gradients), the differentiable forward equals
the inference oracle, the losses reduce to each other the way EaBNet.py:627-650 / GaGNet.py:601-619 define them, and one
backward pass reaches every parameter."""
import torch

from oracle import eabnet_oracle as O
from oracle import gagnet_oracle as G
from oracle import train_oracle as TO


def test_training_oracle_forward_losses_and_gradients():
    torch.manual_seed(0)
    m = TO.TrainableEaBNetWithPostNet()
    n = sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert n == 8789307 and abs(n * 4 / 1e6 - 35.16) < 0.01
    wave, clean = O.make_wave(2, 9, 2400, seed=3)
    spec = O.stft_compress(wave)
    target = O.stft_compress(clean.unsqueeze(1))[..., 0, :].permute(0, 3, 1, 2).contiguous()
    out = m(spec)
    # same function as the inference oracle
    sd = dict(zip(m.names, [v.detach() for v in m.values]))
    ref = G.postnet_forward(sd, spec, m.cfg_e, m.cfg_g, ref_mic=0)
    assert float((out["esti0_stft"].detach() - ref["esti0_stft"]).abs().max()) <= 1e-5
    assert float((out["esti_stft"].detach() - ref["esti_stft"]).abs().max()) <= 1e-4
    T = spec.shape[1]
    frames = [T, T - 3]
    l = TO.loss_fn(out, target, frames)
    assert float((l["final"] - l["eabnet"] - l["postnet"]).abs()) <= 1e-6
    # a single stage of the stagewise loss is the plain loss on transposed tensors
    e = out["esti0_stft"].detach()
    a = TO.com_mag_mse_loss(e, target, frames)
    b = TO.stagewise_com_mag_mse_loss([e.permute(0, 1, 3, 2)], target.permute(0, 1, 3, 2), frames)
    assert abs(float(a) - float(b)) <= 1e-6 * max(1.0, abs(float(a)))
    l["final"].backward()
    missing = [k for k, p in zip(m.names, m.values) if p.requires_grad and (p.grad is None or not torch.isfinite(p.grad).all())]
    assert not missing, missing[:5]
