"""GPU (-m gpu): first slice of the training step (train_distributed.py:214-230) - the hand-written forward and backward
kernels of the head's tail (w_dnn + filter-and-sum) and of com_mag_mse_loss - against torch autograd of the reference
formulas restated in float64 on the CPU (EaBNet.py:593-597, 612-613, 114-117, 627-640)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_head(h2, spec, W1, b1, W2, b2):
    """LSTM_BF.w_dnn + the filter-and-sum of EaBNet.forward"""
    B, T, F, M, _ = spec.shape
    w = torch.relu(h2 @ W1.T + b1) @ W2.T + b2
    w = w.view(B, T, F, M, 2)
    wr, wi = w[..., 0], w[..., -1]
    xr, xi = spec[..., 0], spec[..., -1]
    return torch.stack(((wr * xr - wi * xi).sum(-1), (wr * xi + wi * xr).sum(-1)), dim=1)


def ref_loss(esti, label, frame_list):
    """com_mag_mse_loss, EaBNet.py:627-640"""
    mask = torch.nn.utils.rnn.pad_sequence([torch.ones((n, esti.shape[-1]), dtype=esti.dtype) for n in frame_list], batch_first=True)
    cmask = torch.stack((mask, mask), dim=1)
    me, ml = torch.norm(esti, dim=1), torch.norm(label, dim=1)
    loss1 = (((me - ml) ** 2.0) * mask).sum() / mask.sum()
    loss2 = (((esti - label) ** 2.0) * cmask).sum() / cmask.sum()
    return 0.5 * (loss1 + loss2)


def _case(B, T, F, M, seed):
    g = torch.Generator().manual_seed(seed)
    h2 = torch.tanh(torch.randn(B, T, F, 64, generator=g))
    spec = 0.5 * torch.randn(B, T, F, M, 2, generator=g)
    W1 = torch.randn(64, 64, generator=g) / 8
    b1 = 0.1 * torch.randn(64, generator=g)
    W2 = torch.randn(2 * M, 64, generator=g) / 8
    b2 = 0.1 * torch.randn(2 * M, generator=g)
    label = 0.5 * torch.randn(B, 2, T, F, generator=g)
    return h2, spec, W1, b1, W2, b2, label


@pytest.mark.parametrize("B,T,F,M", [(1, 3, 5, 9), (2, 17, 161, 9), (3, 40, 33, 1), (1, 129, 161, 16)])
def test_head_and_loss_forward_backward_match_autograd(B, T, F, M):
    from eabnet_b200.train import com_mag_mse_loss, head_filter_sum
    h2, spec, W1, b1, W2, b2, label = _case(B, T, F, M, 7 * B + T)
    frames = [T] * B
    if B > 1:
        frames[1] = max(1, T // 2)
    # reference: float64 autograd
    p64 = [t.double().requires_grad_(True) for t in (h2, W1, b1, W2, b2)]
    out64 = ref_head(p64[0], spec.double(), *p64[1:])
    loss64 = ref_loss(out64, label.double(), frames)
    g64 = torch.autograd.grad(loss64, p64)
    # product: hand-written kernels behind autograd Functions
    pg = [t.cuda().requires_grad_(True) for t in (h2, W1, b1, W2, b2)]
    out = head_filter_sum(pg[0], spec.cuda(), *pg[1:])
    loss = com_mag_mse_loss(out, label.cuda(), frames)
    gg = torch.autograd.grad(3.0 * loss, pg)                    # an upstream factor exercises grad_loss_dev
    assert float((out.detach().cpu().double() - out64.detach()).abs().max()) <= 2e-5 * max(1.0, float(out64.abs().max()))
    assert abs(float(loss.detach()) - float(loss64.detach())) <= 1e-5 * max(1.0, abs(float(loss64.detach())))
    for name, a, b in zip(("h2", "W1", "b1", "W2", "b2"), gg, g64):
        ref = 3.0 * b
        assert a.shape == ref.shape
        err = float((a.detach().cpu().double() - ref).abs().max())
        assert err <= 3e-5 * max(1e-3, float(ref.abs().max())) + 1e-9, (name, err, float(ref.abs().max()))


def test_head_backward_is_deterministic_and_loss_default_mask():
    from eabnet_b200.train import com_mag_mse_loss, head_filter_sum
    h2, spec, W1, b1, W2, b2, label = _case(2, 300, 161, 9, 3)
    pg = [t.cuda().requires_grad_(True) for t in (h2, W1, b1, W2, b2)]
    runs = []
    for _ in range(2):
        out = head_filter_sum(pg[0], spec.cuda(), *pg[1:])
        loss = com_mag_mse_loss(out, label.cuda())              # frame_list = [T] * B, train_distributed.py:221
        runs.append(torch.autograd.grad(loss, pg))
    for a, b in zip(*runs):
        assert torch.equal(a, b)
    ref = ref_loss(ref_head(h2.double(), spec.double(), W1.double(), b1.double(), W2.double(), b2.double()), label.double(), [300, 300])
    assert abs(float(loss.detach()) - float(ref)) <= 1e-5 * abs(float(ref))


def test_train_slice_rejects_cpu_and_bad_shapes():
    from eabnet_b200.train import com_mag_mse_loss, head_filter_sum
    h2, spec, W1, b1, W2, b2, label = _case(1, 4, 5, 2, 1)
    with pytest.raises(TypeError):
        head_filter_sum(h2, spec, W1, b1, W2, b2)
    with pytest.raises(ValueError):
        head_filter_sum(h2.cuda(), spec.cuda(), W1.cuda(), b1.cuda(), W2[:2].cuda(), b2.cuda())
    with pytest.raises(ValueError):
        com_mag_mse_loss(label.cuda(), label.cuda(), [2])     # longest entry must be T


def test_wrapper_loss_matches_autograd_of_the_reference_formulas():
    """eabnet_with_postnet_loss (EaBNet.py:642-650): plain loss on the beamformer estimate + stagewise loss on the q post-filter
    estimates ([B,2,F,T], weights 0.1 / 1, GaGNet.py:601-619); forward value and gradients against float64 autograd"""
    from eabnet_b200.train import eabnet_with_postnet_loss
    from oracle import train_oracle as TO
    g = torch.Generator().manual_seed(17)
    B, T, F = 3, 37, 161
    e0 = 0.5 * torch.randn(B, 2, T, F, generator=g)
    e1 = [0.5 * torch.randn(B, 2, F, T, generator=g).contiguous() for _ in range(3)]
    label = 0.5 * torch.randn(B, 2, T, F, generator=g)
    frames = [T, 20, 31]
    r0 = e0.double().requires_grad_(True)
    r1 = [t.double().requires_grad_(True) for t in e1]
    ref = TO.loss_fn({"esti0_stft": r0, "esti1_stft_list": r1}, label.double(), frames)
    gref = torch.autograd.grad(ref["final"], [r0] + r1)
    p0 = e0.cuda().requires_grad_(True)
    p1 = [t.cuda().requires_grad_(True) for t in e1]
    got = eabnet_with_postnet_loss({"esti0_stft": p0, "esti1_stft_list": p1}, label.cuda(), frames)
    ggot = torch.autograd.grad(got["final"], [p0] + p1)
    for k in ("eabnet", "postnet", "final"):
        assert abs(float(got[k].detach()) - float(ref[k].detach())) <= 1e-5 * max(1.0, abs(float(ref[k].detach()))), k
    for a, b in zip(ggot, gref):
        assert float((a.cpu().double() - b).abs().max()) <= 3e-5 * max(1e-6, float(b.abs().max())) + 1e-10
