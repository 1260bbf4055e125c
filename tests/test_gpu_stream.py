"""GPU (-m gpu): causal frame-by-frame inference with carried state (BASELINE configs[2]) through the C ABI
(eab_stream_*), against the CPU oracle run OFFLINE on the whole signal: with is_causal=True and norm_type='BN' frame n
of the offline result depends on frames <= n only, so the streamed frames must reproduce it one for one."""
import pytest
import torch

from oracle import eabnet_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3
EXACT = 5e-5        # the streaming path runs the fp32 CUDA-core kernels


def _net(cfg, variant="B", seed=0):
    from eabnet_b200 import EaBNet
    sd = O.make_weights(cfg, seed, variant)
    net = EaBNet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    return net.cuda(), sd


@pytest.mark.parametrize("extra", [{}, {"is_u2": False, "bf_type": "cnn", "M": 4}, {"intra_connect": "add"}])
def test_stream_spec_frames_match_offline_oracle(extra):
    cfg = O.make_cfg(norm_type="BN", **extra)
    net, sd = _net(cfg, seed=3)
    S, T = 3, 70                                   # 70 frames: deeper than the largest TCM dilation ring (4*32+1 = 129? no: covers d<=16 fully, d=32 partially)
    wave, _ = O.make_wave(S, cfg["M"], 160 * (T - 1), seed=11)
    spec = O.stft_compress(wave)                   # [S,T,161,M,2]
    ref = O.forward(sd, spec, cfg)                 # [S,2,T,161]
    ses = net.stream(S)
    dspec = spec.cuda()
    worst = 0.0
    for t in range(T):
        got = ses.step_spec(dspec[:, t].contiguous()).cpu()
        worst = max(worst, float((got - ref[:, :, t]).abs().max()))
    assert worst <= EXACT * max(1.0, float(ref.abs().max())), worst
    assert net.last_launch_count() > 0


def test_stream_long_history_and_reset():
    """more frames than the deepest dilation ring (4*32+1), then reset() and replay: identical results"""
    cfg = O.make_cfg(norm_type="BN")
    net, sd = _net(cfg, seed=5)
    S, T = 2, 150
    wave, _ = O.make_wave(S, 9, 160 * (T - 1), seed=21)
    spec = O.stft_compress(wave)
    ref = O.forward(sd, spec, cfg)
    ses = net.stream(S)
    dspec = spec.cuda()
    outs = torch.stack([ses.step_spec(dspec[:, t].contiguous()) for t in range(T)], 2).cpu()      # [S,2,T,161]
    assert (outs - ref).abs().max() <= EXACT * max(1.0, float(ref.abs().max()))
    ses.reset()
    again = torch.stack([ses.step_spec(dspec[:, t].contiguous()) for t in range(20)], 2).cpu()
    assert torch.equal(again, outs[:, :, :20])


@pytest.mark.parametrize("graph", [False, True])
def test_stream_wave_hops_match_offline_enhance(graph):
    """hop in -> hop out (delayed by one hop) == the oracle's offline wave -> wave result, and the tensor-core offline
    path of the same module agrees with it inside the contractual tolerance"""
    from eabnet_b200.model import EaBNetStream
    cfg = O.make_cfg(norm_type="BN")
    net, sd = _net(cfg, seed=7)
    S, nh = 4, 40
    wave, _ = O.make_wave(S, 9, 160 * nh, seed=31)
    ref = O.enhance(sd, wave, cfg)                 # [S, 160*nh]
    ses = EaBNetStream(net, S, graph=graph)
    dw = wave.cuda()
    hops = [ses.step(dw[:, :, 160 * k:160 * (k + 1)].contiguous()).clone() for k in range(nh)]
    assert float(hops[0].abs().max()) == 0.0       # nothing can be emitted before the second frame exists
    got = torch.cat(hops[1:], 1).cpu()             # samples [0, 160*(nh-1))
    assert (got - ref[:, :160 * (nh - 1)]).abs().max() <= EXACT
    with torch.no_grad():
        off = net.enhance(dw).cpu()
    assert (off[:, :160 * (nh - 1)] - got).abs().max() <= TOL


def test_stream_rejects_instance_norm_and_noncausal():
    from eabnet_b200 import EaBNet
    for kw in ({}, {"norm_type": "BN", "is_causal": False}):
        net = EaBNet(**kw).eval().cuda()
        with pytest.raises(RuntimeError):
            net.stream(2)


def test_stream_graph_capture_as_first_cuda_work_of_a_process():
    """graph=True must not depend on an earlier eager call having initialised the library's tables (fresh process)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch\n"
            "from eabnet_b200 import EaBNet\n"
            "from eabnet_b200.model import EaBNetStream\n"
            "torch.manual_seed(0)\n"
            "net = EaBNet(norm_type='BN').eval().cuda()\n"
            "ses = EaBNetStream(net, 3, graph=True)\n"
            "hop = 0.1 * torch.randn(3, 9, 160, device='cuda')\n"
            "a = ses.step(hop).clone(); b = ses.step(hop).clone()\n"
            "ref = EaBNetStream(net, 3, graph=False)\n"
            "c = ref.step(hop).clone(); d = ref.step(hop).clone()\n"
            "torch.cuda.synchronize()\n"
            "assert torch.isfinite(b).all() and torch.equal(a, c) and torch.equal(b, d)\n"
            "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
