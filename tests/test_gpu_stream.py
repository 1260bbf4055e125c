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


@pytest.mark.parametrize("opts", [{}, {"stream_umma": 0}, {"stream_lstm": 0}, {"stream_tcm": 0}, {"stream_pair": 0}, {"stream_fuse": 0},
                                  {"stream_pair": 0, "stream_fuse": 0}])
@pytest.mark.parametrize("extra", [{}, {"is_u2": False, "bf_type": "cnn", "M": 4}, {"intra_connect": "add"}])
def test_stream_spec_frames_match_offline_oracle(extra, opts):
    """default: per-layer convs and the LSTM gate GEMM on the tensor cores (conv_umma with ring addressing), fused TCM kernel;
    opts switch each of them back to the CUDA-core / per-layer form"""
    cfg = O.make_cfg(norm_type="BN", **extra)
    net, sd = _net(cfg, seed=3)
    for k, v in opts.items():
        net.set_option(k, v)
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


def test_stream_launch_merging_is_bit_identical():
    """stream_pair (both output parities of a transposed conv in one grid) and stream_fuse (a module's residual sum in the last inner
    deconv's epilogue) change the launch list, not the arithmetic: same bits, fewer launches"""
    cfg = O.make_cfg(norm_type="BN")
    S, T = 5, 12
    wave, _ = O.make_wave(S, 9, 160 * (T - 1), seed=31)
    spec = O.stft_compress(wave).cuda()
    outs, launches = [], []
    for opts in ({}, {"stream_pair": 0}, {"stream_fuse": 0}, {"stream_pair": 0, "stream_fuse": 0}):
        net, _ = _net(cfg, seed=7)
        for k, v in opts.items():
            net.set_option(k, v)
        ses = net.stream(S)
        outs.append(torch.stack([ses.step_spec(spec[:, t].contiguous()) for t in range(T)], 2).cpu())
        launches.append(net.last_launch_count())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert launches[0] < launches[1] < launches[3] and launches[0] < launches[2] < launches[3], launches


def test_stream_many_streams_fill_the_gpu():
    """128 streams: the merged two-parity launches of the wide layers have more tiles than SMs (the grid is split between the
    variants by tile count, CTAs walk several tiles) - checked against the offline oracle"""
    cfg = O.make_cfg(norm_type="BN")
    net, sd = _net(cfg, seed=13)
    S, T = 128, 4
    wave, _ = O.make_wave(S, 9, 160 * (T - 1), seed=51)
    spec = O.stft_compress(wave)
    ref = O.forward(sd, spec, cfg)
    ses = net.stream(S)
    dspec = spec.cuda()
    outs = torch.stack([ses.step_spec(dspec[:, t].contiguous()) for t in range(T)], 2).cpu()
    assert (outs - ref).abs().max() <= EXACT * max(1.0, float(ref.abs().max()))


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


def test_stream_reset_one_restarts_that_stream_only():
    """eab_stream_reset_one: stream 1 is restarted mid-run with new audio; streams 0 and 2 continue bit-identically, and the
    restarted stream reproduces a fresh run of its new audio (causal history, LSTM state, STFT reflection, overlap-add)"""
    from eabnet_b200.model import EaBNetStream
    cfg = O.make_cfg(norm_type="BN")
    net, sd = _net(cfg, seed=11)
    S, nh, k0 = 3, 30, 11
    wave, _ = O.make_wave(S, 9, 160 * nh, seed=51)
    wave2, _ = O.make_wave(1, 9, 160 * (nh - k0), seed=52)
    dw, dw2 = wave.cuda(), wave2.cuda()
    plain = EaBNetStream(net, S)
    ref_hops = [plain.step(dw[:, :, 160 * k:160 * (k + 1)].contiguous()).clone() for k in range(nh)]
    fresh = EaBNetStream(net, 1)
    fresh_hops = [fresh.step(dw2[:, :, 160 * k:160 * (k + 1)].contiguous()).clone() for k in range(nh - k0)]
    ses = EaBNetStream(net, S)
    for k in range(nh):
        hop = dw[:, :, 160 * k:160 * (k + 1)].clone()
        if k == k0:
            ses.reset_stream(1)
        if k >= k0:
            hop[1] = dw2[0, :, 160 * (k - k0):160 * (k - k0 + 1)]
        out = ses.step(hop.contiguous())
        assert torch.equal(out[0], ref_hops[k][0]) and torch.equal(out[2], ref_hops[k][2]), k
        if k < k0:
            assert torch.equal(out[1], ref_hops[k][1])
        else:
            assert (out[1] - fresh_hops[k - k0][0]).abs().max() <= 1e-6, k


@pytest.mark.parametrize("graph", [False, True])
def test_stream_pcm16_hops(graph):
    """the 16-bit PCM hop ABI: int16 hops in / out against the float step on the same samples (<= 1 LSB)"""
    from eabnet_b200.model import EaBNetStream
    cfg = O.make_cfg(norm_type="BN")
    net, _ = _net(cfg, seed=7)
    S, nh = 2, 12
    wave, _ = O.make_wave(S, 9, 160 * nh, seed=61)
    pcm = (wave * 32768.0).round().clamp(-32768, 32767).to(torch.int16).cuda()
    f = EaBNetStream(net, S)
    q = EaBNetStream(net, S, graph=graph)
    for k in range(nh):
        hp = pcm[:, :, 160 * k:160 * (k + 1)].contiguous()
        a = f.step(hp.float() / 32768.0)
        b = q.step(hp)
        exp = (a.clamp(-1, 1) * 32767.0).to(torch.int16)
        assert b.dtype == torch.int16 and int((b.int() - exp.int()).abs().max()) <= 1, k


def test_instancenorm_model_streams_through_batchnorm_calibration():
    """to_batchnorm: the statistics the InstanceNorm layers saw on a calibration utterance become the BN buffers of the same
    network.  (1) the resulting state_dict is a reference-loadable norm_type="BN" model: the oracle run with it equals the
    oracle run of the IN model on that utterance; (2) our BN model equals our IN model there; (3) it streams."""
    from eabnet_b200.model import EaBNetStream
    cfg_in = O.make_cfg()
    net, sd_in = _net(cfg_in, seed=4)
    wave, _ = O.make_wave(1, 9, 160 * 40, seed=71)
    spec = O.stft_compress(wave)
    bn = net.to_batchnorm([spec.cuda()])
    assert bn.norm_type == "BN" and net.norm_type == "IN"
    sd_bn = {k: v.detach().cpu() for k, v in bn.state_dict().items()}
    cfg_bn = O.make_cfg(norm_type="BN")
    assert list(sd_bn) == list(O.param_shapes(cfg_bn))
    ref_in = O.forward(sd_in, spec, cfg_in)
    ref_bn = O.forward(sd_bn, spec, cfg_bn)
    scale = max(1.0, float(ref_in.abs().max()))
    assert (ref_bn - ref_in).abs().max() <= 2e-4 * scale          # (1): statistics measured on fp16-split tensor-core outputs
    with torch.no_grad():
        out_in = net(spec.cuda()).cpu()
        out_bn = bn(spec.cuda()).cpu()
    assert (out_bn - out_in).abs().max() <= 2e-4 * scale          # (2)
    assert (out_bn - ref_bn).abs().max() <= 4e-4 * scale
    # (3) hop by hop == offline enhance of the BN model
    with torch.no_grad():
        off = bn.enhance(wave.cuda()).cpu()
    ses = EaBNetStream(bn, 1)
    dw = wave.cuda()
    hops = [ses.step(dw[:, :, 160 * k:160 * (k + 1)].contiguous()).cpu() for k in range(40)]
    got = torch.cat(hops[1:], dim=1)                             # output delayed by one hop
    assert (got - off[:, :got.shape[1]]).abs().max() <= 1e-4


def test_to_batchnorm_pools_batches_and_rejects_bn_models():
    cfg = O.make_cfg()
    net, _ = _net(cfg, seed=5)
    w1, _ = O.make_wave(2, 9, 4000, seed=1)
    w2, _ = O.make_wave(1, 9, 6400, seed=2)
    bn = net.to_batchnorm([O.stft_compress(w1).cuda(), O.stft_compress(w2).cuda()])
    sd = bn.state_dict()
    rv = [v for k, v in sd.items() if k.endswith("running_var")]
    assert len(rv) > 60 and all(bool((v >= 0).all()) and bool(torch.isfinite(v).all()) for v in rv)
    with pytest.raises(RuntimeError, match="static normalisation"):
        bn.to_batchnorm([O.stft_compress(w2).cuda()])


def test_stream_session_replans_after_an_option_change():
    """kernel-selection options change the state layout: a step after set_option raises until reset() re-plans the session"""
    cfg = O.make_cfg(norm_type="BN")
    net, sd = _net(cfg, seed=9)
    S, T = 2, 12
    wave, _ = O.make_wave(S, 9, 160 * (T - 1), seed=13)
    spec = O.stft_compress(wave)
    ref = O.forward(sd, spec, cfg)
    ses = net.stream(S)
    ses.step_spec(spec[:, 0].cuda().contiguous())
    net.set_option("stream_umma", 0)
    with pytest.raises(RuntimeError, match="reset"):
        ses.step_spec(spec[:, 1].cuda().contiguous())
    ses.reset()
    for t in range(T):
        got = ses.step_spec(spec[:, t].cuda().contiguous()).cpu()
        assert float((got - ref[:, :, t]).abs().max()) <= EXACT * max(1.0, float(ref.abs().max()))
