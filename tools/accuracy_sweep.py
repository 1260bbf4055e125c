"""Accuracy of the precision policies at the full 6 s length against the fp32 CPU oracle (diagnostics)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import eabnet_oracle as O
from eabnet_b200 import EaBNet
torch.set_num_threads(os.cpu_count())
cfg = O.make_cfg()
for seed, variant, L in [(0, "B", 96000), (1, "A", 96000), (2, "B", 64000)]:
    sd = O.make_weights(cfg, seed, variant)
    wave, _ = O.make_wave(1, 9, L, seed=100 + seed)
    spec = O.stft_compress(wave)
    ref = O.forward(sd, spec, cfg)
    net = EaBNet(**cfg).eval(); net.load_state_dict(sd); net.cuda()
    row = []
    for enc, dec, inner in [(3, 3, 3), (3, 1, 3), (3, 1, 1), (1, 1, 1)]:
        net.set_option("enc_passes", enc); net.set_option("dec_passes", dec); net.set_option("inner_passes", inner)
        with torch.no_grad():
            out = net(spec.cuda()).cpu()
        row.append("enc%d/dec%d/inner%d: %.2e" % (enc, dec, inner, float((out - ref).abs().max())))
    net.set_option("enc_passes", 3); net.set_option("dec_passes", 1); net.set_option("first_passes", 1)
    with torch.no_grad():
        out = net(spec.cuda()).cpu()
    net.set_option("first_passes", 3)
    row.append("default + single-pass first layer: %.2e" % float((out - ref).abs().max()))
    net.set_option("round_half", 1)
    with torch.no_grad():
        out = net(spec.cuda()).cpu()
    net.set_option("round_half", 0)
    row.append("default + fp16-rounded decoder activations: %.2e" % float((out - ref).abs().max()))
    print("seed %d variant %s T=%d |out|max %.2f  " % (seed, variant, spec.shape[1], float(ref.abs().max())), "  ".join(row), flush=True)
