"""Diagnostics: tcm_chain forms (1 = cluster per utterance, 3 = cooperative grid barriers, 0 = layer by layer): outputs and times."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
from oracle import eabnet_oracle as O
cfg = O.make_cfg()
sd = O.make_weights(cfg, 2, "B")
net = EaBNet(**cfg).eval(); net.load_state_dict(sd, strict=True); net = net.cuda()
for L in (16000, 4800, 320):
    wave, _ = O.make_wave(2, 9, L, seed=33)
    spec = O.stft_compress(wave).cuda()
    res = {}
    for mode in (3, 1, 0):
        net.set_option("tcm_chain", mode)
        with torch.no_grad():
            res[mode] = net(spec).clone()
        torch.cuda.synchronize()
    print("L %d: |cluster - coop| %.3e  |layerwise - coop| %.3e" % (L, float((res[1] - res[3]).abs().max()), float((res[0] - res[3]).abs().max())), flush=True)
torch.manual_seed(0)
big = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
for mode in (1, 3):
    big.set_option("tcm_chain", mode)
    with torch.no_grad():
        for _ in range(2): big.enhance(wave)
        big.profile(1); big.enhance(wave)
        fam = {f["kernel"]: round(f["ms"], 3) for f in big.profile_summary()}
        big.profile(0)
    print("tcm_chain=%d:" % mode, fam, flush=True)
