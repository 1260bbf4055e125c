"""Diagnostics: conv_raw with a capped grid (every CTA walks many tiles: all the phase logic runs at a small size) and every
work-sharing setting, against the stage + conv_tma pair.  BatchNorm models must agree bit for bit."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
from oracle import eabnet_oracle as O
for over in ({"norm_type": "BN"}, {}):
    cfg = O.make_cfg(**over)
    sd = O.make_weights(cfg, 2, "B")
    net = EaBNet(**cfg).eval(); net.load_state_dict(sd, strict=True); net = net.cuda()
    wave, _ = O.make_wave(2, cfg["M"], 16000, seed=33)
    spec = O.stft_compress(wave).cuda()
    net.set_option("raw", 0)
    with torch.no_grad():
        ref = net(spec).clone()
    net.set_option("raw", 1)
    for grid in (0, 5, 2):
        for share in (0,):
            net.set_option("raw_grid", grid)
            with torch.no_grad():
                out = net(spec)
            torch.cuda.synchronize()
            print("cfg %s grid %d share %d: max diff %.3e" % (over, grid, share, float((out - ref).abs().max())), flush=True)
