"""Diagnostics for conv_raw_kernel (option raw=1, the default) against the stage + conv_tma pair (raw=0):
   1. small case: outputs and every debug tap of both paths, and the oracle
   2. bench configuration (64 x 6 s): per-launch CUDA-event times of both paths (profile detail mode)
Prints nothing that is a bench value."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet  # noqa: E402
from oracle import eabnet_oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=6.0)
ap.add_argument("--skip-small", action="store_true")
ap.add_argument("--skip-big", action="store_true")
ap.add_argument("--detail", action="store_true")
a = ap.parse_args()


def small():
    for over in ({}, {"norm_type": "BN"}, {"is_causal": False}, {"intra_connect": "add"}):
        cfg = O.make_cfg(**over)
        sd = O.make_weights(cfg, 2, "B")
        net = EaBNet(**cfg).eval()
        net.load_state_dict(sd, strict=True)
        net = net.cuda()
        wave, _ = O.make_wave(2, cfg["M"], 16000, seed=33)
        spec = O.stft_compress(wave)
        taps = {}
        ref = O.forward(sd, spec, cfg, taps)
        res = {}
        for raw in (0, 1):
            net.set_option("raw", raw)
            with torch.no_grad():
                out = net(spec.cuda()).cpu()
            res[raw] = (out, {k: net.debug_tap(k, tuple(v.shape)).cpu() for k, v in taps.items()}, net.last_launch_count())
        d01 = float((res[0][0] - res[1][0]).abs().max())
        e0 = float((res[0][0] - ref).abs().max())
        e1 = float((res[1][0] - ref).abs().max())
        print("cfg %s: launches %d -> %d ; |raw0-raw1| %.3e ; err vs oracle raw0 %.3e raw1 %.3e" % (over, res[0][2], res[1][2], d01, e0, e1))
        worst = max(((float((res[0][1][k] - res[1][1][k]).abs().max()), k) for k in taps), default=(0, ""))
        print("   worst tap difference raw0 vs raw1: %.3e at %s" % worst)
        for k in taps:
            d = float((res[1][1][k] - taps[k]).abs().max())
            s = float(taps[k].abs().max())
            if d > 2e-3 * max(1.0, s):
                print("   TAP %s: raw1 vs oracle %.3e (scale %.3e)" % (k, d, s))


def big():
    torch.manual_seed(0)
    net = EaBNet().eval().cuda()
    wave = 0.1 * torch.randn(a.batch, 9, int(a.seconds * 16000), device="cuda")
    for raw in (1, 0):
        net.set_option("raw", raw)
        with torch.no_grad():
            for _ in range(2):
                y = net.enhance(wave)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = net.enhance(wave)
            e1.record()
            torch.cuda.synchronize()
            print("raw=%d: %.3f ms/step, %d launches, finite %s" % (raw, e0.elapsed_time(e1) / 5, net.last_launch_count(), bool(torch.isfinite(y).all())))
            net.profile(1)
            net.enhance(wave)
            fam = net.profile_summary()
            print("  families:", json.dumps([{k: (round(v, 3) if isinstance(v, float) else v) for k, v in f.items() if k in ("kernel", "launches", "ms")} for f in fam]))
            if a.detail:
                net.profile(2)
                net.enhance(wave)
                for f in net.profile_summary():
                    if "conv" in f["kernel"] or "stage" in f["kernel"]:
                        gbs = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0
                        tfs = f["flops"] / (f["ms"] * 1e-3) / 1e12 if f["ms"] > 0 else 0
                        print("   %-22s %8.4f ms  %7.0f GB/s  %6.0f TF/s" % (f["kernel"], f["ms"], gbs, tfs))
            net.profile(0)


if not a.skip_small:
    small()
if not a.skip_big:
    big()
