"""Stage-by-stage parity report on a GPU box (diagnostic; the asserting versions live in tests/).

    python tools/gpu_check.py [--cfg '{"norm_type":"BN"}'] [--B 2] [--L 3200]

Prints max-abs error of every debug tap, of the STFT, the forward output and the iSTFT against the CPU
oracle, without stopping at the first mismatch."""
import argparse
import ast
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import eabnet_oracle as O  # noqa: E402  (diagnostic script: oracle used as the checker)
import eabnet_b200 as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="{}")
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--L", type=int, default=3200)
    ap.add_argument("--variant", default="B")
    ap.add_argument("--opt", action="append", default=[], help="name=value for EaBNet.set_option")
    a = ap.parse_args()
    cfg = O.make_cfg(**ast.literal_eval(a.cfg))
    sd = O.make_weights(cfg, 0, a.variant)
    wave, _ = O.make_wave(a.B, cfg["M"], a.L)
    dev = torch.device("cuda")
    net = E.EaBNet(**cfg).eval()
    net.load_state_dict(sd, strict=True)
    net.to(dev)
    for o in a.opt:
        k, v = o.split("=")
        net.set_option(k, int(v))

    spec_ref = O.stft_compress(wave)
    spec = E.stft_compress(wave.to(dev))
    torch.cuda.synchronize()
    print("stft        max|err| %.3e (absmax %.3f)" % ((spec.cpu() - spec_ref).abs().max(), spec_ref.abs().max()))

    taps = {}
    t0 = time.time()
    out_ref = O.forward(sd, spec_ref, cfg, taps)
    t_cpu = time.time() - t0
    with torch.no_grad():
        out = net(spec_ref.to(dev))
    torch.cuda.synchronize()
    print("launches", net.last_launch_count(), " oracle cpu %.2fs" % t_cpu)
    for name, ref in taps.items():
        try:
            got = net.debug_tap(name, tuple(ref.shape)).cpu()
            print("tap %-7s max|err| %.3e (absmax %.3f) shape %s" % (name, (got - ref).abs().max(), ref.abs().max(), tuple(ref.shape)))
        except Exception as e:  # noqa
            print("tap %-7s FAILED: %s" % (name, e))
    print("forward     max|err| %.3e (absmax %.3f) nan=%s" % ((out.cpu() - out_ref).abs().max(), out_ref.abs().max(),
                                                            bool(torch.isnan(out).any())))
    if out_ref.dim() == 4:
        wav_ref = O.istft(out_ref)
        wav = E.istft(out_ref.to(dev))
        print("istft       max|err| %.3e (absmax %.3f)" % ((wav.cpu() - wav_ref).abs().max(), wav_ref.abs().max()))
        with torch.no_grad():
            e2e = net.enhance(wave.to(dev))
        print("enhance     max|err| %.3e" % (e2e.cpu() - wav_ref).abs().max())
        pin = wave.pin_memory()
        e2h = net.enhance_host(pin)
        print("enhance_host max|err| %.3e" % (e2h - wav_ref).abs().max())


if __name__ == "__main__":
    main()
