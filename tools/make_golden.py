"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only (the reference does not travel to the GPU box):

    python tools/make_golden.py

For every case: weights = oracle.make_weights(cfg, seed, variant) (name-seeded, so they can be regenerated
anywhere), loaded into the reference ``EaBNet(**cfg)`` with ``load_state_dict(strict=True)``; input =
reference ``prepare_data`` (AST-extracted from train_distributed.py:68-95 because that module imports
packages that are not installed) applied to ``oracle.make_wave``; outputs = reference forward and
``torch.istft`` exactly as enhance.py:59-61 calls it.  Stored: the compressed input spectrum, the forward
output and the enhanced waveform (fp32) - a few hundred kB per case.
"""
import argparse
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import eabnet_oracle as O  # noqa: E402

# name, cfg overrides, B, L (samples), weight variant
CASES = [
    ("default_b2_t21", {}, 2, 3200, "B"),
    ("default_b1_t51_initA", {}, 1, 8000, "A"),
    ("bn_b2_t13", {"norm_type": "BN"}, 2, 1920, "B"),
    ("unet_cnn_m8_b1_t17", {"is_u2": False, "bf_type": "cnn", "M": 8}, 1, 2560, "B"),
    ("miso_add_m1_b2_t9", {"topo_type": "miso", "intra_connect": "add", "M": 1}, 2, 1280, "B"),
    ("noncausal_b1_t40", {"is_causal": False}, 1, 6240, "B"),
    ("default_b1_t601", {}, 1, 96000, "B"),         # the BASELINE config-2 length (6 s, T = 601): ~6.5 MB
]


def load_reference():
    sys.path.insert(0, REF)
    from EaBNet import EaBNet  # noqa
    src = open(os.path.join(REF, "train_distributed.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "prepare_data")
    ns = {"torch": torch}
    exec(compile(ast.Module([fn], []), "prepare_data", "exec"), ns)
    return EaBNet, ns["prepare_data"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default="", help="comma-separated case names (default: all)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    EaBNet, prepare_data = load_reference()
    torch.set_num_threads(os.cpu_count())
    for name, over, B, L, variant in CASES:
        if a.only and name not in a.only.split(","):
            continue
        cfg = O.make_cfg(**over)
        net = EaBNet(**cfg).eval()
        ref_shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        mine = O.param_shapes(cfg)
        assert list(ref_shapes.items()) == list(mine.items()), "param table differs from reference: " + name
        sd = O.make_weights(cfg, seed=0, variant=variant)
        net.load_state_dict(sd, strict=True)
        wave, clean = O.make_wave(B, cfg["M"], L, seed=1234)
        args = types.SimpleNamespace(mics=cfg["M"], sr=16000, wav_len=L / 16000, win_size=0.020,
                                     win_shift=0.010, fft_num=320)
        with torch.no_grad():
            spec, _ = prepare_data(wave, clean.unsqueeze(1), "cpu", args)
            y = net(spec)
            rec = {"cfg": repr(over), "B": B, "L": L, "variant": variant, "spec": spec.numpy(),
                   "out": y.numpy()}
            if y.dim() == 4:
                z = torch.view_as_complex(y.permute(0, 3, 2, 1).contiguous())
                rec["wav"] = torch.istft(z, 320, 160, 320, torch.hann_window(320)).numpy()
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **rec)
        print(name, tuple(spec.shape), "->", tuple(y.shape), "absmax %.3f" % float(y.abs().max()))


if __name__ == "__main__":
    main()
