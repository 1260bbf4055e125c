"""Generate tests/golden/gag_*.npz by running the UNMODIFIED reference GaGNet / EaBNetWithPostNet (/root/reference)
on seeded inputs (build container only; see tools/make_golden.py for the scheme).

    python tools/make_golden_gag.py

GaGNet cases: weights = gagnet_oracle.make_gag_weights (name-seeded); inpt = compressed spectrum of microphone 0 of
oracle.make_wave, pre_x = compressed spectrum of microphone 1 scaled by 0.8 (a stand-in beamformer estimate); stored:
both inputs and the q stage outputs [q,B,2,F,T].  Wrapper case: EaBNetWithPostNet(args) with the argparse defaults of
train_distributed.py:277-318 on a 9-mic spectrum; stored: input, esti0_stft, esti_stft.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import eabnet_oracle as O  # noqa: E402
from oracle import gagnet_oracle as G  # noqa: E402

# name, cfg overrides, B, L, variant
CASES = [
    ("gag_default_b2_t21", {}, 2, 3200, "B"),
    ("gag_default_b1_t40_initA", {}, 1, 6240, "A"),
    ("gag_unet_bn_squeezed_tanh_b2_t13", {"is_u2": False, "norm_type": "BN", "is_squeezed": True, "acti_type": "tanh"}, 2, 1920, "B"),
    ("gag_noncausal_add_relu_b1_t17", {"is_causal": False, "intra_connect": "add", "acti_type": "relu", "dilas": (1, 2, 4), "q": 2}, 1, 2560, "B"),
]


def gag_inputs(B, L):
    wave, _ = O.make_wave(B, 2, L, seed=4321)
    spec = O.stft_compress(wave)                                   # [B,T,F,2,2]
    inpt = spec[..., 0, :].permute(0, 3, 1, 2).contiguous()
    pre = (0.8 * spec[..., 1, :]).permute(0, 3, 1, 2).contiguous()
    return inpt, pre


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    sys.path.insert(0, REF)
    from GaGNet import GaGNet
    import EaBNet as RE
    torch.set_num_threads(os.cpu_count())
    for name, over, B, L, variant in CASES:
        cfg = G.make_gag_cfg(**over)
        net = GaGNet(**{**cfg, "dilas": list(cfg["dilas"])}).eval()
        assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == list(G.gag_param_shapes(cfg).items()), name
        net.load_state_dict(G.make_gag_weights(cfg, 0, variant), strict=True)
        inpt, pre = gag_inputs(B, L)
        with torch.no_grad():
            outs = torch.stack(net(inpt, pre))
        np.savez_compressed(os.path.join(a.out, name + ".npz"), cfg=repr(over), B=B, L=L, variant=variant,
                            inpt=inpt.numpy(), pre=pre.numpy(), outs=outs.numpy())
        print(name, tuple(inpt.shape), "->", tuple(outs.shape), "absmax %.3f" % float(outs.abs().max()))
    # the wrapper enhance.py builds (EaBNet.py:127-148); make_gag_net there calls .cuda(), so assemble it by hand
    from eabnet_b200.postnet import default_postnet_args
    args = default_postnet_args()
    RE.make_gag_net = lambda ar: GaGNet(cin=2, k1=ar.gagnet_k1, k2=ar.gagnet_k2, c=ar.gagnet_c, kd1=ar.gagnet_kd1, cd1=ar.gagnet_cd1,
                                        d_feat=ar.gagnet_d_feat, p=ar.gagnet_p, q=ar.gagnet_q, dilas=ar.gagnet_dilas,
                                        fft_num=ar.gagnet_fft_num, is_u2=ar.gagnet_is_u2, is_causal=ar.gagnet_is_causal,
                                        is_squeezed=ar.gagnet_is_squeezed, acti_type=ar.gagnet_acti_type,
                                        intra_connect=ar.gagnet_intra_connect, norm_type=ar.gagnet_norm_type)
    wrap = RE.EaBNetWithPostNet(args).eval()
    sd = G.make_postnet_weights(None, None, 0, "B")
    assert list(wrap.state_dict().keys()) == list(sd.keys())
    wrap.load_state_dict(sd, strict=True)
    wave, _ = O.make_wave(1, 9, 3200, seed=1234)
    spec = O.stft_compress(wave)
    with torch.no_grad():
        r = wrap(spec)
    np.savez_compressed(os.path.join(a.out, "gag_wrapper_default_b1_t21.npz"), cfg=repr({}), B=1, L=3200, variant="B",
                        spec=spec.numpy(), esti0=r["esti0_stft"].numpy(), esti=r["esti_stft"].numpy(),
                        stages=torch.stack(r["esti1_stft_list"]).numpy())
    print("wrapper", tuple(spec.shape), "->", tuple(r["esti_stft"].shape), "absmax %.3f" % float(r["esti_stft"].abs().max()))


if __name__ == "__main__":
    main()
