"""Diagnostics (GPU box): stage-wise comparison of the GaGNet CUDA path with the CPU oracle through the debug taps."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import GaGNet
from oracle import eabnet_oracle as E
from oracle import gagnet_oracle as G

over = eval(sys.argv[1]) if len(sys.argv) > 1 else {}
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2, 21)
cfg = G.make_gag_cfg(**over)
sd = G.make_gag_weights(cfg, 0, "B")
net = GaGNet(**cfg).eval(); net.load_state_dict(sd); net.cuda()
for o in sys.argv[4:]:
    k, v = o.split("="); net.set_option(k, int(v))
g = torch.Generator().manual_seed(1)
x, pre = torch.randn(B, 2, T, 161, generator=g), 0.5 * torch.randn(B, 2, T, 161, generator=g)
with torch.no_grad():
    got = net.forward_time_major(x.cuda(), pre.cuda()).cpu().transpose(-2, -1)
    ref = torch.stack(G.gag_forward(sd, x, pre, cfg))
    print("stages max|err|:", [float(e) for e in (got - ref).abs().flatten(1).max(1).values], "scale", float(ref.abs().max()),
          "launches", net.last_launch_count())
    # stage-wise
    feat4 = G.gag_encoder(sd, torch.cat((x, pre), 1), cfg)                       # [B,64,T,Fb]
    Fb = feat4.shape[-1]
    tap = net.debug_tap("en.4", (B, T, Fb, 64)).cpu()
    print("encoder bottleneck err", float((tap - feat4.permute(0, 2, 3, 1)).abs().max()), "scale", float(feat4.abs().max()))
    feat = feat4.transpose(-2, -1).reshape(B, -1, T)
    prer = pre.transpose(-2, -1).contiguous()
    p = "gags.0.glance_block"
    xin = G._gated_in(sd, p, feat, prer)
    tap = net.debug_tap("g.in_g.0", (B, T, 1, cfg["d_feat"])).cpu()[:, :, 0]
    print("glance in_conv err", float((tap - xin.transpose(1, 2)).abs().max()), "scale", float(xin.abs().max()))
    xt = G._tcm_groups(sd, p + ".tcn_g", xin, cfg)
    tap = net.debug_tap("g.tcn_g.0", (B, T, 1, cfg["d_feat"])).cpu()[:, :, 0]
    print("glance tcn err", float((tap - xt.transpose(1, 2)).abs().max()), "scale", float(xt.abs().max()))
    gl = F.conv1d(xt, sd[p + ".linear_g.0.weight"], sd[p + ".linear_g.0.bias"])
    for C in (256, 161):
        try:
            tap = net.debug_tap("g.gain.0", (B, T, 1, C)).cpu()[:, :, 0, :161]
            print("gain (raw) err", float((tap - gl.transpose(1, 2)).abs().max()), "scale", float(gl.abs().max()))
            break
        except RuntimeError as e:
            pass
