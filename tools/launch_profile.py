"""Per-launch CUDA-event times of one step at the bench configuration (diagnostics)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
for o in sys.argv[1:]:
    k, v = o.split("="); net.set_option(k, int(v))
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    for _ in range(2): net.enhance(wave)
    net.profile(2); net.enhance(wave); prof = net.profile_summary(); net.profile(0)
tot = sum(k["ms"] for k in prof)
print("total %.3f ms over %d launches" % (tot, len(prof)))
for k in prof:
    tf = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0
    gb = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0
    print("%-22s %8.3f ms  %7.1f TFLOP/s %7.0f GB/s  %5.1f%%" % (k["kernel"], k["ms"], tf, gb, 100 * k["ms"] / tot))
