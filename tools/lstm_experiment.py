import os, sys, torch
sys.path.insert(0, "/root/repo")
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    for f in (0, 1, 2, 3):
        net.set_option("lstm_exp", f)
        for _ in range(2): net.enhance(wave)
        net.profile(2); net.enhance(wave); prof = net.profile_summary(); net.profile(0)
        print("flags", f, [round(k["ms"], 3) for k in prof if "lstm" in k["kernel"]])
