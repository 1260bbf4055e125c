"""Diagnostics (build with EAB_NVCC_EXTRA=-DEAB_CONV_EXPERIMENT): conv_tma time with parts of the kernel disabled.
flags: 1 no stores, 2 no statistics, 4 no TMEM loads, 8 no MMA issue."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
res = {}
with torch.no_grad():
    flags = [int(x) for x in sys.argv[1:]] or [0, 7, 8, 15]
    for f in flags:
        net.set_option("conv_exp", f)
        for _ in range(2): net.enhance(wave)
        net.profile(2); net.enhance(wave); prof = net.profile_summary(); net.profile(0)
        res[f] = prof
        print("flags", f, "conv_tma total %.3f" % sum(k["ms"] for k in prof if "conv_tma" in k["kernel"]))
idx = [i for i, k in enumerate(res[flags[0]]) if "conv_tma" in k["kernel"]]
idx.sort(key=lambda i: -res[flags[0]][i]["ms"])
print("launch".ljust(16), " ".join(("f=%d" % f).rjust(8) for f in flags), "   GB(alg)")
for i in idx[:28] + idx[60:64] + idx[-4:]:
    print(res[flags[0]][i]["kernel"].ljust(16), " ".join(("%.3f" % res[f][i]["ms"]).rjust(8) for f in flags), "   %.2f" % (res[flags[0]][i]["bytes"] / 1e9))
