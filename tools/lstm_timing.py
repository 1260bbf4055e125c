import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    net.enhance(wave)
    for l in (0, 1):
        net.set_option("dbg_launch", -100 - l)
        net.enhance(wave)
        buf = (C.c_uint64 * 16)()
        net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
        r = list(buf); T = max(r[4], 1)
        print("layer", l, "per step: total %.0f | cell wait(half 0) %.0f  cell math %.0f (of which wait half 1: %.0f)  fence+arrive %.0f | mma wait(h_ready) %.0f wait(x_ready) %.0f" % (
            r[0] / T, r[1] / T, r[2] / T, r[6] / T, r[3] / T, r[5] / T, r[7] / T))
