import ctypes as C, sys, os, torch
sys.path.insert(0, "/root/repo")
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    net.enhance(wave)
    for idx in [0, 1]:
        net.set_option("dbg_launch", idx)
        net.enhance(wave)
        torch.cuda.synchronize()
        buf = (C.c_uint64 * 16)()
        net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
        print(idx, list(buf))
