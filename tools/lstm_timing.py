"""Diagnostics: per-step cycle counters of the two LSTM layers at the bench configuration (CTA 0, thread 0 / the MMA warp).
usage: [LSTM_KERNEL=umma|pp] python tools/lstm_timing.py [lstm_exp flags ...]"""
import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
kern = os.environ.get("LSTM_KERNEL", "")
if kern:
    net.set_option("lstm_pp", int(kern == "pp"))
with torch.no_grad():
    net.enhance(wave)
    for exp in [int(x) for x in sys.argv[1:]] or [0]:
        net.set_option("lstm_exp", exp)
        for l in (0, 1):
            net.set_option("dbg_launch", -100 - l)
            net.enhance(wave)
            buf = (C.c_uint64 * 16)()
            net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
            r = list(buf); T = max(r[4], 1)
            print("exp", exp, "layer", l, "per step: total %.0f | cell wait(half 0) %.0f  cell math %.0f (of which wait half 1: %.0f) | mma wait(h_ready) %.0f wait(x_ready) %.0f" % (
                r[0] / T, r[1] / T, r[2] / T, r[6] / T, r[5] / T, r[7] / T))
            if r[9]:
                print("      producer per step: wait h_ready %.0f  copy-out %.0f  wait x_free %.0f  publish %.0f" % (r[8] / T, r[9] / T, r[11] / T, r[10] / T))
