"""Diagnostics: cycle counters of the first layer's conv_tma launch at the bench configuration (CTA 0): MMA warp total,
issue / commit / wait-for-weights / wait-for-planes / wait-for-accumulator, epilogue total / wait / busy.
[EAB_TMA_K3=0] [EAB_TMA_NBUF=2] python tools/first_layer_counters.py"""
import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    net.enhance(wave)
    for idx in [0]:
        net.set_option("dbg_launch", idx)
        net.enhance(wave)
        torch.cuda.synchronize()
        buf = (C.c_uint64 * 16)()
        net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
        r = list(buf)
        if idx == 0:
            print("tiles %d | mma warp total %d  issue %d  commit %d  wait weights %d  wait planes %d  wait acc %d | epilogue total %d  wait %d  busy %d | nsb %d nbuf %d resident %d" % (
                r[2], r[0], r[1], r[3], r[7], r[6], r[5], r[8], r[9], r[10], r[13], r[14], r[15]))
        else:
            print(idx, r)
