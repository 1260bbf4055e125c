"""Diagnostics (build with EAB_NVCC_EXTRA=-DEAB_CHAIN_DEBUG): cycle counters of CTA 0 of the last tcm_chain launch of an EaBNet step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
names = ["A load+store", "A mma", "A epilogue", "barrier1", "B load+store", "B mma", "B epilogue", "barrier2", "C load+store", "C mma", "C epilogue", "total", "tiles"]
for mode in [int(x) for x in sys.argv[1:]] or [1, 3]:
    net.set_option("tcm_chain", mode)
    with torch.no_grad():
        for _ in range(2): net.enhance(wave)
        net.set_option("dbg_launch", -200)
        net.enhance(wave)
    d = net.debug_counters()
    print("tcm_chain = %d (1: cluster per utterance, 3: cooperative grid barriers)" % mode)
    for n, v in zip(names, d):
        print("  %-14s %12d cycles  %6.1f%%" % (n, v, 100.0 * v / max(d[11], 1)), flush=True)
