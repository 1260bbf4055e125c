"""Diagnostics: in-kernel cycle counters of one tcgen05 conv launch at the bench configuration."""
import ctypes as C
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
names = ["prod_total", "prod_issue", "prod_consume", "prod_wait_empty", "units", "tiles", "mma_total", "mma_wait_full",
         "mma_wait_acc", "epi_total", "epi_wait_full"]
with torch.no_grad():
    net.enhance(wave)
    for idx in [int(x) for x in sys.argv[1:]] or [73]:
        net.set_option("dbg_launch", idx)
        net.enhance(wave)
        buf = (C.c_uint64 * 16)()
        net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
        d = dict(zip(names, list(buf)))
        u = max(d["units"], 1)
        print("launch", idx, {k: v for k, v in d.items()})
        print("   per unit: total %.0f issue %.0f consume %.0f (wait_empty %.0f) | mma wait_full %.0f wait_acc/tile %.0f | epi busy/tile %.0f" % (
            d["prod_total"] / u, d["prod_issue"] / u, d["prod_consume"] / u, d["prod_wait_empty"] / u, d["mma_wait_full"] / u,
            d["mma_wait_acc"] / max(d["tiles"], 1), (d["epi_total"] - d["epi_wait_full"]) / max(d["tiles"], 1)))
