"""Diagnostics: in-kernel cycle counters of one conv_plane launch at the bench configuration."""
import ctypes as C
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
names = ["prod_total", "prod_wait_empty", "tiles", "rows_total", "mma_total", "mma_wait_acc", "mma_wait_plane", "mma_wait_b",
         "epi_total", "epi_wait_full", "epi_tmem", "epi_store", "epi_stats"]
with torch.no_grad():
    net.enhance(wave)
    for idx in [int(x) for x in sys.argv[1:]]:
        net.set_option("dbg_launch", idx)
        net.enhance(wave)
        buf = (C.c_uint64 * 16)()
        net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
        d = dict(zip(names, list(buf)))
        n = max(d["tiles"], 1)
        print("launch", idx, "tiles", d["tiles"], "rows/buf", d["rows_total"], "| per tile: total %.0f | prod busy %.0f wait %.0f | mma wait acc %.0f plane %.0f b %.0f | epi wait %.0f tmem %.0f store %.0f stats %.0f" % (
            d["prod_total"] / n, (d["prod_total"] - d["prod_wait_empty"]) / n, d["prod_wait_empty"] / n, d["mma_wait_acc"] / n,
            d["mma_wait_plane"] / n, d["mma_wait_b"] / n, d["epi_wait_full"] / n, d["epi_tmem"] / n, d["epi_store"] / n, d["epi_stats"] / n))
