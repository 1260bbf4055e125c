"""Cycle counters of conv_raw_kernel's roles (CTA 0) for chosen layers of a 64 x 6 s forward.  Needs a diagnostics build:
    EAB_NVCC_EXTRA=-DEAB_RAW_DEBUG python -m eabnet_b200.build --force
Layer indices = order of the tensor-core conv launches in a forward (0 = first layer, 1..24 encoder, 25..49 decoder)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet, stft_compress  # noqa: E402

layers = [int(x) for x in sys.argv[1:]] or [1, 2, 9, 40, 41, 48, 49]
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
names = ["xf_total", "xf_wait_opnd", "xf_wait_raw", "xf_coef", "ld_wait_empty", "ld_total", "mma_wait_acc", "mma_wait_opnd",
         "mma_wait_b", "mma_total", "epi_wait_acc", "epi_total", "tiles", "groups"]
with torch.no_grad():
    spec = stft_compress(wave)
    net(spec)
    for k in layers:
        net.set_option("dbg_launch", k)
        net(spec)
        c = net.debug_counters()
        tiles = max(1, c[12])
        print("layer %2d: tiles/CTA %d groups %d | per tile:" % (k, c[12], c[13]),
              " ".join("%s=%d" % (n, c[i] // tiles) for i, n in enumerate(names[:12])), flush=True)
