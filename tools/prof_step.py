"""Run a few wave->wave steps at the bench configuration (for ncu captures; prints nothing that is a bench value)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--seconds", type=float, default=6.0)
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(a.batch, 9, int(a.seconds * 16000), device="cuda")
with torch.no_grad():
    for _ in range(a.steps):
        y = net.enhance(wave)
torch.cuda.synchronize()
print("ok", tuple(y.shape), net.last_launch_count(), "launches/step")
