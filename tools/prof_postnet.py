"""Run a few wave->wave steps WITH the post-filter at the bench configuration (for ncu captures; prints no bench value)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import make_eabnet_with_postnet
from eabnet_b200.postnet import default_postnet_args
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
w = make_eabnet_with_postnet(default_postnet_args()).eval().cuda()
wave = 0.1 * torch.randn(B, 9, 96000, device="cuda")
with torch.no_grad():
    for _ in range(steps):
        y = w.enhance(wave)
torch.cuda.synchronize()
print("ok", tuple(y.shape), w.eabnet.last_launch_count(), "launches/step")
