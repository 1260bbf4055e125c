"""Small end-to-end runs for compute-sanitizer (memcheck): EaBNet and the post-filter wrapper, odd sizes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet, make_eabnet_with_postnet
from eabnet_b200.postnet import default_postnet_args
torch.manual_seed(0)
w = make_eabnet_with_postnet(default_postnet_args()).eval().cuda()
with torch.no_grad():
    for B, L in [(1, 320), (2, 20800), (3, 4805)]:
        y = w.enhance(0.1 * torch.randn(B, 9, L, device="cuda"))
        torch.cuda.synchronize()
        print(B, L, tuple(y.shape), bool(torch.isfinite(y).all()))
print("done")
