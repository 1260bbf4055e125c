"""Diagnostics: A/B of an option at the bench configuration with the step replayed from a CUDA graph (as bench.py times it)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
def run(k=20):
    g = net.graphed_enhance(wave)
    for _ in range(3): g.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k): g.step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
name, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for rep in range(3):
    for v in (a, b):
        net.set_option(name, v)
        print("%s=%d: %.3f ms/step" % (name, v, run()), flush=True)
