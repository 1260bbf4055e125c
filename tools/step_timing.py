"""Diagnostics: device time of whole steps at the bench configuration under option settings given as name=value."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
def run(k=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3): net.enhance(wave)
        torch.cuda.synchronize(); e0.record()
        for _ in range(k): net.enhance(wave)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("default: %.3f ms/step" % run())
for o in sys.argv[1:]:
    k, v = o.split("="); net.set_option(k, int(v))
    print("%s: %.3f ms/step" % (o, run()))
    net.set_option(k, {1: 3, 3: 1}[int(v)] if k.endswith('_passes') else 1 - int(v))
