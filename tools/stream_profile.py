"""Per-kernel-family CUDA-event times of ONE streaming step (256 streams, BN) - diagnostics."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = EaBNet(norm_type="BN").eval().cuda()
for o in sys.argv[2:]:
    k, v = o.split("="); net.set_option(k, int(v))
ses = net.stream(S)
hop = 0.1 * torch.randn(S, 9, 160, device="cuda")
for _ in range(5): ses.step(hop)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ses.step(hop)
e1.record(); torch.cuda.synchronize()
print("eager step: %.3f ms (%d launches)" % (e0.elapsed_time(e1) / 20, net.last_launch_count()))
net.profile(2); ses.step(hop); prof = net.profile_summary(); net.profile(0)
agg = collections.OrderedDict()
for k in prof:
    name = k["kernel"].split(":")[-1]
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += k["ms"]; a[2] += k["flops"]
tot = sum(a[1] for a in agg.values())
print("sum of kernels %.3f ms" % tot)
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-14s %3d launches %7.3f ms %5.1f%%  %6.1f TFLOP/s" % (n, a[0], a[1], 100 * a[1] / tot, a[2] / (a[1] * 1e-3) / 1e12 if a[1] else 0))
top = sorted(prof, key=lambda k: -k["ms"])[:12]
for k in top: print("   %-20s %.3f ms %.1f GFLOP" % (k["kernel"], k["ms"], k["flops"] / 1e9))
