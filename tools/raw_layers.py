"""Per-launch CUDA-event times of the conv_raw launches of one 64 x 6 s step next to their shared-memory plans
(EAB_RAW_VERBOSE=1 makes the launcher print one line per launch to stderr)."""
import os
import sys

os.environ["EAB_RAW_VERBOSE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet  # noqa: E402

torch.manual_seed(0)
net = EaBNet().eval().cuda()
for o in sys.argv[1:]:
    k, v = o.split("=")
    net.set_option(k, int(v))
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
with torch.no_grad():
    for _ in range(2):
        net.enhance(wave)
    torch.cuda.synchronize()
    sys.stderr.write("==== profiled step\n")
    net.profile(2)
    net.enhance(wave)
    prof = net.profile_summary()
    net.profile(0)
tot = sum(k["ms"] for k in prof)
print("total %.3f ms over %d launches" % (tot, len(prof)))
for k in prof:
    gb = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0
    print("%-12s %8.3f ms %7.0f GB/s (algorithmic) %6.1f TFLOP/s" % (k["kernel"], k["ms"], gb, k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0))
