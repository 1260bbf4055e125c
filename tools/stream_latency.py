"""Diagnostics: per-step latency of the streaming path (256 streams x one 10 ms hop), plain launches vs CUDA graph."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
from eabnet_b200.model import EaBNetStream
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
torch.manual_seed(0)
net = EaBNet(norm_type="BN").eval().cuda()
hop_host = (0.1 * torch.randn(S, 9, 160)).pin_memory()
out_host = torch.empty(S, 160).pin_memory()
hop = torch.empty(S, 9, 160, device="cuda")
out = torch.empty(S, 160, device="cuda")
for graph in (False, True):
    ses = EaBNetStream(net, S, graph=graph)
    for _ in range(20):
        ses.step(hop, out)
    torch.cuda.synchronize()
    wall, dev = [], []
    for _ in range(N):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        hop.copy_(hop_host, non_blocking=True)
        e0.record()
        ses.step(hop, out)
        e1.record()
        out_host.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(e0.elapsed_time(e1))
    wall.sort(); dev.sort()
    print("graph=%d S=%d launches/step=%d | host-to-host p50 %.3f p99 %.3f ms | device p50 %.3f p99 %.3f ms" % (
        graph, S, net.last_launch_count(), wall[N // 2], wall[int(N * 0.99)], dev[N // 2], dev[int(N * 0.99)]))
if len(sys.argv) > 3:
    ses = EaBNetStream(net, S)
    for _ in range(3): ses.step(hop, out)
    net.profile(2); ses.step(hop, out); prof = net.profile_summary(); net.profile(0)
    tot = sum(k["ms"] for k in prof)
    print("total %.3f ms over %d launches" % (tot, len(prof)))
    for k in sorted(prof, key=lambda k: -k["ms"])[:25]:
        print("%-22s %8.4f ms %7.1f TFLOP/s" % (k["kernel"], k["ms"], k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0))
