"""Stress: many dual-stream host batches (cooperative chain launches from two streams, graph replays); checks results stay identical."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ins = [(0.1 * torch.randn(64, 9, 96000)).pin_memory() for _ in range(2)]
outs = [torch.empty(64, 96000).pin_memory() for _ in range(K)]
with torch.no_grad():
    ref = [net.enhance(w.cuda()).cpu() for w in ins]
t0 = time.perf_counter()
for rep in range(5):
    net.enhance_host_batches([ins[i % 2] for i in range(K)], outs)
    for i in range(K):
        assert torch.equal(outs[i], ref[i % 2]), (rep, i)
print("ok: %d batches, %.1f s" % (5 * K, time.perf_counter() - t0))
