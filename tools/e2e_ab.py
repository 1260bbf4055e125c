"""Diagnostics: end-to-end (host buffers) ms per batch of eab_enhance_host_batches under an option, A/B in one process."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
K = 10
ins = [(0.1 * torch.randn(64, 9, 96000)).pin_memory() for _ in range(2)]
outs = [torch.empty(64, 96000).pin_memory() for _ in range(2)]
I = [ins[i % 2] for i in range(K)]; O = [outs[i % 2] for i in range(K)]
name, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for rep in range(3):
    for v in (a, b):
        net.set_option(name, v)
        net.enhance_host_batches(I[:4], O[:4])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        net.enhance_host_batches(I, O)
        torch.cuda.synchronize()
        print("%s=%d: %.3f ms/batch" % (name, v, (time.perf_counter() - t0) * 1e3 / K), flush=True)
