"""Cycle stamps of CTA 0 of chosen conv_umma launches of a streaming step (needs EAB_NVCC_EXTRA=-DEAB_UMMA_DEBUG)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet(norm_type="BN").eval().cuda()
ses = net.stream(256)
hop = 0.1 * torch.randn(256, 9, 160, device="cuda")
for _ in range(3): ses.step(hop)
names = ["entry", "prologue", "pdl_wait", "coef", "prod_end", "mma_end", "acc_full", "stored", "sync", "dealloc"]
for k in [int(x) for x in sys.argv[1:]] or [1, 5, 10, 20, 40, 60, 75]:
    net.set_option("dbg_launch", -300 - k)
    ses.step(hop); torch.cuda.synchronize()
    buf = (C.c_uint64 * 16)()
    net._native.lib.eab_debug_counters(net._native.h, C.byref(buf))
    r = list(buf)
    t0 = r[0]
    print("launch %2d:" % k, " ".join("%s=%d" % (n, r[i] - t0) for i, n in enumerate(names) if r[i]))
