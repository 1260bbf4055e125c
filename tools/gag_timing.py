"""Diagnostics (GPU box): device time of the post-filter at the bench configuration (64 x 6 s) + per-kernel-family profile."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import make_eabnet_with_postnet
from eabnet_b200.postnet import default_postnet_args
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = make_eabnet_with_postnet(default_postnet_args()).eval().cuda()
wave = 0.1 * torch.randn(B, 9, 96000, device="cuda")
def run(fn, k=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(2): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
t_all = run(lambda: w.enhance(wave))
t_eab = run(lambda: w.eabnet.enhance(wave))
print("EaBNet + GaGNet wave->wave: %.2f ms/step ; EaBNet alone: %.2f ms ; post-filter: %.2f ms ; launches %d" % (
    t_all, t_eab, t_all - t_eab, w.eabnet.last_launch_count()))
print("workspace GB", w.eabnet._ws.numel() / 1e9)
for o in [x for x in sys.argv[1:] if "=" in x]:
    k, v = o.split("="); w.postnet.set_option(k, int(v))
with torch.no_grad():
    spec = torch.randn(B, 601, 161, 9, 2, device="cuda") * 0.3
    est0 = w.eabnet(spec)
    inpt = spec[..., 0, :].permute(0, 3, 1, 2)
    w.postnet.profile(2 if "detail" in sys.argv else 1)
    w.postnet.forward_time_major(inpt, est0)
    prof = w.postnet.profile_summary()
    w.postnet.profile(0)
tot = sum(k["ms"] for k in prof)
for k in sorted(prof, key=lambda k: -k["ms"])[:40]:
    print("%-28s n=%4d  %8.3f ms  %5.1f%%  %8.1f GB/s  %8.1f TF/s" % (k["kernel"], k["launches"], k["ms"], 100 * k["ms"] / tot,
          k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] else 0, k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] else 0))
print("total %.3f ms" % tot)
if "dbg" in sys.argv:
    w.postnet.set_option("dbg_launch", -200)
    with torch.no_grad():
        w.postnet.forward_time_major(inpt, est0)
    d = w.postnet.debug_counters()
    names = ["A load+store", "A mma", "A epilogue", "barrier1", "B load+store", "B mma", "B epilogue", "barrier2", "C load+store", "C mma", "C epilogue", "total", "tiles"]
    for n, v in zip(names, d):
        print("%-14s %12d cycles  %6.1f%%" % (n, v, 100.0 * v / max(d[11], 1)))
