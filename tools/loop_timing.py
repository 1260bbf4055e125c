import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
net = EaBNet().eval().cuda()
wave = 0.1 * torch.randn(64, 9, 96000, device="cuda")
def run(k, sync_each):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
    for _ in range(k):
        net.enhance(wave)
        if sync_each: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, (time.perf_counter() - t0) * 1e3 / k
with torch.no_grad():
    for _ in range(3): net.enhance(wave)
    t0 = time.perf_counter()
    for _ in range(5): net.enhance(wave)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 5
    torch.cuda.synchronize()
    print("cpu time to enqueue one step: %.2f ms" % cpu_ms)
    for rep in range(3):
        for k in (5, 10, 20):
            a = run(k, False); b = run(k, True)
            print("steps %2d  queued: %.2f ms/step (wall %.2f) | sync each: %.2f ms/step" % (k, a[0], a[1], b[0]))
