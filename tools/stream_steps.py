"""A few eager streaming steps (256 streams) for ncu launch lists - prints nothing that is a bench value."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eabnet_b200 import EaBNet
torch.manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = EaBNet(norm_type="BN").eval().cuda()
ses = net.stream(S)
hop = 0.1 * torch.randn(S, 9, 160, device="cuda")
for _ in range(3):
    ses.step(hop)
torch.cuda.synchronize()
print("ok", net.last_launch_count())
