"""Golden vectors for the sample-rate conversion of enhance.py:36-37: outputs of torchaudio.transforms.Resample itself
(default arguments) on seeded inputs -> tests/golden/resample_<orig>_<new>.npz.  Run where torchaudio is importable."""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import resample_oracle as R  # noqa: E402

for orig, new, L in ((44100, 16000, 2205), (48000, 16000, 1501), (8000, 16000, 700), (22050, 16000, 1000), (32000, 16000, 641)):
    g = torch.Generator().manual_seed(orig + new)
    x = (0.3 * torch.randn(3, L, generator=g)).float()
    y = torchaudio.transforms.Resample(orig, new)(x)
    mine = R.resample(x.numpy(), orig, new)
    print(orig, new, tuple(y.shape), "oracle vs torchaudio: %.3e" % float(np.abs(mine - y.numpy()).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resample_%d_%d.npz" % (orig, new)), x=x.numpy(), y=y.numpy(),
                        orig=orig, new=new)
