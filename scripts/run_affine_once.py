#!/usr/bin/env python
"""GPU: one forward + backward of the closed-form engine on (rays, z) rows at the C2 coarse-pass shape (32,768 rays x 64
samples, chunk 262,144) -- a small target for `ncu -k regex:k_aff` captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcnerf_b200 import ops  # noqa: E402
from pcnerf_b200.nof.networks import NOF_coarse  # noqa: E402

dev = torch.device("cuda:0")
n, S = int(os.environ.get("RAYS", 32768)), int(os.environ.get("S", 64))
gen = torch.Generator().manual_seed(1)
d = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=1)
rays = torch.cat([torch.zeros(n, 3), d, torch.rand(n, 9, generator=gen)], 1).contiguous().to(dev)
z = torch.sort(torch.rand(n, S, generator=gen) * 40.0 + 0.5, dim=1).values.contiguous().to(dev)
torch.manual_seed(0)
m = NOF_coarse().to(dev).train()
m.precision = "affine"
for _ in range(int(os.environ.get("REPS", 2))):
    p = m.forward_encoded(ops.LazyEnc(rays, z), 262144)
    p.sum().backward()
torch.cuda.synchronize()
print("ok", float(p.mean()))
