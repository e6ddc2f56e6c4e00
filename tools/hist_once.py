#!/usr/bin/env python
"""One fused GEMM+histogram sweep over all pairs of 32768 embeddings (for `ncu -k regex:tdnn_gemm`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402

n, d = 32768, 512
g = torch.Generator(device="cuda").manual_seed(3)
spk = torch.randint(0, n // 16, (n,), device="cuda", dtype=torch.int32)
x = ops.center_length_norm(torch.randn(n, d, device="cuda", generator=g), torch.zeros(d, device="cuda"))
for lo, hi in ((-1.0, 1.0), (0.09, 0.11)):
    h = ops.trial_histogram(x, spk, x, spk, lo, hi, 2048, symmetric=True)
torch.cuda.synchronize()
print("counted", int(h.sum()), "expected", n * (n - 1) // 2)
