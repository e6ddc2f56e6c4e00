#!/usr/bin/env python
"""Which 64-channel blocks of K reach the output of a split-K segment layer?  (debug aid)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402


def run(B, Cin, Cout, splitk):
    os.environ["XVB_SPLITK"] = splitk
    rng = np.random.RandomState(11)
    x = rng.standard_normal((B, 1, Cin)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 1)) * np.sqrt(2.0 / Cin)).astype(np.float32)
    xp = ops.split_f32(torch.from_numpy(x).cuda())
    wp = ops.pack_tdnn_weight(torch.from_numpy(w).cuda(), [0])
    _, yf = ops.tdnn_affine(xp, wp, Cout, [0], None, None, None, relu=False, out_planes=False, out_f32=True)
    torch.cuda.synchronize()
    got = yf.cpu().numpy().reshape(B, Cout).astype(np.float64)
    X, W = x[:, 0].astype(np.float64), w[:, :, 0].astype(np.float64)
    nb = (Cin + 63) // 64
    parts = np.stack([(X[:, j * 64:(j + 1) * 64] @ W[:, j * 64:(j + 1) * 64].T).ravel() for j in range(nb)], axis=1)
    coef, *_ = np.linalg.lstsq(parts, got.ravel(), rcond=None)
    print("B=%d Cin=%d Cout=%d splitk=%s blocks=%d coef=%s" % (B, Cin, Cout, splitk, nb, np.round(coef, 2).tolist()))


for args in ((9, 3000, 512, "1"), (9, 3000, 512, "0"), (256, 3000, 512, "1"), (9, 1536, 128, "1"), (130, 3072, 192, "1")):
    run(*args)
