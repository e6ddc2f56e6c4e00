#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 TDNN GEMM: runs a ladder of shapes against the fp32 SIMT layer
and prints where (rows / columns / K blocks) the two disagree.  Not a test; a debugging aid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402

CASES = [
    # B, T, Cin, Cout, context
    (1, 128, 64, 32, [0]),
    (1, 128, 64, 256, [0]),
    (1, 128, 128, 256, [0]),
    (1, 128, 16, 32, [0]),
    (2, 64, 64, 32, [0]),
    (16, 8, 64, 32, [0]),
    (1, 128, 64, 32, [-1, 0, 1]),
    (1, 100, 64, 32, [-2, 0, 2]),
    (3, 37, 80, 512, [-2, -1, 0, 1, 2]),
    (256, 200, 512, 512, [-2, 0, 2]),
]


def main():
    torch.manual_seed(0)
    for (B, T, Cin, Cout, ctx) in CASES:
        left, right, tot = ops.context_span(ctx)
        x = torch.randn(B, T, Cin, device="cuda")
        w = torch.randn(Cout, Cin, tot, device="cuda") / np.sqrt(Cin * len(ctx))
        bias = torch.randn(Cout, device="cuda") * 0.1
        xp = ops.split_f32(x)
        wp = ops.pack_tdnn_weight(w, ctx)
        try:
            _, y = ops.tdnn_affine(xp, wp, Cout, ctx, bias, out_planes=False, out_f32=True)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("CASE", (B, T, Cin, Cout, ctx), "EXCEPTION", e)
            break
        ref = ops.tdnn_affine_simt(x, w, ctx, bias)
        torch.cuda.synchronize()
        d = (y - ref).abs()
        err = (d.max() / ref.abs().max()).item()
        print("CASE B={} T={} Cin={} Cout={} ctx={}: max rel err {:.3e} {}".format(B, T, Cin, Cout, ctx, err,
                                                                                  "OK" if err < 3e-5 else "MISMATCH"))
        if err >= 3e-5:
            bad = d > 1e-3 * ref.abs().max()
            print("   bad fraction {:.4f}; bad rows (b,t) sample {}; bad cols sample {}".format(
                bad.float().mean().item(), bad.any(dim=2).nonzero()[:8].tolist(), bad.any(dim=0).any(dim=0).nonzero()[:16].flatten().tolist()))
            print("   y[0,0,:8]  ", y[0, 0, :8].tolist())
            print("   ref[0,0,:8]", ref[0, 0, :8].tolist())
            print("   y[0,1,:4]  ", y[0, 1, :4].tolist(), " ref[0,1,:4]", ref[0, 1, :4].tolist())
            print("   nan count", torch.isnan(y).sum().item(), "zero frac", (y == 0).float().mean().item())


if __name__ == "__main__":
    main()
