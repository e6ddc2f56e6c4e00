#!/usr/bin/env python
"""Throughput of the pooling kernels added for the snowdar blueprint's variants, at the BASELINE tensor (256 x 200 x 1500):
attention pooling with a head map (attentive / multi-head / xi-vector forms) and LDE.  CUDA events, 3 rotating inputs
(3 x 307 MB >> L2); algorithmic bytes = x read once + logits read once (+ output)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402


def timed(fn, n=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    B, T, C = 256, 200, 1500
    peak = 6582.5
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    xs = [torch.randn(B, T, C, device="cuda") for _ in range(3)]
    out = {"shape": [B, T, C], "hbm_peak_gbs": peak}
    for name, G, gdiv, kw in (("attentive (1 shared logit)", 1, C, {}), ("multi-head share (4 logits)", 4, C // 4, {}),
                              ("multi-head full / xi-vector (C logits)", C, 1, {})):
        lg = [torch.randn(B, T, (G + 7) // 8 * 8, device="cuda") for _ in range(3)]
        ms = timed(lambda i: ops.attn_head_stats_pool(lg[i % 3][..., :G], xs[i % 3], C, gdiv, **kw))
        nbytes = B * T * C * 4 + B * T * G * 4 + B * 2 * C * 4
        out[name] = {"ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    prior = torch.zeros(C, device="cuda")
    lg = [torch.randn(B, T, C, device="cuda") for _ in range(3)]
    ms = timed(lambda i: ops.attn_head_stats_pool(lg[i % 3], xs[i % 3], C, 1, prior_logit=prior, prior_x=prior, softplus2log=True))
    nbytes = 2 * B * T * C * 4
    out["xi-vector (prior + softplus2log)"] = {"ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    for Cc, K in ((1500, 64), (512, 64), (128, 8)):
        x2 = [torch.randn(B, T, Cc, device="cuda") * 0.6 for _ in range(3)]
        mu = torch.randn(Cc, K, device="cuda") * 0.6
        nb = -torch.full((K,), 0.01, device="cuda")
        ms = timed(lambda i: ops.lde_pool(x2[i % 3], mu, nb), n=6)
        flop = 2 * B * T * Cc * K * 3            # (sub, fma) for the distances + fma for the encoding
        out["lde C=%d K=%d" % (Cc, K)] = {"ms": ms, "fp32_tflops": flop / ms / 1e9, "x_read_twice_gbs": 2 * B * T * Cc * 4 / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
