#!/usr/bin/env python
"""Where the time of ONE 200-frame, 23-dim utterance goes (BASELINE configs[0] on the GPU path)."""
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200.model.xvector import Xvector  # noqa: E402
from oracle import nnet as onn  # noqa: E402


def med(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6


def main():
    sd = onn.make_state_dict(onn.xvector_spec(23), 101)
    m = Xvector(23, 10, training=False, extracted_embedding="far")
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    feats = onn.synthetic_feats(1, 200, 23, 5)[0]
    ex = m.extractor()
    x = torch.from_numpy(feats).cuda().unsqueeze(0).contiguous()
    out = {}
    out["plugin extract_embedding(ndarray) -> cpu tensor"] = med(lambda: m.extract_embedding(feats))
    out["ex.extract(device tensor) + synchronize"] = med(lambda: (ex.extract(x), torch.cuda.synchronize()))
    out["ex.extract(device tensor), launch only (async)"] = med(lambda: ex.extract(x))
    torch.cuda.synchronize()
    out["ex.extract_host(ndarray) (H2D + stack + D2H + sync in C)"] = med(lambda: ex.extract_host(feats[None]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        ex.extract(x)
    e1.record()
    torch.cuda.synchronize()
    out["device time per call when launches are queued back to back"] = e0.elapsed_time(e1) * 10.0
    xh = torch.from_numpy(feats)
    out["torch H2D of (200,23) from pageable + sync"] = med(lambda: (xh.to("cuda"), torch.cuda.synchronize()))
    y = torch.empty(512, device="cuda")
    out["torch D2H .cpu() of 512 floats"] = med(lambda: y.cpu())
    for k, v in out.items():
        print("%-70s %8.1f us" % (k, v))


if __name__ == "__main__":
    main()
