#!/usr/bin/env python
"""ECAPA-TDNN c1024 throughput (BASELINE configs[2]: 80-d fbank, 300-frame chunks, batch 128) --
a side measurement, not the bench.py contract line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200.model.ecapa_tdnn_xvector import ECAPA_TDNN  # noqa: E402
from oracle import nnet as onn  # noqa: E402

CANON = dict(training=False, extracted_embedding="near",
             ecapa_params={"channels": 1024, "embd_dim": 192, "mfa_conv": 1536,
                           "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}},
             fc2_params={"nonlinearity": "", "bn": True, "bn_params": {"momentum": 0.5, "affine": False,
                                                                        "track_running_stats": True}})


def main():
    B, T, F = 128, 300, 80
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    m = ECAPA_TDNN(F, 10, **CANON)
    m.load_state_dict(onn.make_state_dict(onn.ecapa_spec(F), 201), strict=True)
    m.cuda().eval()
    if "--profile" in sys.argv:
        os.environ["XVB_ECAPA_NATIVE"] = "0"   # per-op CUDA events live in the op-by-op Python twin
    ex = m.extractor()
    xs = [torch.randn(B, T, F, device="cuda") for _ in range(4)]
    for i in range(3):
        ex.extract(xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = ex.extract(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flop = 25701908 * B * T
    if "--profile" in sys.argv:
        import collections
        from asv_subtools_b200.model import ecapa_tdnn_xvector as mod
        agg = collections.OrderedDict()
        for rep in range(5):
            mod._PROFILE = []
            ex.extract(xs[rep % 4])
            torch.cuda.synchronize()
            evs = mod._PROFILE
            for (l0, e0_), (l1, e1_) in zip(evs[:-1], evs[1:]):
                agg.setdefault(l1, []).append(e0_.elapsed_time(e1_))
        mod._PROFILE = None
        tot = 0.0
        for k, v in agg.items():
            per_call = sorted(v)[len(v) // 2]
            n = len(v) // 5
            tot += per_call * n
            print("%-28s x%-3d %8.1f us each %9.1f us total" % (k, n, per_call * 1e3, per_call * n * 1e3))
        print("sum %.1f us" % (tot * 1e3))
    print(json.dumps({"workload": "ECAPA-TDNN c1024, 80-d fbank, 300-frame chunks, batch 128",
                      "ms_per_step": ms, "frames_per_s": B * T / (ms * 1e-3),
                      "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "finite": bool(torch.isfinite(out).all()),
                      "extractor": type(ex).__name__, "launches": getattr(ex, "last_launches", None)}))


if __name__ == "__main__":
    main()
