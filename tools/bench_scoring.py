#!/usr/bin/env python
"""Back-end scoring throughput on one B200 (BASELINE configs 4/5 scaled to one GPU's share):
all-pairs cosine row blocks and the PLDA score matrix, plus the per-trial kernels."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200 import ops  # noqa: E402
from asv_subtools_b200.score.backend import PldaModel  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = {}
    g = torch.Generator(device="cuda").manual_seed(1024)
    # --- config 4 share: 16384-row block of enroll x 131072 test, D=512 (row block of the 1M x 1M matrix)
    D = 512
    e = torch.nn.functional.normalize(torch.randn(16384, D, device="cuda", generator=g))
    t = torch.nn.functional.normalize(torch.randn(131072, D, device="cuda", generator=g))
    S = torch.empty(16384, 131072, device="cuda")
    import ctypes as C
    from asv_subtools_b200._lib import check, lib

    def cos():
        check(lib.xvb_cosine_matrix(C.c_void_p(e.data_ptr()), e.shape[0], C.c_void_p(t.data_ptr()), t.shape[0], D,
                                    C.c_void_p(S.data_ptr()), t.shape[0], None))
    ms = timed(cos, 3)
    n = e.shape[0] * t.shape[0]
    out["cosine_matrix_16384x131072x512"] = {"ms": ms, "scores_per_s": n / ms * 1e3, "algorithmic_tflops": 2 * D * n / ms * 1e-9,
                                             "write_gbs": 4 * n / ms * 1e-6}
    del S
    # --- config 5 share: PLDA, 131072 enroll x 10000 test, D=192
    D = 192
    rng = np.random.RandomState(0)
    a = rng.standard_normal((D, D))
    model = PldaModel(rng.standard_normal(D) * 0.1, a @ a.T / D + np.eye(D), np.eye(D) + 0.1 * (a + a.T) / np.sqrt(D))
    E = torch.randn(131072, D, device="cuda", generator=g)
    T = torch.randn(10000, D, device="cuda", generator=g)
    ms = timed(lambda: model.score_matrix(E, T), 3)
    n = E.shape[0] * T.shape[0]
    out["plda_matrix_131072x10000x192"] = {"ms": ms, "scores_per_s": n / ms * 1e3, "write_gbs": 4 * n / ms * 1e-6}
    # --- per-trial kernels: 10M listed trials
    te = torch.randint(0, E.shape[0], (10_000_000,), device="cuda", dtype=torch.int32)
    tt = torch.randint(0, T.shape[0], (10_000_000,), device="cuda", dtype=torch.int32)
    ms = timed(lambda: model.score_trials(E, T, te, tt), 3)
    out["plda_trials_10M"] = {"ms": ms, "trials_per_s": 1e7 / ms * 1e3}
    x = torch.randn(1_000_000, 512, device="cuda", generator=g)
    ms = timed(lambda: ops.center_length_norm(x, ops.column_mean(x)), 3)
    out["submean_norm_1Mx512"] = {"ms": ms, "gbs": 3 * x.numel() * 4 / ms * 1e-6}
    # --- config 4, fused consumer: all pairs of one set -> trial histogram -> EER; scores never stored
    import time
    from asv_subtools_b200.score import trial_histogram as th

    def wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t0) * 1e3

    D = 512
    for n in (131072, 262144, 1 << 20):
        spk = torch.randint(0, n // 16, (n,), device="cuda", dtype=torch.int32)
        base = torch.randn(n // 16, D, device="cuda", generator=g)
        x = base[spk.long()] + 2.0 * torch.randn(n, D, device="cuda", generator=g)
        del base
        x = ops.center_length_norm(x, ops.column_mean(x))
        trials = n * (n - 1) // 2
        key = "cosine_hist_sym_%d" % n
        out[key] = {"trials": trials}
        if n <= 262144:   # every score lands in a shared-memory counter: the expensive kind of pass
            ms = timed(lambda: ops.trial_histogram(x, spk, x, spk, -1.0, 1.0, 2048, symmetric=True), 2)
            out[key]["wide_window_pass_ms"] = ms
        th.zoom_eer(x[:4096].contiguous(), spk[:4096].contiguous(), passes=1, group=False)   # warm
        r, ms_total = wall(lambda: th.zoom_eer(x, spk, passes=2, pilot=32, group=False))
        lo, hi = r["lo"], r["hi"]
        ms = timed(lambda: ops.trial_histogram(x, spk, x, spk, lo, hi, 2048, symmetric=True), 2)
        out[key].update({"narrow_window_pass_ms": ms, "trials_per_s": trials / ms * 1e3,
                         "algorithmic_tflops": 2 * D * trials / ms * 1e-9, "executed_tflops": 3 * 2 * D * trials / ms * 1e-9,
                         "eer_pilot32_plus_2_passes": r["eer"], "eer_wall_ms": ms_total, "window": [lo, hi],
                         "counted": int(r["hist"].sum()), "count_ok": int(r["hist"].sum()) == trials})
        del x
    # --- config 5, fused: PLDA 131072 x 10000 -> histogram
    D = 192
    es = torch.randint(0, 5000, (E.shape[0],), device="cuda", dtype=torch.int32)
    ts = torch.randint(0, 5000, (T.shape[0],), device="cuda", dtype=torch.int32)
    ms = timed(lambda: model.score_histogram(E, es, T, ts, -200.0, 200.0), 3)
    hh = model.score_histogram(E, es, T, ts, -200.0, 200.0)
    Sm = model.score_matrix(E[:4096], T)
    out["plda_hist_131072x10000x192"] = {"ms": ms, "scores_per_s": E.shape[0] * T.shape[0] / ms * 1e3,
                                         "count_ok": int(hh.sum().item()) == E.shape[0] * T.shape[0],
                                         "score_range_sample": [float(Sm.min()), float(Sm.max())]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
