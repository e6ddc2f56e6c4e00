"""Fused score->histogram consumer (xvb_trial_histogram) and the EER read from it.

CPU: eer_from_histogram == the reference's DET-interpolated EER on the quantised scores; the zoom
driver converges to the exact value; two gloo ranks shard 256-row units and all-reduce counters.
GPU: counters bit-equal to a histogram of the materialised score matrix of the same kernel, equal to
the float64 oracle except for scores within rounding of a bin edge; symmetric/sharded/PLDA modes."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from asv_subtools_b200.score import metrics
from asv_subtools_b200.score import trial_histogram as th
from oracle import scoring as osc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _speakers(num_spk, per, dim, seed, noise=1.6):
    emb, lab = osc.synthetic_speakers(num_spk, per, dim, seed, noise=noise)
    emb = osc.length_norm(osc.subtract_global_mean(emb, osc.global_mean(emb))).astype(np.float32)
    perm = np.random.RandomState(seed + 1).permutation(emb.shape[0])
    return emb[perm], lab[perm].astype(np.int32)


def oracle_histogram(enroll, enroll_spk, test, test_spk, lo, hi, nbins, row_term=None, col_term=None, symmetric=False,
                     unit_first=0, unit_stride=1):
    """CPU stand-in with the signature of ops.trial_histogram (float64 scores, oracle bin rule)."""
    e, t = np.asarray(enroll, dtype=np.float64), np.asarray(test, dtype=np.float64)
    s = e @ t.T
    if row_term is not None:
        s = s + np.asarray(row_term, dtype=np.float64)[:, None]
    if col_term is not None:
        s = s + np.asarray(col_term, dtype=np.float64)[None, :]
    tgt = np.asarray(enroll_spk)[:, None] == np.asarray(test_spk)[None, :]
    rows = np.arange(e.shape[0])
    keep = ((rows // 256 - unit_first) % unit_stride == 0) & (rows // 256 >= unit_first)
    mask = np.broadcast_to(keep[:, None], s.shape).copy()
    if symmetric:
        mask &= np.arange(t.shape[0])[None, :] > rows[:, None]
    return osc.trial_histogram(s[mask], tgt[mask], lo, hi, nbins)


def test_eer_from_histogram_is_det_eer_of_quantised_scores():
    rng = np.random.RandomState(5)
    s = np.concatenate([rng.standard_normal(3000) * 0.1 + 0.25, rng.standard_normal(60000) * 0.1]).astype(np.float32)
    lab = np.r_[np.ones(3000, bool), np.zeros(60000, bool)]
    for nbins, lo, hi in ((2048, -1.0, 1.0), (64, -0.2, 0.4), (300, 0.1, 0.15)):
        hist = osc.trial_histogram(s, lab, lo, hi, nbins)
        assert hist.sum() == s.size and hist[1].sum() == 3000
        edges = th.bin_edges(lo, hi, nbins)
        w = edges[1] - edges[0]
        inv_w = np.float32(np.float32(nbins - 2) / (np.float32(hi) - np.float32(lo)))
        x = (s - np.float32(lo)) * inv_w
        b = np.where(x < 0, 0, np.where(x < nbins - 2, 1 + np.floor(np.maximum(x, 0)), nbins - 1)).astype(np.int64)
        q = np.r_[edges[0] - w, edges][b]                      # every score replaced by its bin's representative
        e_ref, t_ref = metrics.eer_det(q, lab)
        e, t, (b0, b1) = th.eer_from_histogram(hist, lo, hi)
        assert abs(e - e_ref) < 1e-12 and abs(t - t_ref) < 1e-9
        assert b0 <= t <= b1
        assert abs(th.min_dcf_from_histogram(hist, lo, hi) - metrics.min_dcf(q, lab)) < 1e-12
    # a coarse histogram is already close; the exact value needs the zoom
    e_exact, _ = metrics.eer_det(s, lab)
    assert abs(th.eer_from_histogram(osc.trial_histogram(s, lab, -1.0, 1.0, 2048), -1.0, 1.0)[0] - e_exact) < 2e-3


def test_zoom_converges_to_exact_eer():
    emb, spk = _speakers(40, 12, 32, 3)
    s = emb.astype(np.float64) @ emb.astype(np.float64).T
    iu = np.triu_indices(emb.shape[0], 1)
    e_exact, thr_exact = metrics.eer_det(s[iu].astype(np.float32), (spk[:, None] == spk[None, :])[iu])
    r1 = th.zoom_eer(emb, spk, passes=1, group=False, _histogram=oracle_histogram)
    r4 = th.zoom_eer(emb, spk, passes=4, group=False, _histogram=oracle_histogram)
    assert r1["hist"].sum() == iu[0].size
    assert abs(r1["eer"] - e_exact) < 5e-3
    assert abs(r4["eer"] - e_exact) < 1e-9 and abs(r4["threshold"] - thr_exact) < 1e-6, (r4["eer"], e_exact)
    rp = th.zoom_eer(emb, spk, passes=3, pilot=2, group=False, _histogram=oracle_histogram)   # locating pass on half the rows
    assert abs(rp["eer"] - e_exact) < 1e-9 and rp["hist"].sum() == iu[0].size
    assert r4["hi"] - r4["lo"] < 1e-2 and r4["hist"].sum() == iu[0].size   # out-of-window counts stay exact
    # enroll x test form
    r = th.zoom_eer(emb[:200], spk[:200], emb[200:], spk[200:], passes=4, group=False, _histogram=oracle_histogram)
    e2, _ = metrics.eer_det(s[:200, 200:].astype(np.float32).ravel(), (spk[:200, None] == spk[None, 200:]).ravel())
    assert abs(r["eer"] - e2) < 1e-9


@pytest.mark.timeout(120)
def test_world_size_2_gloo_histogram_allreduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
from asv_subtools_b200.score import trial_histogram as th
from test_trial_histogram import oracle_histogram, _speakers
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
emb, spk = _speakers(60, 10, 16, 9)        # 600 rows = 3 units of 256: rank 0 gets units 0 and 2, rank 1 unit 1
def hist(*a, **k):
    return torch.from_numpy(oracle_histogram(*a, **k))
one = th.zoom_eer(emb, spk, passes=3, group=False, _histogram=oracle_histogram)
both = th.zoom_eer(emb, spk, passes=3, rank=r, world=w, _histogram=hist)
assert np.array_equal(one["hist"], both["hist"]) and one["eer"] == both["eer"], (one["eer"], both["eer"])
print("rank", r, "ok", both["eer"])
dist.destroy_process_group()
''' % (ROOT, ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                         capture_output=True, text=True, timeout=110)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


# ------------------------------------------------------------------------------------------ GPU
def _cuda(*arrs):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


@pytest.mark.gpu
@pytest.mark.parametrize("n,dim", [(1000, 64), (2304, 512), (300, 192), (4700, 128)])
def test_gpu_symmetric_histogram_matches_own_matrix_and_oracle(n, dim):
    from asv_subtools_b200 import ops
    emb, spk = _speakers(n // 8, 8, dim, 21)     # a multiple of 4 rows (xvb_cosine_matrix)
    n = emb.shape[0]
    e, s = _cuda(emb, spk)
    for lo, hi, nbins in ((-1.0, 1.0, 2048), (0.05, 0.25, 512), (-0.01, 0.01, 4)):
        h = ops.trial_histogram(e, s, e, s, lo, hi, nbins, symmetric=True).cpu().numpy()
        assert h.sum() == n * (n - 1) // 2
        tgt = spk[:, None] == spk[None, :]
        iu = np.triu_indices(n, 1)
        assert h[1].sum() == tgt[iu].sum()
        # (1) same kernel arithmetic -> the materialised matrix bins identically, bit for bit
        S = ops.cosine_matrix(e, e).cpu().numpy()
        assert np.array_equal(h, osc.trial_histogram(S[iu], tgt[iu], lo, hi, nbins))
        # (2) float64 oracle: only scores within fp32 rounding of an edge may change bin
        S64 = emb.astype(np.float64) @ emb.astype(np.float64).T
        ho = osc.trial_histogram(S64[iu], tgt[iu], lo, hi, nbins)
        w = (hi - lo) / (nbins - 2)
        pos = (S64[iu] - lo) / w
        near = int((np.abs(pos - np.round(pos)) * w < 2e-5).sum())   # scores within rounding distance of an edge
        moved = np.abs(np.cumsum(h, 1) - np.cumsum(ho, 1)).max()
        assert moved <= near, (moved, near)


@pytest.mark.gpu
def test_gpu_zoom_eer_equals_exact_eer_of_the_score_matrix():
    from asv_subtools_b200 import ops
    emb, spk = _speakers(150, 12, 512, 33)                       # 1800 embeddings, 1.6 M trials
    e, s = _cuda(emb, spk)
    S = ops.cosine_matrix(e, e).cpu().numpy()
    iu = np.triu_indices(emb.shape[0], 1)
    lab = (spk[:, None] == spk[None, :])[iu]
    e_exact, thr = metrics.eer_det(S[iu], lab)
    r = th.zoom_eer(e, s, passes=4, group=False)
    assert abs(r["eer"] - e_exact) < 1e-9, (r["eer"], e_exact)
    e_b, _ = metrics.eer_bosaris(S[iu], lab)
    assert round(r["eer"] * 100, 3) == round(e_exact * 100, 3) and abs(r["eer"] - e_b) < 1e-4
    # float64 oracle scores: identical to 3 decimals (north star)
    S64 = (emb.astype(np.float64) @ emb.astype(np.float64).T)[iu]
    e64, _ = metrics.eer_det(S64, lab)
    assert abs(r["eer"] - e64) < 5e-6


@pytest.mark.gpu
def test_gpu_histogram_row_units_shard_and_accumulate():
    from asv_subtools_b200 import ops
    emb, spk = _speakers(130, 10, 128, 41)                       # 1300 rows -> 6 units of 256
    e, s = _cuda(emb, spk)
    full = ops.trial_histogram(e, s, e, s, -1.0, 1.0, 1024, symmetric=True)
    acc = torch.zeros_like(full)
    for r in range(4):
        ops.trial_histogram(e, s, e, s, -1.0, 1.0, 1024, symmetric=True, unit_first=r, unit_stride=4, out=acc)
    assert torch.equal(full, acc)
    # a shard with no rows is a no-op
    z = ops.trial_histogram(e, s, e, s, -1.0, 1.0, 1024, symmetric=True, unit_first=9, unit_stride=16)
    assert int(z.sum()) == 0


@pytest.mark.gpu
def test_gpu_histogram_enroll_by_test_with_plda_terms():
    from asv_subtools_b200 import ops
    rng = np.random.RandomState(7)
    E = rng.standard_normal((777, 192)).astype(np.float32)
    T = rng.standard_normal((1000, 192)).astype(np.float32)
    es = rng.randint(0, 50, 777).astype(np.int32)
    ts = rng.randint(0, 50, 1000).astype(np.int32)
    row = rng.standard_normal(777).astype(np.float32)
    col = rng.standard_normal(1000).astype(np.float32)
    e, t, esd, tsd, r, c = _cuda(E, T, es, ts, row, col)
    lo, hi, nbins = -60.0, 60.0, 2048
    h = ops.trial_histogram(e, esd, t, tsd, lo, hi, nbins, row_term=r, col_term=c).cpu().numpy()
    assert h.sum() == 777 * 1000 and h[1].sum() == (es[:, None] == ts[None, :]).sum()
    S = ops.plda_matrix(e, t, torch.eye(192, device="cuda"), r, c).cpu().numpy()   # L2 = I: same bilinear form
    ho = osc.trial_histogram(S.ravel(), (es[:, None] == ts[None, :]).ravel(), lo, hi, nbins)
    assert np.abs(np.cumsum(h, 1) - np.cumsum(ho, 1)).max() <= 16   # E.I re-rounds the rows: edge cases only
    with pytest.raises(RuntimeError):
        ops.trial_histogram(e, esd, t, tsd, lo, hi, 4096)
    with pytest.raises(RuntimeError):
        ops.trial_histogram(e, esd, t, tsd, lo, hi, nbins, symmetric=True)
