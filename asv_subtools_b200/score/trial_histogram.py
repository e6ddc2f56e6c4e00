"""EER over score matrices too large to store: BASELINE config 4 (10^6 x 10^6 cosine trials) and
config 5 (10^6 x 10^4 PLDA trials).  The reference's flow -- write `enroll test score` lines
(score/score.sh:82-97), paste them against the trials file, sort (computeEER.sh:21-22 /
subtools2/egrecho/score/binary_metrics.py:11-120) -- cannot exist at 10^12 trials; here every score
is binned by trial class inside the GEMM epilogue (xvb_trial_histogram) and the EER is read from the
counters.  Rows shard over GPUs with one all-reduce of the (2, nbins) counters ("EER
needs an all-reduce of target/non-target histograms").

The histogram is a quantisation of the scores to bin edges, so `eer_from_histogram` is the
reference's DET-interpolated EER (binary_metrics.py `det_curve` + `eer_processor`) of the quantised
scores.  `zoom_eer` repeats the pass with the window narrowed to the bins around the crossing; the
counts outside the window stay exact (edge bins), so once the bins around the crossing hold at most
one distinct score each the value equals the reference's on the unquantised scores.
"""
import numpy as np
import torch

from .. import ops


def bin_edges(lo, hi, nbins):
    """Lower edge of bins 1..nbins-1 (bin 0 is everything below lo): edges[k] = lo + k*w."""
    w = (float(hi) - float(lo)) / (nbins - 2)
    return float(lo) + w * np.arange(nbins - 1, dtype=np.float64)


def det_points_from_histogram(hist, lo, hi):
    """DET points of the scores quantised to their bin: for every non-empty bin, in ascending order,
    (fpr, fnr, threshold) with fnr = targets strictly below the bin, fpr = nontargets in or above it --
    binary_metrics.py:37-75 on the quantised scores, including its trimming of the flat ends."""
    hist = np.asarray(hist, dtype=np.int64)
    non, tar = hist[0].astype(np.float64), hist[1].astype(np.float64)
    nbins = hist.shape[1]
    edges = bin_edges(lo, hi, nbins)
    w = edges[1] - edges[0]
    thr_all = np.r_[edges[0] - w, edges]            # a representative value for bin 0 (below the window)
    keep = (non + tar) > 0
    n_tar, n_non = tar.sum(), non.sum()
    if n_tar == 0 or n_non == 0:
        raise ValueError("need both target and nontarget trials")
    fns = (np.cumsum(tar) - tar)[keep]              # targets below the bin
    fps = (n_non - (np.cumsum(non) - non))[keep]    # nontargets at or above it
    tps = n_tar - fns
    thr = thr_all[keep]
    # descending-threshold view, trimmed like det_curve: drop leading points that repeat the first fps,
    # stop at the first point where all positives are counted
    fps_d, fns_d, tps_d, thr_d = fps[::-1], fns[::-1], tps[::-1], thr[::-1]
    r = fps_d.searchsorted(fps_d[0], side="right")
    first = r - 1 if r > 0 else None
    last = tps_d.searchsorted(tps_d[-1]) + 1
    sl = slice(first, last)
    return fps_d[sl][::-1] / n_non, fns_d[sl][::-1] / n_tar, thr_d[sl][::-1]


def eer_from_histogram(hist, lo, hi):
    """(eer, threshold, (lo_edge, hi_edge)) -- the bracket is the score interval that contains the
    FNR/FPR crossing, for the next zoom pass."""
    fpr, fnr, thr = det_points_from_histogram(hist, lo, hi)
    le = np.flatnonzero(fnr - fpr <= 0)
    gt = np.flatnonzero(fnr - fpr > 0)
    if le.size == 0 or gt.size == 0:
        raise ValueError("no FNR/FPR crossing inside the histogram")
    i0, i1 = le[-1], gt[0]
    d0, d1 = fnr[i0] - fpr[i0], fnr[i1] - fpr[i1]
    s = abs(d0) / (d1 - d0)
    eer = fnr[i0] + s * (fnr[i1] - fnr[i0])
    return float(eer), float(thr[i0] + s * (thr[i1] - thr[i0])), (float(thr[i0]), float(thr[i1]))


def min_dcf_from_histogram(hist, lo, hi, p_target=0.01, c_miss=1.0, c_fa=1.0):
    fpr, fnr, _ = det_points_from_histogram(hist, lo, hi)
    return float(np.min(c_miss * fnr * p_target + c_fa * fpr * (1 - p_target)) / min(c_miss * p_target, c_fa * (1 - p_target)))


def _reduce(hist, group):
    import torch.distributed as dist
    if group is not False and dist.is_available() and dist.is_initialized() and dist.get_world_size(group or None) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group or None)
    return hist


def zoom_eer(enroll, enroll_spk, test=None, test_spk=None, lo=-1.0, hi=1.0, nbins=2048, passes=2, row_term=None,
             col_term=None, rank=0, world=1, group=None, pilot=0, pilot_margin=8, _histogram=None):
    """EER of all enroll x test trials (test=None: all unordered pairs of `enroll`, each counted once).

    Every pass is one fused GEMM+histogram sweep over this rank's 256-row units (rank, rank+world, ...)
    followed by an all-reduce of the counters; pass k+1 narrows [lo, hi) to the crossing bracket of
    pass k (one bin of margin either side).  Returns dict(eer, threshold, hist, lo, hi, passes) where
    hist/lo/hi describe the last pass.

    pilot = P > 0 runs a cheap locating pass first on every P-th row unit only (the shared-memory
    counters are the bottleneck of a wide window, where every score lands in a bin; in a narrow window
    almost all scores are counted in registers): its crossing bracket, widened by `pilot_margin` bins
    either side, becomes the window of the first full pass.  A full pass whose crossing falls outside
    its window raises (ValueError from eer_from_histogram) -- rerun without the pilot."""
    histogram = _histogram or ops.trial_histogram
    symmetric = test is None
    if symmetric:
        test, test_spk = enroll, enroll_spk
    result = None
    if pilot and pilot > 1:
        h = histogram(enroll, enroll_spk, test, test_spk, lo, hi, nbins, row_term=row_term, col_term=col_term,
                      symmetric=symmetric, unit_first=rank * pilot, unit_stride=world * pilot)
        h = _reduce(h, group)
        hist = h.cpu().numpy() if isinstance(h, torch.Tensor) else np.asarray(h)
        _, _, (b0, b1) = eer_from_histogram(hist, lo, hi)
        w = (hi - lo) / (nbins - 2)
        lo, hi = max(b0, lo) - pilot_margin * w, min(b1, hi) + pilot_margin * w
    for k in range(passes):
        h = histogram(enroll, enroll_spk, test, test_spk, lo, hi, nbins, row_term=row_term, col_term=col_term,
                      symmetric=symmetric, unit_first=rank, unit_stride=world)
        h = _reduce(h, group)
        hist = h.cpu().numpy() if isinstance(h, torch.Tensor) else np.asarray(h)
        eer, thr, (b0, b1) = eer_from_histogram(hist, lo, hi)
        result = dict(eer=eer, threshold=thr, hist=hist, lo=lo, hi=hi, passes=k + 1)
        w = (hi - lo) / (nbins - 2)
        min_w = 6e-8 * max(abs(lo), abs(hi), 1e-3)        # about one fp32 ulp of the scores: no finer bins
        if w <= 1.01 * min_w:
            break
        nlo, nhi = max(b0, lo) - w, min(b1, hi) + w
        if (nhi - nlo) / (nbins - 2) < min_w:
            c, half = 0.5 * (nlo + nhi), 0.5 * min_w * (nbins - 2)
            nlo, nhi = c - half, c + half
        if not (nhi - nlo < hi - lo):
            break
        lo, hi = nlo, nhi
    return result
