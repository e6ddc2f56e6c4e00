"""Oracle (NumPy) restatement of the reference's back-end scoring arithmetic.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Pinned pieces (replayed against the imported reference in tests/golden):
  * two-covariance PLDA (score/pyplda/gaussian-plda-scoring.py:23-50, :65)
  * Bosaris-like EER (computeEER-like-Bosaris.py:50-107)
  * DET-interpolated EER (subtools2/egrecho/score/binary_metrics.py:11-120)
  * global mean / per-speaker mean / mean subtraction + length normalisation + dot product -- pinned through the
    reference tree's OWN restatement of those Kaldi steps, subtools2/egrecho/score/{utils.py:11-24, score.py:160-175,
    asnorm.py:30-98} (tests/golden/make_golden_egrecho.py -> egrecho_backend.npz): `compute_mean_stats` ==
    `global_mean` / `speaker_mean`; torch cosine_similarity of mean-subtracted vectors ==
    `cosine_trials(length_norm(subtract_global_mean(.)))`
  * AS-norm in both forms of the tree: pandas / ddof = 1 (score/ScoreNormalization.py) and NumPy / ddof = 0
    (subtools2/egrecho/score/asnorm.py:101-142, :283-352)
Still unpinned ("parity unpinned": the arithmetic lives in Kaldi binaries which the reference neither vendors nor
pins -- README.md:195-196 -- and which are absent here; restated from Kaldi's published semantics and anchored on
the call sites).  The egrecho pin above covers their arithmetic, not Kaldi's own rounding / file conventions:
  * ivector-normalize-length --scaleup=false  (score/process.sh:194-203)  [scaleup=true would multiply by sqrt(D)]
  * ivector-mean / ivector-subtract-global-mean (score/process.sh:156-192)
  * ivector-compute-dot-products              (score/score.sh:82-97)
  * compute-eer                               (computeEER.sh:21-22)
  * ivector-plda-scoring                      (score/score.sh:99-121; arithmetic pinned via plda_base.py instead)
  * apply-cmvn-sliding                        (extract_xvectors_for_pytorch.sh:106-111; oracle/frontend.py)
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------- pre-processing (Kaldi)
def length_norm(x):
    """ivector-normalize-length --scaleup=false: x / ||x||_2 per row (process.sh:194-203)."""
    x = np.asarray(x)
    n = np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=1, keepdims=True))
    return (x / n).astype(x.dtype)


def global_mean(x):
    """ivector-mean <vectors> <mean-out>: arithmetic mean of all rows (process.sh:169-179)."""
    return np.mean(np.asarray(x, dtype=np.float64), axis=0).astype(np.asarray(x).dtype)


def subtract_global_mean(x, mean):
    """ivector-subtract-global-mean (process.sh:181-192)."""
    return np.asarray(x) - np.asarray(mean)[None, :]


def speaker_mean(x, spk_index, num_spk):
    """ivector-mean ark:spk2utt ...: per-speaker average + utterance counts (process.sh:156-167).
    spk_index[i] = speaker id of row i.  Returns (means (S,D), num_utts (S,))."""
    x = np.asarray(x)
    sums = np.zeros((num_spk, x.shape[1]), dtype=np.float64)
    np.add.at(sums, spk_index, x.astype(np.float64))
    counts = np.bincount(spk_index, minlength=num_spk)
    return (sums / counts[:, None]).astype(x.dtype), counts


# ---------------------------------------------------------------- cosine
def cosine_trials(enroll, test, trial_e, trial_t):
    """ivector-compute-dot-products on already length-normalised vectors, one dot per listed
    trial (score.sh:82-97)."""
    return np.einsum("ij,ij->i", enroll[trial_e].astype(np.float64), test[trial_t].astype(np.float64))


def cosine_matrix(enroll, test):
    """All-pairs form of the same dot products (BASELINE config 4)."""
    return enroll.astype(np.float64) @ test.astype(np.float64).T


# ---------------------------------------------------------------- PLDA (pyplda)
def plda_calculate_var(between_var, within_var, mean):
    """CalculateVar, score/pyplda/gaussian-plda-scoring.py:31-50 (k = 0, :47-48).
    mean is a column (D,1), float64 throughout."""
    total_var_inv = np.linalg.inv(between_var + within_var)
    wc_add_2ac_inv = np.linalg.inv(within_var + 2 * between_var)
    wc_inv = np.linalg.inv(within_var)
    gamma = (-1 / 4) * (wc_add_2ac_inv + wc_inv) + (1 / 2) * total_var_inv
    lam = (-1 / 4) * (wc_add_2ac_inv - wc_inv)
    c = np.matmul(wc_add_2ac_inv - total_var_inv, mean)
    return gamma, lam, c, 0


def plda_smooth_within(within_var):
    """main(): within_var += 5e-5 * I before CalculateVar (gaussian-plda-scoring.py:65)."""
    return within_var + 5e-5 * np.eye(within_var.shape[0])


def plda_score_pair(e, t, gamma, lam, c, k):
    """PLDAScoring, gaussian-plda-scoring.py:23-29; e, t are (D,1) columns."""
    s = e.T @ lam @ t + t.T @ lam @ e + e.T @ gamma @ e + t.T @ gamma @ t + (e + t).T @ c + k
    return s[0][0]


def plda_score_matrix(enroll, test, gamma, lam, c, k=0):
    """The same bilinear form for all pairs at once:
    S = E(L+L^T)T^T + (diag(E G E^T) + E c) 1^T + 1 (diag(T G T^T) + T c)^T + k."""
    enroll = np.asarray(enroll, dtype=np.float64)
    test = np.asarray(test, dtype=np.float64)
    cvec = np.asarray(c, dtype=np.float64).reshape(-1)
    row = np.einsum("ij,jk,ik->i", enroll, gamma, enroll) + enroll @ cvec
    col = np.einsum("ij,jk,ik->i", test, gamma, test) + test @ cvec
    return enroll @ (lam + lam.T) @ test.T + row[:, None] + col[None, :] + k


# ---------------------------------------------------------------- EER
def eer_bosaris_like(scores, labels):
    """computeEER-like-Bosaris.py:50-107.  labels: 1 target / 0 nontarget.  Returns
    (eer, threshold).  The reference sorts [score, label] lists ascending (ties broken by
    label) and walks thresholds upward; restated with the same tie order."""
    scores = np.asarray(scores, dtype=np.float64)
    labels = np.asarray(labels).astype(np.int64)
    order = np.lexsort((labels, scores))
    s, l = scores[order], labels[order]
    num_p = int(l.sum())
    num_n = int(l.shape[0] - num_p)
    num_fa, num_fr = num_n, 0
    memory = None
    for i in range(s.shape[0]):
        if l[i] == 1:
            num_fr += 1
        else:
            num_fa -= 1
        far = num_fa * 1.0 / num_n
        frr = num_fr * 1.0 / num_p
        if far <= frr:
            lnow = abs(far - frr)
            lmem = abs(memory[0] - memory[1])
            if lnow <= lmem:
                return (far + frr) / 2, s[i]
            return (memory[0] + memory[1]) / 2, memory[2]
        memory = (far, frr, s[i])
    raise ValueError("no crossing")


def det_curve(scores, labels):
    """det_curve/_binary_clf_curve, subtools2/egrecho/score/binary_metrics.py:37-75,:151-210
    (sklearn-style thresholds at distinct scores)."""
    y_score = np.asarray(scores, dtype=np.float64)
    y_true = np.asarray(labels) == 1
    desc = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[desc], y_true[desc]
    distinct = np.where(y_score[1:] - y_score[:-1])[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true.astype(np.float64))[idx]
    fps = 1 + idx - tps
    thr = y_score[idx]
    fns = tps[-1] - tps
    p_count, n_count = tps[-1], fps[-1]
    first = fps.searchsorted(fps[0], side="right") - 1 if fps.searchsorted(fps[0], side="right") > 0 else None
    last = tps.searchsorted(tps[-1]) + 1
    sl = slice(first, last)
    return fps[sl][::-1] / n_count, fns[sl][::-1] / p_count, thr[sl][::-1]


def eer_det_interp(scores, labels):
    """compute_metrics -> eer_processor, binary_metrics.py:11-34, :77-112: linear
    interpolation of the FNR/FPR crossing.  Returns (eer, threshold)."""
    fprs, fnrs, thr = det_curve(scores, labels)
    i0 = np.flatnonzero(fnrs - fprs <= 0)[-1]
    i1 = np.flatnonzero(fnrs - fprs > 0)[0]
    d0 = fnrs[i0] - fprs[i0]
    d1 = fnrs[i1] - fprs[i1]
    scale = abs(d0) / (d1 - d0)
    return fnrs[i0] + scale * (fnrs[i1] - fnrs[i0]), thr[i0] + scale * (thr[i1] - thr[i0])


def min_dcf(scores, labels, p_target=0.01, c_miss=1.0, c_fa=1.0):
    """min_dcf_processor, binary_metrics.py:115-148."""
    fprs, fnrs, _ = det_curve(scores, labels)
    cost = c_miss * fnrs * p_target + c_fa * fprs * (1 - p_target)
    return float(np.min(cost) / min(c_miss * p_target, c_fa * (1 - p_target)))


def eer_kaldi(scores, labels):
    """Kaldi compute-eer step rule (computeEER.sh:21-22).  PARITY UNPINNED (restated from
    Kaldi's ComputeEer: sort both lists; first target index i with
    nontarget[N_non-1-floor(N_non*i/N_tar)] < target[i]; EER = i/N_tar)."""
    scores = np.asarray(scores, dtype=np.float64)
    labels = np.asarray(labels)
    tar = np.sort(scores[labels == 1])
    non = np.sort(scores[labels != 1])
    nt, nn = tar.shape[0], non.shape[0]
    i = 0
    for i in range(nt):
        ni = int(nn * i * 1.0 / nt)
        ni = min(max(nn - 1 - ni, 0), nn - 1)
        if non[ni] < tar[i]:
            break
    return i / nt, tar[i]


# ---------------------------------------------------------------- synthetic trials
def synthetic_speakers(num_spk, utts_per_spk, dim, seed, noise=0.5):
    """emb = spk_s + noise * eps so EER is non-trivial."""
    rng = np.random.RandomState(seed)
    spk = rng.standard_normal((num_spk, dim)).astype(np.float32)
    lab = np.repeat(np.arange(num_spk), utts_per_spk)
    emb = spk[lab] + noise * rng.standard_normal((lab.shape[0], dim)).astype(np.float32)
    return emb.astype(np.float32), lab


# ---------------------------------------------------------------- S-norm / AS-norm
def snorm_stats(cohort_scores, top_n=0, ddof=1):
    """Per-row mean and std (ddof=1, pandas' default) of the top_n largest cohort scores; top_n <= 0
    means all of them.  score/ScoreNormalization.py:93-98 (S-norm), :151-166 (AS-norm, cross_select
    false): sort descending, groupby(key).head(top_n), .mean()/.std().  ddof=0 is the other AS-norm of the tree,
    subtools2/egrecho/score/asnorm.py:128-140 (np.partition top-n, np.mean / np.std)."""
    s = np.sort(np.asarray(cohort_scores, dtype=np.float64), axis=1)[:, ::-1]
    if top_n and top_n > 0:
        s = s[:, :top_n]
    return s.mean(axis=1), s.std(axis=1, ddof=ddof)


def snorm_apply(scores, trial_e, trial_t, mean_e, std_e, mean_t, std_t):
    """normed = 0.5*((s-mu_e)/sd_e + (s-mu_t)/sd_t), ScoreNormalization.py:101-104 / :172-173."""
    s = np.asarray(scores, dtype=np.float64)
    return 0.5 * ((s - mean_e[trial_e]) / std_e[trial_e] + (s - mean_t[trial_t]) / std_t[trial_t])


# ---------------------------------------------------------------- trial histogram (fused consumer)
def trial_histogram(scores, is_target, lo, hi, nbins):
    """Bin rule of include/xvb200.h xvb_trial_histogram, in the same fp32 arithmetic:
    bin = 1 + floor((s - lo) * ((nbins-2)/(hi-lo))); below lo -> 0; at/above hi -> nbins-1.
    Returns (2, nbins) int64 [nontarget | target]."""
    s = np.asarray(scores, dtype=np.float32).reshape(-1)
    t = np.asarray(is_target).reshape(-1).astype(bool)
    inv_w = np.float32(np.float32(nbins - 2) / (np.float32(hi) - np.float32(lo)))
    x = (s - np.float32(lo)) * inv_w
    b = np.where(x < 0, 0, np.where(x < np.float32(nbins - 2), 1 + np.floor(np.maximum(x, 0)).astype(np.int64), nbins - 1))
    out = np.zeros((2, nbins), dtype=np.int64)
    np.add.at(out[0], b[~t], 1)
    np.add.at(out[1], b[t], 1)
    return out


def snorm_cross_apply(scores, trial_e, trial_t, enroll_cohort, test_cohort, top_n):
    """AS-norm with cross selection, score/ScoreNormalization.py:146-160,:171-173: for trial (e, t) the enroll
    statistics are taken over the cohort utterances that are the top_n of the TEST side, and vice versa;
    std is pandas' (ddof = 1)."""
    ec = np.asarray(enroll_cohort, dtype=np.float64)
    tc = np.asarray(test_cohort, dtype=np.float64)
    top_e = np.argsort(-ec, axis=1, kind="stable")[:, :top_n]
    top_t = np.argsort(-tc, axis=1, kind="stable")[:, :top_n]
    out = np.empty(len(scores), dtype=np.float64)
    for k, (e, t, s) in enumerate(zip(trial_e, trial_t, np.asarray(scores, dtype=np.float64))):
        ge = ec[e, top_t[t]]
        gt = tc[t, top_e[e]]
        out[k] = 0.5 * ((s - ge.mean()) / ge.std(ddof=1) + (s - gt.mean()) / gt.std(ddof=1))
    return out


# ---------------------------------------------------------------- back-end transforms (process.sh trainlda / trainwhiten / trainpcawhiten)
def zca_whitening(x, regularization=1e-6):
    """score/whiten/train_ZCA_Whitening.py ZCA.fit (:34-52) + write_matrix (:68-76): cov = X^T X / (n-1) with NO mean
    removal, SVD, whiten = U diag(1/sqrt(clip(S))) U^T, a zero bias column appended.  Pinned by tests/golden/whiten.npz
    (the script itself, run as process.sh:235-248 runs it)."""
    x = np.asarray(x, dtype=np.float64)
    cov = x.T @ x / (x.shape[0] - 1)
    u, s, _ = np.linalg.svd(cov)
    w = u @ np.diag(1.0 / np.sqrt(s.clip(regularization))) @ u.T
    return np.concatenate([w, np.zeros((w.shape[0], 1))], axis=1)


def lda_transform(x, spk, dim, total_covariance_factor=0.1, covariance_floor=1e-6):
    """Kaldi ivector-compute-lda --dim --total-covariance-factor (process.sh:218-228), restated from Kaldi's published
    source semantics -- PARITY UNPINNED (Kaldi absent): subtract the global mean; per speaker accumulate
    tot += X^T X, between += n * avg avg^T; total = tot/N, within = total - between/N; T whitens
    tcf*total + (1-tcf)*within (eigenvalues floored at floor * largest); eigenvectors of T between T^T, largest
    first, top `dim`; A = U_part^T T; last column = -A mean.  Returns (dim, D+1)."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=0)
    xc = x - mean
    d = x.shape[1]
    tot, btw, n = np.zeros((d, d)), np.zeros((d, d)), 0
    for s in np.unique(spk):
        rows = xc[np.asarray(spk) == s]
        tot += rows.T @ rows
        avg = rows.mean(axis=0)
        btw += rows.shape[0] * np.outer(avg, avg)
        n += rows.shape[0]
    total, between = tot / n, btw / n
    within = total - between
    s, u = np.linalg.eigh(total_covariance_factor * total + (1 - total_covariance_factor) * within)
    s, u = s[::-1], u[:, ::-1]
    s = np.maximum(s, covariance_floor * s[0])
    t = np.diag(s ** -0.5) @ u.T
    e, v = np.linalg.eigh(t @ between @ t.T)
    order = np.argsort(-e, kind="stable")
    a = v[:, order[:dim]].T @ t
    return np.concatenate([a, -(a @ mean)[:, None]], axis=1)


def pca_transform(x, dim=-1, normalize_variance=False, normalize_mean=True):
    """Kaldi est-pca --read-vectors=true with its defaults (process.sh:250-260) -- PARITY UNPINNED: eigenvectors of the
    centred covariance (divided by N), largest eigenvalue first, as rows; offset column = -P mean."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=0)
    cov = x.T @ x / x.shape[0] - np.outer(mean, mean)
    s, p = np.linalg.eigh(cov)
    order = np.argsort(-s, kind="stable")
    k = x.shape[1] if dim < 0 else dim
    a = p[:, order[:k]].T
    if normalize_variance:
        a = a / np.sqrt(s[order[:k]])[:, None]
    return np.concatenate([a, -(a @ mean)[:, None]], axis=1) if normalize_mean else a


def apply_affine(x, mat):
    """ivector-transform (process.sh:205-216): A x, plus the last column as offset when mat is (., D+1)."""
    x = np.asarray(x, dtype=np.float64)
    d = x.shape[1]
    return x @ mat[:, :d].T + (mat[:, d] if mat.shape[1] == d + 1 else 0.0)
