"""Oracle (NumPy float64) restatement of the reference's PLDA training.  TEST INFRASTRUCTURE ONLY.

  plda_stats      : PldaStats.add_samples, score/pyplda/plda_base.py:50-66 (per class: n, mean, weighted
                    scatter about the class mean; classes sorted by n as PldaEstimation requires, :68-81,:233-236)
  plda_estimate   : PldaEstimation.estimate / estimate_one_iter, :248-300 -- the literal per-class loop with one
                    D x D inverse per class
  plda_estimate_grouped : the same EM step written the way the GPU path computes it -- in the basis that
                    makes within_var the identity and between_var diagonal (get_output's transform, :302-335)
                    every per-class matrix is diagonal and depends on n only; used to check the algebra.
Pinned by tests/golden/plda_train.npz, produced by the imported reference (tests/golden/make_golden_plda.py)."""
import numpy as np


def plda_stats(emb, spk, weights=None):
    """emb (N, D), spk (N,) int labels -> dict(n (S,), mean (S, D), weight (S,), offset_scatter (D, D), sum (D,),
    class_weight, example_weight), classes in ascending order of n (stable), as the reference sorts them."""
    emb = np.asarray(emb, dtype=np.float64)
    ids = np.unique(spk)
    groups = [emb[spk == s] for s in ids]
    w = np.ones(len(ids)) if weights is None else np.asarray(weights, dtype=np.float64)
    order = np.argsort([g.shape[0] for g in groups], kind="stable")
    groups = [groups[i] for i in order]
    w = w[order]
    D = emb.shape[1]
    scatter, total = np.zeros((D, D)), np.zeros(D)
    n, means = [], []
    for g, wk in zip(groups, w):
        m = g.mean(axis=0)
        scatter += wk * (g.T @ g) - g.shape[0] * wk * np.outer(m, m)
        total += wk * m
        n.append(g.shape[0])
        means.append(m)
    n = np.asarray(n, dtype=np.float64)
    return dict(n=n, mean=np.asarray(means), weight=w, offset_scatter=scatter, sum=total, class_weight=float(w.sum()),
                example_weight=float((w * n).sum()))


def plda_estimate(stats, num_em_iters=10):
    """-> (mean (D,), within_var, between_var), the three vectors plda_write stores (:337-342)."""
    D = stats["mean"].shape[1]
    within, between = np.eye(D), np.eye(D)
    gmean = stats["sum"] / stats["class_weight"]
    for _ in range(num_em_iters):
        w_stats = stats["offset_scatter"].copy()
        w_count = stats["example_weight"] - stats["class_weight"]
        b_stats, b_count = np.zeros((D, D)), 0.0
        w_inv, b_inv = np.linalg.inv(within), np.linalg.inv(between)
        for n, mk, wk in zip(stats["n"], stats["mean"], stats["weight"]):
            mix = np.linalg.inv(b_inv + n * w_inv)
            m = mk - gmean
            w = mix @ (n * (w_inv @ m))
            r = m - w
            b_stats += wk * mix + wk * np.outer(w, w)
            b_count += wk
            w_stats += wk * n * mix + wk * n * np.outer(r, r)
            w_count += wk
        within, between = w_stats / w_count, b_stats / b_count
    return gmean, within, between


def diagonalising_transform(within, between):
    """T with T W T^T = I and T B T^T = diag(psi)  (compute_normalizing_transform + eigh, :302-335)."""
    t1 = np.linalg.inv(np.linalg.cholesky(within))
    psi, u = np.linalg.eigh(t1 @ between @ t1.T)
    return u.T @ t1, psi


def plda_estimate_grouped(stats, num_em_iters=10):
    D = stats["mean"].shape[1]
    within, between = np.eye(D), np.eye(D)
    gmean = stats["sum"] / stats["class_weight"]
    mc = stats["mean"] - gmean
    n, wk = stats["n"], stats["weight"]
    for _ in range(num_em_iters):
        t, psi = diagonalising_transform(within, between)
        tinv = np.linalg.inv(t)
        u = mc @ t.T                                         # class means in the diagonal basis
        mixd = psi[None, :] / (1.0 + n[:, None] * psi[None, :])   # diag of (B^-1 + n W^-1)^-1 there
        what = n[:, None] * mixd * u
        r = u - what
        b_t = np.diag((wk[:, None] * mixd).sum(0)) + (what * wk[:, None]).T @ what
        w_t = np.diag((wk[:, None] * n[:, None] * mixd).sum(0)) + (r * (wk * n)[:, None]).T @ r
        between = tinv @ b_t @ tinv.T / wk.sum()
        within = (tinv @ w_t @ tinv.T + stats["offset_scatter"]) / (stats["example_weight"] - stats["class_weight"] + wk.sum())
    return gmean, within, between


def synthetic_plda_data(num_spk, dim, seed, min_utts=3, max_utts=9, spread=1.5, conditioned=False):
    """conditioned=True: within / between factors with singular values in [0.6, 1.4] (a square Gaussian matrix has
    a condition number in the thousands, which turns fp32-level differences of the covariances into visible
    score differences through the inverses of the scoring formula)."""
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    b = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    if conditioned:
        qa, _ = np.linalg.qr(a)
        qb, _ = np.linalg.qr(b)
        a = qa * rng.uniform(0.6, 1.4, dim)[None, :]
        b = qb * rng.uniform(0.6, 1.4, dim)[None, :]
    counts = rng.randint(min_utts, max_utts + 1, num_spk)
    spk = np.repeat(np.arange(num_spk), counts)
    centres = rng.standard_normal((num_spk, dim)) @ b.T * spread + 0.3
    emb = centres[spk] + rng.standard_normal((spk.shape[0], dim)) @ a.T
    perm = rng.permutation(spk.shape[0])
    return emb[perm].astype(np.float32), spk[perm].astype(np.int32)


def synthetic_adaptation_data(n, dim, seed):
    """Unlabelled in-domain vectors: another covariance, another mean."""
    rng = np.random.RandomState(seed)
    a = rng.standard_normal((dim, dim)) / np.sqrt(dim) * 1.7
    return (rng.standard_normal((n, dim)) @ a.T + 0.8).astype(np.float32)


def coral_adapt(mean, within, between, adapt, mean_diff_scale=1.0):
    """CORAL.update_plda, score/pyplda/ivector-adapt-plda-coral.py:40-84 (eigh is ascending already, so its
    sort_svd is a no-op): A = Var^(1/2) (W + B)^(-1/2); W, B -> A W A^T, A B A^T; mean -> data mean."""
    x = np.asarray(adapt, dtype=np.float64)
    m = x.mean(axis=0).reshape(-1, 1)
    var = x.T @ x / x.shape[0] - m @ m.T
    d = m - np.asarray(mean, dtype=np.float64).reshape(-1, 1)
    var = var + mean_diff_scale * (d @ d.T)
    eo, qo = np.linalg.eigh(within + between)
    ei, qi = np.linalg.eigh(var)
    a = (qi @ np.diag(np.sqrt(ei)) @ qi.T) @ (qo @ np.diag(1.0 / np.sqrt(eo)) @ qo.T)
    return m.reshape(-1), a @ within @ a.T, a @ between @ a.T


def covariance_regulariser(base, target):
    """The term shared by CORAL+ / LIP-reg / CIP-reg (ivector-adapt-plda-coralplus.py:76-84, -lip-reg.py:35-41,
    -cip-reg.py:113-121): with B the simultaneous diagonaliser of (base, target) -- B^T base B = I,
    B^T target B = diag(E) -- return B^-T max(0, diag(E) - I) B^-1, i.e. the part of `target` that exceeds `base`."""
    base = np.asarray(base, dtype=np.float64)
    lam, q = np.linalg.eigh(base)
    t = np.diag(1.0 / np.sqrt(lam)) @ q.T                       # base -> I
    e, p = np.linalg.eigh(t @ np.asarray(target, dtype=np.float64) @ t.T)
    b_inv = np.linalg.inv(q @ np.diag(1.0 / np.sqrt(lam)) @ p)
    return b_inv.T @ np.maximum(0.0, np.diag(e) - np.eye(base.shape[0])) @ b_inv


def coralplus_adapt(mean, within, between, adapt, within_scale=0.8, between_scale=0.8, mean_diff_scale=1.0):
    """CORALPlus.update_plda, ivector-adapt-plda-coralplus.py:40-96: the CORAL pseudo in-domain covariances only ADD
    their excess over the out-of-domain ones."""
    m, s_w, s_b = coral_adapt(mean, within, between, adapt, mean_diff_scale)
    return (m, within + within_scale * covariance_regulariser(within, s_w),
            between + between_scale * covariance_regulariser(between, s_b))


def lip_adapt(out_model, in_model, weight=0.4):
    """LIP.interpolation, ivector-adapt-plda-lip.py:25-34.  Models are (mean, within, between); mean of the in-domain one."""
    return (in_model[0], weight * out_model[1] + (1 - weight) * in_model[1], weight * out_model[2] + (1 - weight) * in_model[2])


def lipreg_adapt(out_model, in_model, weight=0.6):
    """LIPReg.interpolation, ivector-adapt-plda-lip-reg.py:26-49."""
    return (in_model[0], in_model[1] + (1 - weight) * covariance_regulariser(in_model[1], out_model[1]),
            in_model[2] + (1 - weight) * covariance_regulariser(in_model[2], out_model[2]))


def cip_adapt(out_model, in_model, adapt, weight=0.5):
    """CORAL.update_plda + CIP.interpolation, ivector-adapt-plda-cip.py:38-77, :113-121."""
    _, s_w, s_b = coral_adapt(out_model[0], out_model[1], out_model[2], adapt)
    return (in_model[0], weight * s_w + (1 - weight) * in_model[1], weight * s_b + (1 - weight) * in_model[2])


def cipreg_adapt(out_model, in_model, adapt, weight=0.5):
    """CORAL.update_plda + CIPReg.interpolation, ivector-adapt-plda-cip-reg.py:109-128."""
    _, s_w, s_b = coral_adapt(out_model[0], out_model[1], out_model[2], adapt)
    return (in_model[0], in_model[1] + weight * covariance_regulariser(in_model[1], s_w),
            in_model[2] + weight * covariance_regulariser(in_model[2], s_b))


def unsupervised_adapt(mean, within, between, adapt, within_scale=0.3, between_scale=0.7, mean_diff_scale=1.0):
    """PldaUnsupervisedAdaptor.update_plda, plda_base.py:368-485 (the reference's restatement of Kaldi's
    ivector-adapt-plda): in the space where the model's total covariance is I, every direction in which the adaptation
    data's covariance exceeds 1 hands its excess to the within / between covariances in the given proportions.
    Returns (mean, within, between) of the adapted model in the original space."""
    x = np.asarray(adapt, dtype=np.float64)
    m = x.mean(axis=0).reshape(-1, 1)
    var = x.T @ x / x.shape[0] - m @ m.T
    d = m - np.asarray(mean, dtype=np.float64).reshape(-1, 1)
    var = var + mean_diff_scale * (d @ d.T)
    c_inv = np.linalg.inv(np.linalg.cholesky(within))
    psi, u = np.linalg.eigh(c_inv @ between @ c_inv.T)
    t = u.T @ c_inv                                             # within -> I, between -> diag(psi)
    tm = t / np.sqrt(1.0 + psi)[:, None]                        # total -> I
    s, p = np.linalg.eigh(tm @ var @ tm.T)
    w2 = p.T @ np.diag(1.0 / (1.0 + psi)) @ p
    b2 = p.T @ np.diag(psi / (1.0 + psi)) @ p
    excess = np.maximum(s - 1.0, 0.0)
    w2 = w2 + np.diag(within_scale * excess)
    b2 = b2 + np.diag(between_scale * excess)
    back = np.linalg.inv(p.T @ tm)
    return m.reshape(-1), back @ w2 @ back.T, back @ b2 @ back.T


# ---------------------------------------------------------------- Kaldi-style PLDA scoring (plda_base.py PLDA)
def plda_transform(x, transform, offset, psi, num_examples=1, normalize_length=True, simple_length_norm=False,
                   reference_dim_quirk=False):
    """PLDA.transform_ivector, plda_base.py:93-107 (+ get_normalization_factor :151-158).  The reference only accepts a
    column vector and then sets self.dim = ivector.shape[-1] = 1 (:95), so it scales by sqrt(1/...) where Kaldi's
    ivector-plda-scoring (the binary score.sh actually calls) uses sqrt(D/...); reference_dim_quirk=True reproduces
    the reference as written, False is the Kaldi semantics the GPU path implements."""
    u = transform @ np.asarray(x, dtype=np.float64) + np.asarray(offset, dtype=np.float64).reshape(-1)
    d = 1 if reference_dim_quirk else u.shape[0]
    if normalize_length:
        f = np.sqrt(d) / np.linalg.norm(u) if simple_length_norm else np.sqrt(d / np.dot(1.0 / (psi + 1.0 / num_examples), u ** 2))
        u = f * u
    return u


def plda_llr(train_u, num_utts, test_u, psi):
    """PLDA.log_likelihood_ratio, plda_base.py:109-136."""
    mean = num_utts * psi / (num_utts * psi + 1.0) * train_u
    var = 1.0 + psi / (num_utts * psi + 1.0)
    given = -0.5 * (np.sum(np.log(var)) + np.sum((test_u - mean) ** 2 / var))
    without = -0.5 * (np.sum(np.log(psi + 1.0)) + np.sum(test_u ** 2 / (psi + 1.0)))
    return given - without
