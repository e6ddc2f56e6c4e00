"""Scoring back-end on device tensors: embedding pre-processing, cosine and two-covariance PLDA.
All per-vector / per-trial arithmetic runs in csrc/scoring.cu (+ the tcgen05 layer kernel for the
projections and score matrices); the D x D PLDA model algebra (three matrix inverses,
gaussian-plda-scoring.py:31-50) is parameter preparation and stays in float64 on the host."""
import numpy as np
import torch

from .. import kaldi_io, ops


def load_vectors(spec, device="cuda"):
    """ark/scp of float vectors -> (keys, (N, D) float32 CUDA tensor)."""
    keys, vecs = [], []
    for k, v in kaldi_io.read_vectors(spec):
        keys.append(k)
        vecs.append(np.asarray(v, dtype=np.float32))
    if not keys:
        raise ValueError("no vectors in {}".format(spec))
    return keys, torch.from_numpy(np.stack(vecs)).to(device)


def read_trials(path):
    """3 columns: enroll test target|nontarget (getTrials.sh:132-160); the label column is optional."""
    e, t, lab = [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            e.append(p[0])
            t.append(p[1])
            lab.append(p[2] if len(p) > 2 else "")
    return e, t, lab


def index_trials(trial_e, trial_t, enroll_keys, test_keys, device="cuda"):
    ei = {k: i for i, k in enumerate(enroll_keys)}
    ti = {k: i for i, k in enumerate(test_keys)}
    try:
        a = np.fromiter((ei[k] for k in trial_e), dtype=np.int32, count=len(trial_e))
        b = np.fromiter((ti[k] for k in trial_t), dtype=np.int32, count=len(trial_t))
    except KeyError as err:
        raise KeyError("trial refers to a vector that is not in the table: {}".format(err))
    return torch.from_numpy(a).to(device), torch.from_numpy(b).to(device)


def preprocess(x, submean=None, norm=True):
    """`submean` then `norm` of score/process.sh:181-203 in one pass."""
    if submean is None and not norm:
        return x
    if not norm:
        return x - submean[None, :]
    return ops.center_length_norm(x, submean)


def cosine_score(enroll, test, te, tt):
    """score/score.sh:82-97: one dot product per listed trial."""
    return ops.cosine_trials(enroll, test, te, tt)


class PldaModel:
    """Two-covariance PLDA scorer (score/pyplda/gaussian-plda-scoring.py).  `mean` (D,), `within`,
    `between` (D,D) as stored by plda_base.plda_write (:337-342)."""

    def __init__(self, mean, within, between, smoothing=5e-5, device="cuda"):
        mean = np.asarray(mean, dtype=np.float64).reshape(-1, 1)
        d = mean.shape[0]
        self.mean64, self.within64 = mean.copy(), np.asarray(within, dtype=np.float64).reshape(d, d).copy()
        within = self.within64 + smoothing * np.eye(d)                                        # :65
        between = np.asarray(between, dtype=np.float64).reshape(d, d)
        self.between64 = between.copy()
        tot_inv = np.linalg.inv(between + within)                                             # :33-35
        w2b_inv = np.linalg.inv(within + 2 * between)
        w_inv = np.linalg.inv(within)
        self.gamma = (-1 / 4) * (w2b_inv + w_inv) + (1 / 2) * tot_inv                        # :38
        self.lam = (-1 / 4) * (w2b_inv - w_inv)                                               # :41
        self.c = np.matmul(w2b_inv - tot_inv, mean).reshape(-1)                               # :44
        self.dim = d
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)  # noqa: E731
        self.gamma_d = f32(0.5 * (self.gamma + self.gamma.T))
        self.l2_d = f32(self.lam + self.lam.T)
        self.c_d = f32(self.c)

    @classmethod
    def read(cls, spec, **kw):
        parts = dict(kaldi_io.read_vec_flt_ark(spec))
        return cls(parts["mean"], parts["within_var"], parts["between_var"], **kw)

    def write(self, spec):
        """plda_base.plda_write layout (plda_base.py:337-342): three float vectors keyed mean /
        within_var / between_var (the un-smoothed within-class covariance is stored)."""
        with kaldi_io.open_or_fd(spec, "wb") as f:
            kaldi_io.write_vec_flt(f, self.mean64.reshape(-1), key="mean")
            kaldi_io.write_vec_flt(f, self.within64.reshape(-1), key="within_var")
            kaldi_io.write_vec_flt(f, self.between64.reshape(-1), key="between_var")

    def terms(self, x):
        """x^T Gamma x + x^T c per row."""
        return ops.plda_terms(x, self.gamma_d, self.c_d)

    def score_trials(self, enroll, test, te, tt):
        """PLDAScoring (:23-29) for each listed trial (k = 0)."""
        proj = ops.project(enroll, self.l2_d)  # E.(Lambda + Lambda^T); l2 is symmetric
        return ops.bilinear_trials(proj, test, te, tt, self.terms(enroll), self.terms(test))

    def score_matrix(self, enroll, test):
        return ops.plda_matrix(enroll, test, self.l2_d, self.terms(enroll), self.terms(test))


    def score_histogram(self, enroll, enroll_spk, test, test_spk, lo, hi, nbins=2048, **shard):
        """The same scores binned by trial class without ever being stored (BASELINE config 5 at
        10^6 x 10^4 is a 40 GB matrix): (2, nbins) int64 counters, see ops.trial_histogram."""
        proj = ops.project(enroll, self.l2_d)
        return ops.trial_histogram(proj, enroll_spk, test, test_spk, lo, hi, nbins, row_term=self.terms(enroll),
                                   col_term=self.terms(test), **shard)


def write_scores(path, trial_e, trial_t, scores):
    s = scores.detach().cpu().numpy() if isinstance(scores, torch.Tensor) else np.asarray(scores)
    with open(path, "w") as f:
        for a, b, v in zip(trial_e, trial_t, s):
            f.write("{} {} {}\n".format(a, b, repr(float(v))))
