"""PLDA training on the GPU: mirror of `PldaStats` / `PldaEstimation` of the reference
(score/pyplda/plda_base.py:37-81, :227-342).

The reference runs one D x D matrix inverse per class per EM iteration in a Python loop.  Here every
per-class quantity is evaluated in the basis that makes within_var the identity and between_var
diagonal (the transform of `get_output`, :302-335), where it is diagonal and depends on the class size
n only; what remains per iteration is one (S, D) x (D, D) projection of the class means and two Gram
products (D, S) x (S, D), all on the tcgen05 layer kernel (fp32-grade bf16x3), plus float64 D x D algebra
on the host.  Same fixed point and, iteration by iteration, the same (within_var, between_var) as the
reference up to fp32 rounding of the Gram sums (tests: 1e-5 relative after 10 iterations).
"""
import numpy as np
import torch

from .. import kaldi_io, ops


class PldaStats:
    """add_samples(weight, group) like the reference, or from_matrix(emb, spk) for a whole set at once."""

    def __init__(self, dim):
        self.dim_ = int(dim)
        self._groups, self._weights = [], []
        self._emb = self._spk = None

    def add_samples(self, weight, group):
        g = np.asarray(group, dtype=np.float32)
        if g.ndim != 2 or g.shape[1] != self.dim_:
            raise ValueError("add_samples: expected (n, {}) rows".format(self.dim_))
        self._groups.append(g)
        self._weights.append(float(weight))

    @classmethod
    def from_matrix(cls, emb, spk, weights=None):
        """emb (N, D) float32 (CUDA tensor or ndarray), spk (N,) integer labels; weights per class in
        the order of np.unique(spk)."""
        self = cls(emb.shape[1])
        self._emb = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32))
        self._spk = np.asarray(spk.cpu() if isinstance(spk, torch.Tensor) else spk)
        self._w_unique = None if weights is None else np.asarray(weights, dtype=np.float64)
        return self

    # the reference insists on classes sorted by size; order does not matter here
    def sort(self):
        return

    def is_sorted(self):
        return True

    def _finalize(self, device="cuda"):
        if self._emb is None:
            emb = torch.from_numpy(np.concatenate(self._groups, axis=0))
            cls_of = np.repeat(np.arange(len(self._groups)), [g.shape[0] for g in self._groups])
            w = np.asarray(self._weights, dtype=np.float64)
        else:
            emb = self._emb
            ids, cls_of = np.unique(self._spk, return_inverse=True)
            w = np.ones(ids.shape[0]) if self._w_unique is None else self._w_unique
        x = emb.to(device=device, dtype=torch.float32).contiguous()
        n = np.bincount(cls_of).astype(np.float64)
        order = np.argsort(cls_of, kind="stable")
        rows = np.split(order.astype(np.int32), np.cumsum(n.astype(np.int64))[:-1])
        means, _ = ops.speaker_mean(x, rows)                                   # (S, D) class means
        spk_d = torch.from_numpy(cls_of.astype(np.int32)).to(device)
        sw = torch.from_numpy(np.sqrt(w).astype(np.float32)).to(device)
        ct = ops.center_rows_transposed(x, spk_d, means, sw)                   # (D, N)
        self.offset_scatter = ops.matmul_nt(ct, ct).double().cpu().numpy()     # sum_k w_k sum_i (x-m_k)(x-m_k)^T
        self.offset_scatter = 0.5 * (self.offset_scatter + self.offset_scatter.T)
        self.n, self.weight, self.means = n, w, means
        self.num_classes, self.num_example = int(n.shape[0]), int(n.sum())
        self.class_weight, self.example_weight = float(w.sum()), float((w * n).sum())
        m64 = means.double().cpu().numpy()
        self.sum = (w[:, None] * m64).sum(axis=0)
        return self


class PldaEstimation:
    def __init__(self, stats):
        self.stats = stats if hasattr(stats, "offset_scatter") else stats._finalize()
        self.dim = self.stats.dim_
        self.between_var = np.eye(self.dim)
        self.within_var = np.eye(self.dim)
        self.mean = self.stats.sum / self.stats.class_weight

    def estimate(self, num_em_iters=10):
        st = self.stats
        dev = st.means.device
        n_d = torch.from_numpy(st.n.astype(np.float32)).to(dev)
        w_d = torch.from_numpy(st.weight.astype(np.float32)).to(dev)
        for _ in range(num_em_iters):
            t1 = np.linalg.inv(np.linalg.cholesky(self.within_var))
            psi, u = np.linalg.eigh(t1 @ self.between_var @ t1.T)
            t = u.T @ t1                                                       # T W T^T = I, T B T^T = diag(psi)
            tinv = np.linalg.inv(t)
            t_d = torch.from_numpy(t.astype(np.float32)).to(dev)
            shift = torch.from_numpy((-(t @ self.mean)).astype(np.float32)).to(dev)
            proj = ops.matmul_nt(st.means, t_d, col_bias=shift)                # (S, D): T (m_k - mean)
            what_t, resid_t = ops.plda_em_rows(proj, n_d, w_d, torch.from_numpy(psi.astype(np.float32)).to(dev))
            gb = ops.matmul_nt(what_t, what_t).double().cpu().numpy()
            gw = ops.matmul_nt(resid_t, resid_t).double().cpu().numpy()
            mixd = psi[None, :] / (1.0 + st.n[:, None] * psi[None, :])         # diagonal of (B^-1 + n W^-1)^-1
            b_t = np.diag((st.weight[:, None] * mixd).sum(0)) + 0.5 * (gb + gb.T)
            w_t = np.diag((st.weight[:, None] * st.n[:, None] * mixd).sum(0)) + 0.5 * (gw + gw.T)
            self.between_var = tinv @ b_t @ tinv.T / st.class_weight
            self.within_var = (tinv @ w_t @ tinv.T + st.offset_scatter) / st.example_weight
        return self

    def plda_write(self, plda):
        """Same file as the reference's plda_write (:337-342): ark of mean / within_var / between_var."""
        with kaldi_io.open_or_fd(plda, "wb") as f:
            kaldi_io.write_vec_flt(f, self.mean.reshape(-1), key="mean")
            kaldi_io.write_vec_flt(f, self.within_var.reshape(-1), key="within_var")
            kaldi_io.write_vec_flt(f, self.between_var.reshape(-1), key="between_var")

    def model(self):
        from .backend import PldaModel
        return PldaModel(self.mean, self.within_var, self.between_var)


class PLDA:
    """The diagonalised model `PldaEstimation.get_output()` returns (plda_base.py:302-335, :84-224): `transform`
    with transform.W.transform^T = I and transform.B.transform^T = diag(psi) (psi descending like Kaldi's),
    `offset = -transform.mean`; written in Kaldi's text format by `plda_trans_write` (:211-224)."""

    def __init__(self, mean, within_var, between_var):
        self.mean = np.asarray(mean, dtype=np.float64).reshape(-1, 1)
        self.dim = self.mean.shape[0]
        t1 = np.linalg.inv(np.linalg.cholesky(np.asarray(within_var, dtype=np.float64)))
        s, u = np.linalg.eigh(t1 @ np.asarray(between_var, dtype=np.float64) @ t1.T)
        order = np.argsort(s)[::-1]
        s, u = s[order], u[:, order]
        if s.min() <= 0:
            raise ValueError("between-class covariance is not positive definite")
        self.transform = u.T @ t1
        self.psi = s
        self.offset = -1.0 * (self.transform @ self.mean)

    # ---- Kaldi-style scoring (ivector-plda-scoring as restated by plda_base.py :93-136, :151-158) on the GPU
    def _dev(self, device):
        if getattr(self, "_d", None) is None or self._d[0] != str(device):
            f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)  # noqa: E731
            self._d = (str(device), f32(self.transform), f32(self.offset.reshape(-1)), f32(self.psi))
        return self._d[1:]

    def smooth_within_class_covariance(self, smoothing_factor):
        """Kaldi's SmoothWithinClassCovariance (ivector-copy-plda --smoothing): within (the identity here) grows by
        smoothing_factor * psi; psi and the rows of `transform` are rescaled so that it is the identity again.  (The
        reference's restatement, plda_base.py:138-149, multiplies `np.diag(.) * transform` elementwise, which zeroes
        the transform's off-diagonal entries; it has no caller.  This is the Kaldi operation.)"""
        w = 1.0 + smoothing_factor * self.psi
        self.psi = self.psi / w
        self.transform = self.transform * (w ** -0.5)[:, None]
        self.offset = -1.0 * (self.transform @ self.mean)
        self._d = None

    def transform_ivectors(self, x, num_examples=None, normalize_length=True, simple_length_norm=False):
        """x (N, D) fp32 CUDA -> transformed (N, D); num_examples (N,) (enroll averages of n utterances) or None = 1."""
        t, off, psi = self._dev(x.device)
        u = ops.matmul_nt(x, t, col_bias=off)
        if normalize_length:
            n = None if num_examples is None else torch.as_tensor(num_examples, dtype=torch.float32, device=x.device).contiguous()
            ops.plda_normalize_rows(u, psi, n, simple_length_norm)
        return u

    def _operands(self, enroll_u, num_examples, test_u):
        _, _, psi = self._dev(enroll_u.device)
        n = None if num_examples is None else torch.as_tensor(num_examples, dtype=torch.float32, device=enroll_u.device).contiguous()
        return ops.plda_llr_operands(enroll_u, psi, n, 0) + ops.plda_llr_operands(test_u, psi, None, 1)

    def log_likelihood_ratio_matrix(self, enroll_u, num_examples, test_u):
        """(Ne, Nt) log-likelihood ratios of transformed vectors: one GEMM with K = 2D plus row / column terms."""
        a, row, b, col = self._operands(enroll_u, num_examples, test_u)
        return ops.matmul_nt(a, b, row_bias=row, col_bias=col)

    def log_likelihood_ratio_trials(self, enroll_u, num_examples, test_u, trial_e, trial_t):
        a, row, b, col = self._operands(enroll_u, num_examples, test_u)
        return ops.bilinear_trials(a, b, trial_e, trial_t, row, col)

    @classmethod
    def read_trans(cls, path):
        """From the Kaldi-text file plda_trans_write produces: <Plda> [ mean ] [ transform rows ] [ psi ] </Plda>."""
        txt = open(path).read()
        if not txt.lstrip().startswith("<Plda>"):
            raise ValueError("{} is not a Kaldi-text PLDA file".format(path))
        body = txt.replace("<Plda>", " ").replace("</Plda>", " ")
        groups = [g.split() for g in body.replace("]", "[").split("[") if g.split()]
        mean = np.array(groups[0], dtype=np.float64)
        d = mean.shape[0]
        self = cls.__new__(cls)
        self.mean, self.dim = mean.reshape(-1, 1), d
        self.transform = np.array(groups[1], dtype=np.float64).reshape(d, d)
        self.psi = np.array(groups[2], dtype=np.float64)
        self.offset = -1.0 * (self.transform @ self.mean)
        return self

    @classmethod
    def read_ori(cls, path):
        """From the three-vector file PldaEstimation.plda_write produces (`<plda>.ori`)."""
        parts = dict(kaldi_io.read_vec_flt_ark(path))
        d = np.asarray(parts["mean"]).shape[0]
        return cls(parts["mean"], np.asarray(parts["within_var"], dtype=np.float64).reshape(d, d),
                   np.asarray(parts["between_var"], dtype=np.float64).reshape(d, d))

    def plda_trans_write(self, plda):
        with open(plda, "w") as f:
            f.write("<Plda>  [ " + " ".join(map(str, self.mean.reshape(-1))) + " ]\n")
            f.write(" [")
            for row in self.transform:
                f.write("\n  " + " ".join(map(str, row)))
            f.write(" ]")
            f.write("\n [ " + " ".join(map(str, self.psi.reshape(-1))) + " ]\n")
            f.write("</Plda> ")


def _get_output(self):
    return PLDA(self.mean, self.within_var, self.between_var)


PldaEstimation.get_output = _get_output


class Coral:
    """CORAL adaptation of a two-covariance PLDA to unlabelled in-domain vectors -- mirror of
    score/pyplda/ivector-adapt-plda-coral.py (CORAL.add_stats :30-38, update_plda :40-84): the adaptation
    data's total covariance (its Gram product, on the GPU) is matched by the linear map A = C_i . C_o with
    C_o = (W + B)^(-1/2), C_i = Var^(1/2); W and B become A W A^T and A B A^T; the mean moves to the data's."""

    def __init__(self, mean_diff_scale=1.0, within_covar_scale=0.8, between_covar_scale=0.8):
        self.mean_diff_scale = mean_diff_scale
        self.within_covar_scale, self.between_covar_scale = within_covar_scale, between_covar_scale
        self._rows = []

    def plda_read(self, plda):
        parts = dict(kaldi_io.read_vec_flt_ark(plda))
        self.mean = np.asarray(parts["mean"], dtype=np.float64).reshape(-1, 1)
        self.dim = self.mean.shape[0]
        self.within_var = np.asarray(parts["within_var"], dtype=np.float64).reshape(self.dim, self.dim)
        self.between_var = np.asarray(parts["between_var"], dtype=np.float64).reshape(self.dim, self.dim)

    def add_stats(self, weight, ivector):
        if weight != 1:
            raise NotImplementedError("Coral.add_stats: the reference CLI only ever passes weight 1")
        self._rows.append(np.asarray(ivector, dtype=np.float32).reshape(-1))

    def add_matrix(self, emb):
        self._emb = emb

    def _adaptation_covariance(self, device="cuda"):
        """(variance + mean_diff_scale * diff diff^T, data mean): the adaptation set's total covariance as one centred
        Gram product on the GPU (E[x x^T] - E[x] E[x]^T without the cancellation)."""
        x = getattr(self, "_emb", None)
        if x is None:
            x = torch.from_numpy(np.stack(self._rows))
        x = (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)))
        x = x.to(device=device, dtype=torch.float32).contiguous()
        n = x.shape[0]
        mean_d = ops.column_mean(x)
        zero = torch.zeros(n, dtype=torch.int32, device=device)
        ct = ops.center_rows_transposed(x, zero, mean_d.view(1, -1).contiguous())
        variance = ops.matmul_nt(ct, ct).double().cpu().numpy() / n
        variance = 0.5 * (variance + variance.T)
        mean = mean_d.double().cpu().numpy().reshape(-1, 1)
        diff = mean - self.mean
        return variance + self.mean_diff_scale * (diff @ diff.T), mean

    def update_plda(self, device="cuda"):
        variance, self.mean = self._adaptation_covariance(device)
        eo, qo = np.linalg.eigh(self.within_var + self.between_var)
        ei, qi = np.linalg.eigh(variance)
        c_o = qo @ np.diag(1.0 / np.sqrt(eo)) @ qo.T
        c_i = qi @ np.diag(np.sqrt(ei)) @ qi.T
        self.A = c_i @ c_o
        self.within_var = self.A @ self.within_var @ self.A.T
        self.between_var = self.A @ self.between_var @ self.A.T

    def plda_write(self, plda):
        with kaldi_io.open_or_fd(plda, "wb") as f:
            kaldi_io.write_vec_flt(f, self.mean.reshape(-1), key="mean")
            kaldi_io.write_vec_flt(f, self.within_var.reshape(-1), key="within_var")
            kaldi_io.write_vec_flt(f, self.between_var.reshape(-1), key="between_var")


def read_ori(plda):
    """(mean (D,1), within (D,D), between (D,D)) float64 from the three-vector ark the pyplda scripts exchange
    (`plda_read` of every ivector-adapt-plda-*.py; written by plda_base.py:337-342)."""
    parts = dict(kaldi_io.read_vec_flt_ark(plda))
    mean = np.asarray(parts["mean"], dtype=np.float64).reshape(-1, 1)
    d = mean.shape[0]
    return (mean, np.asarray(parts["within_var"], dtype=np.float64).reshape(d, d),
            np.asarray(parts["between_var"], dtype=np.float64).reshape(d, d))


def _excess_over(base, target):
    """B^-T max(0, E - I) B^-1 where B diagonalises the pair: B^T base B = I, B^T target B = diag(E) -- the part of
    `target` that exceeds `base`, the regulariser shared by CORAL+ / LIP-reg / CIP-reg
    (ivector-adapt-plda-coralplus.py:76-84).  D x D float64 algebra on the host, like the EM's."""
    lam, q = np.linalg.eigh(base)
    scale = q / np.sqrt(lam)[None, :]                              # Q diag(lam^-1/2): base -> I
    e, p = np.linalg.eigh(scale.T @ target @ scale)
    b_inv = p.T @ (q * np.sqrt(lam)[None, :]).T                    # (Q diag(lam^-1/2) P)^-1 = P^T diag(lam^1/2) Q^T
    return b_inv.T @ (np.maximum(0.0, e - 1.0)[:, None] * b_inv)


class _Adapter:
    def plda_write(self, plda):
        Coral.plda_write(self, plda)

    def get_output(self):
        return PLDA(self.mean, self.within_var, self.between_var)


class CoralPlus(Coral):
    """CORAL+ (ivector-adapt-plda-coralplus.py, CORALPlus.update_plda :40-96): the pseudo in-domain covariances of
    CORAL only add their excess over the out-of-domain ones, scaled by within/between_covar_scale."""

    def update_plda(self, device="cuda"):
        w0, b0 = self.within_var, self.between_var
        super().update_plda(device)                                # adaptation-set covariance: Gram product on the GPU
        self.within_var = w0 + self.within_covar_scale * _excess_over(w0, self.within_var)
        self.between_var = b0 + self.between_covar_scale * _excess_over(b0, self.between_var)


class Lip(_Adapter):
    """Linear interpolation of an out-of-domain and an in-domain model (ivector-adapt-plda-lip.py :25-34)."""

    def __init__(self, interpolation_weight=0.4):
        self.interpolation_weight = interpolation_weight

    def interpolation(self, plda_out_domain, plda_in_domain):
        a = self.interpolation_weight
        _, w_out, b_out = read_ori(plda_out_domain)
        self.mean, w_in, b_in = read_ori(plda_in_domain)
        self.within_var, self.between_var = a * w_out + (1 - a) * w_in, a * b_out + (1 - a) * b_in


class LipReg(_Adapter):
    """ivector-adapt-plda-lip-reg.py :26-49: in-domain model + (1 - weight) x the out-of-domain model's excess over it."""

    def __init__(self, interpolation_weight=0.6):
        self.interpolation_weight = interpolation_weight

    def interpolation(self, plda_out_domain, plda_in_domain):
        a = 1.0 - self.interpolation_weight
        _, w_out, b_out = read_ori(plda_out_domain)
        self.mean, w_in, b_in = read_ori(plda_in_domain)
        self.within_var, self.between_var = w_in + a * _excess_over(w_in, w_out), b_in + a * _excess_over(b_in, b_out)


class Cip(_Adapter):
    """ivector-adapt-plda-cip.py :113-121: interpolation of a CORAL-adapted out-of-domain model (a `Coral` after
    update_plda) with the in-domain model."""

    def __init__(self, interpolation_weight=0.5):
        self.interpolation_weight = interpolation_weight

    def interpolation(self, coral, plda_in_domain):
        a = self.interpolation_weight
        self.mean, w_in, b_in = read_ori(plda_in_domain)
        self.within_var, self.between_var = a * coral.within_var + (1 - a) * w_in, a * coral.between_var + (1 - a) * b_in


class CipReg(_Adapter):
    """ivector-adapt-plda-cip-reg.py :109-128: plda_read(in-domain), then add weight x the CORAL model's excess."""

    def __init__(self, interpolation_weight=0.5):
        self.interpolation_weight = interpolation_weight

    def plda_read(self, plda):
        self.mean, self.within_var, self.between_var = read_ori(plda)
        self.dim = self.mean.shape[0]

    def interpolation(self, coral):
        a = self.interpolation_weight
        self.within_var = self.within_var + a * _excess_over(self.within_var, coral.within_var)
        self.between_var = self.between_var + a * _excess_over(self.between_var, coral.between_var)


class PldaUnsupervisedAdaptor(Coral):
    """Kaldi's ivector-adapt-plda as the reference restates it (score/pyplda/plda_base.py PldaUnsupervisedAdaptor
    :344-485; the `trainaplda` step of score/process.sh:280-292): in the space where the model's total covariance is the
    identity, each direction in which the adaptation data's covariance exceeds 1 hands that excess to the within- and
    between-class covariances in the proportions within_covar_scale / between_covar_scale.  add_stats / add_matrix and the
    adaptation-set covariance (one Gram product on the GPU) are Coral's; update_plda(plda) rewrites plda.mean / transform /
    psi / offset in place (psi descending, like PLDA.get_output) and returns the adapted (within, between)."""

    def __init__(self, mean_diff_scale=1.0, within_covar_scale=0.3, between_covar_scale=0.7):
        super().__init__(mean_diff_scale, within_covar_scale, between_covar_scale)

    def update_plda(self, plda, device="cuda"):
        self.mean = np.asarray(plda.mean, dtype=np.float64).reshape(-1, 1)
        variance, mean = self._adaptation_covariance(device)
        psi = np.asarray(plda.psi, dtype=np.float64)
        tm = np.asarray(plda.transform, dtype=np.float64) / np.sqrt(1.0 + psi)[:, None]      # total covariance -> I
        s, p = np.linalg.eigh(tm @ variance @ tm.T)
        excess = np.maximum(s - 1.0, 0.0)
        w2 = p.T @ (p / (1.0 + psi)[:, None]) + np.diag(self.within_covar_scale * excess)
        b2 = p.T @ (p * (psi / (1.0 + psi))[:, None]) + np.diag(self.between_covar_scale * excess)
        back = np.linalg.inv(p.T @ tm)
        self.within_var, self.between_var = back @ w2 @ back.T, back @ b2 @ back.T
        self.mean = mean
        new = PLDA(mean, self.within_var, self.between_var)
        plda.mean, plda.transform, plda.psi, plda.offset = new.mean, new.transform, new.psi, new.offset
        return self.within_var, self.between_var
