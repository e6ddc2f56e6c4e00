"""Back-end pre-processing steps of score/process.sh as one CLI (same step names, same positionals):

    python -m asv_subtools_b200.score.process mean            <spk2utt> <infile> <outfile> <num_utts>      (:156-167)
    python -m asv_subtools_b200.score.process getmean         <infile> <outmean>                           (:169-179)
    python -m asv_subtools_b200.score.process submean         <mean> <infile> <outfile>                    (:181-192)
    python -m asv_subtools_b200.score.process norm            <infile> <outfile>                           (:194-203)
    python -m asv_subtools_b200.score.process transform|lda|whiten  <mat> <infile> <outfile>               (:205-216, :230-233, :262-265)
    python -m asv_subtools_b200.score.process trainlda [--dim 10] [--total-covariance-factor 0.1] <infile> <utt2spk> <outfile>   (:218-228)
    python -m asv_subtools_b200.score.process trainwhiten     <trainfile> <outmat>                         (:235-248, score/whiten/train_ZCA_Whitening.py)
    python -m asv_subtools_b200.score.process trainpcawhiten  <trainfile> <outmat>                         (:250-260)

`infile` is an ark or scp of float vectors (the extension decides, like the shell functions); outputs are binary
vector arks / Kaldi matrices.  The reference runs one Kaldi binary per step with an ark file in between; here the
vectors are one (N, D) device tensor and every O(N) piece is a kernel: centred Gram products (total / between-class
covariance, X^T X) on the tcgen05 layer kernel, speaker means, length normalisation and the affine transforms
in csrc/scoring.cu.  What stays on the host is D x D float64 algebra (eigendecompositions), as in PLDA training.

Semantics: `trainwhiten` is pinned by the reference's own train_ZCA_Whitening.py (tests/golden/make_golden_whiten.py).
`trainlda` restates Kaldi's ivector-compute-lda (global-mean subtraction; total / within covariance; normalise
tcf*total + (1-tcf)*within with a relative eigenvalue floor of 1e-6; keep the top `dim` eigenvectors of the projected
between-class covariance; offset column = -A.mean) and `trainpcawhiten` Kaldi's est-pca with its defaults
(--normalize-mean=true, --normalize-variance=false, full dimension): Kaldi is absent here -- parity unpinned for
these two, properties tested instead."""
import argparse
import sys
import traceback

import numpy as np
import torch

from .. import kaldi_io, ops
from . import backend


# ------------------------------------------------------------------------------------ device pieces
def _pad4(x):
    """The Gram / projection kernels produce rows of a multiple of 4 floats: zero-pad odd embedding sizes."""
    pad = (-x.shape[1]) % 4
    return x if pad == 0 else torch.nn.functional.pad(x, (0, pad))


def centred_gram(x, mean):
    """sum_i (x_i - mean)(x_i - mean)^T as float64 (D, D): one transposed centring pass + one Gram product."""
    d = x.shape[1]
    xp = _pad4(x)
    mp = _pad4(mean.view(1, -1))
    zero = torch.zeros(xp.shape[0], dtype=torch.int32, device=x.device)
    ct = ops.center_rows_transposed(xp.contiguous(), zero, mp.contiguous())
    g = ops.matmul_nt(ct, ct).double().cpu().numpy()[:d, :d]
    return 0.5 * (g + g.T)


def speaker_groups(keys, utt2spk):
    """rows of every speaker, in sorted speaker order; utterances without a speaker are dropped like ivector-mean does."""
    rows = {}
    for i, k in enumerate(keys):
        s = utt2spk.get(k)
        if s is not None:
            rows.setdefault(s, []).append(i)
    names = sorted(rows)
    return names, [np.asarray(rows[s], dtype=np.int32) for s in names]


def lda_statistics(x, groups):
    """(mean (D,), total (D,D), between (D,D)) float64 -- Kaldi CovarianceStats on globally centred vectors:
    total = sum x x^T / N, between = sum_s n_s m_s m_s^T / N (m_s = centred speaker mean)."""
    used = np.concatenate(groups)
    xs = x[torch.from_numpy(used.astype(np.int64)).to(x.device)].contiguous()
    n = xs.shape[0]
    mean = ops.column_mean(xs)
    total = centred_gram(xs, mean) / n
    offs = np.concatenate([[0], np.cumsum([len(g) for g in groups])])
    local = [np.arange(offs[i], offs[i + 1], dtype=np.int32) for i in range(len(groups))]
    smeans, counts = ops.speaker_mean(xs, local)
    d = x.shape[1]
    sp = _pad4(smeans)
    gm = _pad4(mean.view(1, -1)).expand(sp.shape[0], -1).contiguous()
    idx = torch.arange(sp.shape[0], dtype=torch.int32, device=x.device)
    sw = torch.from_numpy(np.sqrt(counts.astype(np.float32))).to(x.device)
    ct = ops.center_rows_transposed(sp.contiguous(), idx, gm, sw)
    between = ops.matmul_nt(ct, ct).double().cpu().numpy()[:d, :d] / n     # K of this Gram product = number of speakers
    return mean.double().cpu().numpy(), total, 0.5 * (between + between.T)


def apply_transform(x, mat):
    """ivector-transform: y = A x (+ b when mat has D+1 columns, the offset being the last one)."""
    mat = np.asarray(mat, dtype=np.float32)
    d = x.shape[1]
    if mat.shape[1] not in (d, d + 1):
        raise ValueError("transform is {}x{} but the vectors have dimension {}".format(mat.shape[0], mat.shape[1], d))
    rows = mat.shape[0]
    pad = (-rows) % 4
    a = np.zeros((rows + pad, d + ((-d) % 4)), dtype=np.float32)
    a[:rows, :d] = mat[:, :d]
    b = None
    if mat.shape[1] == d + 1:
        bb = np.zeros(rows + pad, dtype=np.float32)
        bb[:rows] = mat[:, d]
        b = torch.from_numpy(bb).to(x.device)
    y = ops.matmul_nt(_pad4(x).contiguous(), torch.from_numpy(a).to(x.device), col_bias=b)
    return y[:, :rows].contiguous() if pad else y


# ------------------------------------------------------------------------------------ host algebra (D x D, float64)
def _normalizing_transform(covar, floor):
    s, u = np.linalg.eigh(covar)
    s, u = s[::-1], u[:, ::-1]
    s = np.maximum(s, floor * s[0])
    return (u / np.sqrt(s)[None, :]).T                          # diag(s^-1/2) U^T:  T covar T^T = I


def lda_from_statistics(mean, total, between, dim, total_covariance_factor=0.1, covariance_floor=1e-6):
    """(dim, D+1) affine LDA, Kaldi ivector-compute-lda semantics (see the module docstring)."""
    within = total - between
    t = _normalizing_transform(total_covariance_factor * total + (1.0 - total_covariance_factor) * within, covariance_floor)
    s, u = np.linalg.eigh(t @ between @ t.T)
    order = np.argsort(-s, kind="stable")
    a = u[:, order[:dim]].T @ t
    return np.concatenate([a, -(a @ mean).reshape(-1, 1)], axis=1)


def zca_from_gram(gram, n, regularization=1e-6):
    """train_ZCA_Whitening.py ZCA.fit (:34-52): cov = X^T X / (n-1) WITHOUT mean removal, U S U^T = cov,
    whiten = U diag(1/sqrt(clip(S, reg))) U^T; written with a zero bias column (:68-76)."""
    s, u = np.linalg.eigh(gram / (n - 1))
    w = (u / np.sqrt(np.clip(s, regularization, None))[None, :]) @ u.T
    return np.concatenate([w, np.zeros((w.shape[0], 1))], axis=1)


def pca_from_statistics(mean, covar, dim=-1, normalize_variance=False, normalize_mean=True):
    s, p = np.linalg.eigh(covar)
    order = np.argsort(-s, kind="stable")
    s, p = s[order], p[:, order]
    k = covar.shape[0] if dim is None or dim < 0 else dim
    a = p[:, :k].T
    if normalize_variance:
        a = a / np.sqrt(s[:k])[:, None]
    if normalize_mean:
        a = np.concatenate([a, -(a @ mean).reshape(-1, 1)], axis=1)
    return a


# ------------------------------------------------------------------------------------ steps
def _spec(path):
    return ("scp:" if path.endswith(".scp") else "ark:") + path if ":" not in path else path


def _write_vectors(path, keys, x):
    x = x.cpu().numpy()
    with kaldi_io.open_or_fd(path, "wb") as f:
        for k, v in zip(keys, x):
            kaldi_io.write_vec_flt(f, v, key=k)


def read_map(path, many=False):
    out = {}
    with open(path) as f:
        for line in f:
            p = line.split()
            if p:
                out[p[0]] = p[1:] if many else p[1]
    return out


def step_mean(spk2utt, infile, outfile, num_utts):
    keys, x = backend.load_vectors(_spec(infile))
    index = {k: i for i, k in enumerate(keys)}
    names, groups = [], []
    for spk, utts in read_map(spk2utt, many=True).items():
        rows = [index[u] for u in utts if u in index]
        if rows:
            names.append(spk)
            groups.append(np.asarray(rows, dtype=np.int32))
    means, counts = ops.speaker_mean(x, groups)
    _write_vectors(outfile, names, means)
    with open(num_utts, "w") as f:
        for s, c in zip(names, counts):
            f.write("{} {}\n".format(s, int(c)))


def step_getmean(infile, outmean):
    _, x = backend.load_vectors(_spec(infile))
    kaldi_io.write_vec_flt(outmean, ops.column_mean(x).cpu().numpy())


def step_submean(mean, infile, outfile):
    keys, x = backend.load_vectors(_spec(infile))
    m = torch.from_numpy(np.array(kaldi_io.read_vec_flt(mean), dtype=np.float32)).to(x.device)
    _write_vectors(outfile, keys, backend.preprocess(x, m, norm=False))


def step_norm(infile, outfile):
    keys, x = backend.load_vectors(_spec(infile))
    _write_vectors(outfile, keys, ops.center_length_norm(x, None))


def step_transform(mat, infile, outfile):
    keys, x = backend.load_vectors(_spec(infile))
    _write_vectors(outfile, keys, apply_transform(x, kaldi_io.read_mat(mat)))


def step_trainlda(infile, utt2spk, outfile, dim=10, total_covariance_factor=0.1):
    keys, x = backend.load_vectors(_spec(infile))
    _, groups = speaker_groups(keys, read_map(utt2spk))
    if not groups:
        raise ValueError("no vector of {} has a speaker in {}".format(infile, utt2spk))
    mean, total, between = lda_statistics(x, groups)
    if dim > x.shape[1]:
        raise ValueError("--dim {} exceeds the vector dimension {}".format(dim, x.shape[1]))
    kaldi_io.write_mat(outfile, lda_from_statistics(mean, total, between, dim, total_covariance_factor).astype(np.float32))


def step_trainwhiten(trainfile, outmat):
    _, x = backend.load_vectors(_spec(trainfile))
    gram = centred_gram(x, torch.zeros(x.shape[1], device=x.device))
    kaldi_io.write_mat(outmat, zca_from_gram(gram, x.shape[0]).astype(np.float32))


def step_trainpcawhiten(trainfile, outmat):
    _, x = backend.load_vectors(_spec(trainfile))
    mean = ops.column_mean(x)
    covar = centred_gram(x, mean) / x.shape[0]
    kaldi_io.write_mat(outmat, pca_from_statistics(mean.double().cpu().numpy(), covar).astype(np.float32))


STEPS = {"mean": (step_mean, 4), "getmean": (step_getmean, 2), "submean": (step_submean, 3), "norm": (step_norm, 2),
         "transform": (step_transform, 3), "lda": (step_transform, 3), "whiten": (step_transform, 3),
         "trainlda": (step_trainlda, 3), "trainwhiten": (step_trainwhiten, 2), "trainpcawhiten": (step_trainpcawhiten, 2)}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("step", choices=sorted(STEPS))
    ap.add_argument("--dim", type=int, default=10, help="trainlda: output dimension (clda of scoreSets.sh)")
    ap.add_argument("--total-covariance-factor", type=float, default=0.1)
    ap.add_argument("paths", nargs="+")
    args = ap.parse_args(argv)
    fn, arity = STEPS[args.step]
    if len(args.paths) != arity:
        ap.error("{} takes {} positionals".format(args.step, arity))
    try:
        if args.step == "trainlda":
            fn(*args.paths, dim=args.dim, total_covariance_factor=args.total_covariance_factor)
        else:
            fn(*args.paths)
    except BaseException as err:  # same contract as the reference's shell functions: message, non-zero exit
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
