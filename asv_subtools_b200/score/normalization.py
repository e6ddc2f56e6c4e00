"""S-norm / AS-norm on the GPU -- CLI twin of score/ScoreNormalization.py (:17-57):

    python -m asv_subtools_b200.score.normalization [--method asnorm|snorm] [--top-n 300]
        [--second-cohort true|false] <enroll-test-score> <enroll-cohort-score> <test-cohort-score> <out-score>

The reference groups pandas frames and loops over trials in Python (its own recipe notes the step
"could be further optimized with using matrix", recipe/voxcelebSRC/gather_results_from_epochs.sh:30-33);
here the two cohort score tables become dense matrices, one CTA per row sorts its cohort scores on chip
and takes mean / std(ddof=1) of the top n (csrc/scoring.cu topn_mean_std_kernel), and one elementwise
kernel normalises the listed trials.  `asnorm_embeddings` skips the score files altogether (cohort
scores = two cosine GEMMs)."""
import argparse
import sys
import traceback

import numpy as np
import torch

from .. import ops


def normalize(scores, trial_e, trial_t, enroll_cohort, test_cohort, top_n=0, cross_select=False, ddof=1):
    """scores (n,) fp32; trial_e/trial_t int32 indices into the rows of the two cohort matrices
    (Ne, Nc) / (Nt, Nc), all CUDA.  top_n <= 0: S-norm (all cohort scores).  cross_select: AS-norm where each
    side's statistics use the OTHER side's top-n cohort (ScoreNormalization.py:146-160).  ddof: 1 = pandas .std() of
    ScoreNormalization.py, 0 = np.std of subtools2/egrecho/score/asnorm.py:137-140."""
    if cross_select:
        if top_n < 2:
            raise ValueError("cross selection needs top_n >= 2")
        if enroll_cohort.shape[1] != test_cohort.shape[1]:
            # one side's top-n column indices address the other side's matrix: both must share ONE cohort order
            raise ValueError("cross selection needs the same cohort (same columns, same order) on both sides: "
                             "{} vs {}".format(enroll_cohort.shape[1], test_cohort.shape[1]))
        top_n = min(top_n, enroll_cohort.shape[1])    # groupby().head(top_n) semantics: a small cohort is used whole
        return ops.snorm_cross_trials(scores, trial_e, trial_t, enroll_cohort, test_cohort,
                                      ops.topn_indices(enroll_cohort, top_n), ops.topn_indices(test_cohort, top_n))
    me, se = ops.topn_mean_std(enroll_cohort, top_n, ddof)
    mt, st = ops.topn_mean_std(test_cohort, top_n, ddof)
    return ops.snorm_trials(scores, trial_e, trial_t, me, se, mt, st)


def asnorm_embeddings(enroll, test, cohort, trial_e, trial_t, top_n=300, ddof=1):
    """AS-norm straight from length-normalised embeddings: cohort scores are two cosine GEMMs.  With ddof = 0 this is
    `ScoreNorm.norm` of subtools2/egrecho/score/asnorm.py:283-352 (cosine GEMM -> top-n -> np.mean / np.std)."""
    pad = (-cohort.shape[0]) % 4
    if pad:  # the score-matrix kernel wants a multiple of 4 columns; duplicate-free padding with -inf scores is
        raise ValueError("cohort size must be a multiple of 4 (got {})".format(cohort.shape[0]))
    s = ops.cosine_trials(enroll, test, trial_e, trial_t)
    return normalize(s, trial_e, trial_t, ops.cosine_matrix(enroll, cohort), ops.cosine_matrix(test, cohort), top_n, ddof=ddof)


def _load(path):
    a, b, s = [], [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if p:
                a.append(p[0])
                b.append(p[1])
                s.append(float(p[2]))
    return a, b, np.asarray(s, dtype=np.float64)


def cohort_index(*cohort_lists):
    """ONE cohort -> column map for every score table of a run (first-appearance order over all lists): the
    reference joins on the cohort name (ScoreNormalization.py:146-160), so a column index must mean the same
    utterance in the enroll-cohort and in the test-cohort matrix whatever order the files list them in."""
    cidx = {}
    for cohort in cohort_lists:
        for c in cohort:
            cidx.setdefault(c, len(cidx))
    return cidx


def _dense(keys, cohort, vals, cidx=None):
    """(key, cohort, score) triples -> dense (num_keys, num_cohort) matrix + key index.  `cidx`: shared
    cohort -> column map (cohort_index); a table that lacks a score for any of its columns raises."""
    kidx = {}
    for k in keys:
        kidx.setdefault(k, len(kidx))
    if cidx is None:
        cidx = cohort_index(cohort)
    m = np.full((len(kidx), len(cidx)), -np.inf, dtype=np.float32)
    m[[kidx[k] for k in keys], [cidx[c] for c in cohort]] = vals
    if np.isinf(m).any():
        raise ValueError("cohort score table is not complete (every key needs a score against every cohort utterance)")
    return m, kidx


def main(argv=None):
    ap = argparse.ArgumentParser(description="Score Normalization (B200).")
    ap.add_argument("--method", default="asnorm", choices=["snorm", "asnorm"])
    ap.add_argument("--top-n", type=int, default=300)
    ap.add_argument("--second-cohort", default="true", choices=["true", "false"])
    ap.add_argument("--cross-select", default="false", choices=["true", "false"])
    ap.add_argument("input_score")
    ap.add_argument("enroll_cohort_score")
    ap.add_argument("test_cohort_score")
    ap.add_argument("output_score")
    print(" ".join(sys.argv))
    args = ap.parse_args(argv)
    try:
        if args.cross_select == "true" and args.method != "asnorm":
            raise ValueError("--cross-select applies to asnorm")
        te, tt, s = _load(args.input_score)
        a, b, v = _load(args.enroll_cohort_score)
        ek, ec = (a, b) if args.second_cohort == "true" else (b, a)
        a, b, v2 = _load(args.test_cohort_score)
        tk, tc = (a, b) if args.second_cohort == "true" else (b, a)
        if args.cross_select == "true":
            if set(ec) != set(tc):
                raise ValueError("--cross-select needs the same cohort set in both cohort score files "
                                 "({} vs {} utterances)".format(len(set(ec)), len(set(tc))))
            cidx_e = cidx_t = cohort_index(ec, tc)
        else:
            cidx_e, cidx_t = cohort_index(ec), cohort_index(tc)
        em, eidx = _dense(ek, ec, v, cidx_e)
        tm, tidx = _dense(tk, tc, v2, cidx_t)
        dev = "cuda"
        ie = torch.tensor([eidx[k] for k in te], dtype=torch.int32, device=dev)
        it = torch.tensor([tidx[k] for k in tt], dtype=torch.int32, device=dev)
        out = normalize(torch.from_numpy(s.astype(np.float32)).to(dev), ie, it, torch.from_numpy(em).to(dev),
                        torch.from_numpy(tm).to(dev), args.top_n if args.method == "asnorm" else 0,
                        cross_select=args.cross_select == "true").cpu().numpy()
        with open(args.output_score, "w") as f:
            for x, y, z in zip(te, tt, out):
                f.write("{} {} {}\n".format(x, y, repr(float(z))))
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
