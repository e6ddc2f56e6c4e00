"""PLDA scoring CLI.

Two-covariance form (score/pyplda/gaussian-plda-scoring.py :87-101):
    python -m asv_subtools_b200.score.plda <trials> <plda.ori> <enroll.ark|scp> <test.ark|scp> <out.score>
Kaldi form, the positionals of score.sh `plda` (:99-121; there: ivector-plda-scoring --normalize-length=true
--num-utts=ark:<num_utts> "ivector-copy-plda --smoothing=S <plda> - |"):
    python -m asv_subtools_b200.score.plda --kaldi [--smoothing S] [--normalize-length true|false]
        <trials> <num_utts.ark | ""> <plda (.ori vectors or Kaldi text)> <enroll> <test> <out.score>
"""
import argparse
import sys
import traceback

import numpy as np

from . import backend


def read_num_utts(path):
    """Text ark `spk n` as written by ivector-mean (process.sh:156-167)."""
    out = {}
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) >= 2:
                out[p[0]] = float(p[1])
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--kaldi", action="store_true", help="Kaldi-style LLR in the diagonalised space (6 positionals)")
    ap.add_argument("--smoothing", type=float, default=0.0)
    ap.add_argument("--normalize-length", default="true", choices=["true", "false"])
    ap.add_argument("args", nargs="+")
    a = ap.parse_args(argv)
    try:
        if not a.kaldi:
            if len(a.args) != 5:
                raise ValueError("expected <trials> <plda> <enroll> <test> <out.score>")
            trials, plda, enroll, test, out_score = a.args
            model = backend.PldaModel.read(plda)
            ek, e = backend.load_vectors(enroll)
            tk, t = backend.load_vectors(test)
            tr_e, tr_t, _ = backend.read_trials(trials)
            ie, it = backend.index_trials(tr_e, tr_t, ek, tk)
            backend.write_scores(out_score, tr_e, tr_t, model.score_trials(e, t, ie, it))
            return
        from .plda_train import PLDA
        if len(a.args) != 6:
            raise ValueError("expected <trials> <num_utts> <plda> <enroll> <test> <out.score>")
        trials, num_utts, plda, enroll, test, out_score = a.args
        try:
            model = PLDA.read_trans(plda)
        except (ValueError, UnicodeDecodeError):
            model = PLDA.read_ori(plda)
        if a.smoothing:
            model.smooth_within_class_covariance(a.smoothing)
        ek, e = backend.load_vectors(enroll)
        tk, t = backend.load_vectors(test)
        n = None
        if num_utts:
            table = read_num_utts(num_utts)
            n = np.array([table[k] for k in ek], dtype=np.float32)
        norm = a.normalize_length == "true"
        eu = model.transform_ivectors(e, n, normalize_length=norm)
        tu = model.transform_ivectors(t, None, normalize_length=norm)
        tr_e, tr_t, _ = backend.read_trials(trials)
        ie, it = backend.index_trials(tr_e, tr_t, ek, tk)
        backend.write_scores(out_score, tr_e, tr_t, model.log_likelihood_ratio_trials(eu, n, tu, ie, it))
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
