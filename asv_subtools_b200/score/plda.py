"""CLI twin of score/pyplda/gaussian-plda-scoring.py (:87-101) / score.sh `plda`:
<trials> <plda.ark> <enroll.ark|scp> <test.ark|scp> <out.score>."""
import argparse
import sys
import traceback

from . import backend


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("trials")
    ap.add_argument("plda")
    ap.add_argument("enroll")
    ap.add_argument("test")
    ap.add_argument("out_score")
    args = ap.parse_args(argv)
    try:
        model = backend.PldaModel.read(args.plda)
        ek, e = backend.load_vectors(args.enroll)
        tk, t = backend.load_vectors(args.test)
        tr_e, tr_t, _ = backend.read_trials(args.trials)
        ie, it = backend.index_trials(tr_e, tr_t, ek, tk)
        backend.write_scores(args.out_score, tr_e, tr_t, model.score_trials(e, t, ie, it))
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
