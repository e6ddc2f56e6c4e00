"""CLI twin of score/score.sh `cosine` (:82-97): <trials> <enroll.ark|scp> <test.ark|scp> <out.score>.
Optional --submean / --norm fold the `submean`/`norm` steps of score/process.sh in."""
import argparse
import sys
import traceback

import numpy as np
import torch

from .. import kaldi_io
from . import backend


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("trials")
    ap.add_argument("enroll")
    ap.add_argument("test")
    ap.add_argument("out_score")
    ap.add_argument("--submean", default=None, help="global mean vector file (ivector-mean output)")
    ap.add_argument("--norm", action="store_true", help="length-normalise both sides (ivector-normalize-length)")
    args = ap.parse_args(argv)
    try:
        ek, e = backend.load_vectors(args.enroll)
        tk, t = backend.load_vectors(args.test)
        mean = None
        if args.submean:
            mean = torch.from_numpy(np.asarray(kaldi_io.read_vec_flt(args.submean), dtype=np.float32)).cuda()
        e = backend.preprocess(e, mean, args.norm)
        t = backend.preprocess(t, mean, args.norm)
        tr_e, tr_t, _ = backend.read_trials(args.trials)
        ie, it = backend.index_trials(tr_e, tr_t, ek, tk)
        backend.write_scores(args.out_score, tr_e, tr_t, backend.cosine_score(e, t, ie, it))
    except BaseException as err:  # same contract as the reference CLIs: traceback, exit 1
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
