"""CLI twin of score/pyplda/ivector-compute-plda.py (:26-73): <spk2utt-rspecifier> <ivector-rspecifier> <plda>.
Writes `<plda>.ori` (mean / within_var / between_var vectors, for adaptation and for `score.plda`) and `<plda>`
(Kaldi text format of the diagonalised model), like the reference; the EM runs on the GPU (plda_train.py).
`--adapt-coral <vectors>` additionally writes `<plda>.coral.ori` (ivector-adapt-plda-coral.py)."""
import argparse
import sys
import traceback

import numpy as np

from . import backend
from .plda_train import Coral, PldaEstimation, PldaStats


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("spk2utt")
    ap.add_argument("ivectors")
    ap.add_argument("plda")
    ap.add_argument("--num-em-iters", type=int, default=10)
    ap.add_argument("--adapt-coral", default=None, help="unlabelled in-domain vectors (ark/scp) for CORAL adaptation")
    args = ap.parse_args(argv)
    try:
        utt2spk = {}
        with open(args.spk2utt) as f:
            for line in f:
                parts = line.split()
                for utt in parts[1:]:
                    utt2spk[utt] = parts[0]
        keys, emb = backend.load_vectors(args.ivectors)
        spk_names = sorted(set(utt2spk[k] for k in keys))
        index = {s: i for i, s in enumerate(spk_names)}
        spk = np.array([index[utt2spk[k]] for k in keys], dtype=np.int32)
        est = PldaEstimation(PldaStats.from_matrix(emb, spk)).estimate(args.num_em_iters)
        est.plda_write(args.plda + ".ori")
        est.get_output().plda_trans_write(args.plda)
        if args.adapt_coral:
            _, adapt = backend.load_vectors(args.adapt_coral)
            coral = Coral()
            coral.plda_read(args.plda + ".ori")
            coral.add_matrix(adapt)
            coral.update_plda()
            coral.plda_write(args.plda + ".coral.ori")
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
