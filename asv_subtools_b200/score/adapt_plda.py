"""CLI twin of the score/pyplda domain-adaptation scripts (one entry point, `--method` picks the script):

    coral      <plda> <adapt-ivector-rspecifier> <plda-adapt>                       ivector-adapt-plda-coral.py
    coralplus  <plda> <adapt-ivector-rspecifier> <plda-adapt>                       ivector-adapt-plda-coralplus.py
    lip        <plda-out-domain> <plda-in-domain> <plda-adapt>                      ivector-adapt-plda-lip.py
    lip-reg    <plda-out-domain> <plda-in-domain> <plda-adapt>                      ivector-adapt-plda-lip-reg.py
    cip        <plda-out-domain> <adapt-ivector-rspecifier> <plda-in-domain> <plda-adapt>   ivector-adapt-plda-cip.py
    cip-reg    <plda-out-domain> <adapt-ivector-rspecifier> <plda-in-domain> <plda-adapt>   ivector-adapt-plda-cip-reg.py
    kaldi      <plda> <adapt-ivector-rspecifier> <plda-adapt>       ivector-adapt-plda.py / plda_base.PldaUnsupervisedAdaptor
               (Kaldi's ivector-adapt-plda, `trainaplda` of score/process.sh:280-292; takes --within-covar-scale,
               --between-covar-scale, --mean-diff-scale like the binary)

Inputs are the `.ori` three-vector arks (mean / within_var / between_var); like the reference's main() the result
is written as the diagonalised Kaldi text model (`PLDA.get_output` + `plda_trans_write`), plus `<plda-adapt>.ori`
for `score.plda` / further adaptation.  The adaptation set's covariance is a Gram product on the GPU; the D x D
algebra is float64 on the host."""
import argparse
import sys
import traceback

from . import backend
from .plda_train import PLDA, Cip, CipReg, Coral, CoralPlus, Lip, LipReg, PldaUnsupervisedAdaptor

ARITY = {"coral": 3, "coralplus": 3, "lip": 3, "lip-reg": 3, "cip": 4, "cip-reg": 4, "kaldi": 3}


def adapt(method, paths, scales=None):
    if method == "kaldi":
        try:
            plda = PLDA.read_trans(paths[0])
        except (ValueError, UnicodeDecodeError):
            plda = PLDA.read_ori(paths[0])
        m = PldaUnsupervisedAdaptor(**(scales or {}))
        m.add_matrix(backend.load_vectors(paths[1])[1])
        m.update_plda(plda)
        return m
    if method in ("coral", "coralplus"):
        m = (Coral if method == "coral" else CoralPlus)()
        m.plda_read(paths[0])
        m.add_matrix(backend.load_vectors(paths[1])[1])
        m.update_plda()
        return m
    if method in ("lip", "lip-reg"):
        m = (Lip if method == "lip" else LipReg)()
        m.interpolation(paths[0], paths[1])
        return m
    coral = Coral()
    coral.plda_read(paths[0])
    coral.add_matrix(backend.load_vectors(paths[1])[1])
    coral.update_plda()
    if method == "cip":
        m = Cip()
        m.interpolation(coral, paths[2])
    else:
        m = CipReg()
        m.plda_read(paths[2])
        m.interpolation(coral)
    return m


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--method", required=True, choices=sorted(ARITY))
    ap.add_argument("--within-covar-scale", type=float, default=0.3, help="--method kaldi")
    ap.add_argument("--between-covar-scale", type=float, default=0.7, help="--method kaldi")
    ap.add_argument("--mean-diff-scale", type=float, default=1.0, help="--method kaldi")
    ap.add_argument("paths", nargs="+")
    args = ap.parse_args(argv)
    if len(args.paths) != ARITY[args.method]:
        ap.error("--method {} takes {} positionals".format(args.method, ARITY[args.method]))
    try:
        m = adapt(args.method, args.paths, dict(mean_diff_scale=args.mean_diff_scale, within_covar_scale=args.within_covar_scale,
                                                between_covar_scale=args.between_covar_scale))
        PLDA(m.mean, m.within_var, m.between_var).plda_trans_write(args.paths[-1])
        Coral.plda_write(m, args.paths[-1] + ".ori")
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
