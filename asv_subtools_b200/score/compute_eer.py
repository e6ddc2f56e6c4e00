"""CLI twin of computeEER-like-Bosaris.py / computeEER.sh: <trials> <scores> [--method]."""
import argparse
import sys
import traceback

from . import metrics


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("trials_path")
    ap.add_argument("score_path")
    ap.add_argument("--method", default="bosaris", choices=sorted(metrics.METHODS))
    args = ap.parse_args(argv)
    try:
        labels = {}
        with open(args.trials_path) as f:
            for line in f:
                p = line.split()
                if len(p) != 3:
                    raise ValueError("trials need 3 fields: {}".format(line.strip()))
                labels[(p[0], p[1])] = p[2]
        s, lab = [], []
        with open(args.score_path) as f:
            for line in f:
                p = line.split()
                s.append(float(p[2]))
                lab.append(labels[(p[0], p[1])])
        eer, thr = metrics.METHODS[args.method](s, lab)
        print("EER% {:.3f} (threshold = {:.5f})".format(eer * 100, thr))
    except BaseException as err:
        if not isinstance(err, KeyboardInterrupt):
            traceback.print_exc()
        sys.exit(1)


if __name__ == "__main__":
    main()
