#!/usr/bin/env python
"""AS-norm with --cross-select true from the REFERENCE's score/ScoreNormalization.py (:146-160), on the seeded
tables already stored in tests/golden/score_norm.npz -- build container only.  -> score_norm_cross.npz"""
import argparse
import importlib.util
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("score_normalization", "/root/reference/score/ScoreNormalization.py")
    sn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sn)
    g = np.load(os.path.join(HERE, "score_norm.npz"))
    ec, tc = g["sn_enroll_cohort"], g["sn_test_cohort"]
    out = {}
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in"), "w") as f:
            for i, j, v in zip(g["sn_trial_e"], g["sn_trial_t"], g["sn_scores"]):
                f.write("e{} t{} {}\n".format(i, j, repr(float(v))))
        for name, m, p in (("ec", ec, "e"), ("tc", tc, "t")):
            with open(os.path.join(d, name), "w") as f:
                for i in range(m.shape[0]):
                    for c in range(m.shape[1]):
                        f.write("{}{} c{} {}\n".format(p, i, c, repr(float(m[i, c]))))
        for topn in (7, 19):
            ns = argparse.Namespace(method="asnorm", top_n=topn, second_cohort="true", cross_select="true",
                                    input_score=os.path.join(d, "in"), enroll_cohort_score=os.path.join(d, "ec"),
                                    test_cohort_score=os.path.join(d, "tc"), output_score=os.path.join(d, "out"))
            sn.asnorm(ns)
            out["cross_top%d" % topn] = np.array([float(l.split()[2]) for l in open(ns.output_score)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "score_norm_cross.npz"), **out)
    print("score_norm_cross.npz ok", {k: v[:3] for k, v in out.items()})


if __name__ == "__main__":
    main()
