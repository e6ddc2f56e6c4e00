#!/usr/bin/env python
"""Golden PLDA-training fixture from the REFERENCE's own score/pyplda/plda_base.py (PldaStats,
PldaEstimation) -- build container only:   python tests/golden/make_golden_plda.py

plda_base.py imports a misspelt `scipye` and the pip package `kaldi_io`: both are stubbed
(the latter aliased to the reference's own libs/support/kaldi_io.py).  Inputs are the seeded
oracle.plda_train.synthetic_plda_data sets, so only the outputs are stored: tests/golden/plda_train.npz."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import plda_train as opt  # noqa: E402

CASES = {"d16": (40, 16, 5), "d24w": (25, 24, 6)}


def main():
    sys.modules["scipye"] = types.ModuleType("scipye")
    sys.path.insert(0, "/root/reference/pytorch")
    import libs.support.kaldi_io as kio
    sys.modules["kaldi_io"] = kio
    spec = importlib.util.spec_from_file_location("plda_base", "/root/reference/score/pyplda/plda_base.py")
    pb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pb)
    out = {}
    for name, (ns, dim, seed) in CASES.items():
        emb, spk = opt.synthetic_plda_data(ns, dim, seed)
        weights = None if not name.endswith("w") else np.random.RandomState(seed).uniform(0.5, 2.0, ns)
        stats = pb.PldaStats(dim)
        for i, s in enumerate(np.unique(spk)):
            stats.add_samples(1.0 if weights is None else float(weights[i]), emb[spk == s].astype(np.float64))
        stats.sort()
        est = pb.PldaEstimation(stats)
        est.estimate(num_em_iters=10)
        out[name + "_mean"] = np.asarray(est.mean).reshape(-1)
        out[name + "_within"] = est.within_var
        out[name + "_between"] = est.between_var
        out[name + "_scatter"] = stats.offset_scatter
        plda = est.get_output()
        out[name + "_psi"] = np.asarray(plda.psi)
        if name == "d16":
            est_d16 = est
    # Kaldi-style scoring with the diagonalised d16 model: transform_ivector + log_likelihood_ratio (1-D vectors)
    plda = est_d16.get_output()
    rng = np.random.RandomState(31)
    ev, tv = rng.standard_normal((5, 16)) * 1.3 + 0.2, rng.standard_normal((7, 16)) * 1.3 + 0.2
    nu = np.array([1, 3, 2, 5, 1])
    # column vectors: the only shape transform_ivector accepts -- and for which it sets self.dim = 1 (:95), so its
    # length normalisation is sqrt(1/...) instead of Kaldi's sqrt(D/...); stored as the reference computes it
    eu = np.stack([np.asarray(plda.transform_ivector(ev[i].reshape(-1, 1), int(nu[i]))).reshape(-1) for i in range(5)])
    tu = np.stack([np.asarray(plda.transform_ivector(tv[j].reshape(-1, 1), 1)).reshape(-1) for j in range(7)])
    llr = np.array([[float(np.asarray(plda.log_likelihood_ratio(eu[i].reshape(-1, 1), int(nu[i]), tu[j].reshape(-1, 1))).reshape(-1)[0])
                     for j in range(7)] for i in range(5)])
    out.update(kaldi_enroll=ev, kaldi_test=tv, kaldi_num_utts=nu, kaldi_enroll_u=eu, kaldi_test_u=tu, kaldi_llr=llr,
               kaldi_transform=plda.transform, kaldi_psi=np.asarray(plda.psi), kaldi_offset=np.asarray(plda.offset).reshape(-1))
    # CORAL adaptation (ivector-adapt-plda-coral.py) of the d16 model to a shifted, rescaled domain
    spec = importlib.util.spec_from_file_location("coral", "/root/reference/score/pyplda/ivector-adapt-plda-coral.py")
    cm = importlib.util.module_from_spec(spec)
    sys.modules["plda_base"] = pb
    spec.loader.exec_module(cm)
    coral = cm.CORAL()
    coral.mean = out["d16_mean"].reshape(-1, 1).copy()
    coral.dim = 16
    coral.within_var, coral.between_var = out["d16_within"].copy(), out["d16_between"].copy()
    adapt = opt.synthetic_adaptation_data(500, 16, 77)
    for v in adapt:
        coral.add_stats(1, v.astype(np.float64))
    coral.update_plda()
    out["coral_mean"], out["coral_within"], out["coral_between"] = coral.mean.reshape(-1), coral.within_var, coral.between_var
    np.savez_compressed(os.path.join(HERE, "plda_train.npz"), **out)
    print("plda_train.npz ok", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
