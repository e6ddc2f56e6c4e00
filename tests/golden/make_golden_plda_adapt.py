#!/usr/bin/env python
"""Golden fixture for the PLDA domain-adaptation scripts of score/pyplda, produced by the REFERENCE's own classes
(build container only):   python tests/golden/make_golden_plda_adapt.py  ->  tests/golden/plda_adapt.npz

  ivector-adapt-plda-coralplus.py  CORALPlus.update_plda          (:40-96)
  ivector-adapt-plda-lip.py        LIP.interpolation              (:25-34)
  ivector-adapt-plda-lip-reg.py    LIPReg.interpolation           (:26-49)
  ivector-adapt-plda-cip.py        CORAL.update_plda + CIP.interpolation      (:38-77, :113-121)
  ivector-adapt-plda-cip-reg.py    CORAL.update_plda + CIPReg.interpolation   (:109-128)
  plda_base.py                     PldaUnsupervisedAdaptor.update_plda        (:344-485; Kaldi's ivector-adapt-plda,
                                   the `trainaplda` step of score/process.sh:280-292)

Out-of-domain model = the d16 model of plda_train.npz (reference PldaEstimation, 10 EM iterations); in-domain model =
the reference's PldaEstimation on a second seeded set; adaptation vectors = oracle.plda_train.synthetic_adaptation_data.
The models travel through the reference's own ark reader/writer (plda_read), as in the scripts' main()."""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import plda_train as opt  # noqa: E402

PY = "/root/reference/score/pyplda"


def load(name, fname):
    spec = importlib.util.spec_from_file_location(name, os.path.join(PY, fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    sys.modules["scipye"] = types.ModuleType("scipye")
    sys.path.insert(0, "/root/reference/pytorch")
    import libs.support.kaldi_io as kio
    sys.modules["kaldi_io"] = kio
    pb = load("plda_base", "plda_base.py")
    sys.modules["plda_base"] = pb
    g = np.load(os.path.join(HERE, "plda_train.npz"))
    out_mean, out_w, out_b = g["d16_mean"], g["d16_within"], g["d16_between"]
    # in-domain model: reference EM on a second labelled set (other seed, shifted by the adaptation data's offset)
    emb, spk = opt.synthetic_plda_data(30, 16, 9)
    emb = emb * 1.4 + 0.5
    stats = pb.PldaStats(16)
    for s in np.unique(spk):
        stats.add_samples(1.0, emb[spk == s].astype(np.float64))
    stats.sort()
    est = pb.PldaEstimation(stats)
    est.estimate(num_em_iters=10)
    in_mean, in_w, in_b = np.asarray(est.mean).reshape(-1), est.within_var, est.between_var
    adapt = opt.synthetic_adaptation_data(500, 16, 77)
    out = dict(in_emb_seed=np.int64(9), in_mean=in_mean, in_within=in_w, in_between=in_b)

    def write_model(path, m, w, b):
        with kio.open_or_fd(path, "wb") as f:
            kio.write_vec_flt(f, np.asarray(m, dtype=np.float64).reshape(-1), key="mean")
            kio.write_vec_flt(f, np.asarray(w, dtype=np.float64).reshape(-1), key="within_var")
            kio.write_vec_flt(f, np.asarray(b, dtype=np.float64).reshape(-1), key="between_var")

    with tempfile.TemporaryDirectory() as d:
        p_out, p_in = os.path.join(d, "out.plda"), os.path.join(d, "in.plda")
        write_model(p_out, out_mean, out_w, out_b)
        write_model(p_in, in_mean, in_w, in_b)
        cp = load("coralplus", "ivector-adapt-plda-coralplus.py")
        c = cp.CORALPlus()
        c.plda_read(p_out)
        for v in adapt:
            c.add_stats(1, v.astype(np.float64))
        c.update_plda()
        out.update(coralplus_mean=c.mean.reshape(-1), coralplus_within=c.within_var, coralplus_between=c.between_var)
        lip = load("lip", "ivector-adapt-plda-lip.py").LIP()
        lip.interpolation(p_out, p_in)
        out.update(lip_mean=lip.mean.reshape(-1), lip_within=lip.within_var, lip_between=lip.between_var)
        lr = load("lipreg", "ivector-adapt-plda-lip-reg.py").LIPReg()
        lr.interpolation(p_out, p_in)
        out.update(lipreg_mean=lr.mean.reshape(-1), lipreg_within=lr.within_var, lipreg_between=lr.between_var)
        cm = load("cip", "ivector-adapt-plda-cip.py")
        coral = cm.CORAL()
        coral.plda_read(p_out)
        for v in adapt:
            coral.add_stats(1, v.astype(np.float64))
        coral.update_plda()
        cip = cm.CIP()
        cip.interpolation(coral, p_in)
        out.update(cip_mean=cip.mean.reshape(-1), cip_within=cip.within_var, cip_between=cip.between_var)
        crm = load("cipreg", "ivector-adapt-plda-cip-reg.py")
        coral = crm.CORAL()
        coral.plda_read(p_out)
        for v in adapt:
            coral.add_stats(1, v.astype(np.float64))
        coral.update_plda()
        cr = crm.CIPReg()
        cr.plda_read(p_in)
        cr.interpolation(coral)
        out.update(cipreg_mean=cr.mean.reshape(-1), cipreg_within=cr.within_var, cipreg_between=cr.between_var)
    # Kaldi's ivector-adapt-plda as restated by the reference itself: plda_base.PldaUnsupervisedAdaptor (:344-485);
    # stored as the adapted covariances reconstructed from (transform, psi): eigenvector order / sign free
    for tag, (ws, bs) in (("default", (0.3, 0.7)), ("scoresets", (0.70, 0.30))):
        p = pb.PLDA()
        p.mean, p.dim = out_mean.reshape(-1, 1).copy(), 16
        p.within_var, p.between_var = out_w.copy(), out_b.copy()
        p.get_output()
        ad = pb.PldaUnsupervisedAdaptor(mean_diff_scale=1.0, within_covar_scale=ws, between_covar_scale=bs)
        for v in adapt:
            ad.add_stats(1, v.astype(np.float64))
        ad.update_plda(p)
        tinv = np.linalg.inv(np.real(p.transform))
        out["unsup_%s_mean" % tag] = np.asarray(p.mean).reshape(-1)
        out["unsup_%s_within" % tag] = tinv @ tinv.T
        out["unsup_%s_between" % tag] = tinv @ np.diag(np.real(p.psi)) @ tinv.T
        out["unsup_%s_psi_sorted" % tag] = np.sort(np.real(p.psi))[::-1]
    np.savez_compressed(os.path.join(HERE, "plda_adapt.npz"), **out)
    print("plda_adapt.npz ok", sorted(out))


if __name__ == "__main__":
    main()
