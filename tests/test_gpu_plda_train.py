"""GPU PLDA training (score/plda_train.py: Gram products on the tcgen05 kernel, D x D algebra in float64
on the host) against the reference's own PldaEstimation outputs (tests/golden/plda_train.npz) and the
float64 oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import plda_train as opt
from oracle import scoring as osc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(np.max(np.abs(b)), 1e-30))


def test_gpu_plda_training_matches_reference_golden(golden):
    from asv_subtools_b200.score.plda_train import PldaEstimation, PldaStats
    spec = importlib.util.spec_from_file_location("mgp", os.path.join(HERE, "golden", "make_golden_plda.py"))
    mgp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgp)
    g = golden("plda_train")
    for name, (ns, dim, seed) in mgp.CASES.items():
        emb, spk = opt.synthetic_plda_data(ns, dim, seed)
        weights = None if not name.endswith("w") else np.random.RandomState(seed).uniform(0.5, 2.0, ns)
        # (i) the reference's call pattern: one add_samples per class
        stats = PldaStats(dim)
        for i, s in enumerate(np.unique(spk)):
            stats.add_samples(1.0 if weights is None else float(weights[i]), emb[spk == s])
        stats.sort()
        est = PldaEstimation(stats).estimate(num_em_iters=10)
        assert rel(est.stats.offset_scatter, g[name + "_scatter"]) < 2e-5      # bf16x3 Gram product
        assert rel(est.mean, g[name + "_mean"]) < 1e-6
        assert rel(est.within_var, g[name + "_within"]) < 1e-5, name
        assert rel(est.between_var, g[name + "_between"]) < 1e-5, name
        # (ii) whole matrix at once
        est2 = PldaEstimation(PldaStats.from_matrix(emb, spk, weights)).estimate(10)
        assert rel(est2.within_var, g[name + "_within"]) < 1e-5 and rel(est2.between_var, g[name + "_between"]) < 1e-5


def test_gpu_plda_training_larger_set_then_scoring_eer(tmp_path):
    from asv_subtools_b200 import kaldi_io
    from asv_subtools_b200.score import metrics
    from asv_subtools_b200.score.backend import PldaModel
    from asv_subtools_b200.score.plda_train import PldaEstimation, PldaStats
    all_emb, all_spk = opt.synthetic_plda_data(760, 64, 17, min_utts=2, max_utts=12, spread=0.7, conditioned=True)   # one population: 600 train / 160 held out
    tr = all_spk < 600
    emb, spk = all_emb[tr], all_spk[tr]
    est = PldaEstimation(PldaStats.from_matrix(torch.from_numpy(emb).cuda(), spk)).estimate(10)
    mean, within, between = opt.plda_estimate(opt.plda_stats(emb, spk), 10)
    assert rel(est.within_var, within) < 1e-5 and rel(est.between_var, between) < 1e-5 and rel(est.mean, mean) < 1e-6
    path = str(tmp_path / "plda")
    est.plda_write(path)
    parts = dict(kaldi_io.read_vec_flt_ark(path))
    assert list(parts) == ["mean", "within_var", "between_var"] and parts["within_var"].shape == (64 * 64,)
    # held-out speakers scored with the GPU-trained and with the oracle-trained model
    te_emb, te_spk = all_emb[~tr], all_spk[~tr]
    half = te_emb.shape[0] // 8 * 4                       # xvb_plda_matrix wants a multiple of 4 columns
    te_emb, te_spk = te_emb[:2 * half], te_spk[:2 * half]
    e, t = torch.from_numpy(te_emb[:half]).cuda(), torch.from_numpy(te_emb[half:]).cuda()
    lab = (te_spk[:half, None] == te_spk[None, half:]).ravel()
    s_gpu = est.model().score_matrix(e, t).cpu().numpy().ravel()
    s_ora = PldaModel(mean, within, between).score_matrix(e, t).cpu().numpy().ravel()
    e_gpu, e_ora = metrics.eer_det(s_gpu, lab)[0], metrics.eer_det(s_ora, lab)[0]
    # covariances equal to 1e-5 and well conditioned (cond ~ 10): the EERs of the two models agree to ~1e-4
    # (a square Gaussian factor instead would have cond ~ 1e4 and amplify fp32-level differences a hundredfold)
    assert 0.005 < e_ora < 0.2 and abs(e_gpu - e_ora) < 2e-4, (e_gpu, e_ora)
    G, L, c, k = osc.plda_calculate_var(between, osc.plda_smooth_within(within), mean.reshape(-1, 1))
    want = osc.plda_score_matrix(te_emb[:half], te_emb[half:], G, L, c, k).ravel()
    assert np.max(np.abs(s_gpu - want)) / np.max(np.abs(want)) < 1e-4


def test_gpu_coral_adaptation_and_diagonalised_output(golden, tmp_path):
    from asv_subtools_b200 import kaldi_io
    from asv_subtools_b200.score.plda_train import PLDA, Coral
    g = golden("plda_train")
    path = str(tmp_path / "plda.ori")
    with open(path, "wb") as f:
        kaldi_io.write_vec_flt(f, g["d16_mean"], key="mean")
        kaldi_io.write_vec_flt(f, g["d16_within"].reshape(-1), key="within_var")
        kaldi_io.write_vec_flt(f, g["d16_between"].reshape(-1), key="between_var")
    c = Coral()
    c.plda_read(path)
    for v in opt.synthetic_adaptation_data(500, 16, 77):
        c.add_stats(1, v)
    c.update_plda()
    assert rel(c.mean.reshape(-1), g["coral_mean"]) < 1e-6
    assert rel(c.within_var, g["coral_within"]) < 2e-5 and rel(c.between_var, g["coral_between"]) < 2e-5
    c.plda_write(str(tmp_path / "plda.adapt"))
    assert list(dict(kaldi_io.read_vec_flt_ark(str(tmp_path / "plda.adapt")))) == ["mean", "within_var", "between_var"]
    # get_output: T W T^T = I, T B T^T = diag(psi) with the reference's psi
    p = PLDA(g["d16_mean"], g["d16_within"], g["d16_between"])
    assert rel(np.sort(p.psi), np.sort(g["d16_psi"])) < 1e-9
    assert rel(p.transform @ g["d16_within"] @ p.transform.T, np.eye(16)) < 1e-9
    p.plda_trans_write(str(tmp_path / "plda.txt"))
    txt = open(str(tmp_path / "plda.txt")).read()
    assert txt.startswith("<Plda>  [ ") and txt.rstrip().endswith("</Plda>")


def test_gpu_plda_adaptation_family_matches_reference_golden(golden, tmp_path):
    """CORAL+ / LIP / LIP-reg / CIP / CIP-reg (score/pyplda/ivector-adapt-plda-*.py) against the reference's own classes
    (tests/golden/make_golden_plda_adapt.py), through the classes and through the CLI twin."""
    from asv_subtools_b200 import kaldi_io
    from asv_subtools_b200.score import adapt_plda
    from asv_subtools_b200.score.plda_train import Cip, CipReg, Coral, CoralPlus, Lip, LipReg, read_ori
    g, ga = golden("plda_train"), golden("plda_adapt")

    def write(path, m, w, b):
        with open(path, "wb") as f:
            kaldi_io.write_vec_flt(f, np.asarray(m).reshape(-1), key="mean")
            kaldi_io.write_vec_flt(f, np.asarray(w).reshape(-1), key="within_var")
            kaldi_io.write_vec_flt(f, np.asarray(b).reshape(-1), key="between_var")
    p_out, p_in, p_vec = str(tmp_path / "out.ori"), str(tmp_path / "in.ori"), str(tmp_path / "adapt.ark")
    write(p_out, g["d16_mean"], g["d16_within"], g["d16_between"])
    write(p_in, ga["in_mean"], ga["in_within"], ga["in_between"])
    adapt = opt.synthetic_adaptation_data(500, 16, 77)
    with open(p_vec, "wb") as f:
        for i, v in enumerate(adapt):
            kaldi_io.write_vec_flt(f, v, key="a%04d" % i)

    def coral(cls=Coral):
        c = cls()
        c.plda_read(p_out)
        c.add_matrix(adapt)
        c.update_plda()
        return c
    got = {"coralplus": coral(CoralPlus)}
    for name, cls in (("lip", Lip), ("lipreg", LipReg)):
        m = cls()
        m.interpolation(p_out, p_in)
        got[name] = m
    m = Cip()
    m.interpolation(coral(), p_in)
    got["cip"] = m
    m = CipReg()
    m.plda_read(p_in)
    m.interpolation(coral())
    got["cipreg"] = m
    for name, m in got.items():
        assert rel(m.mean.reshape(-1), ga[name + "_mean"]) < 1e-6, name
        tol = 1e-9 if name.startswith("lip") else 5e-5              # LIP*: pure float64; the rest carry the fp32-grade Gram product
        assert rel(m.within_var, ga[name + "_within"]) < tol and rel(m.between_var, ga[name + "_between"]) < tol, name
    from asv_subtools_b200.score.plda_train import PLDA, PldaUnsupervisedAdaptor
    for tag, (ws, bs) in (("default", (0.3, 0.7)), ("scoresets", (0.70, 0.30))):
        p = PLDA(g["d16_mean"], g["d16_within"], g["d16_between"])
        ad = PldaUnsupervisedAdaptor(1.0, ws, bs)
        for v in adapt:
            ad.add_stats(1, v)
        w, b = ad.update_plda(p)
        assert rel(w, ga["unsup_%s_within" % tag]) < 5e-5 and rel(b, ga["unsup_%s_between" % tag]) < 5e-5, tag
        assert rel(p.psi, ga["unsup_%s_psi_sorted" % tag]) < 5e-4 and rel(p.mean.reshape(-1), ga["unsup_%s_mean" % tag]) < 1e-6
        tinv = np.linalg.inv(p.transform)
        assert rel(tinv @ tinv.T, w) < 1e-9                        # the rewritten transform diagonalises the adapted model
    dst = str(tmp_path / "adapt_kaldi")
    adapt_plda.main(["--method", "kaldi", "--within-covar-scale", "0.70", "--between-covar-scale", "0.30", p_out, p_vec, dst])
    assert rel(read_ori(dst + ".ori")[1], ga["unsup_scoresets_within"]) < 5e-5
    for method, paths in (("coralplus", [p_out, p_vec]), ("lip-reg", [p_out, p_in]), ("cip-reg", [p_out, p_vec, p_in])):
        dst = str(tmp_path / ("adapt_" + method))
        adapt_plda.main(["--method", method] + paths + [dst])
        _, w, b = read_ori(dst + ".ori")
        key = method.replace("-", "")
        assert rel(w, ga[key + "_within"]) < 5e-5 and rel(b, ga[key + "_between"]) < 5e-5, method
        assert open(dst).read().startswith("<Plda>  [ ")


def test_gpu_kaldi_style_plda_scoring(golden):
    """PLDA.transform_ivectors / log_likelihood_ratio_matrix|trials (Kaldi semantics) against the float64 oracle, and
    against the reference's own numbers where they coincide (its vectors are ours / sqrt(D), see the oracle)."""
    from asv_subtools_b200.score.plda_train import PLDA
    g = golden("plda_train")
    p = PLDA(g["d16_mean"], g["d16_within"], g["d16_between"])
    assert rel(np.sort(p.psi), np.sort(g["kaldi_psi"])) < 1e-9
    T, off, psi = p.transform, p.offset.reshape(-1), p.psi
    ev, tv, nu = g["kaldi_enroll"].astype(np.float32), g["kaldi_test"].astype(np.float32), g["kaldi_num_utts"]
    eu = p.transform_ivectors(torch.from_numpy(ev).cuda(), nu)
    tu = p.transform_ivectors(torch.from_numpy(tv).cuda())
    want_e = np.stack([opt.plda_transform(ev[i], T, off, psi, int(nu[i])) for i in range(5)])
    want_t = np.stack([opt.plda_transform(tv[j], T, off, psi, 1) for j in range(7)])
    assert rel(eu.cpu().numpy(), want_e) < 2e-5 and rel(tu.cpu().numpy(), want_t) < 2e-5
    # the rows agree with the reference's up to the sign of each eigenvector, the order of psi (PldaEstimation.get_output
    # keeps eigh's ascending order, this class sorts descending like plda_base.PLDA.get_output) and the sqrt(D) of its dim quirk
    assert rel(np.abs(eu.cpu().numpy())[:, ::-1] / 4.0, np.abs(g["kaldi_enroll_u"])) < 2e-5
    pad = torch.zeros(1, 16, device="cuda")                                   # matmul_nt wants a multiple of 4 columns
    tu8 = torch.cat([tu, pad]).contiguous()
    S = p.log_likelihood_ratio_matrix(eu, nu, tu8).cpu().numpy()[:, :7]
    want = np.array([[opt.plda_llr(want_e[i], int(nu[i]), want_t[j], psi) for j in range(7)] for i in range(5)])
    assert np.max(np.abs(S - want)) < 2e-4 * max(1.0, np.max(np.abs(want)))
    te = torch.tensor([0, 4, 2, 2], dtype=torch.int32, device="cuda")
    tt = torch.tensor([6, 0, 3, 1], dtype=torch.int32, device="cuda")
    s = p.log_likelihood_ratio_trials(eu, nu, tu, te, tt).cpu().numpy()
    assert np.max(np.abs(s - want[[0, 4, 2, 2], [6, 0, 3, 1]])) < 2e-4 * max(1.0, np.max(np.abs(want)))
    # smoothing keeps T W T^T = I structure: psi shrinks, rows of the transform rescale
    p.smooth_within_class_covariance(0.1)
    assert np.all(p.psi < psi + 1e-12) and rel(p.transform @ g["d16_within"] @ p.transform.T, np.diag(1.0 / (1.0 + 0.1 * psi))) < 1e-9


def test_kaldi_form_plda_cli(golden, tmp_path):
    """score.sh `plda` positionals: <trials> <num_utts> <plda> <enroll> <test> <out>, Kaldi-text model file."""
    import subprocess
    import sys
    from asv_subtools_b200 import kaldi_io
    from asv_subtools_b200.score.plda_train import PLDA
    g = golden("plda_train")
    root = os.path.dirname(HERE)
    p = PLDA(g["d16_mean"], g["d16_within"], g["d16_between"])
    p.plda_trans_write(str(tmp_path / "plda"))
    q = PLDA.read_trans(str(tmp_path / "plda"))
    assert rel(q.transform, p.transform) < 1e-12 and rel(q.psi, p.psi) < 1e-12
    ev, tv, nu = g["kaldi_enroll"].astype(np.float32), g["kaldi_test"].astype(np.float32), g["kaldi_num_utts"]
    with open(tmp_path / "enroll.ark", "wb") as f:
        for i in range(5):
            kaldi_io.write_vec_flt(f, ev[i], key="spk{}".format(i))
    with open(tmp_path / "test.ark", "wb") as f:
        for j in range(7):
            kaldi_io.write_vec_flt(f, tv[j], key="utt{}".format(j))
    (tmp_path / "num_utts.ark").write_text("".join("spk{} {}\n".format(i, int(nu[i])) for i in range(5)))
    (tmp_path / "trials").write_text("".join("spk{} utt{} {}\n".format(i, j, "target" if (i + j) % 3 == 0 else "nontarget")
                                             for i in range(5) for j in range(7)))
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.score.plda", "--kaldi", str(tmp_path / "trials"),
                        str(tmp_path / "num_utts.ark"), str(tmp_path / "plda"), str(tmp_path / "enroll.ark"),
                        str(tmp_path / "test.ark"), str(tmp_path / "out.score")], capture_output=True, text=True, cwd=root,
                       env=dict(os.environ, PYTHONPATH=root))
    assert r.returncode == 0, r.stdout + r.stderr
    T, off, psi = p.transform, p.offset.reshape(-1), p.psi
    lines = [l.split() for l in open(tmp_path / "out.score")]
    assert len(lines) == 35
    for e_key, t_key, sc in lines:
        i, j = int(e_key[3:]), int(t_key[3:])
        want = opt.plda_llr(opt.plda_transform(ev[i], T, off, psi, int(nu[i])), int(nu[i]), opt.plda_transform(tv[j], T, off, psi, 1), psi)
        assert abs(float(sc) - want) < 2e-4 * max(1.0, abs(want))
