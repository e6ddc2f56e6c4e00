"""Back-end scoring parity on the GPU: pre-processing, cosine, PLDA and "EER identical to 3
decimals" against the oracle / golden fixtures, plus the CLI twins end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import nnet as onn
from oracle import scoring as osc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.fixture(scope="module")
def ops():
    from asv_subtools_b200 import ops as _ops
    return _ops


def cuda(a, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).cuda()


def test_center_length_norm_and_column_mean(ops):
    emb, _ = osc.synthetic_speakers(50, 7, 192, 3)
    emb = emb * 3 + 0.7
    mean = ops.column_mean(cuda(emb))
    assert rel(mean.cpu().numpy(), osc.global_mean(emb)) < 1e-6
    y = ops.center_length_norm(cuda(emb), mean)
    ref = osc.length_norm(osc.subtract_global_mean(emb, osc.global_mean(emb)))
    assert rel(y.cpu().numpy(), ref) < 1e-6
    y2 = ops.center_length_norm(cuda(emb))
    assert rel(y2.cpu().numpy(), osc.length_norm(emb)) < 1e-6


def test_speaker_mean(ops):
    emb, lab = osc.synthetic_speakers(9, 6, 192, 8)
    perm = np.random.RandomState(1).permutation(emb.shape[0])
    emb, lab = emb[perm], lab[perm]
    spk2rows = [np.flatnonzero(lab == s).tolist() for s in range(9)]
    spk2rows[3] = spk2rows[3][:2]                       # ragged speakers
    means, counts = ops.speaker_mean(cuda(emb), spk2rows)
    for s, rows in enumerate(spk2rows):
        assert counts[s] == len(rows)
        assert np.allclose(means[s].cpu().numpy(), emb[rows].astype(np.float64).mean(0), atol=1e-6)
    ref, cnt = osc.speaker_mean(emb, lab, 9)
    assert np.allclose(means[0].cpu().numpy(), ref[0], atol=1e-6) and cnt[0] == 6


def test_cosine_trials_and_matrix(ops):
    e, _ = osc.synthetic_speakers(40, 3, 512, 4)
    t, _ = osc.synthetic_speakers(30, 4, 512, 5)
    e, t = osc.length_norm(e), osc.length_norm(t)
    rng = np.random.RandomState(0)
    te = rng.randint(0, e.shape[0], 500).astype(np.int32)
    tt = rng.randint(0, t.shape[0], 500).astype(np.int32)
    s = ops.cosine_trials(cuda(e), cuda(t), cuda(te, np.int32), cuda(tt, np.int32))
    assert np.max(np.abs(s.cpu().numpy() - osc.cosine_trials(e, t, te, tt))) < 2e-6
    S = ops.cosine_matrix(cuda(e), cuda(t))
    assert S.shape == (120, 120)
    assert np.max(np.abs(S.cpu().numpy() - osc.cosine_matrix(e, t))) < 2e-5  # bf16x3 GEMM, |s| <= 1


def test_plda_matches_reference_golden(golden):
    from asv_subtools_b200.score.backend import PldaModel
    g = golden("scoring")
    m = PldaModel(g["plda_mean"], g["plda_within"], g["plda_between"])
    assert rel(m.gamma, g["plda_gamma"]) < 1e-12 and rel(m.lam, g["plda_lambda"]) < 1e-12
    assert rel(m.c, g["plda_c"].reshape(-1)) < 1e-12
    E, T = cuda(g["plda_E"]), cuda(g["plda_T"][:8])   # Nt % 4 == 0 for the matrix form
    S = m.score_matrix(E, T).cpu().numpy()
    assert rel(S, g["plda_S"][:, :8]) < 5e-5
    te = np.repeat(np.arange(12), 9).astype(np.int32)
    tt = np.tile(np.arange(9), 12).astype(np.int32)
    s = m.score_trials(E, cuda(g["plda_T"]), cuda(te, np.int32), cuda(tt, np.int32)).cpu().numpy()
    assert rel(s.reshape(12, 9), g["plda_S"]) < 5e-5


def _trials(lab_e, lab_t, rng, n):
    te = rng.randint(0, lab_e.shape[0], n).astype(np.int32)
    tt = rng.randint(0, lab_t.shape[0], n).astype(np.int32)
    # make ~10% of the trials targets
    k = n // 10
    for i in range(k):
        cand = np.flatnonzero(lab_t == lab_e[te[i]])
        tt[i] = cand[rng.randint(cand.size)]
    return te, tt, (lab_e[te] == lab_t[tt]).astype(np.int64)


def test_eer_identical_to_three_decimals_cosine_and_plda(ops):
    """North-star criterion on synthetic trials: EER(new scores) == EER(reference-arithmetic scores)
    to 3 decimals (in %), under each in-repo EER definition."""
    from asv_subtools_b200.score import metrics
    from asv_subtools_b200.score.backend import PldaModel
    emb, lab = osc.synthetic_speakers(300, 8, 192, 21, noise=1.6)
    rng = np.random.RandomState(9)
    te, tt, y = _trials(lab, lab, rng, 200000)
    # --- cosine with submean + norm (score/process.sh + score.sh)
    ref_x = osc.length_norm(osc.subtract_global_mean(emb, osc.global_mean(emb)))
    ref_s = osc.cosine_trials(ref_x, ref_x, te, tt)
    x = cuda(emb)
    xn = ops.center_length_norm(x, ops.column_mean(x))
    s = ops.cosine_trials(xn, xn, cuda(te, np.int32), cuda(tt, np.int32)).cpu().numpy()
    for name, fn_ref, fn_new in (("bosaris", osc.eer_bosaris_like, metrics.eer_bosaris),
                                 ("det", osc.eer_det_interp, metrics.eer_det)):
        e_ref, e_new = fn_ref(ref_s, y)[0], fn_new(s, y)[0]
        assert 0.01 < e_ref < 0.3, e_ref
        assert round(e_ref * 100, 3) == round(e_new * 100, 3), (name, e_ref, e_new)
    # --- PLDA with a synthetic two-covariance model
    d = emb.shape[1]
    a = rng.standard_normal((d, d))
    within = a @ a.T / d * 2.0 + np.eye(d)
    between = np.eye(d) * 1.0 + 0.05 * (a + a.T) / np.sqrt(d)
    between = between @ between.T
    mean = emb.mean(0).astype(np.float64).reshape(-1, 1)
    G, L, c, k = osc.plda_calculate_var(between, osc.plda_smooth_within(within), mean)
    row = np.einsum("ij,jk,ik->i", emb.astype(np.float64), G, emb.astype(np.float64)) + emb.astype(np.float64) @ c.reshape(-1)
    ref_p = np.einsum("ij,jk,ik->i", emb[te].astype(np.float64), L + L.T, emb[tt].astype(np.float64)) + row[te] + row[tt]
    m = PldaModel(mean, within, between)
    p = m.score_trials(x, x, cuda(te, np.int32), cuda(tt, np.int32)).cpu().numpy()
    assert rel(p, ref_p) < 1e-4
    for fn_ref, fn_new in ((osc.eer_bosaris_like, metrics.eer_bosaris), (osc.eer_det_interp, metrics.eer_det)):
        assert round(fn_ref(ref_p, y)[0] * 100, 3) == round(fn_new(p, y)[0] * 100, 3)


def test_score_normalization_matches_reference_golden(ops, golden, tmp_path):
    """S-norm / AS-norm kernels against score/ScoreNormalization.py run on the same tables; then the
    embedding-level AS-norm against the oracle on a larger cohort; then the CLI twin."""
    from asv_subtools_b200.score import normalization as norm
    g = golden("score_norm")
    ec, tc = cuda(g["sn_enroll_cohort"]), cuda(g["sn_test_cohort"])
    te, tt = cuda(g["sn_trial_e"], np.int32), cuda(g["sn_trial_t"], np.int32)
    for method, topn in (("snorm", 0), ("asnorm", 7)):
        out = norm.normalize(cuda(g["sn_scores"]), te, tt, ec, tc, topn).cpu().numpy()
        assert np.max(np.abs(out - g["sn_" + method]) / (1 + np.abs(g["sn_" + method]))) < 2e-5, method
    # embeddings -> cohort GEMMs -> top-300 of a 2000-utterance cohort
    emb, lab = osc.synthetic_speakers(60, 5, 192, 77, noise=1.2)
    coh, _ = osc.synthetic_speakers(400, 5, 192, 78, noise=1.2)
    x = osc.length_norm(emb)
    c = osc.length_norm(coh)
    rng = np.random.RandomState(3)
    ie = rng.randint(0, 300, 5000).astype(np.int32)
    it = rng.randint(0, 300, 5000).astype(np.int32)
    got = norm.asnorm_embeddings(cuda(x), cuda(x), cuda(c), cuda(ie, np.int32), cuda(it, np.int32), top_n=300).cpu().numpy()
    sc = osc.cosine_matrix(x, c)
    me, se = osc.snorm_stats(sc, 300)
    ref = osc.snorm_apply(osc.cosine_trials(x, x, ie, it), ie, it, me, se, me, se)
    assert np.max(np.abs(got - ref) / (1 + np.abs(ref))) < 1e-3
    y = (lab[ie] == lab[it]).astype(np.int64)
    from asv_subtools_b200.score import metrics
    assert round(metrics.eer_bosaris(got, y)[0] * 100, 2) == round(osc.eer_bosaris_like(ref, y)[0] * 100, 2)
    # CLI twin on score files
    keys_e = ["e{}".format(i) for i in range(6)]
    keys_t = ["t{}".format(j) for j in range(9)]
    with open(tmp_path / "in", "w") as f:
        for i, j, v in zip(g["sn_trial_e"], g["sn_trial_t"], g["sn_scores"]):
            f.write("{} {} {}\n".format(keys_e[i], keys_t[j], repr(float(v))))
    for name, keys, tab in (("ec", keys_e, g["sn_enroll_cohort"]), ("tc", keys_t, g["sn_test_cohort"])):
        with open(tmp_path / name, "w") as f:
            for i, k in enumerate(keys):
                for cidx in range(tab.shape[1]):
                    f.write("{} c{} {}\n".format(k, cidx, repr(float(tab[i, cidx]))))
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.score.normalization", "--method", "asnorm", "--top-n", "7",
                        str(tmp_path / "in"), str(tmp_path / "ec"), str(tmp_path / "tc"), str(tmp_path / "out")],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    vals = np.array([float(l.split()[2]) for l in open(tmp_path / "out")])
    assert np.max(np.abs(vals - g["sn_asnorm"]) / (1 + np.abs(g["sn_asnorm"]))) < 2e-5


def test_kaldi_boundary_steps_match_egrecho_golden(ops, golden):
    """SURVEY 8a row a12 on the GPU against the reference tree's own restatement (subtools2/egrecho/score, run
    unmodified by tests/golden/make_golden_egrecho.py): column mean, speaker mean, submean + length-norm, per-trial
    cosine (5-decimal score file), cohort top-n statistics with ddof = 0, AS-norm end to end."""
    from asv_subtools_b200.score import normalization as norm
    g = golden("egrecho_backend")
    emb, mean = cuda(g["emb"]), g["mean"]
    assert np.max(np.abs(ops.column_mean(emb).cpu().numpy() - mean)) < 1e-6
    rows = [np.flatnonzero(g["cohort_spk"] == s) for s in range(g["cohort_mean"].shape[0])]
    sm, cnt = ops.speaker_mean(cuda(g["cohort_utt"]), rows)
    assert np.all(cnt == 5) and np.max(np.abs(sm.cpu().numpy() - g["cohort_mean"])) < 1e-6
    x = ops.center_length_norm(emb, cuda(mean))
    te, tt = cuda(g["trial_e"], np.int32), cuda(g["trial_t"], np.int32)
    cos = ops.cosine_trials(x, x, te, tt).cpu().numpy()
    assert np.max(np.abs(cos - g["cosine_5dp"])) < 6e-6
    x0 = ops.center_length_norm(emb, None)
    assert np.max(np.abs(ops.cosine_trials(x0, x0, te, tt).cpu().numpy() - g["cosine_nosub_5dp"])) < 6e-6
    c = ops.center_length_norm(cuda(g["cohort_mean"]), cuda(mean))
    top_n = int(g["top_n"])
    ie, it = torch.from_numpy(g["stats_e_idx"].astype(np.int64)).cuda(), torch.from_numpy(g["stats_t_idx"].astype(np.int64)).cuda()
    me, se = ops.topn_mean_std(ops.cosine_matrix(x[ie].contiguous(), c), top_n, ddof=0)
    mt, st = ops.topn_mean_std(ops.cosine_matrix(x[it].contiguous(), c), top_n, ddof=0)
    for got, want in ((me, g["e_mean"]), (se, g["e_std"]), (mt, g["t_mean"]), (st, g["t_std"])):
        assert np.max(np.abs(got.cpu().numpy() - want)) < 2e-5      # bf16x3 cohort GEMM: ~1e-5 on a cosine
    sc = cuda(g["cosine_5dp"].astype(np.float32))
    out = norm.normalize(sc, te, tt, ops.cosine_matrix(x, c), ops.cosine_matrix(x, c), top_n, ddof=0).cpu().numpy()
    assert np.max(np.abs(out - g["asnorm_5dp"]) / (1 + np.abs(g["asnorm_5dp"]))) < 2e-4   # scores / std ~ 10x amplification
    got = norm.asnorm_embeddings(x, x, c, te, tt, top_n=top_n, ddof=0).cpu().numpy()
    assert np.max(np.abs(got - g["asnorm_5dp"]) / (1 + np.abs(g["asnorm_5dp"]))) < 5e-4


def test_process_steps_cli_matches_oracle_and_reference_whitening(ops, golden, tmp_path):
    """score/process.sh twin (asv_subtools_b200.score.process) through files, step by step: getmean / submean / norm /
    mean against the oracle, trainwhiten against the reference's own ZCA script (tests/golden/whiten.npz), trainlda /
    trainpcawhiten against the oracle's restatement of the Kaldi binaries (rows compared up to sign), and the
    affine `transform` with an output dimension that is not a multiple of 4."""
    from asv_subtools_b200 import kaldi_io
    from asv_subtools_b200.score import process as proc
    emb, lab = osc.synthetic_speakers(40, 6, 16, 5, noise=0.9)
    emb = (emb + 0.7).astype(np.float32)
    keys = ["u%03d" % i for i in range(emb.shape[0])]
    ark = str(tmp_path / "xv.ark")
    with open(ark, "wb") as f:
        for k, v in zip(keys, emb):
            kaldi_io.write_vec_flt(f, v, key=k)
    with open(tmp_path / "utt2spk", "w") as f:
        for k, s in zip(keys, lab):
            f.write("%s s%02d\n" % (k, s))
    with open(tmp_path / "spk2utt", "w") as f:
        for s in np.unique(lab):
            f.write("s%02d %s\n" % (s, " ".join(k for k, l in zip(keys, lab) if l == s)))
    P = lambda n: str(tmp_path / n)  # noqa: E731

    def vecs(path):
        d = dict(kaldi_io.read_vec_flt_ark(path))
        return np.stack([d[k] for k in keys if k in d]) if keys[0] in d else d
    proc.main(["getmean", ark, P("mean.vec")])
    mean = kaldi_io.read_vec_flt(P("mean.vec"))
    assert np.max(np.abs(mean - osc.global_mean(emb))) < 1e-6
    proc.main(["submean", P("mean.vec"), ark, P("sub.ark")])
    assert np.max(np.abs(vecs(P("sub.ark")) - osc.subtract_global_mean(emb, osc.global_mean(emb)))) < 1e-6
    proc.main(["norm", P("sub.ark"), P("norm.ark")])
    assert np.max(np.abs(vecs(P("norm.ark")) - osc.length_norm(osc.subtract_global_mean(emb, osc.global_mean(emb))))) < 1e-6
    proc.main(["mean", P("spk2utt"), ark, P("spk.ark"), P("num_utts.ark")])
    sm, cnt = osc.speaker_mean(emb, lab, 40)
    got = vecs(P("spk.ark"))
    assert np.max(np.abs(np.stack([got["s%02d" % s] for s in range(40)]) - sm)) < 1e-6
    assert [l.split() for l in open(P("num_utts.ark"))] == [["s%02d" % s, str(int(c))] for s, c in enumerate(cnt)]

    def same_rows_up_to_sign(a, b, tol):
        d = a.shape[1] - 1
        sign = np.sign(np.sum(a[:, :d] * b[:, :d], axis=1))[:, None]
        return np.max(np.abs(a * sign - b)) / np.max(np.abs(b)) < tol
    proc.main(["trainlda", "--dim", "6", ark, P("utt2spk"), P("lda.mat")])
    lda = kaldi_io.read_mat(P("lda.mat"))
    assert lda.shape == (6, 17) and same_rows_up_to_sign(lda.astype(np.float64), osc.lda_transform(emb, lab, 6), 2e-4)
    proc.main(["lda", P("lda.mat"), ark, P("lda.ark")])
    assert rel(vecs(P("lda.ark")), osc.apply_affine(emb, lda.astype(np.float64))) < 1e-4
    proc.main(["trainpcawhiten", ark, P("pca.mat")])
    assert same_rows_up_to_sign(kaldi_io.read_mat(P("pca.mat")).astype(np.float64), osc.pca_transform(emb), 2e-4)
    g = golden("whiten")
    zark = str(tmp_path / "z.ark")
    with open(zark, "wb") as f:
        for i, v in enumerate(g["emb"]):
            kaldi_io.write_vec_flt(f, v, key="z%03d" % i)
    proc.main(["trainwhiten", zark, P("zca.mat")])
    zca = kaldi_io.read_mat(P("zca.mat"))
    assert np.max(np.abs(zca - g["zca"])) / np.max(np.abs(g["zca"])) < 1e-4
    proc.main(["whiten", P("zca.mat"), zark, P("zw.ark")])
    zd = dict(kaldi_io.read_vec_flt_ark(P("zw.ark")))
    zw = np.stack([zd["z%03d" % i] for i in range(g["emb"].shape[0])])
    assert rel(zw, osc.apply_affine(g["emb"], g["zca"])) < 1e-4


def test_asnorm_cross_select_matches_reference_golden(ops, golden, tmp_path):
    """--cross-select true: top-n (score, index) sort + per-trial gather statistics against the reference's pandas
    merge (tests/golden/make_golden_snorm_cross.py), a larger random case against the oracle, and the CLI."""
    from asv_subtools_b200.score import normalization as norm
    g, gc = golden("score_norm"), golden("score_norm_cross")
    ec, tc = cuda(g["sn_enroll_cohort"]), cuda(g["sn_test_cohort"])
    te, tt = cuda(g["sn_trial_e"], np.int32), cuda(g["sn_trial_t"], np.int32)
    for topn in (7, 19):
        out = norm.normalize(cuda(g["sn_scores"]), te, tt, ec, tc, topn, cross_select=True).cpu().numpy()
        ref = gc["cross_top%d" % topn]
        assert np.max(np.abs(out - ref) / (1 + np.abs(ref))) < 2e-5, topn
    idx = ops.topn_indices(ec, 7).cpu().numpy()
    assert np.array_equal(idx, np.argsort(-g["sn_enroll_cohort"], axis=1, kind="stable")[:, :7])
    rng = np.random.RandomState(9)
    E, T = rng.standard_normal((50, 1000)).astype(np.float32), rng.standard_normal((70, 1000)).astype(np.float32)
    ie, it = rng.randint(0, 50, 3000).astype(np.int32), rng.randint(0, 70, 3000).astype(np.int32)
    sc = rng.standard_normal(3000).astype(np.float32)
    got = norm.normalize(cuda(sc), cuda(ie, np.int32), cuda(it, np.int32), cuda(E), cuda(T), 300, cross_select=True).cpu().numpy()
    ref = osc.snorm_cross_apply(sc, ie, it, E, T, 300)
    assert np.max(np.abs(got - ref) / (1 + np.abs(ref))) < 2e-5
    with open(tmp_path / "in", "w") as f:
        for i, j, v in zip(g["sn_trial_e"], g["sn_trial_t"], g["sn_scores"]):
            f.write("e{} t{} {}\n".format(i, j, repr(float(v))))
    for name, p, tab in (("ec", "e", g["sn_enroll_cohort"]), ("tc", "t", g["sn_test_cohort"])):
        with open(tmp_path / name, "w") as f:
            for i in range(tab.shape[0]):
                for cidx in range(tab.shape[1]):
                    f.write("{}{} c{} {}\n".format(p, i, cidx, repr(float(tab[i, cidx]))))
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.score.normalization", "--method", "asnorm", "--top-n", "19",
                        "--cross-select", "true", str(tmp_path / "in"), str(tmp_path / "ec"), str(tmp_path / "tc"),
                        str(tmp_path / "out")], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT), cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    vals = np.array([float(l.split()[2]) for l in open(tmp_path / "out")])
    assert np.max(np.abs(vals - gc["cross_top19"]) / (1 + np.abs(gc["cross_top19"]))) < 2e-5


def test_extract_and_score_clis_end_to_end(tmp_path):
    """feats.ark + checkpoint + nnet.config -> extract CLI -> xvector.ark -> cosine CLI -> EER CLI,
    compared with the oracle running the reference arithmetic on the same files."""
    from asv_subtools_b200 import kaldi_io
    dim, n_spk, per = 24, 12, 4
    sd = onn.make_state_dict(onn.xvector_spec(dim), 31)
    torch.save(sd, str(tmp_path / "final.params"))
    bp = os.path.join(ROOT, "asv_subtools_b200", "model", "xvector.py")
    (tmp_path / "nnet.config").write_text(
        'model_blueprint;{}\nmodel_creation;"Xvector({},10,training=False,extracted_embedding=""far"")"\n'.format(bp, dim))
    rng = np.random.RandomState(1)
    spk_dir = rng.standard_normal((n_spk, dim)).astype(np.float32)
    feats, labels = {}, {}
    for s in range(n_spk):
        for u in range(per):
            T = [60, 60, 75, 90][u]
            key = "spk{:02d}-utt{}".format(s, u)
            feats[key] = (rng.standard_normal((T, dim)) + 0.8 * spk_dir[s]).astype(np.float32)
            labels[key] = s
    with open(tmp_path / "feats.ark", "wb") as f:
        for k, v in feats.items():
            kaldi_io.write_mat(f, v, key=k)
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.pipeline.extract_embeddings", "--nnet-config",
                        str(tmp_path / "nnet.config"), "--batch-size", "8", str(tmp_path / "final.params"),
                        "ark:" + str(tmp_path / "feats.ark"), "ark:" + str(tmp_path / "xvector.ark")],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Process utterance for key") == len(feats)
    got = dict(kaldi_io.read_vec_flt_ark(str(tmp_path / "xvector.ark")))
    assert set(got) == set(feats)
    ref = {}
    for k, v in feats.items():
        ref[k] = onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, "far"), v).numpy()
        assert got[k].dtype == np.float32 and rel(got[k], ref[k]) < 1e-4
    keys = sorted(feats)
    with open(tmp_path / "trials", "w") as f:
        for a in keys[::3]:
            for b in keys:
                f.write("{} {} {}\n".format(a, b, "target" if labels[a] == labels[b] else "nontarget"))
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.score.cosine", "--norm", str(tmp_path / "trials"),
                        str(tmp_path / "xvector.ark"), str(tmp_path / "xvector.ark"), str(tmp_path / "cos.score")],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([sys.executable, "-m", "asv_subtools_b200.score.compute_eer", str(tmp_path / "trials"),
                        str(tmp_path / "cos.score")], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0 and "EER%" in r.stdout, r.stdout + r.stderr
    eer_cli = float(r.stdout.split("EER%")[1].split()[0])
    X = osc.length_norm(np.stack([ref[k] for k in keys]))
    idx = {k: i for i, k in enumerate(keys)}
    s, y = [], []
    for line in open(tmp_path / "trials"):
        a, b, l = line.split()
        s.append(float(X[idx[a]].astype(np.float64) @ X[idx[b]].astype(np.float64)))
        y.append(1 if l == "target" else 0)
    assert round(osc.eer_bosaris_like(s, y)[0] * 100, 3) == round(eer_cli, 3)
