"""Pin the CPU oracle against fixtures produced by the imported reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import nnet as onn
from oracle import scoring as osc

RTOL = 2e-5  # oracle and reference run the same ATen ops; only thread/blocking order differs


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("dim,seed", [(23, 101), (80, 102)])
def test_xvector_embeddings(golden, dim, seed):
    g = golden("xvector")
    sd = onn.make_state_dict(onn.xvector_spec(dim), seed)
    feats = onn.synthetic_feats(4, 200, dim, seed + 1000)
    for pos in ("far", "near"):
        emb = np.stack([onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, pos), feats[i]).numpy()
                        for i in range(4)])
        assert emb.shape == (4, 512)
        assert rel(emb, g["xv{}_{}_emb".format(dim, pos)]) < RTOL


@pytest.mark.parametrize("dim,seed", [(23, 101), (80, 102)])
def test_xvector_layers_and_edges(golden, dim, seed):
    g = golden("xvector")
    sd = onn.make_state_dict(onn.xvector_spec(dim), seed)
    x = torch.from_numpy(onn.synthetic_feats(2, 50, dim, seed + 2000)).transpose(1, 2)
    with torch.no_grad():
        _, inter = onn.xvector_forward(sd, x, "far", return_intermediates=True)
    for name, v in inter.items():
        ref = g["xv{}_inter_{}".format(dim, name)]
        got = v.numpy() if v.shape[2] == 1 else v.numpy()[:, :8]
        assert rel(got, ref) < RTOL, name
        assert abs(np.abs(v.numpy()).mean() - g["xv{}_absmean_{}".format(dim, name)]) < 1e-4
    for T in (1, 3, 7):
        f = onn.synthetic_feats(1, T, dim, seed + 3000 + T)[0]
        e = onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, "far"), f).numpy()
        assert rel(e, g["xv{}_far_T{}".format(dim, T)]) < RTOL


def test_xvector_chunked(golden):
    g = golden("xvector")
    sd = onn.make_state_dict(onn.xvector_spec(23), 101)
    f = onn.synthetic_feats(1, 10050, 23, 4242)[0]
    e = onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, "far"), f).numpy()
    assert rel(e, g["xv23_far_T10050"]) < RTOL
    # chunking changes the answer: a single pass must differ measurably
    e1 = onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, "far"), f, max_chunk=10 ** 9).numpy()
    assert rel(e1, g["xv23_far_T10050"]) > 1e-6


def test_ecapa(golden):
    g = golden("ecapa")
    sd = onn.make_state_dict(onn.ecapa_spec(80), 201)
    feats = onn.synthetic_feats(2, 300, 80, 1201)
    for pos in ("near", "near_affine"):
        emb = np.stack([onn.extract_embedding(lambda x: onn.ecapa_forward(sd, x, pos), feats[i]).numpy()
                        for i in range(2)])
        assert rel(emb, g["ecapa80_{}_emb".format(pos)]) < RTOL
    x = torch.from_numpy(onn.synthetic_feats(2, 60, 80, 2201)).transpose(1, 2)
    with torch.no_grad():
        _, inter = onn.ecapa_forward(sd, x, "near", return_intermediates=True)
        h = inter["layer1"]
        r1 = onn.relu_bn_tdnn_layer(h, sd, "layer2.conv_relu_bn1", [0])
        r2 = onn.res2net_block(r1, sd, "layer2.res2net_block", 2)
        r3 = onn.relu_bn_tdnn_layer(r2, sd, "layer2.conv_relu_bn2", [0])
        r4 = onn.se_connect(r3, sd, "layer2.se")
        inter.update(l2_bn1=r1, l2_res2=r2, l2_bn2=r3, l2_se=r4)
        inter["stats"] = inter["stats"].unsqueeze(2)
        inter["bn_stats"] = onn.batchnorm_eval(inter["stats"].squeeze(2), sd, "bn_stats").unsqueeze(2)
    for name, v in inter.items():
        ref = g["ecapa80_inter_" + name]
        got = v.numpy() if v.shape[2] == 1 else v.numpy()[:, :8]
        assert rel(got, ref) < RTOL, name
    sd2 = onn.make_state_dict(onn.ecapa_spec(80, fc2_bn_affine=True), 202)
    for T in (2, 40):
        f = onn.synthetic_feats(1, T, 80, 3201 + T)[0]
        e = onn.extract_embedding(lambda x: onn.ecapa_forward(sd2, x, "near", fc2_relu=True), f).numpy()
        assert rel(e, g["ecapa80_default_T{}".format(T)]) < RTOL


def test_plda(golden):
    g = golden("scoring")
    within = osc.plda_smooth_within(g["plda_within"])
    G, L, c, k = osc.plda_calculate_var(g["plda_between"], within, g["plda_mean"])
    assert rel(G, g["plda_gamma"]) < 1e-12 and rel(L, g["plda_lambda"]) < 1e-12 and rel(c, g["plda_c"]) < 1e-12
    E, T = g["plda_E"], g["plda_T"]
    S = osc.plda_score_matrix(E, T, G, L, c, k)
    assert rel(S, g["plda_S"]) < 1e-12
    assert abs(osc.plda_score_pair(E[3].reshape(-1, 1), T[4].reshape(-1, 1), G, L, c, k) - g["plda_S"][3, 4]) < 1e-10


def _eer_scores(seed):
    rng = np.random.RandomState(seed)
    tar = rng.standard_normal(2000) + 2.0
    non = rng.standard_normal(50000)
    scores = np.concatenate([tar, non])
    labels = np.concatenate([np.ones(2000, dtype=np.int64), np.zeros(50000, dtype=np.int64)])
    perm = rng.permutation(scores.shape[0])
    return scores[perm], labels[perm]


def test_eer_definitions(golden):
    g = golden("scoring")
    scores, labels = _eer_scores(int(g["eer_scores_seed"]))
    eb, tb = osc.eer_bosaris_like(scores, labels)
    assert abs(eb - g["eer_bosaris"]) < 1e-12 and abs(tb - g["eer_bosaris_thr"]) < 1e-12
    ed, td = osc.eer_det_interp(scores, labels)
    assert abs(ed - g["eer_det"]) < 1e-12 and abs(td - g["eer_det_thr"]) < 1e-9
    assert abs(osc.min_dcf(scores, labels, 0.01) - g["mindcf_det"]) < 1e-12
    ek, _ = osc.eer_kaldi(scores, labels)  # unpinned definition: sanity only
    assert abs(ek - eb) < 2e-3


def test_length_norm_and_means():
    emb, lab = osc.synthetic_speakers(7, 5, 16, 11)
    n = osc.length_norm(emb)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-6)
    m, cnt = osc.speaker_mean(emb, lab, 7)
    assert np.all(cnt == 5) and np.allclose(m[2], emb[lab == 2].mean(0), atol=1e-6)
    gm = osc.global_mean(emb)
    assert np.allclose(osc.subtract_global_mean(emb, gm).mean(0), 0, atol=1e-6)


def test_score_normalization(golden):
    g = golden("score_norm")
    for method, topn in (("snorm", 0), ("asnorm", 7)):
        me, se = osc.snorm_stats(g["sn_enroll_cohort"], topn)
        mt, st = osc.snorm_stats(g["sn_test_cohort"], topn)
        out = osc.snorm_apply(g["sn_scores"], g["sn_trial_e"], g["sn_trial_t"], me, se, mt, st)
        assert np.max(np.abs(out - g["sn_" + method])) < 1e-9, method


def test_kaldi_boundary_steps_match_the_reference_trees_own_restatement(golden):
    """Row a12 (submean / norm / mean / cosine) and the ddof = 0 AS-norm against subtools2/egrecho/score run
    unmodified (tests/golden/make_golden_egrecho.py).  The reference writes scores with 5 decimals."""
    g = golden("egrecho_backend")
    emb, mean = g["emb"], g["mean"]
    assert np.max(np.abs(osc.global_mean(emb) - mean)) < 1e-7
    sm, cnt = osc.speaker_mean(g["cohort_utt"], g["cohort_spk"], g["cohort_mean"].shape[0])
    assert np.all(cnt == 5) and np.max(np.abs(sm - g["cohort_mean"])) < 1e-7
    x = osc.length_norm(osc.subtract_global_mean(emb, mean))
    cos = osc.cosine_trials(x, x, g["trial_e"], g["trial_t"])
    assert np.max(np.abs(cos - g["cosine_5dp"])) < 5.5e-6
    x0 = osc.length_norm(emb)
    assert np.max(np.abs(osc.cosine_trials(x0, x0, g["trial_e"], g["trial_t"]) - g["cosine_nosub_5dp"])) < 5.5e-6
    c = osc.length_norm(osc.subtract_global_mean(g["cohort_mean"], mean))
    top_n = int(g["top_n"])
    me, se = osc.snorm_stats(osc.cosine_matrix(x[g["stats_e_idx"]], c), top_n, ddof=0)
    mt, st = osc.snorm_stats(osc.cosine_matrix(x[g["stats_t_idx"]], c), top_n, ddof=0)
    for got, want in ((me, g["e_mean"]), (se, g["e_std"]), (mt, g["t_mean"]), (st, g["t_std"])):
        assert np.max(np.abs(got - want)) < 1e-6
    # end to end on the reference's 5-decimal cosine file, statistics indexed by utterance
    me_all, se_all = osc.snorm_stats(osc.cosine_matrix(x, c), top_n, ddof=0)
    out = osc.snorm_apply(g["cosine_5dp"], g["trial_e"], g["trial_t"], me_all, se_all, me_all, se_all)
    assert np.max(np.abs(out - g["asnorm_5dp"])) < 1e-5              # 5 decimals written by the reference


def test_zca_whitening_oracle_matches_reference_script(golden):
    g = golden("whiten")
    assert np.max(np.abs(osc.zca_whitening(g["emb"]) - g["zca"])) < 1e-6      # the script writes %f (6 decimals)


def test_lda_and_pca_oracle_properties():
    """Kaldi semantics (unpinned): the LDA output has unit tcf-mixed covariance and diagonal, descending between-class
    covariance; PCA rows are orthonormal and decorrelate the data; both centre it."""
    emb, lab = osc.synthetic_speakers(40, 6, 16, 5, noise=0.9)
    emb = emb + 0.7
    lda = osc.lda_transform(emb, lab, 6)
    y = osc.apply_affine(emb, lda)
    assert lda.shape == (6, 17) and np.max(np.abs(y.mean(0))) < 1e-9
    means = np.stack([y[lab == s].mean(0) for s in np.unique(lab)])
    btw = (means.T * np.bincount(lab)) @ means / y.shape[0]
    tot = y.T @ y / y.shape[0]
    mix = 0.1 * tot + 0.9 * (tot - btw)
    assert np.max(np.abs(mix - np.eye(6))) < 1e-9
    assert np.max(np.abs(btw - np.diag(np.diag(btw)))) < 1e-9 and np.all(np.diff(np.diag(btw)) <= 1e-12)
    pca = osc.pca_transform(emb)
    z = osc.apply_affine(emb, pca)
    cov = z.T @ z / z.shape[0]
    assert np.max(np.abs(pca[:, :16] @ pca[:, :16].T - np.eye(16))) < 1e-9
    assert np.max(np.abs(cov - np.diag(np.diag(cov)))) < 1e-9 and np.all(np.diff(np.diag(cov)) <= 1e-12)


def test_extended_xvector(golden):
    g = golden("xvector")
    sd = onn.make_state_dict(onn.extended_xvector_spec(80), 103)
    feats = onn.synthetic_feats(3, 150, 80, 1103)
    for pos in ("far", "near"):
        emb = np.stack([onn.extract_embedding(lambda x: onn.extended_xvector_forward(sd, x, pos), feats[i]).numpy()
                        for i in range(3)])
        assert rel(emb, g["ext80_{}_emb".format(pos)]) < RTOL


def test_kaldi_fbank_mfcc_oracle_matches_reference_kaldifeature(golden):
    """oracle.frontend.kaldi_fbank / kaldi_mfcc / sequence_normalize vs the reference's KaldiFeature
    (tests/golden/make_golden_fbank.py).  The reference computes in fp32 (torch FFT), the oracle in
    float64: agreement is bounded by fp32 rounding of log-mel energies."""
    import importlib.util
    import os
    from oracle import frontend as ofe
    spec = importlib.util.spec_from_file_location("mgf", os.path.join(os.path.dirname(__file__), "golden", "make_golden_fbank.py"))
    mgf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgf)
    g = golden("fbank")
    seen = 0
    for cname, (ftype, featset, mv) in mgf.CONFIGS.items():
        for wname, (n, seed) in mgf.WAVES.items():
            key = "{}_{}".format(cname, wname)
            if key not in g.files:
                continue
            wave = ofe.synthetic_wave(n, seed, sample_frequency=featset.get("sample_frequency", 16000.0))
            got = (ofe.kaldi_mfcc if ftype == "mfcc" else ofe.kaldi_fbank)(wave, **featset)
            if mv:
                if got.shape[0] == 1 and mv.get("std_norm"):
                    continue
                got = ofe.sequence_normalize(got, **mv)
            ref = g[key].astype(np.float64)
            assert got.shape == ref.shape, key
            if mv.get("std_norm") and got.shape[0] < 3:
                continue     # std over two frames amplifies fp32 noise
            tol = 2e-4 if featset.get("use_log_fbank", True) else 2e-5 * max(1.0, np.abs(ref).max())
            if ftype == "mfcc":   # the lifter scales cepstra (and their fp32 noise) by up to 1 + lifter/2
                tol *= 1.0 + 0.5 * featset.get("cepstral_lifter", 22.0)
            assert np.max(np.abs(got - ref)) < tol, (key, np.max(np.abs(got - ref)))
            seen += 1
    assert seen >= 16


def test_plda_training_oracle_matches_reference(golden):
    """oracle.plda_train vs the reference's PldaStats / PldaEstimation (10 EM iterations), and the
    diagonal-basis form of the EM step (what the GPU path computes) vs the literal per-class loop."""
    import importlib.util
    import os
    from oracle import plda_train as opt
    spec = importlib.util.spec_from_file_location("mgp", os.path.join(os.path.dirname(__file__), "golden", "make_golden_plda.py"))
    mgp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgp)
    g = golden("plda_train")
    for name, (ns, dim, seed) in mgp.CASES.items():
        emb, spk = opt.synthetic_plda_data(ns, dim, seed)
        weights = None if not name.endswith("w") else np.random.RandomState(seed).uniform(0.5, 2.0, ns)
        st = opt.plda_stats(emb, spk, weights)
        assert rel(st["offset_scatter"], g[name + "_scatter"]) < 1e-12
        mean, within, between = opt.plda_estimate(st, 10)
        assert rel(mean, g[name + "_mean"]) < 1e-12 and rel(within, g[name + "_within"]) < 1e-10
        assert rel(between, g[name + "_between"]) < 1e-10
        _, w2, b2 = opt.plda_estimate_grouped(st, 10)
        assert rel(w2, within) < 1e-9 and rel(b2, between) < 1e-9
        _, psi = opt.diagonalising_transform(within, between)
        assert rel(np.sort(psi), np.sort(g[name + "_psi"])) < 1e-9


def test_snowdar_xvector_oracle_matches_reference(golden):
    g = golden("snowdar")
    for cname, extend, seed in (("std", False, 301), ("ext", True, 302)):
        sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, extend=extend), seed)
        feats = onn.synthetic_feats(3, 120, 40, seed + 1000)
        for pos in ("far", "near_affine", "near"):
            emb = np.stack([onn.extract_embedding(lambda x: onn.snowdar_xvector_forward(sd, x, pos, extend), feats[i]).numpy()
                            for i in range(3)])
            assert rel(emb, g["{}_{}".format(cname, pos)]) < RTOL, (cname, pos)


SNOWDAR_POOLING_CASES = {
    "attn1": ("attentive", {}, 311),
    "attn2": ("attentive", {"affine_layers": 2, "hidden_size": 64}, 312),
    "mha_share": ("multi-head", {"num_head": 4}, 313),
    "mha_full": ("multi-head", {"num_head": 4, "share": False, "affine_layers": 2}, 314),
    "mres": ("multi-resolution", {"num_head": 4, "temperature": True, "affine_layers": 2}, 315),
    "lde": ("lde", {"num_head": 12, "num_nodes": 200}, 316),                        # LDEPooling(200, c_num=12): 2400-d encoding
    "xi_mean": ("xi-postmean-softplus2", {"hidden_size": 64, "num_nodes": 200}, 319),  # xi-vector, posterior mean
    "xi_dist": ("xi-postdist-softplus2", {"hidden_size": 64, "num_nodes": 200}, 320),  # ... mean | spread
}


def test_snowdar_attention_poolings_oracle_matches_reference(golden):
    """AttentiveStatisticsPooling / MultiHeadAttentionPooling / MultiResolutionMultiHeadAttentionPooling behind the
    snowdar blueprint's `pooling` switch (pooling.py:214-587, snowdar_xvector.py:119-136)."""
    g = golden("snowdar")
    for cname, (pooling, pp, seed) in SNOWDAR_POOLING_CASES.items():
        sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, pooling=pooling, pooling_params=pp), seed)
        feats = onn.synthetic_feats(3, 120, 40, seed + 1000)
        for pos in ("far", "near"):
            fwd = lambda x: onn.snowdar_xvector_forward(sd, x, pos, pooling=pooling, pooling_params=pp)  # noqa: E731
            emb = np.stack([onn.extract_embedding(fwd, feats[i]).numpy() for i in range(3)])
            assert rel(emb, g["{}_{}".format(cname, pos)]) < RTOL, (cname, pos)


def test_snowdar_bn_relu_order_oracle_matches_reference(golden):
    """tdnn_layer_params={"bn-relu": True}: affine -> BatchNorm -> ReLU (components.py:386-403)."""
    g = golden("snowdar")
    sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, bn_affine=True), 317)
    feats = onn.synthetic_feats(3, 120, 40, 1317)
    for pos in ("far", "near_affine", "near"):
        emb = np.stack([onn.extract_embedding(lambda x: onn.snowdar_xvector_forward(sd, x, pos, bn_relu=True), feats[i]).numpy()
                        for i in range(3)])
        assert rel(emb, g["bnrelu_{}".format(pos)]) < RTOL, pos


def test_ecapa_with_fc1_oracle_matches_reference(golden):
    """ECAPA_TDNN(fc1=True): far = fc1.affine, near_affine = fc1 -> fc2.affine, near = fc1 -> fc2 (ecapa_tdnn_xvector.py:412-422)."""
    g = golden("ecapa_fc1")
    sd = onn.make_state_dict(onn.ecapa_spec(80, fc1=True, fc2_bn_affine=True), 203)
    feats = onn.synthetic_feats(2, 120, 80, 1203)
    for pos in ("far", "near_affine", "near"):
        fwd = lambda x: onn.ecapa_forward(sd, x, pos, fc2_relu=True, fc1=True)  # noqa: E731
        emb = np.stack([onn.extract_embedding(fwd, feats[i]).numpy() for i in range(2)])
        assert rel(emb, g["fc1_" + pos]) < RTOL, pos


def test_coral_adaptation_oracle_matches_reference(golden):
    from oracle import plda_train as opt
    g = golden("plda_train")
    m, w, b = opt.coral_adapt(g["d16_mean"], g["d16_within"], g["d16_between"], opt.synthetic_adaptation_data(500, 16, 77))
    assert rel(m, g["coral_mean"]) < 1e-10 and rel(w, g["coral_within"]) < 1e-9 and rel(b, g["coral_between"]) < 1e-9


def test_plda_adaptation_family_oracle_matches_reference(golden):
    """CORAL+ / LIP / LIP-reg / CIP / CIP-reg restatements against the reference's own classes
    (tests/golden/make_golden_plda_adapt.py)."""
    from oracle import plda_train as opt
    g, ga = golden("plda_train"), golden("plda_adapt")
    out_m = (g["d16_mean"], g["d16_within"], g["d16_between"])
    in_m = (ga["in_mean"], ga["in_within"], ga["in_between"])
    adapt = opt.synthetic_adaptation_data(500, 16, 77)
    got = {"coralplus": opt.coralplus_adapt(*out_m, adapt), "lip": opt.lip_adapt(out_m, in_m),
           "lipreg": opt.lipreg_adapt(out_m, in_m), "cip": opt.cip_adapt(out_m, in_m, adapt),
           "cipreg": opt.cipreg_adapt(out_m, in_m, adapt)}
    for name, (m, w, b) in got.items():
        assert rel(m, ga[name + "_mean"]) < 1e-10, name
        assert rel(w, ga[name + "_within"]) < 1e-8 and rel(b, ga[name + "_between"]) < 1e-8, name
    for tag, (ws, bs) in (("default", (0.3, 0.7)), ("scoresets", (0.70, 0.30))):
        m, w, b = opt.unsupervised_adapt(*out_m, adapt, ws, bs)
        assert rel(m, ga["unsup_%s_mean" % tag]) < 1e-10
        assert rel(w, ga["unsup_%s_within" % tag]) < 1e-8 and rel(b, ga["unsup_%s_between" % tag]) < 1e-8, tag


def test_score_normalization_cross_select(golden):
    g, gc = golden("score_norm"), golden("score_norm_cross")
    for topn in (7, 19):
        out = osc.snorm_cross_apply(g["sn_scores"], g["sn_trial_e"], g["sn_trial_t"], g["sn_enroll_cohort"], g["sn_test_cohort"], topn)
        assert np.max(np.abs(out - gc["cross_top%d" % topn])) < 1e-9


def test_factored_xvector_oracle_matches_reference(golden):
    g = golden("ftdnn")
    sd = onn.make_state_dict(onn.factored_xvector_spec(40), 401)
    feats = onn.synthetic_feats(2, 90, 40, 1401)
    for pos in ("far", "near"):
        emb = np.stack([onn.extract_embedding(lambda x: onn.factored_xvector_forward(sd, x, pos), feats[i]).numpy() for i in range(2)])
        assert rel(emb, g[pos]) < RTOL, pos


def test_kaldi_style_plda_scoring_oracle_matches_reference(golden):
    """oracle.plda_train.plda_transform / plda_llr vs the reference's PLDA.transform_ivector / log_likelihood_ratio."""
    from oracle import plda_train as opt
    g = golden("plda_train")
    T, off, psi = g["kaldi_transform"], g["kaldi_offset"], g["kaldi_psi"]
    for i in range(5):
        u = opt.plda_transform(g["kaldi_enroll"][i], T, off, psi, int(g["kaldi_num_utts"][i]), reference_dim_quirk=True)
        assert rel(u, g["kaldi_enroll_u"][i]) < 1e-12
        for j in range(7):
            t = opt.plda_transform(g["kaldi_test"][j], T, off, psi, 1, reference_dim_quirk=True)
            assert abs(opt.plda_llr(u, int(g["kaldi_num_utts"][i]), t, psi) - g["kaldi_llr"][i, j]) < 1e-10
    # Kaldi semantics = the reference's vectors times sqrt(D)
    u = opt.plda_transform(g["kaldi_enroll"][1], T, off, psi, 3)
    assert rel(u, g["kaldi_enroll_u"][1] * 4.0) < 1e-12
