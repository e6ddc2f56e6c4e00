#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

It imports the reference's own modules (libs.nnet, model/xvector.py,
model/ecapa_tdnn_xvector.py, libs/support/kaldi_io.py, score/pyplda/gaussian-plda-scoring.py,
computeEER-like-Bosaris.py, subtools2/egrecho/score/binary_metrics.py), feeds them seeded
synthetic checkpoints/inputs produced by ``oracle.nnet.make_state_dict`` /
``synthetic_feats`` (so the fixtures only have to store the *outputs*), and writes small
``.npz`` files.  ``tests/test_oracle_golden.py`` replays them against the oracle;
the ``-m gpu`` tests replay them against the CUDA path.
"""
import importlib.util
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import nnet as onn  # noqa: E402
from oracle import scoring as osc  # noqa: E402


def import_reference():
    # libs/nnet/transformer imports tkinter/turtle by accident.
    for name, attrs in (("tkinter", {"N": "n"}), ("tkinter.messagebox", {"NO": "no"}), ("turtle", {"xcor": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
    sys.path.insert(0, os.path.join(REF, "pytorch"))
    import libs.support.utils as utils
    import libs.support.kaldi_io as kaldi_io
    return utils, kaldi_io


def load_file_module(name, path, pre=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    for k, v in (pre or {}).items():
        sys.modules[k] = v
    spec.loader.exec_module(mod)
    return mod


ECAPA_CANON = ('ECAPA_TDNN({dim},10,training=False,extracted_embedding="{pos}",'
               'ecapa_params={{"channels":1024,"embd_dim":192,"mfa_conv":1536,'
               '"bn_params":{{"momentum":0.5,"affine":True,"track_running_stats":True}}}},'
               'pooling="ecpa-attentive",pooling_params={{"hidden_size":128,"time_attention":True,"stddev":True}},'
               'fc1=False,fc2_params={{"nonlinearity":"","nonlinearity_params":{{"inplace":True}},"bn-relu":False,'
               '"bn":True,"bn_params":{{"momentum":0.5,"affine":False,"track_running_stats":True}}}})')


def slim(inter, nch=8):
    """Keep fixtures small: first `nch` channels of every intermediate + its mean |x|."""
    out = {}
    for k, v in inter.items():
        v = v.detach().numpy()
        out["inter_" + k] = v.copy() if v.shape[2] == 1 else v[:, :nch].copy()
        out["absmean_" + k] = np.float64(np.abs(v).mean())
    return out


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    utils, kaldi_io = import_reference()

    # ------------------------------------------------------------------ x-vector
    def ref_xvector(dim, pos):
        m = utils.create_model_from_py(os.path.join(REF, "pytorch/model/xvector.py"),
                                       'Xvector({},10,training=False,extracted_embedding="{}")'.format(dim, pos))
        m.eval()
        return m

    out = {}
    for dim, seed in ((23, 101), (80, 102)):
        sd = onn.make_state_dict(onn.xvector_spec(dim), seed)
        for pos in ("far", "near"):
            m = ref_xvector(dim, pos)
            m.load_state_dict(sd, strict=True)  # asserts oracle spec == reference keys/shapes
            feats = onn.synthetic_feats(4, 200, dim, seed + 1000)
            emb = np.stack([m.extract_embedding(feats[i]).numpy() for i in range(feats.shape[0])])
            out["xv{}_{}_emb".format(dim, pos)] = emb
        # intermediates through the reference's own layers (batched call, B=2, T=50)
        m = ref_xvector(dim, "far")
        m.load_state_dict(sd, strict=True)
        x = torch.from_numpy(onn.synthetic_feats(2, 50, dim, seed + 2000)).transpose(1, 2)
        inter = {}
        with torch.no_grad():
            h = x
            for name in ("tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5"):
                h = getattr(m, name)(h)
                inter[name] = h
            h = m.stats(h)
            inter["stats"] = h
        for k, v in slim(inter).items():
            out["xv{}_{}".format(dim, k)] = v
        # edge lengths: T=1, T=3, T=7 (zero padding dominates)
        for T in (1, 3, 7):
            f = onn.synthetic_feats(1, T, dim, seed + 3000 + T)[0]
            out["xv{}_far_T{}".format(dim, T)] = m.extract_embedding(f).numpy()
    # chunked extraction: T=10050 -> two chunks of 5025 (framework.py:34-47)
    sd = onn.make_state_dict(onn.xvector_spec(23), 101)
    m = ref_xvector(23, "far")
    m.load_state_dict(sd, strict=True)
    f = onn.synthetic_feats(1, 10050, 23, 4242)[0]
    out["xv23_far_T10050"] = m.extract_embedding(f).numpy()
    # extended x-vector (pytorch/model/extended_xvector.py)
    sd = onn.make_state_dict(onn.extended_xvector_spec(80), 103)
    for pos in ("far", "near"):
        m = utils.create_model_from_py(os.path.join(REF, "pytorch/model/extended_xvector.py"),
                                       'ExtendedXvector(80,10,training=False,extracted_embedding="{}")'.format(pos))
        m.eval()
        m.load_state_dict(sd, strict=True)
        feats = onn.synthetic_feats(3, 150, 80, 1103)
        out["ext80_{}_emb".format(pos)] = np.stack([m.extract_embedding(feats[i]).numpy() for i in range(3)])
    np.savez_compressed(os.path.join(HERE, "xvector.npz"), **out)
    print("xvector.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

    # ------------------------------------------------------------------ ECAPA
    out = {}
    sd = onn.make_state_dict(onn.ecapa_spec(80), 201)
    for pos in ("near", "near_affine"):
        m = utils.create_model_from_py(os.path.join(REF, "pytorch/model/ecapa_tdnn_xvector.py"),
                                       ECAPA_CANON.format(dim=80, pos=pos))
        m.eval()
        m.load_state_dict(sd, strict=True)
        feats = onn.synthetic_feats(2, 300, 80, 1201)
        out["ecapa80_{}_emb".format(pos)] = np.stack([m.extract_embedding(feats[i]).numpy() for i in range(2)])
    x = torch.from_numpy(onn.synthetic_feats(2, 60, 80, 2201)).transpose(1, 2)
    inter = {}
    with torch.no_grad():
        h = m.layer1(x)
        inter["layer1"] = h
        x1 = m.layer2(h)
        x2 = m.layer3(h + x1)
        x3 = m.layer4(h + x1 + x2)
        inter["layer2"], inter["layer3"], inter["layer4"] = x1, x2, x3
        # inside layer2 (for the Res2Net / SE unit tests)
        r = m.layer2.conv_relu_bn1(h)
        inter["l2_bn1"] = r
        r = m.layer2.res2net_block(r)
        inter["l2_res2"] = r
        r = m.layer2.conv_relu_bn2(r)
        inter["l2_bn2"] = r
        inter["l2_se"] = m.layer2.se(r)
        hh = m.mfa(torch.cat([x1, x2, x3], dim=1))
        inter["mfa"] = hh
        st = m.stats(hh)
        inter["stats"] = st.unsqueeze(2)
        inter["bn_stats"] = m.bn_stats(st).unsqueeze(2)
    for k, v in slim(inter).items():
        out["ecapa80_" + k] = v
    # blueprint-default fc2 (ReLU + affine BN) variant, T=2 edge (T=1 is NaN in the reference)
    sd2 = onn.make_state_dict(onn.ecapa_spec(80, fc2_bn_affine=True), 202)
    m2 = utils.create_model_from_py(os.path.join(REF, "pytorch/model/ecapa_tdnn_xvector.py"),
                                    'ECAPA_TDNN(80,10,training=False)')
    m2.eval()
    m2.load_state_dict(sd2, strict=True)
    for T in (2, 40):
        f = onn.synthetic_feats(1, T, 80, 3201 + T)[0]
        out["ecapa80_default_T{}".format(T)] = m2.extract_embedding(f).numpy()
    np.savez_compressed(os.path.join(HERE, "ecapa.npz"), **out)
    print("ecapa.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

    # ------------------------------------------------------------------ PLDA + EER
    out = {}
    gps = load_file_module("gaussian_plda_scoring", os.path.join(REF, "score/pyplda/gaussian-plda-scoring.py"))
    rng = np.random.RandomState(301)
    D = 24
    a = rng.standard_normal((D, D))
    b = rng.standard_normal((D, D))
    within = a @ a.T / D + 0.5 * np.eye(D)
    between = b @ b.T / D + 0.1 * np.eye(D)
    mean = rng.standard_normal((D, 1)) * 0.1
    within_s = within + 5e-5 * np.eye(D)  # gaussian-plda-scoring.py:65
    G, L, c, k = gps.CalculateVar(between, within_s, mean)
    E = rng.standard_normal((12, D))
    T = rng.standard_normal((9, D))
    S = np.zeros((12, 9))
    for i in range(12):
        for j in range(9):
            S[i, j] = gps.PLDAScoring(E[i].reshape(-1, 1), T[j].reshape(-1, 1), G, L, c, k)
    out.update(plda_within=within, plda_between=between, plda_mean=mean, plda_gamma=G, plda_lambda=L,
               plda_c=c, plda_E=E, plda_T=T, plda_S=S)

    bos = load_file_module("compute_eer_bosaris", os.path.join(REF, "computeEER-like-Bosaris.py"))
    eg = types.ModuleType("egrecho"); eg.__path__ = []
    egs = types.ModuleType("egrecho.score"); egs.__path__ = []
    egu = load_file_module("egrecho.score.utils", os.path.join(REF, "subtools2/egrecho/score/utils.py"),
                           pre={"egrecho": eg, "egrecho.score": egs})
    sys.modules["egrecho.score.utils"] = egu
    bm = load_file_module("egrecho.score.binary_metrics",
                          os.path.join(REF, "subtools2/egrecho/score/binary_metrics.py"))
    rng = np.random.RandomState(302)
    tar = rng.standard_normal(2000) + 2.0
    non = rng.standard_normal(50000)
    scores = np.concatenate([tar, non])
    labels = np.concatenate([np.ones(2000, dtype=np.int64), np.zeros(50000, dtype=np.int64)])
    perm = rng.permutation(scores.shape[0])
    scores, labels = scores[perm], labels[perm]
    eer_b, thr_b = bos.compute_eer([[float(s), "target" if l else "nontarget"] for s, l in zip(scores, labels)])
    eer_d, dcf_d, thr_d = bm.compute_metrics(scores, labels, p_target=0.01)
    out.update(eer_scores_seed=np.int64(302), eer_bosaris=np.float64(eer_b), eer_bosaris_thr=np.float64(thr_b),
               eer_det=np.float64(eer_d), eer_det_thr=np.float64(thr_d), mindcf_det=np.float64(dcf_d))
    np.savez_compressed(os.path.join(HERE, "scoring.npz"), **out)
    print("scoring.npz eer", eer_b, eer_d, dcf_d)

    # ------------------------------------------------------------------ S-norm / AS-norm (score/ScoreNormalization.py)
    import argparse
    import tempfile
    sn = load_file_module("score_normalization", os.path.join(REF, "score/ScoreNormalization.py"))
    rng = np.random.RandomState(501)
    ne, nt, nc = 6, 9, 40
    ec = rng.standard_normal((ne, nc)).astype(np.float32)
    tc = rng.standard_normal((nt, nc)).astype(np.float32)
    trials = [(i, j, float(np.float32(rng.standard_normal()))) for i in range(ne) for j in range(nt) if (i + j) % 2 == 0]
    out = {"sn_enroll_cohort": ec, "sn_test_cohort": tc,
           "sn_trial_e": np.array([t[0] for t in trials], dtype=np.int32),
           "sn_trial_t": np.array([t[1] for t in trials], dtype=np.int32),
           "sn_scores": np.array([t[2] for t in trials], dtype=np.float32)}
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "in"), "w") as f:
            for i, j, v in trials:
                f.write("e{} t{} {}\n".format(i, j, repr(v)))
        with open(os.path.join(d, "ec"), "w") as f:
            for i in range(ne):
                for c in range(nc):
                    f.write("e{} c{} {}\n".format(i, c, repr(float(ec[i, c]))))
        with open(os.path.join(d, "tc"), "w") as f:
            for j in range(nt):
                for c in range(nc):
                    f.write("t{} c{} {}\n".format(j, c, repr(float(tc[j, c]))))
        for method, topn in (("snorm", 0), ("asnorm", 7)):
            ns = argparse.Namespace(method=method, top_n=topn, second_cohort="true", cross_select="false",
                                    input_score=os.path.join(d, "in"), enroll_cohort_score=os.path.join(d, "ec"),
                                    test_cohort_score=os.path.join(d, "tc"), output_score=os.path.join(d, "out_" + method))
            getattr(sn, method)(ns)
            vals = [float(l.split()[2]) for l in open(ns.output_score)]
            out["sn_" + method] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "score_norm.npz"), **out)
    print("score_norm.npz", out["sn_snorm"][:3], out["sn_asnorm"][:3])

    # ------------------------------------------------------------------ Kaldi ark bytes
    out = {}
    rng = np.random.RandomState(401)
    mat32 = rng.standard_normal((5, 7)).astype(np.float32)
    mat64 = rng.standard_normal((3, 4)).astype(np.float64)
    vec32 = rng.standard_normal(11).astype(np.float32)
    vec64 = rng.standard_normal(6).astype(np.float64)

    class Sink(io.BytesIO):  # kaldi_io asserts fd.mode == 'wb'
        mode = "wb"

    s = Sink()
    kaldi_io.write_mat(s, mat32, key="utt-a")
    kaldi_io.write_mat(s, mat64, key="utt_b")
    out["ark_mats_bytes"] = np.frombuffer(s.getvalue(), dtype=np.uint8)
    s = Sink()
    kaldi_io.write_vec_flt(s, vec32, key="spk1")
    kaldi_io.write_vec_flt(s, vec64, key="spk2")
    out["ark_vecs_bytes"] = np.frombuffer(s.getvalue(), dtype=np.uint8)
    out.update(mat32=mat32, mat64=mat64, vec32=vec32, vec64=vec64)
    # hand-built CM (compressed) matrix: global header, per-column uint16 percentiles, uint8 col-major data
    rows, cols = 6, 3
    gh = np.zeros(1, dtype=np.dtype([("minvalue", "<f4"), ("range", "<f4"), ("num_rows", "<i4"), ("num_cols", "<i4")]))
    gh["minvalue"], gh["range"], gh["num_rows"], gh["num_cols"] = -3.0, 7.5, rows, cols
    ph = np.sort(rng.randint(0, 65535, size=(cols, 4)).astype(np.uint16), axis=1)
    data = rng.randint(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :3] = (0, 64, 65)
    data[1, :3] = (192, 193, 255)
    cm = b"cmutt \0BCM " + gh.tobytes() + ph.astype("<u2").tobytes() + data.tobytes()

    class Src(io.BytesIO):
        mode = "rb"

    r = Src(cm)
    assert kaldi_io.read_key(r) == "cmutt"
    out["ark_cm_bytes"] = np.frombuffer(cm, dtype=np.uint8)
    out["ark_cm_decoded"] = kaldi_io.read_mat(r)
    np.savez_compressed(os.path.join(HERE, "kaldi_ark.npz"), **out)
    print("kaldi_ark.npz ok")


if __name__ == "__main__":
    main()
