"""Python-free deployment path: .xvbm model files, xvb_extractor_load, and the
`bin/xvb-extract` binary (native ark reader -> batched extraction -> FV ark/scp writer)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from asv_subtools_b200 import kaldi_io, ops
from oracle import frontend as ofe
from oracle import nnet as onn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "asv_subtools_b200", "bin", "xvb-extract")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(np.max(np.abs(b)), 1e-30))


def test_binary_is_built_and_fails_loudly_without_arguments():
    assert os.path.exists(BIN), "make in asv_subtools_b200/csrc builds bin/xvb-extract"
    out = subprocess.run([BIN, "--help"], capture_output=True, text=True)
    assert out.returncode == 0 and "feats-rspecifier" in out.stdout
    out = subprocess.run([BIN, "only-one"], capture_output=True, text=True)
    assert out.returncode == 1 and "ERROR" in out.stderr


def _model(dim, seed, pos):
    from asv_subtools_b200.model.xvector import Xvector
    sd = onn.make_state_dict(onn.xvector_spec(dim), seed)
    m = Xvector(dim, 10, training=False, extracted_embedding=pos)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.gpu
@pytest.mark.parametrize("pos", ["far", "near"])
def test_model_file_roundtrip_is_bit_exact(tmp_path, pos):
    m, _ = _model(23, 101, pos)
    path = str(tmp_path / "xv.xvbm")
    m.extractor().save(path)
    ex = ops.Extractor.load(path)
    assert ex.feat_dim == 23 and ex.embed_dim == 512
    x = torch.from_numpy(onn.synthetic_feats(6, 120, 23, 5)).cuda()
    assert torch.equal(ex.extract(x), m.extractor().extract(x))
    with open(path, "r+b") as f:
        f.truncate(os.path.getsize(path) - 100)
    with pytest.raises(RuntimeError):
        ops.Extractor.load(path)
    with pytest.raises(RuntimeError):
        ops.Extractor.load(str(tmp_path / "missing.xvbm"))


@pytest.mark.gpu
def test_xvb_extract_binary_matches_oracle_and_plugin(tmp_path):
    m, sd = _model(80, 102, "far")
    model = str(tmp_path / "xv80.xvbm")
    m.extractor().save(model)
    lengths = [200, 200, 57, 200, 450, 57, 1, 333]           # 450 and 333 exceed --max-chunk 200: 3 and 2 chunks
    feats = {"utt{:02d}".format(i): onn.synthetic_feats(1, t, 80, 900 + i)[0] for i, t in enumerate(lengths)}
    ark = str(tmp_path / "feats.ark")
    with open(ark, "wb") as f:
        for k, v in feats.items():
            kaldi_io.write_mat(f, v, key=k)
    out_ark, out_scp = str(tmp_path / "xv.ark"), str(tmp_path / "xv.scp")
    run = subprocess.run([BIN, "--batch", "3", "--max-chunk", "200", model, "ark:" + ark,
                          "ark,scp:{},{}".format(out_ark, out_scp)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert run.stdout.count("Process utterance for key") == len(lengths)
    got = dict(kaldi_io.read_vec_flt_ark(out_ark))
    assert sorted(got) == sorted(feats) and sorted(dict(kaldi_io.read_vectors("scp:" + out_scp))) == sorted(feats)
    fwd = lambda x: onn.xvector_forward(sd, x, "far")
    for k, v in feats.items():
        want = onn.extract_embedding(fwd, v, max_chunk=200).numpy()
        assert got[k].shape == (512,) and rel(got[k], want) < 1e-4, k      # north-star tolerance vs the oracle
        if v.shape[0] <= 200:                                               # same kernels as the plugin call
            assert rel(got[k], m.extract_embedding(v).numpy()) < 1e-5, k
    # per-utterance CMN inside the binary == oracle CMN upstream of the model (torch_asv_extractor.cc:99-101)
    run = subprocess.run([BIN, "--cmn", "utt", model, ark, "ark:" + out_ark], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    got = dict(kaldi_io.read_vec_flt_ark(out_ark))
    for k in ("utt00", "utt02", "utt07"):
        want = onn.extract_embedding(fwd, ofe.cmn_utterance(feats[k])).numpy()
        assert rel(got[k], want) < 1e-4, k
    # wrong feature dimension: message with ERROR, exit 1 (extract_xvectors_for_pytorch.sh:144-145 greps for it)
    bad = str(tmp_path / "bad.ark")
    with open(bad, "wb") as f:
        kaldi_io.write_mat(f, np.zeros((10, 23), np.float32), key="b")
    run = subprocess.run([BIN, model, bad, "ark:" + out_ark], capture_output=True, text=True, timeout=120)
    assert run.returncode == 1 and "ERROR" in run.stderr


@pytest.mark.gpu
def test_xvb_extract_wav_scp_mode_matches_oracle_chain(tmp_path):
    """wav.scp -> GPU fbank (runtime/test/feat_conf.yaml options) -> mean norm -> x-vector, no Python at run
    time; against the all-CPU oracle chain (kaldi_fbank -> sequence_normalize -> xvector_forward)."""
    import wave as wavmod
    m, sd = _model(80, 102, "far")
    model = str(tmp_path / "xv80.xvbm")
    m.extractor().save(model)
    scp = tmp_path / "wav.scp"
    waves = {}
    with open(scp, "w") as f:
        for i, n in enumerate([16000, 24000, 16000, 5000]):
            pcm = np.clip(np.round(ofe.synthetic_wave(n, 700 + i)), -32768, 32767).astype(np.int16)
            path = tmp_path / "u{}.wav".format(i)
            with wavmod.open(str(path), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(16000)
                w.writeframes(pcm.tobytes())
            waves["u{}".format(i)] = pcm.astype(np.float32)
            f.write("u{} {}\n".format(i, path))
    out_ark = str(tmp_path / "xv.ark")
    run = subprocess.run([BIN, "--wav", "fbank", "--num-mel-bins", "80", "--low-freq", "40", "--high-freq", "-200",
                          "--energy-floor", "0", "--cmn", "utt", model, str(scp), "ark:" + out_ark],
                         capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    got = dict(kaldi_io.read_vec_flt_ark(out_ark))
    featset = dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200, low_freq=40, num_mel_bins=80)
    fwd = lambda x: onn.xvector_forward(sd, x, "far")
    for k, wv in waves.items():
        feats = ofe.sequence_normalize(ofe.kaldi_fbank(wv, **featset)).astype(np.float32)
        want = onn.extract_embedding(fwd, feats).numpy()
        assert rel(got[k], want) < 1e-4, k
    # a model/feature mismatch is an ERROR, not a silent resize
    run = subprocess.run([BIN, "--wav", "fbank", "--num-mel-bins", "40", model, str(scp), "ark:" + out_ark],
                         capture_output=True, text=True, timeout=120)
    assert run.returncode == 1 and "ERROR" in run.stderr


@pytest.mark.gpu
def test_online_cli_wav_scp_matches_oracle_chain(tmp_path):
    """python -m asv_subtools_b200.pipeline.extract_embeddings_online: nnet.config + feat_conf.yaml + wav.scp."""
    import sys
    import wave as wavmod
    import yaml
    _, sd = _model(80, 102, "far")
    torch.save(sd, str(tmp_path / "final.params"))
    blueprint = os.path.join(ROOT, "asv_subtools_b200", "model", "xvector.py")
    (tmp_path / "nnet.config").write_text('model_blueprint;{}\nmodel_creation;Xvector(80,10,training=False,extracted_embedding="far")\n'.format(blueprint))
    conf = dict(feature_type="fbank", kaldi_featset=dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200,
                                                         low_freq=40, num_mel_bins=80, use_energy=False),
                mean_var_conf=dict(mean_norm=True, std_norm=False))
    (tmp_path / "feat_conf.yaml").write_text(yaml.safe_dump(conf))
    waves = {}
    with open(tmp_path / "wav.scp", "w") as f:
        for i, n in enumerate([16000, 16000, 9000]):
            pcm = np.clip(np.round(ofe.synthetic_wave(n, 800 + i)), -32768, 32767).astype(np.int16)
            path = tmp_path / "o{}.wav".format(i)
            with wavmod.open(str(path), "wb") as w:
                w.setnchannels(1)
                w.setsampwidth(2)
                w.setframerate(16000)
                w.writeframes(pcm.tobytes())
            waves["o{}".format(i)] = pcm.astype(np.float32)
            f.write("o{} {}\n".format(i, path))
    out = str(tmp_path / "xv.ark")
    run = subprocess.run([sys.executable, "-m", "asv_subtools_b200.pipeline.extract_embeddings_online", "--nnet-config",
                          str(tmp_path / "nnet.config"), "--feat-config", str(tmp_path / "feat_conf.yaml"),
                          str(tmp_path / "final.params"), str(tmp_path / "wav.scp"), "ark:" + out],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert run.returncode == 0 and "RTF:" in run.stdout, run.stdout + run.stderr
    got = dict(kaldi_io.read_vec_flt_ark(out))
    fwd = lambda x: onn.xvector_forward(sd, x, "far")
    for k, wv in waves.items():
        feats = ofe.sequence_normalize(ofe.kaldi_fbank(wv, **conf["kaldi_featset"])).astype(np.float32)
        assert rel(got[k], onn.extract_embedding(fwd, feats).numpy()) < 1e-4, k


@pytest.mark.gpu
def test_xvb_extract_binary_runs_an_ecapa_model_file(tmp_path):
    """XVBE0001 model file -> bin/xvb-extract: ECAPA-TDNN without Python, against the oracle forward."""
    from asv_subtools_b200.model.ecapa_tdnn_xvector import ECAPA_TDNN
    canon = dict(training=False, extracted_embedding="near",
                 ecapa_params={"channels": 1024, "embd_dim": 192, "mfa_conv": 1536,
                               "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}},
                 fc2_params={"nonlinearity": "", "bn": True, "bn_params": {"momentum": 0.5, "affine": False, "track_running_stats": True}})
    sd = onn.make_state_dict(onn.ecapa_spec(80), 201)
    m = ECAPA_TDNN(80, 10, **canon)
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    model = str(tmp_path / "ecapa.xvbm")
    m.extractor().save(model)
    feats = {"e{}".format(i): onn.synthetic_feats(1, t, 80, 300 + i)[0] for i, t in enumerate([120, 120, 75])}
    ark = str(tmp_path / "feats.ark")
    with open(ark, "wb") as f:
        for k, v in feats.items():
            kaldi_io.write_mat(f, v, key=k)
    out = str(tmp_path / "xv.ark")
    run = subprocess.run([BIN, "--batch", "4", model, ark, "ark:" + out], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    got = dict(kaldi_io.read_vec_flt_ark(out))
    for k, v in feats.items():
        want = onn.extract_embedding(lambda x: onn.ecapa_forward(sd, x, "near"), v).numpy()
        assert got[k].shape == (192,) and rel(got[k], want) < 1e-4, k
