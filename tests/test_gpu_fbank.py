"""xvb_fbank_compute (GPU fbank / MFCC) against the golden outputs of the reference's own KaldiFeature
(tests/golden/fbank.npz) and the float64 oracle, through the KaldiFeature mirror."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import frontend as ofe

HERE = os.path.dirname(os.path.abspath(__file__))


def _mgf():
    spec = importlib.util.spec_from_file_location("mgf", os.path.join(HERE, "golden", "make_golden_fbank.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.gpu
def test_gpu_fbank_mfcc_match_reference_kaldifeature(golden):
    from asv_subtools_b200.frontend import KaldiFeature
    mgf, g = _mgf(), golden("fbank")
    seen = 0
    for cname, (ftype, featset, mv) in mgf.CONFIGS.items():
        if mv.get("std_norm"):
            with pytest.raises(NotImplementedError):
                KaldiFeature(ftype, featset, mv)
            mv = {}
        kf = KaldiFeature(ftype, featset, mv)
        sf = featset.get("sample_frequency", 16000.0)
        names = [w for w in mgf.WAVES if "{}_{}".format(cname, w) in g.files]
        waves = [ofe.synthetic_wave(mgf.WAVES[w][0], mgf.WAVES[w][1], sample_frequency=sf) for w in names]
        outs = kf(waves)                                   # one ragged launch for all four utterances
        for w, wave, got in zip(names, waves, outs):
            got = got.cpu().numpy()
            want = (ofe.kaldi_mfcc if ftype == "mfcc" else ofe.kaldi_fbank)(wave, **featset)
            if mv:
                want = ofe.sequence_normalize(want, **mv)
            tol = 2e-4 if featset.get("use_log_fbank", True) else 2e-5 * max(1.0, np.abs(want).max())
            if ftype == "mfcc":
                tol *= 1.0 + 0.5 * featset.get("cepstral_lifter", 22.0)
            assert got.shape == want.shape and got.shape[1] == kf.dim, (cname, w)
            assert np.max(np.abs(got - want)) < tol, (cname, w, np.max(np.abs(got - want)))       # float64 oracle
            if "std_norm" not in mgf.CONFIGS[cname][2] or not mgf.CONFIGS[cname][2]["std_norm"]:
                ref = g["{}_{}".format(cname, w)]
                assert np.max(np.abs(got - ref)) < 2 * tol, (cname, w, np.max(np.abs(got - ref)))  # reference output
            seen += 1
    assert seen >= 16


@pytest.mark.gpu
def test_gpu_fbank_batched_tensor_call_lengths_and_edges():
    from asv_subtools_b200.frontend import KaldiFeature
    featset = dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200, low_freq=40, num_mel_bins=80)
    kf = KaldiFeature("fbank", featset, dict(mean_norm=True, std_norm=False))
    wav = torch.from_numpy(np.stack([ofe.synthetic_wave(8000, 3), ofe.synthetic_wave(8000, 4)]))
    rel = torch.tensor([1.0, 0.5])
    outs = kf(wav, rel)
    assert outs[0].shape == (48, 80) and outs[1].shape == (23, 80)
    want = ofe.sequence_normalize(ofe.kaldi_fbank(wav[1, :4000].numpy(), **featset))
    assert np.max(np.abs(outs[1].cpu().numpy() - want)) < 2e-4
    assert kf.num_frames(399) == 0 and kf.num_frames(400) == 1 and kf.num_frames(560) == 2
    feats, fo = kf.compute([ofe.synthetic_wave(399, 5), ofe.synthetic_wave(1000, 6)])     # a too-short utterance yields 0 frames
    assert fo.cpu().tolist() == [0, 0, 4] and feats.shape == (4, 80)
    with pytest.raises(NotImplementedError):
        KaldiFeature("fbank", dict(dither=1.0))
    with pytest.raises(RuntimeError):
        KaldiFeature("fbank", dict(num_mel_bins=2))
    # silence: log of the floor, finite
    z = KaldiFeature("fbank", dict(num_mel_bins=40, dither=0.0)).compute([np.zeros(800, np.float32)])[0]
    assert torch.isfinite(z).all() and abs(float(z[0, 0]) - np.log(np.finfo(np.float32).eps)) < 1e-5


@pytest.mark.gpu
def test_wave_to_embedding_online_path_matches_oracle():
    """The reference's online extraction (extract_embeddings_online.py): wav -> KaldiFeature(fbank80, mean_norm)
    -> model.extract_embedding, here entirely on the GPU, against the all-CPU oracle chain."""
    from asv_subtools_b200.frontend import KaldiFeature
    from asv_subtools_b200.model.xvector import Xvector
    from oracle import nnet as onn
    featset = dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200, low_freq=40, num_mel_bins=80)
    sd = onn.make_state_dict(onn.xvector_spec(80), 102)
    m = Xvector(80, 10, training=False, extracted_embedding="far")
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    wave = ofe.synthetic_wave(32000, 21, scale=0.1)          # 2 s; torchaudio-style [-1,1]-range amplitude
    feats = KaldiFeature("fbank", featset, dict(mean_norm=True, std_norm=False))([wave])[0]
    emb = m.extract_embedding(feats.cpu().numpy()).numpy()
    f64 = ofe.sequence_normalize(ofe.kaldi_fbank(wave, **featset)).astype(np.float32)
    want = onn.extract_embedding(lambda x: onn.xvector_forward(sd, x, "far"), f64).numpy()
    assert np.max(np.abs(emb - want)) / np.max(np.abs(want)) < 1e-4
