#!/usr/bin/env python
"""Golden fbank / MFCC fixtures from the REFERENCE's own feature class.

    python tests/golden/make_golden_fbank.py        (build container only: needs /root/reference)

Imports pytorch/libs/egs/kaldi_features.py `KaldiFeature` (which calls torchaudio.compliance.kaldi, the
reference's feature dependency) and runs it on the seeded `oracle.frontend.synthetic_wave` signals with
the configurations the reference ships (runtime/test/feat_conf.yaml) plus option coverage.  Stores only
the outputs: tests/golden/fbank.npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import frontend as ofe  # noqa: E402

CONFIGS = {
    # runtime/test/feat_conf.yaml (the reference's deployed ECAPA front-end)
    "fbank80": ("fbank", dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200, low_freq=40,
                              num_mel_bins=80, use_energy=False), dict(mean_norm=True, std_norm=False)),
    # 23-dim MFCC of the x-vector recipes (BASELINE configs[0] feature type), 8 kHz telephone band
    "mfcc23": ("mfcc", dict(dither=0.0, sample_frequency=8000.0, frame_length=25, frame_shift=10, low_freq=20, high_freq=3700,
                            num_mel_bins=23, num_ceps=23, use_energy=True, energy_floor=0.0), {}),
    "fbank40_energy": ("fbank", dict(dither=0.0, num_mel_bins=40, use_energy=True, raw_energy=False, energy_floor=1.0,
                                     window_type="hamming", preemphasis_coefficient=0.9, htk_compat=True), {}),
    "fbank24_lin": ("fbank", dict(dither=0.0, num_mel_bins=24, use_log_fbank=False, use_power=False, remove_dc_offset=False,
                                  window_type="hanning", frame_length=20, frame_shift=5), dict(mean_norm=True, std_norm=True)),
    "mfcc13_htk": ("mfcc", dict(dither=0.0, num_mel_bins=30, num_ceps=13, htk_compat=True, cepstral_lifter=0.0,
                                window_type="rectangular", preemphasis_coefficient=0.0), {}),
}
WAVES = {"a": (16000, 11), "b": (5243, 12), "c": (400, 13), "d": (48000, 14)}


def main():
    for name, attrs in (("tkinter", {"N": "n"}), ("tkinter.messagebox", {"NO": "no"}), ("turtle", {"xcor": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
    sys.path.insert(0, "/root/reference/pytorch")
    import libs.egs.kaldi_features as kf
    out = {}
    for cname, (ftype, featset, mv) in CONFIGS.items():
        f = kf.KaldiFeature(ftype, featset, mv)
        for wname, (n, seed) in WAVES.items():
            sf = featset.get("sample_frequency", 16000.0)
            wave = torch.from_numpy(ofe.synthetic_wave(n, seed, sample_frequency=sf))
            if n < int(sf * featset.get("frame_length", 25) * 0.001):
                continue
            feat = f(wave.unsqueeze(0))[0]
            out["{}_{}".format(cname, wname)] = feat.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "fbank.npz"), **out)
    print("fbank.npz ok:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
