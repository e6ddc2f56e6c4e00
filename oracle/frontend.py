"""Oracle (NumPy) restatement of the feature-side front-end.  TEST INFRASTRUCTURE ONLY.

  vad_energy      : TorchAsvExtractor::ComputeVadEnergy, runtime/extractor/torch_asv_extractor.cc:14-62
  cmn_utterance   : `input_feats - input_feats.mean(0)`, torch_asv_extractor.cc:99-101
  select_voiced   : index_select(0, nonzero(vad)), torch_asv_extractor.cc:103-107
  cmn_sliding     : Kaldi apply-cmvn-sliding --norm-vars=false --center=true (call site
                    pytorch/pipeline/extract_xvectors_for_pytorch.sh:105-111).  PARITY UNPINNED: Kaldi is not
                    vendored; the window rule is restated from Kaldi's SlidingWindowCmn.
The C++ runtime cannot be built here (libtorch/gflags are fetched from the network), so these are
pinned by code reading only; the functions are a handful of lines each."""
import numpy as np


def vad_energy(feats, energy_threshold=5.0, energy_mean_scale=0.5, frames_context=0, proportion_threshold=0.6):
    T = feats.shape[0]
    log_energy = feats[:, 0].astype(np.float32)
    thr = np.float32(energy_threshold)
    if energy_mean_scale != 0.0:
        thr = np.float32(thr + np.float32(energy_mean_scale) * np.float32(log_energy.sum(dtype=np.float32)) / np.float32(T))
    out = np.zeros(T, dtype=np.uint8)
    for t in range(T):
        num = den = 0
        for t2 in range(t - frames_context, t + frames_context + 1):
            if 0 <= t2 < T:
                den += 1
                if log_energy[t2] > thr:
                    num += 1
        out[t] = 1 if num >= den * proportion_threshold else 0
    return out


def cmn_utterance(feats):
    return feats - feats.mean(axis=0, dtype=np.float64).astype(np.float32)


def cmn_sliding(feats, window=300):
    T = feats.shape[0]
    out = np.empty_like(feats)
    csum = np.concatenate([np.zeros((1, feats.shape[1])), np.cumsum(feats.astype(np.float64), axis=0)])
    for t in range(T):
        b = t - window // 2
        e = b + window
        if b < 0:
            e -= b
            b = 0
        if e > T:
            b -= e - T
            e = T
            if b < 0:
                b = 0
        out[t] = feats[t] - ((csum[e] - csum[b]) / (e - b)).astype(np.float32)
    return out


def select_voiced(feats, voiced):
    return feats[np.flatnonzero(voiced)]
