"""Feature-side front-end on the GPU for a ragged batch of utterances: energy VAD,
CMN (per-utterance or Kaldi-style sliding window) and voiced-frame selection -- the steps between
`feats.scp` and the extractor in pytorch/pipeline/extract_xvectors_for_pytorch.sh:105-118 and in the
reference's C++ runtime (runtime/extractor/torch_asv_extractor.cc:71-108)."""
import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack(utts, device="cuda"):
    """list of (T_i, F) float32 arrays -> ((sum_T, F) CUDA tensor, (U+1) int32 CUDA offsets)."""
    lens = np.array([u.shape[0] for u in utts], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    x = torch.from_numpy(np.ascontiguousarray(np.concatenate(utts, axis=0), dtype=np.float32)).to(device)
    return x, torch.from_numpy(off).to(device)


def unpack(x, offsets):
    off = offsets.cpu().numpy()
    return [x[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def vad_energy(x, offsets, energy_threshold=5.0, energy_mean_scale=0.5, frames_context=0, proportion_threshold=0.6):
    """-> (voiced (sum_T,) uint8, counts (U,) int32).  Defaults are Kaldi's VadEnergyOptions."""
    u = offsets.shape[0] - 1
    voiced = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
    counts = torch.empty(u, dtype=torch.int32, device=x.device)
    check(lib.xvb_vad_energy(x.data_ptr(), offsets.data_ptr(), u, x.shape[1], energy_threshold, energy_mean_scale,
                             frames_context, proportion_threshold, voiced.data_ptr(), counts.data_ptr(), _s()),
          "xvb_vad_energy")
    return voiced, counts


def cmn(x, offsets, window=0):
    """window = 0: per-utterance mean subtraction; window = 300: apply-cmvn-sliding --center --cmn-window=300."""
    y = torch.empty_like(x)
    check(lib.xvb_cmn(x.data_ptr(), offsets.data_ptr(), offsets.shape[0] - 1, x.shape[1], int(window), y.data_ptr(), _s()),
          "xvb_cmn")
    return y


def select_frames(x, offsets, voiced, counts):
    """-> (packed voiced frames, new offsets)."""
    new_off = torch.zeros(offsets.shape[0], dtype=torch.int32, device=x.device)
    new_off[1:] = torch.cumsum(counts, 0)
    total = int(new_off[-1].item())
    y = torch.empty(total, x.shape[1], dtype=torch.float32, device=x.device)
    if total:
        check(lib.xvb_select_frames(x.data_ptr(), offsets.data_ptr(), voiced.data_ptr(), new_off.data_ptr(),
                                    offsets.shape[0] - 1, x.shape[1], y.data_ptr(), _s()), "xvb_select_frames")
    return y, new_off


# ------------------------------------------------------------------ waveform -> fbank / MFCC
_WINDOWS = {"povey": 0, "hamming": 1, "hanning": 2, "rectangular": 3, "blackman": 4}
_UNSUPPORTED = {"dither": 0.0, "snip_edges": True, "round_to_power_of_two": True, "vtln_warp": 1.0, "subtract_mean": False,
                "min_duration": 0.0, "channel": -1}


class KaldiFeature:
    """Mirror of the reference's `KaldiFeature` (pytorch/libs/egs/kaldi_features.py:69-135): same constructor
    (`feature_type` in {'fbank','mfcc'}, `kaldi_featset` with torchaudio.compliance.kaldi keyword names,
    `mean_var_conf`), same call (`waveforms [batch, time]` (+ relative `lengths`) or a list of 1-D
    waveforms) -> list of (frames, dim) feature tensors -- computed by xvb_fbank_compute on the GPU for the
    whole ragged batch in one launch, mean normalisation by xvb_cmn.  Options torchaudio applies randomly
    or that the extraction recipes never set (dither != 0, snip_edges=False, VTLN) raise."""

    def __init__(self, feature_type="mfcc", kaldi_featset={}, mean_var_conf={}):
        from ._lib import FbankOpts
        assert feature_type in ("mfcc", "fbank")
        self.feat_type = feature_type
        self.kaldi_featset = dict(kaldi_featset)
        o = FbankOpts()
        lib.xvb_fbank_default_opts(C.byref(o))
        if feature_type == "mfcc":
            o.num_ceps = 13
        for k, v in self.kaldi_featset.items():
            if k in _UNSUPPORTED:
                if v != _UNSUPPORTED[k]:
                    raise NotImplementedError("KaldiFeature: {}={!r} is not supported on the GPU path".format(k, v))
            elif k == "window_type":
                o.window_type = _WINDOWS[v]
            elif k in ("frame_length", "frame_shift"):
                setattr(o, k + "_ms", float(v))
            elif k in ("vtln_high", "vtln_low"):
                pass
            elif hasattr(o, k):
                setattr(o, k, type(getattr(o, k))(v))
            else:
                raise TypeError("KaldiFeature: unknown option {}".format(k))
        if feature_type == "fbank":
            o.num_ceps = 0
        self.mean_norm = bool(mean_var_conf.get("mean_norm", True)) if mean_var_conf else False
        if mean_var_conf and mean_var_conf.get("std_norm", False):
            raise NotImplementedError("KaldiFeature: std_norm is not supported on the GPU path (the recipes use mean_norm only)")
        self._opts = o
        self._h = C.c_void_p()
        check(lib.xvb_fbank_create(C.byref(self._h), C.byref(o)), "xvb_fbank_create")
        self.dim = lib.xvb_fbank_dim(self._h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        try:
            if h and lib is not None:
                lib.xvb_fbank_destroy(h)
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def num_frames(self, num_samples):
        return int(lib.xvb_fbank_num_frames(self._h, int(num_samples)))

    def compute(self, waves, device="cuda"):
        """list of 1-D float32 waveforms -> ((sum_frames, dim) CUDA tensor, (U+1) int32 CUDA frame offsets)."""
        waves = [w.detach().cpu().numpy() if isinstance(w, torch.Tensor) else np.asarray(w) for w in waves]
        lens = np.array([w.shape[0] for w in waves], dtype=np.int64)
        soff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        frames = np.array([self.num_frames(n) for n in lens], dtype=np.int64)
        foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)
        x = torch.from_numpy(np.ascontiguousarray(np.concatenate(waves), dtype=np.float32)).to(device)
        so, fo = torch.from_numpy(soff).to(device), torch.from_numpy(foff).to(device)
        total = int(foff[-1])
        feats = torch.empty(total, self.dim, dtype=torch.float32, device=device)
        check(lib.xvb_fbank_compute(self._h, x.data_ptr(), so.data_ptr(), fo.data_ptr(), len(waves), total, feats.data_ptr(),
                                    _s()), "xvb_fbank_compute")
        if self.mean_norm and total:
            feats = cmn(feats, fo, 0)
        return feats, fo

    def __call__(self, waveforms, lengths=None):
        if isinstance(waveforms, torch.Tensor) and waveforms.dim() >= 2:
            if torch.any(torch.isnan(waveforms)):
                raise ValueError("feats:{}".format(waveforms))
            ws = []
            for i, w in enumerate(waveforms):
                w = w if w.dim() == 1 else w.transpose(0, 1)[0]     # [time, channel] -> channel 0
                if lengths is not None:
                    w = w[: int((lengths[i] * waveforms.shape[1]).long())]
                ws.append(w)
        else:
            ws = list(waveforms)
        feats, fo = self.compute(ws)
        return unpack(feats, fo)
