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
