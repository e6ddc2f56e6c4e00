#!/usr/bin/env python
"""GPU fbank throughput on a ragged batch (256 utterances of 8-12 s, 16 kHz, 80 mel bins) -- a side measurement,
also the target of `ncu -k regex:fbank`."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200.frontend import KaldiFeature  # noqa: E402
from asv_subtools_b200 import frontend as fe  # noqa: E402
from asv_subtools_b200._lib import check, lib  # noqa: E402

featset = dict(dither=0.0, energy_floor=0.0, frame_length=25, frame_shift=10, high_freq=-200, low_freq=40, num_mel_bins=80)
kf = KaldiFeature("fbank", featset, {})
rng = np.random.RandomState(0)
lens = rng.randint(8 * 16000, 12 * 16000, 256)
soff = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
frames = np.array([kf.num_frames(n) for n in lens])
foff = np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)
wave = torch.randn(int(soff[-1]), device="cuda") * 3000.0
so, fo = torch.from_numpy(soff).cuda(), torch.from_numpy(foff).cuda()
total = int(foff[-1])
out = torch.empty(total, kf.dim, device="cuda")


def run():
    check(lib.xvb_fbank_compute(kf._h, wave.data_ptr(), so.data_ptr(), fo.data_ptr(), 256, total, out.data_ptr(), fe._s()))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"workload": "fbank80, 256 utterances of 8-12 s at 16 kHz", "frames": total, "ms": ms, "frames_per_s": total / ms * 1e3,
                  "bytes_per_frame": 160 * 4 + 80 * 4, "gbs": total * (160 * 4 + 80 * 4) / ms * 1e-6,
                  "finite": bool(torch.isfinite(out).all())}))
