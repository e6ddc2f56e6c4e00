#!/usr/bin/env python
"""Per-kernel CUDA-event times of one x-vector step (median of 10) -- tuning aid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asv_subtools_b200.model.xvector import Xvector  # noqa: E402
from oracle import nnet as onn  # noqa: E402

m = Xvector(80, 10, training=False)
m.load_state_dict(onn.make_state_dict(onn.xvector_spec(80), 102), strict=True)
m.cuda().eval()
ex = m.extractor()
xs = [torch.randn(256, 200, 80, device="cuda") for _ in range(4)]
for i in range(3):
    ex.extract(xs[i % 4])
ex.set_profiling(True)
per = []
for i in range(10):
    ex.extract(xs[i % 4])
    per.append(ex.kernel_times_ms())
per = np.median(np.array(per), axis=0) * 1e3
names = ["split", "tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5", "pool", "tdnn6"]
print(" ".join("{}={:.0f}us".format(n, v) for n, v in zip(names, per)), "total={:.0f}us".format(per.sum()))
