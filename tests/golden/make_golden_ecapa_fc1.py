#!/usr/bin/env python
"""Golden embeddings of the REFERENCE's ECAPA_TDNN with fc1=True (pytorch/model/ecapa_tdnn_xvector.py:286-287, :326-333;
positions far / near_affine / near of extract_embedding :412-422) -- build container only:
    python tests/golden/make_golden_ecapa_fc1.py   ->  tests/golden/ecapa_fc1.npz
Seeded checkpoint from oracle.nnet.make_state_dict(ecapa_spec(80, fc1=True, fc2_bn_affine=True)); only outputs are stored."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nnet as onn  # noqa: E402


def main():
    for name, attrs in (("tkinter", {"N": "n"}), ("tkinter.messagebox", {"NO": "no"}), ("turtle", {"xcor": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
    sys.path.insert(0, "/root/reference/pytorch")
    import libs.support.utils as utils
    sd = onn.make_state_dict(onn.ecapa_spec(80, fc1=True, fc2_bn_affine=True), 203)
    feats = onn.synthetic_feats(2, 120, 80, 1203)
    out = {}
    for pos in ("far", "near_affine", "near"):
        model = utils.create_model_from_py("/root/reference/pytorch/model/ecapa_tdnn_xvector.py",
                                           'ECAPA_TDNN(80,10,training=False,fc1=True,extracted_embedding="{}")'.format(pos))
        model.load_state_dict(sd, strict=True)
        model.eval()
        out["fc1_" + pos] = np.stack([model.extract_embedding(feats[i]).numpy() for i in range(2)])
    np.savez_compressed(os.path.join(HERE, "ecapa_fc1.npz"), **out)
    print("ecapa_fc1.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
