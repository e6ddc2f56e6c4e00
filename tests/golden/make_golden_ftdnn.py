#!/usr/bin/env python
"""Golden embeddings of the REFERENCE's factored x-vector blueprint (pytorch/model/factored_xvector.py) -- build
container only:   python tests/golden/make_golden_ftdnn.py      -> tests/golden/ftdnn.npz"""
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
    sd = onn.make_state_dict(onn.factored_xvector_spec(40), 401)
    feats = onn.synthetic_feats(2, 90, 40, 1401)
    out = {}
    for pos in ("far", "near"):
        model = utils.create_model_from_py("/root/reference/pytorch/model/factored_xvector.py",
                                           'Xvector(40,10,training=False,extracted_embedding="{}")'.format(pos))
        model.load_state_dict(sd, strict=True)
        model.eval()
        out[pos] = np.stack([model.extract_embedding(feats[i]).numpy() for i in range(2)])
    np.savez_compressed(os.path.join(HERE, "ftdnn.npz"), **out)
    print("ftdnn.npz ok", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
