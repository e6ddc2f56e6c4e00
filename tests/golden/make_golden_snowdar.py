#!/usr/bin/env python
"""Golden embeddings of the REFERENCE's snowdar x-vector blueprint (pytorch/model/snowdar_xvector.py) --
build container only:   python tests/golden/make_golden_snowdar.py
Seeded checkpoints come from oracle.nnet.make_state_dict(snowdar_xvector_spec), inputs from synthetic_feats;
only the reference's outputs are stored (tests/golden/snowdar.npz)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nnet as onn  # noqa: E402

CASES = {"std": dict(extend=False, seed=301), "ext": dict(extend=True, seed=302)}
# attention poolings of libs/nnet/pooling.py behind the blueprint's `pooling` switch (snowdar_xvector.py:119-136)
POOLING_CASES = {
    "attn1": ("attentive", {}, 311),                                                  # snowdar default: one affine
    "attn2": ("attentive", {"affine_layers": 2, "hidden_size": 64}, 312),
    "mha_share": ("multi-head", {"num_head": 4}, 313),                                # the paper's form: shared weights
    "mha_full": ("multi-head", {"num_head": 4, "share": False, "affine_layers": 2}, 314),
    "mres": ("multi-resolution", {"num_head": 4, "temperature": True, "affine_layers": 2}, 315),
    "lde": ("lde", {"num_head": 12, "num_nodes": 200}, 316),                        # LDEPooling(200, c_num=12): 2400-d encoding
    "xi_mean": ("xi-postmean-softplus2", {"hidden_size": 64, "num_nodes": 200}, 319),  # xi-vector, posterior mean
    "xi_dist": ("xi-postdist-softplus2", {"hidden_size": 64, "num_nodes": 200}, 320),  # ... mean | spread
}


def main():
    for name, attrs in (("tkinter", {"N": "n"}), ("tkinter.messagebox", {"NO": "no"}), ("turtle", {"xcor": None})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
    sys.path.insert(0, "/root/reference/pytorch")
    import libs.support.utils as utils
    out = {}
    for cname, c in CASES.items():
        sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, extend=c["extend"]), c["seed"])
        feats = onn.synthetic_feats(3, 120, 40, c["seed"] + 1000)
        for pos in ("far", "near_affine", "near"):
            model = utils.create_model_from_py("/root/reference/pytorch/model/snowdar_xvector.py",
                                               'Xvector(40,10,extend={},training=False,extracted_embedding="{}")'.format(c["extend"], pos))
            missing = model.load_state_dict(sd, strict=True)
            model.eval()
            emb = np.stack([model.extract_embedding(feats[i]).numpy() for i in range(3)])
            out["{}_{}".format(cname, pos)] = emb
    for cname, (pooling, pp, seed) in POOLING_CASES.items():
        sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, pooling=pooling, pooling_params=pp), seed)
        feats = onn.synthetic_feats(3, 120, 40, seed + 1000)
        for pos in ("far", "near"):
            model = utils.create_model_from_py(
                "/root/reference/pytorch/model/snowdar_xvector.py",
                'Xvector(40,10,training=False,extracted_embedding="{}",pooling="{}",pooling_params={!r})'.format(pos, pooling, pp))
            model.load_state_dict(sd, strict=True)
            model.eval()
            out["{}_{}".format(cname, pos)] = np.stack([model.extract_embedding(feats[i]).numpy() for i in range(3)])
    # the other BatchNorm order: tdnn_layer_params={"bn-relu": True} (components.py:386-403), BatchNorm with affine parameters
    tlp = {"bn-relu": True, "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}}
    sd = onn.make_state_dict(onn.snowdar_xvector_spec(40, bn_affine=True), 317)
    feats = onn.synthetic_feats(3, 120, 40, 1317)
    for pos in ("far", "near_affine", "near"):
        model = utils.create_model_from_py(
            "/root/reference/pytorch/model/snowdar_xvector.py",
            'Xvector(40,10,training=False,extracted_embedding="{}",tdnn_layer_params={!r})'.format(pos, tlp))
        model.load_state_dict(sd, strict=True)
        model.eval()
        out["bnrelu_{}".format(pos)] = np.stack([model.extract_embedding(feats[i]).numpy() for i in range(3)])
    np.savez_compressed(os.path.join(HERE, "snowdar.npz"), **out)
    print("snowdar.npz ok", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
