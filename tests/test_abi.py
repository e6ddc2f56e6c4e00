"""The C-ABI library loads and exports every symbol include/xvb200.h declares; without a GPU the
compute entry points fail loudly instead of falling back.  CPU only (no compute calls)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "xvb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xvb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from asv_subtools_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), "libxvb200.so does not export " + n
        assert n in _lib.SIGNATURES, "ctypes binding missing for " + n
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert _lib.lib.xvb_version() == 100


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "xvb200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code and "at::" not in code and "std::" not in code and "#include <cuda" not in code


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_a_gpu():
    from asv_subtools_b200 import _lib
    assert _lib.lib.xvb_device_check() == -3  # XVB_ENODEVICE
    h = C.c_void_p()
    assert _lib.lib.xvb_extractor_create(C.byref(h), 23) == -3
    assert "no CUDA device" in _lib.last_error() or "fallback" in _lib.last_error()
    with pytest.raises(_lib.XvbError):
        _lib.check(_lib.lib.xvb_split_f32(None, 1, 8, 8, None, None, 8, None), "xvb_split_f32")


def test_model_blueprint_contract_on_cpu():
    """Plugin surface: constructor signature, creation string, state_dict keys ==
    the oracle's spec of the reference keys, strict=False tolerance for loss.* keys."""
    from asv_subtools_b200.model.xvector import Xvector
    from oracle import nnet as onn
    m = Xvector(23, 1211, training=False, extracted_embedding="near")
    assert m.get_model_creation().startswith("Xvector(23,1211,")
    spec = {k: tuple(s) for k, s, _ in onn.xvector_spec(23)}
    sd = m.state_dict()
    assert set(sd) == set(spec)
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k], k
    ck = onn.make_state_dict(onn.xvector_spec(23), 1)
    ck["loss.weight"] = torch.zeros(1211, 512, 1)       # training checkpoints carry loss.* keys
    m.load_state_dict(ck, strict=False)
    assert m.extracted_embedding == "near" and hasattr(m, "extract_embedding")
    with pytest.raises(RuntimeError):                     # no CPU path
        m.extract_embedding(onn.synthetic_feats(1, 10, 23, 0)[0])


def test_ecapa_blueprint_contract_on_cpu():
    """ECAPA plugin surface: canonical c1024 creation string -> the reference's keys and shapes."""
    from asv_subtools_b200.model.ecapa_tdnn_xvector import ECAPA_TDNN
    from oracle import nnet as onn
    canon = dict(training=False, extracted_embedding="near",
                 ecapa_params={"channels": 1024, "embd_dim": 192, "mfa_conv": 1536,
                               "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}},
                 pooling="ecpa-attentive", pooling_params={"hidden_size": 128, "time_attention": True, "stddev": True},
                 fc1=False, fc2_params={"nonlinearity": "", "nonlinearity_params": {"inplace": True}, "bn-relu": False,
                                        "bn": True, "bn_params": {"momentum": 0.5, "affine": False,
                                                                  "track_running_stats": True}})
    m = ECAPA_TDNN(80, 10, **canon)
    spec = {k: tuple(s) for k, s, _ in onn.ecapa_spec(80)}
    sd = m.state_dict()
    assert set(sd) == set(spec), set(sd) ^ set(spec)
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k], k
    m.load_state_dict(onn.make_state_dict(onn.ecapa_spec(80), 201), strict=True)
    assert not m.fc2.relu and m.fc2.batchnorm.weight is None
    d = ECAPA_TDNN(80, 10, training=False)                # blueprint defaults: fc2 = affine + ReLU + affine BN
    assert d.fc2.relu and set(d.state_dict()) == {k for k, _, _ in onn.ecapa_spec(80, fc2_bn_affine=True)}
    assert d.layer3.res2net_block.context == [-3, 0, 3] and d.layer4.res2net_block.blocks[0].affine.weight.shape == (128, 128, 9)
