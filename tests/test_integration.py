"""The drop-in boundary exercised from the REFERENCE's side: its own blueprint loader building the B200
blueprints and loading reference-made state_dicts, and its shell function files with integration/score_b200.sh sourced
on top.  Needs /root/reference (build container only): skipped on the GPU box."""
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine")

ECAPA_ARGS = ('80,10,training=False,extracted_embedding="near",'
              'ecapa_params={"channels":1024,"embd_dim":192,"mfa_conv":1536,'
              '"bn_params":{"momentum":0.5,"affine":True,"track_running_stats":True}},'
              'pooling="ecpa-attentive",pooling_params={"hidden_size":128,"time_attention":True,"stddev":True},'
              'fc1=False,fc2_params={"nonlinearity":"","nonlinearity_params":{"inplace":True},"bn-relu":False,'
              '"bn":True,"bn_params":{"momentum":0.5,"affine":False,"track_running_stats":True}}')


@pytest.fixture(scope="module")
def ref_utils():
    for name, attrs in (("tkinter", {"N": "n"}), ("tkinter.messagebox", {"NO": "no"}), ("turtle", {"xcor": None})):
        m = types.ModuleType(name)           # libs/nnet/transformer imports these by accident
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules.setdefault(name, m)
    sys.path.insert(0, os.path.join(REF, "pytorch"))
    import libs.support.utils as utils
    yield utils
    sys.path.remove(os.path.join(REF, "pytorch"))


@pytest.mark.parametrize("blueprint,creation", [
    ("xvector.py", 'Xvector(23,10,training=False,extracted_embedding="far")'),
    ("ecapa_tdnn_xvector.py", "ECAPA_TDNN(" + ECAPA_ARGS + ")"),
])
def test_reference_loader_builds_b200_blueprints_and_loads_reference_state_dicts(ref_utils, blueprint, creation):
    """utils.create_model_from_py (utils.py:163-186) on the B200 blueprint file with the reference's creation string, then
    extract_embeddings.py:63's load_state_dict(strict=False) of a state_dict made by the REFERENCE's own class: same keys,
    same shapes, nothing missing, nothing unexpected; the plugin surface of framework.py is there."""
    ref_model = ref_utils.create_model_from_py(os.path.join(REF, "pytorch/model", blueprint), creation)
    sd = ref_model.state_dict()
    for name in [m for m in sys.modules if m == blueprint[:-3]]:
        del sys.modules[name]                 # same module name, other directory: let the loader import ours
    model = ref_utils.create_model_from_py(os.path.join(ROOT, "asv_subtools_b200/model", blueprint), creation)
    assert type(model).__module__ == blueprint[:-3] and "asv_subtools_b200" in sys.modules[type(model).__module__].__file__
    ours = model.state_dict()
    assert set(ours) == set(sd)
    assert all(tuple(ours[k].shape) == tuple(sd[k].shape) for k in sd)
    res = model.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    k = next(k for k in sd if k.endswith("affine.weight"))
    assert torch.equal(model.state_dict()[k], sd[k])
    for attr in ("extract_embedding", "extracted_embedding", "eval", "train", "cuda", "cpu", "parameters"):
        assert hasattr(model, attr), attr
    assert next(model.parameters()).device.type == "cpu"       # utils.to_device reads this (utils.py:105-114)
    with pytest.raises(Exception):                             # no CPU path: extraction on a CPU-resident model raises
        model.eval().extract_embedding(np.zeros((50, 23 if "xvector.py" == blueprint else 80), dtype=np.float32))
    for name in [m for m in sys.modules if m == blueprint[:-3]]:
        del sys.modules[name]


def test_shell_shadows_cover_the_reference_functions_they_replace():
    """After `. score/process.sh; . score/score.sh; . integration/score_b200.sh` (scoreSets.sh:133-134 + one line) every
    function the B200 file defines existed before under the same name, now runs the B200 CLI, and the functions it does
    not shadow (e.g. get_params, process) are still the reference's."""
    script = r'''
set -e
. {ref}/score/process.sh
. {ref}/score/score.sh
before=$(declare -F | awk '{{print $3}}' | sort)
. {root}/integration/score_b200.sh
for f in $(grep -o '^function [a-z_]*' {root}/integration/score_b200.sh | awk '{{print $2}}' | grep -v '^_'); do
  echo "$before" | grep -qx "$f" || {{ echo "NOT-IN-REFERENCE $f"; exit 3; }}
  declare -f $f | grep -q _xvb200 || {{ echo "NOT-SHADOWED $f"; exit 4; }}
done
declare -f process | grep -q the_process || exit 5
echo OK $(grep -c '^function [a-z]' {root}/integration/score_b200.sh)
'''.format(ref=REF, root=ROOT)
    r = subprocess.run(["bash", "-c", script], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
    assert int(r.stdout.split()[1]) >= 12


def test_extraction_wrapper_rewrites_only_the_hard_coded_extractor_command(tmp_path):
    env = dict(os.environ, XVB200_DRYRUN="1", XVB200_REF=os.path.join(REF, "pytorch/pipeline/extract_xvectors_for_pytorch.sh"))
    r = subprocess.run(["bash", os.path.join(ROOT, "integration/extract_xvectors_b200.sh"), "m", "d", "o"], capture_output=True,
                       text=True, env=env, cwd=str(tmp_path))
    lines = [l for l in r.stdout.splitlines() if l.startswith(">")]
    assert r.returncode == 0 and len(lines) == 2, r.stdout + r.stderr        # the --use-gpu and the CPU branch (:128, :139)
    for l in lines:
        assert "-m asv_subtools_b200.pipeline.extract_embeddings --batch-size 256 --blueprint-dir" in l
        assert "--use-gpu" in l and "onestep/extract_embeddings.py" not in l


@pytest.mark.parametrize("pooling,pp", [
    ("multi-head", {"num_head": 4, "share": False}),
    ("multi-head", {"num_head": 2, "affine_layers": 2, "hidden_size": 32}),
    ("multi-resolution", {"num_head": 4, "temperature": True, "affine_layers": 1, "share": False}),
    ("multi-resolution", {"num_head": 3, "temperature": True, "affine_layers": 2, "share": False, "fixed": False}),
    ("attentive", {"affine_layers": 2, "context": [-1, 0, 1]}),
    ("lde", {"num_head": 5, "num_nodes": 64}),
    ("xi-postdist-softplus2", {"hidden_size": 32, "num_nodes": 64}),
])
def test_snowdar_pooling_variants_register_the_reference_parameters(ref_utils, pooling, pp):
    """Every pooling option of the snowdar blueprint's switch (snowdar_xvector.py:119-136): the B200 blueprint registers the
    same parameter / buffer names and shapes as the reference's (grouped attention affines, temperatures, LDE dictionary,
    xi-vector prior), so reference checkpoints of those configurations load with strict=True."""
    creation = 'Xvector(40,10,training=False,pooling="{}",pooling_params={!r})'.format(pooling, pp)
    ref = ref_utils.create_model_from_py(os.path.join(REF, "pytorch/model/snowdar_xvector.py"), creation)
    want = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    for name in [m for m in sys.modules if m == "snowdar_xvector"]:
        del sys.modules[name]
    ours = ref_utils.create_model_from_py(os.path.join(ROOT, "asv_subtools_b200/model/snowdar_xvector.py"), creation)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == want
    assert not ours.load_state_dict(ref.state_dict(), strict=True).missing_keys
    for name in [m for m in sys.modules if m == "snowdar_xvector"]:
        del sys.modules[name]
