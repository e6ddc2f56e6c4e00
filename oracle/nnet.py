"""Oracle (torch-CPU, fp32) restatement of the reference's embedding-extraction forward.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Layout follows the reference:
activations are ``(B, C, T)`` (time contiguous), features arrive as ``(T, F)``.

The arithmetic goes through the same ATen CPU ops the reference calls (``F.pad`` +
``weight*mask`` + ``F.conv1d`` including the masked taps, ``F.batch_norm`` in eval
mode ...), so timing this file on host cores is a faithful stand-in for "the
reference's CPU PyTorch path" (``bench.py`` ``cpu_baseline``, kind ``"port"``).
Parity is pinned by ``tests/golden/*.npz`` produced from the imported reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, used everywhere in the reference


# --------------------------------------------------------------------------------------
# Layer restatements
# --------------------------------------------------------------------------------------
def context_span(context):
    """left/right/total context exactly as TdnnAffine.__init__ computes them
    (pytorch/libs/nnet/components.py:50-53)."""
    left = context[0] if context[0] < 0 else 0
    right = context[-1] if context[-1] > 0 else 0
    return left, right, right - left + 1


def tdnn_affine(x, weight, bias, context, groups=1):
    """TdnnAffine.forward, pytorch/libs/nnet/components.py:107-149.

    x (B, Cin, T); weight (Cout, Cin / groups, tot_context) *including* the masked taps; zero pad
    (-left, right) (:117); taps not in ``context`` are multiplied by 0 (:133-138); dense
    conv1d, stride 1 (:147), `groups` as given to the constructor (:68-76)."""
    left, right, tot = context_span(context)
    assert weight.shape[2] == tot, (weight.shape, context)
    x = F.pad(x, (-left, right), mode="constant", value=0.0)
    if len(context) != tot:
        mask = torch.tensor([[[1.0 if i in context else 0.0 for i in range(left, right + 1)]]],
                            dtype=weight.dtype)
        weight = weight * mask
    return F.conv1d(x, weight, bias, stride=1, padding=0, dilation=1, groups=groups)


def batchnorm_eval(x, sd, prefix):
    """BatchNorm1d in eval mode (running stats, eps 1e-5); affine optional
    (components.py:374-378; ECAPA fc2 uses affine=False, runEcapaXvector_online.py:243-247)."""
    w = sd.get(prefix + ".weight")
    b = sd.get(prefix + ".bias")
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b,
                        training=False, momentum=0.0, eps=BN_EPS)


def relu_bn_tdnn_layer(x, sd, prefix, context, relu=True, bn=True, bn_relu=False, norm_w=False):
    """ReluBatchNormTdnnLayer.forward = affine -> ReLU -> BatchNorm (order: ReLU first),
    components.py:410-431, :434-461; bn_relu: the other order, affine -> BatchNorm -> ReLU (:386-396, :398-403);
    norm_w: F.normalize(filters, dim=1) inside the affine (:139-140)."""
    w = sd[prefix + ".affine.weight"]
    if norm_w:
        left, right, tot = context_span(context)
        mask = torch.tensor([[[1.0 if i in context else 0.0 for i in range(left, right + 1)]]], dtype=w.dtype)
        w = F.normalize(w * mask, dim=1)
    y = tdnn_affine(x, w, sd[prefix + ".affine.bias"], context)
    if bn_relu:
        if bn:
            y = batchnorm_eval(y, sd, prefix + ".batchnorm")
        return F.relu(y) if relu else y
    if relu:
        y = F.relu(y)
    if bn:
        y = batchnorm_eval(y, sd, prefix + ".batchnorm")
    return y


def statistics_pooling(x, eps=1.0e-10):
    """StatisticsPooling.forward, no-lengths branch, pytorch/libs/nnet/pooling.py:58-67:
    mean over T; biased variance sum((x-mean)^2)/T; std = sqrt(clamp(var, eps)); cat."""
    counts = x.shape[2]
    mean = x.mean(dim=2, keepdim=True)
    var = torch.sum((x - mean) ** 2, dim=2, keepdim=True) / counts
    std = torch.sqrt(var.clamp(min=eps))
    return torch.cat((mean, std), dim=1)


# --------------------------------------------------------------------------------------
# Standard x-vector (pytorch/model/xvector.py)
# --------------------------------------------------------------------------------------
XVECTOR_LAYERS = (  # name, context  (xvector.py:26-33)
    ("tdnn1", [-2, -1, 0, 1, 2]),
    ("tdnn2", [-2, 0, 2]),
    ("tdnn3", [-3, 0, 3]),
    ("tdnn4", [0]),
    ("tdnn5", [0]),
)


def xvector_forward(sd, x, extracted_embedding="far", return_intermediates=False):
    """Xvector.extract_embedding body, pytorch/model/xvector.py:84-98.  x (B, F, T)."""
    inter = OrderedDict()
    for name, ctx in XVECTOR_LAYERS:
        x = relu_bn_tdnn_layer(x, sd, name, ctx)
        inter[name] = x
    x = statistics_pooling(x)
    inter["stats"] = x
    if extracted_embedding == "far":
        x = tdnn_affine(x, sd["tdnn6.affine.weight"], sd["tdnn6.affine.bias"], [0])
    elif extracted_embedding == "near":
        x = relu_bn_tdnn_layer(x, sd, "tdnn6", [0])
        x = tdnn_affine(x, sd["tdnn7.affine.weight"], sd["tdnn7.affine.bias"], [0])
    else:
        raise TypeError("Expected far or near position, but got {}".format(extracted_embedding))
    return (x, inter) if return_intermediates else x


EXTENDED_LAYERS = (  # pytorch/model/extended_xvector.py:25-38 (extend=True), in forward order (:101-110)
    ("tdnn1", [-2, -1, 0, 1, 2]), ("ex_tdnn1", [0]), ("tdnn2", [-2, 0, 2]), ("ex_tdnn2", [0]),
    ("tdnn3", [-3, 0, 3]), ("ex_tdnn3", [0]), ("ex_tdnn4", [-4, 0, 4]), ("ex_tdnn5", [0]),
    ("tdnn4", [0]), ("tdnn5", [0]),
)


def extended_xvector_forward(sd, x, extracted_embedding="far"):
    """ExtendedXvector.extract_embedding body, pytorch/model/extended_xvector.py:99-116."""
    for name, ctx in EXTENDED_LAYERS:
        x = relu_bn_tdnn_layer(x, sd, name, ctx)
    x = statistics_pooling(x)
    if extracted_embedding == "far":
        return tdnn_affine(x, sd["tdnn6.affine.weight"], sd["tdnn6.affine.bias"], [0])
    x = relu_bn_tdnn_layer(x, sd, "tdnn6", [0])
    return tdnn_affine(x, sd["tdnn7.affine.weight"], sd["tdnn7.affine.bias"], [0])


def extended_xvector_spec(inputs_dim):
    """Keys/shapes of ExtendedXvector(inputs_dim, N, training=False).state_dict() in registration order."""
    reg = [("tdnn1", inputs_dim, 512, [-2, -1, 0, 1, 2]), ("ex_tdnn1", 512, 512, [0]), ("tdnn2", 512, 512, [-2, 0, 2]),
           ("ex_tdnn2", 512, 512, [0]), ("tdnn3", 512, 512, [-3, 0, 3]), ("ex_tdnn3", 512, 512, [0]),
           ("ex_tdnn4", 512, 512, [-4, 0, 4]), ("ex_tdnn5", 512, 512, [0]), ("tdnn4", 512, 512, [0]),
           ("tdnn5", 512, 1500, [0]), ("tdnn6", 3000, 512, [0]), ("tdnn7", 512, 512, [0])]
    spec = []
    for name, cin, cout, ctx in reg:
        spec += _affine_entries(name, cin, cout, ctx) + _bn_entries(name + ".batchnorm", cout)
    return spec


# --------------------------------------------------------------------------------------
# Snowdar x-vector (pytorch/model/snowdar_xvector.py), the TDNN options: extend, tdnn_layer_params
# (default BatchNorm affine=False), positions far / near_affine / near
# --------------------------------------------------------------------------------------
def snowdar_layers(extend):
    """Module order of Xvector.extract_embedding, snowdar_xvector.py:262-283 (SE / skip connection off)."""
    if not extend:
        return XVECTOR_LAYERS
    return [("tdnn1", [-2, -1, 0, 1, 2]), ("ex_tdnn1", [0]), ("tdnn2", [-2, 0, 2]), ("ex_tdnn2", [0]), ("tdnn3", [-3, 0, 3]),
            ("ex_tdnn3", [0]), ("ex_tdnn4", [-4, 0, 4]), ("ex_tdnn5", [0]), ("tdnn4", [0]), ("tdnn5", [0])]


ATTENTION_DEFAULTS = {"num_head": 1, "share": True, "affine_layers": 1, "hidden_size": 64, "context": [0],
                      "temperature": False, "fixed": True}            # snowdar_xvector.py:44-55


def attention_layout(input_dim, num_head=1, split_input=True, share=True, affine_layers=2, hidden_size=64, bias=True):
    """Shapes / groups of AttentionAlphaComponent's affines, pooling.py:262-298:
    -> (first (cin, cout, groups) | None, last (cin, cout, groups), final_dim)."""
    final_dim = 1 if share else (input_dim // num_head if split_input else input_dim)
    first_groups = last_groups = 1
    if affine_layers == 1:
        last_in = input_dim
        if num_head > 1 and split_input:
            last_groups = num_head
        first = None
    else:
        last_in = hidden_size * num_head
        if num_head > 1:
            last_groups = num_head
            if split_input:
                first_groups = num_head
        first = (input_dim, last_in, first_groups)
    return first, (last_in, final_dim * num_head, last_groups), final_dim


def attention_alpha(x, sd, prefix, input_dim, num_head=1, split_input=True, share=True, affine_layers=2, hidden_size=64,
                    context=(0,), bias=True, temperature=False, fixed=True):
    """AttentionAlphaComponent.forward, pooling.py:300-319: alpha (B, final_dim * num_head, T) = softmax over T of
    last_affine(relu(first_affine(x))), per-head temperature t (fixed: max(1, (i // 2) * 5), :245-249; learnt:
    1 + t^2, :311-312) dividing the logits."""
    first, last, _ = attention_layout(input_dim, num_head, split_input, share, affine_layers, hidden_size, bias)
    context = list(context)
    if first is not None:
        x = F.relu(tdnn_affine(x, sd[prefix + ".first_affine.weight"], sd.get(prefix + ".first_affine.bias"), context, first[2]))
    logits = tdnn_affine(x, sd[prefix + ".last_affine.weight"], sd.get(prefix + ".last_affine.bias"), context, last[2])
    if num_head > 1 and temperature:
        b, _, t = logits.shape
        temp = sd[prefix + ".t"].float() if fixed else 1 + sd[prefix + ".t"] ** 2
        logits = (logits.reshape(b, num_head, -1, t) / temp).reshape(b, -1, t)
    return torch.softmax(logits, dim=2)


def attention_pooling(x, alpha, num_head, global_heads, stddev_attention=True, eps=1.0e-10):
    """The shared tail of AttentiveStatisticsPooling (:347-362), MultiHeadAttentionPooling (:407-437) and the Global /
    MultiResolution variants (:482-512, :553-583): alpha reshaped to (B, head, -1, T) weights x reshaped to
    (B, head, -1, T) (split heads) or (B, 1, -1, T) (global heads); mean = sum_T alpha x; std =
    sqrt(clamp(sum_T alpha x^2 - mean^2, eps)) or, without stddev_attention, sqrt(clamp(mean_T (x - mean)^2, eps))."""
    b, c, t = x.shape
    a = alpha.reshape(b, num_head, -1, t)
    xs = x.reshape(b, 1, -1, t) if global_heads else x.reshape(b, num_head, -1, t)
    mean = torch.sum((a * xs).reshape(b, -1, t), dim=2, keepdim=True)
    if stddev_attention:
        var = torch.sum((a * xs ** 2).reshape(b, -1, t), dim=2, keepdim=True) - mean ** 2
    else:
        var = torch.mean((x - mean) ** 2, dim=2, keepdim=True)
    return torch.cat((mean, torch.sqrt(var.clamp(min=eps))), dim=1)


def lde_pooling(x, mu, s, eps=1.0e-10):
    """LDEPooling.forward, pooling.py:148-159: r = x^T[..., None] - mu; w = softmax over clusters of
    -(s^2 + eps) * sum_c r^2; e = mean over T of w * r; (B, C * c_num, 1)."""
    r = x.transpose(1, 2).unsqueeze(3) - mu
    w = torch.softmax(-(s ** 2 + eps) * torch.sum(r ** 2, dim=2, keepdim=True), dim=3)
    e = torch.mean(w * r, dim=1)
    return e.reshape(-1, mu.shape[0] * mu.shape[1], 1)


def xi_vector_pooling(x, sd, prefix, stddev):
    """xivec_stdinit_softplus2_prec_pooling.forward, pooling.py:188-207."""
    h = relu_bn_tdnn_layer(x, sd, prefix + ".lin1_relu_bn", [0])
    logprec = F.softplus(tdnn_affine(h, sd[prefix + ".lin2.weight"], sd[prefix + ".lin2.bias"], [0]), beta=1, threshold=20)
    logprec = 2.0 * torch.log(logprec)
    b = x.shape[0]
    pl = sd[prefix + ".prior_logprec"].repeat(b, 1).unsqueeze(2)
    pm = sd[prefix + ".prior_mean"].repeat(b, 1).unsqueeze(2)
    w = torch.softmax(torch.cat((logprec, pl), 2), dim=2)
    xx = torch.cat((x, pm), 2)
    phi = torch.sum(xx * w, dim=2)
    if not stddev:
        return phi.unsqueeze(2)
    sigma = torch.sqrt(torch.clamp(torch.sum(xx.pow(2) * w, dim=2) - phi ** 2, min=1.0e-10))
    return torch.cat((phi, sigma), dim=1).unsqueeze(2)


def snowdar_pooling(x, sd, pooling, params, num_nodes):
    """Xvector.init's pooling switch, snowdar_xvector.py:119-136, for statistics / attentive / multi-head /
    multi-resolution."""
    if pooling == "statistics":
        return statistics_pooling(x)
    if pooling == "lde":             # :121-122 -> LDEPooling(num_nodes, c_num=num_head)
        return lde_pooling(x, sd["stats.mu"], sd["stats.s"])
    if pooling.startswith("xi-"):    # :131-134
        return xi_vector_pooling(x, sd, "stats", pooling == "xi-postdist-softplus2")
    p = dict(ATTENTION_DEFAULTS)
    p.update(params)
    if pooling == "attentive":       # :124-126 -> AttentiveStatisticsPooling(:327-337): one head, shared weight, bias
        alpha = attention_alpha(x, sd, "stats.attention", num_nodes, 1, True, True, p["affine_layers"], p["hidden_size"], p["context"])
        return attention_pooling(x, alpha, 1, False)
    if pooling == "multi-head":      # :127-128 -> MultiHeadAttentionPooling(:377-396): split input, no bias
        alpha = attention_alpha(x, sd, "stats.attention", num_nodes, p["num_head"], True, p["share"], p["affine_layers"],
                                p["hidden_size"], p["context"], False, p["temperature"], p["fixed"])
        return attention_pooling(x, alpha, p["num_head"], False)
    if pooling == "multi-resolution":  # :129-130 -> MultiResolutionMultiHeadAttentionPooling(:520-545): global heads, temperature
        alpha = attention_alpha(x, sd, "stats.attention", num_nodes, p["num_head"], False, p["share"], p["affine_layers"],
                                p["hidden_size"], p["context"], True, True, p["fixed"])
        return attention_pooling(x, alpha, p["num_head"], True)
    raise ValueError(pooling)


def snowdar_xvector_forward(sd, x, extracted_embedding="far", extend=False, pooling="statistics", pooling_params=None,
                            bn_relu=False):
    """snowdar_xvector.py:262-294: far = tdnn6.affine; near_affine = tdnn6 -> tdnn7.affine; near = tdnn6 -> tdnn7
    (the whole layer, ReLU and BatchNorm included).  bn_relu: tdnn_layer_params={"bn-relu": True}."""
    for name, ctx in snowdar_layers(extend):
        x = relu_bn_tdnn_layer(x, sd, name, ctx, bn_relu=bn_relu)
    x = snowdar_pooling(x, sd, pooling, pooling_params or {}, x.shape[1])
    if extracted_embedding == "far":
        return tdnn_affine(x, sd["tdnn6.affine.weight"], sd["tdnn6.affine.bias"], [0])
    x = relu_bn_tdnn_layer(x, sd, "tdnn6", [0], bn_relu=bn_relu)
    if extracted_embedding == "near_affine":
        return tdnn_affine(x, sd["tdnn7.affine.weight"], sd["tdnn7.affine.bias"], [0])
    return relu_bn_tdnn_layer(x, sd, "tdnn7", [0], bn_relu=bn_relu)


def snowdar_pooling_spec(pooling, params, num_nodes=1500):   # num_nodes may also come inside params, like the blueprint's
    """(state_dict entries of `stats`, output dim) for the attention poolings (registration order: [t], first_affine,
    last_affine -- pooling.py:245-298)."""
    if pooling == "statistics":
        return [], 2 * num_nodes
    p = dict(ATTENTION_DEFAULTS)
    p.update(params)
    if pooling.startswith("xi-"):    # prior_mean, prior_logprec, lin1_relu_bn (BN affine=True default), lin2 -- pooling.py:179-186
        h = p["hidden_size"]
        spec = [("stats.prior_mean", (1, num_nodes), ("b", 0)), ("stats.prior_logprec", (1, num_nodes), ("b", 0))]
        spec += _affine_entries("stats.lin1_relu_bn", num_nodes, h, [0]) + _bn_entries("stats.lin1_relu_bn.batchnorm", h)
        spec += _affine_entries("stats.lin2", h, num_nodes, [0], "conv")
        return spec, num_nodes * (2 if pooling == "xi-postdist-softplus2" else 1)
    if pooling == "lde":             # parameters mu (C, c_num) ~ randn, s (c_num,) ~ ones (pooling.py:143-144)
        return [("stats.mu", (num_nodes, p["num_head"]), ("lde_mu", 0)), ("stats.s", (p["num_head"],), ("lde_s", 0))], \
            num_nodes * p["num_head"]
    if pooling == "attentive":
        heads, split, share, bias, temp = 1, True, True, True, False
    elif pooling == "multi-head":
        heads, split, share, bias, temp = p["num_head"], True, p["share"], False, p["temperature"]
    else:
        heads, split, share, bias, temp = p["num_head"], False, p["share"], True, True
    first, last, _ = attention_layout(num_nodes, heads, split, share, p["affine_layers"], p["hidden_size"], bias)
    _, _, tot = context_span(p["context"])
    spec = []
    if heads > 1 and temp:       # buffer of fixed temperatures max(1, (i // 2) * 5) (:245-249) or the learnt parameter (:253)
        spec.append(("stats.attention.t", (1, heads, 1, 1), ("temp_fixed" if p["fixed"] else "b", 0)))
    for name, lay in (("first_affine", first), ("last_affine", last)):
        if lay is None:
            continue
        cin, cout, groups = lay
        spec.append(("stats.attention.{}.weight".format(name), (cout, cin // groups, tot), ("w", cin // groups * len(p["context"]))))
        if bias:
            spec.append(("stats.attention.{}.bias".format(name), (cout,), ("b", 0)))
    out_dim = 2 * num_nodes * (heads if pooling == "multi-resolution" else 1)
    return spec, out_dim


def snowdar_xvector_spec(inputs_dim, extend=False, bn_affine=False, pooling="statistics", pooling_params=None):
    """Keys/shapes of snowdar Xvector(inputs_dim, N, extend=..., training=False).state_dict() in registration
    order (:104-152): tdnn1, [ex_tdnn1], tdnn2, [ex_tdnn2], tdnn3, [ex_tdnn3, ex_tdnn4, ex_tdnn5], tdnn4, tdnn5,
    tdnn6, tdnn7; default tdnn_layer_params have BatchNorm affine=False (:45-48)."""
    reg = [("tdnn1", inputs_dim, 512, [-2, -1, 0, 1, 2])] + ([("ex_tdnn1", 512, 512, [0])] if extend else []) + \
          [("tdnn2", 512, 512, [-2, 0, 2])] + ([("ex_tdnn2", 512, 512, [0])] if extend else []) + \
          [("tdnn3", 512, 512, [-3, 0, 3])] + \
          ([("ex_tdnn3", 512, 512, [0]), ("ex_tdnn4", 512, 512, [-4, 0, 4]), ("ex_tdnn5", 512, 512, [0])] if extend else []) + \
          [("tdnn4", 512, 512, [0]), ("tdnn5", 512, (pooling_params or {}).get("num_nodes", 1500), [0])]
    pool_spec, stats_dim = snowdar_pooling_spec(pooling, pooling_params or {}, (pooling_params or {}).get("num_nodes", 1500))
    spec = []
    for name, cin, cout, ctx in reg:
        spec += _affine_entries(name, cin, cout, ctx) + _bn_entries(name + ".batchnorm", cout, affine=bn_affine)
    spec += pool_spec
    for name, cin, cout, ctx in [("tdnn6", stats_dim, 512, [0]), ("tdnn7", 512, 512, [0])]:
        spec += _affine_entries(name, cin, cout, ctx) + _bn_entries(name + ".batchnorm", cout, affine=bn_affine)
    return spec


# --------------------------------------------------------------------------------------
# Factored (F-TDNN) x-vector (pytorch/model/factored_xvector.py; FTdnnBlock components.py:168-212)
# --------------------------------------------------------------------------------------
FTDNN_BLOCKS = {2: (512, 1024, 2, 0.0), 3: (1024, 1024, 0, 0.66), 4: (1024, 1024, 3, 0.66), 5: (1024, 1024, 0, 0.66),
                6: (1024, 1024, 3, 0.66), 7: (2048, 1024, 3, 0.0), 8: (1024, 1024, 3, 0.66), 9: (3072, 1024, 0, 0.0)}


def ftdnn_block(x, sd, prefix, context_size, bypass_scale):
    """FTdnnBlock.forward: factor (no bias, [-c,0]) -> affine ([0,c]) -> ReLU -> BN -> + bypass_scale * input."""
    c1, c2 = ([-context_size, 0], [0, context_size]) if context_size > 0 else ([0], [0])
    out = tdnn_affine(x, sd[prefix + ".factor.weight"], None, c1)
    out = tdnn_affine(out, sd[prefix + ".affine.weight"], sd[prefix + ".affine.bias"], c2)
    out = batchnorm_eval(F.relu(out), sd, prefix + ".bn")
    return out + bypass_scale * x if bypass_scale != 0 else out


def factored_xvector_forward(sd, x, extracted_embedding="far"):
    """Xvector.extract_embedding, factored_xvector.py:99-122."""
    blk = lambda i, v: ftdnn_block(v, sd, "layer{:02d}".format(i), FTDNN_BLOCKS[i][2], FTDNN_BLOCKS[i][3])  # noqa: E731
    x1 = relu_bn_tdnn_layer(x, sd, "layer01", [-2, -1, 0, 1, 2])
    x2 = blk(2, x1)
    x3 = blk(3, x2)
    x4 = blk(4, x3)
    x5 = blk(5, x3)
    x6 = blk(6, x5)
    x7 = blk(7, torch.cat((x2, x4), 1))
    x8 = blk(8, x7)
    x9 = blk(9, torch.cat((x4, x6, x8), 1))
    v = statistics_pooling(relu_bn_tdnn_layer(x9, sd, "layer10", [0]))
    if extracted_embedding == "far":
        return tdnn_affine(v, sd["embedding1.affine.weight"], sd["embedding1.affine.bias"], [0])
    v = relu_bn_tdnn_layer(v, sd, "embedding1", [0])
    return tdnn_affine(v, sd["embedding2.affine.weight"], sd["embedding2.affine.bias"], [0])


def factored_xvector_spec(inputs_dim, embd_dim=512):
    """Keys/shapes of factored Xvector(inputs_dim, N, training=False).state_dict() in registration order."""
    spec = _affine_entries("layer01", inputs_dim, 512, [-2, -1, 0, 1, 2]) + _bn_entries("layer01.batchnorm", 512)
    for i in range(2, 10):
        cin, cout, c, _ = FTDNN_BLOCKS[i]
        p = "layer{:02d}".format(i)
        c1, c2 = ([-c, 0], [0, c]) if c > 0 else ([0], [0])
        spec += [e for e in _affine_entries(p + ".factor", cin, 256, c1, key_style="plain") if not e[0].endswith(".bias")]
        spec += _affine_entries(p + ".affine", 256, cout, c2, key_style="plain")
        spec += _bn_entries(p + ".bn", cout)
    spec += _affine_entries("layer10", 1024, 2048, [0]) + _bn_entries("layer10.batchnorm", 2048)
    spec += _affine_entries("embedding1", 4096, embd_dim, [0]) + _bn_entries("embedding1.batchnorm", embd_dim)
    spec += _affine_entries("embedding2", embd_dim, embd_dim, [0]) + _bn_entries("embedding2.batchnorm", embd_dim)
    return spec


# --------------------------------------------------------------------------------------
# ECAPA-TDNN c1024 (pytorch/model/ecapa_tdnn_xvector.py)
# --------------------------------------------------------------------------------------
def res2net_block(x, sd, prefix, dilation, scale=8):
    """Res2NetBlock.forward, ecapa_tdnn_xvector.py:61-75 (context from :50-51)."""
    ctx = [-dilation, 0, dilation]
    spx = torch.chunk(x, scale, dim=1)
    y = [spx[0]]
    sp = None
    for i in range(scale - 1):
        sp = spx[i + 1] if i == 0 else sp + spx[i + 1]
        sp = relu_bn_tdnn_layer(sp, sd, "{}.blocks.{}".format(prefix, i), ctx)
        y.append(sp)
    return torch.cat(y, dim=1)


def se_connect(x, sd, prefix):
    """SE_Connect.forward, ecapa_tdnn_xvector.py:97-111: avgpool(T) -> conv1x1 -> ReLU ->
    conv1x1 -> sigmoid; x * gate."""
    s = x.mean(dim=2, keepdim=True)
    s = F.relu(F.conv1d(s, sd[prefix + ".se.1.weight"], sd[prefix + ".se.1.bias"]))
    s = torch.sigmoid(F.conv1d(s, sd[prefix + ".se.3.weight"], sd[prefix + ".se.3.bias"]))
    return x * s


def se_res2block(x, sd, prefix, dilation):
    """SE_Res2Block.forward, ecapa_tdnn_xvector.py:139-149 (in==out so no shortcut conv)."""
    residual = x
    x = relu_bn_tdnn_layer(x, sd, prefix + ".conv_relu_bn1", [0])
    x = res2net_block(x, sd, prefix + ".res2net_block", dilation)
    x = relu_bn_tdnn_layer(x, sd, prefix + ".conv_relu_bn2", [0])
    x = se_connect(x, sd, prefix + ".se")
    return x + residual


def attentive_stats_pool(x, sd, prefix="stats"):
    """AttentiveStatsPool.forward with time_attention=True, ecapa_tdnn_xvector.py:173-188.
    global std uses torch.var default (unbiased) + 1e-5 (:177-178); attention =
    conv -> ReLU -> BN -> tanh -> conv -> softmax over T (:164-171)."""
    g_mean = torch.mean(x, dim=2, keepdim=True).expand_as(x)
    g_std = torch.sqrt(torch.var(x, dim=-1, keepdim=True) + 1e-5).expand_as(x)
    x_in = torch.cat((x, g_mean, g_std), dim=1)
    a = F.conv1d(x_in, sd[prefix + ".attention.0.weight"], sd[prefix + ".attention.0.bias"])
    a = F.relu(a)
    a = batchnorm_eval(a, sd, prefix + ".attention.2")
    a = torch.tanh(a)
    a = F.conv1d(a, sd[prefix + ".attention.4.weight"], sd[prefix + ".attention.4.bias"])
    alpha = torch.softmax(a, dim=2)
    mean = torch.sum(alpha * x, dim=2)
    residuals = torch.sum(alpha * (x ** 2), dim=2) - mean ** 2
    std = torch.sqrt(residuals.clamp(min=1e-5))
    return torch.cat([mean, std], dim=1)


def ecapa_forward(sd, x, extracted_embedding="near", fc2_relu=False, return_intermediates=False, fc1=False):
    """ECAPA_TDNN.extract_embedding body, ecapa_tdnn_xvector.py:403-426 with the canonical
    c1024 parameters of pytorch/launcher/runEcapaXvector_online.py:221-263 (fc1=False; fc2
    nonlinearity '' and BN affine=False -> ``fc2_relu=False`` and no fc2.batchnorm.weight
    key).  ``fc2_relu=True`` reproduces the blueprint's constructor defaults (:219-224)."""
    inter = OrderedDict()
    x = relu_bn_tdnn_layer(x, sd, "layer1", [-2, -1, 0, 1, 2])
    inter["layer1"] = x
    x1 = se_res2block(x, sd, "layer2", 2)
    x2 = se_res2block(x + x1, sd, "layer3", 3)
    x3 = se_res2block(x + x1 + x2, sd, "layer4", 4)
    inter["layer2"], inter["layer3"], inter["layer4"] = x1, x2, x3
    x = torch.cat([x1, x2, x3], dim=1)
    x = relu_bn_tdnn_layer(x, sd, "mfa", [0])
    inter["mfa"] = x
    x = attentive_stats_pool(x, sd, "stats")
    inter["stats"] = x
    x = batchnorm_eval(x, sd, "bn_stats")
    x = x.unsqueeze(2)
    if extracted_embedding == "far":                     # :414-416 (fc1=True models only; fc1 with the constructor defaults)
        assert fc1
        x = tdnn_affine(x, sd["fc1.affine.weight"], sd["fc1.affine.bias"], [0])
    elif extracted_embedding in ("near_affine", "near"):
        if fc1:
            x = relu_bn_tdnn_layer(x, sd, "fc1", [0])
        if extracted_embedding == "near_affine":
            x = tdnn_affine(x, sd["fc2.affine.weight"], sd["fc2.affine.bias"], [0])
        else:
            x = relu_bn_tdnn_layer(x, sd, "fc2", [0], relu=fc2_relu)
    else:
        raise TypeError("Expected far or near position, but got {}".format(extracted_embedding))
    return (x, inter) if return_intermediates else x


# --------------------------------------------------------------------------------------
# Whole-utterance wrapper
# --------------------------------------------------------------------------------------
def extract_embedding(forward_fn, feats, max_chunk=10000):
    """for_extract_embedding wrapper, pytorch/libs/nnet/framework.py:18-52.

    feats: (T, F) float32 ndarray -> (1, F, T); num_split = ceil(T/maxChunk);
    split_size = T // num_split; the last chunk takes the remainder; embedding =
    sum(len_i * emb_i) / T; returns a 1-D CPU tensor."""
    with torch.no_grad():
        x = torch.tensor(feats).unsqueeze(0).transpose(1, 2)
        num_frames = x.shape[2]
        num_split = (num_frames + max_chunk - 1) // max_chunk
        split_size = num_frames // num_split
        offset = 0
        stats = 0.0
        for _ in range(num_split - 1):
            stats = stats + split_size * forward_fn(x[:, :, offset:offset + split_size])
            offset += split_size
        last = forward_fn(x[:, :, offset:])
        emb = (stats + (num_frames - offset) * last) / num_frames
        return torch.squeeze(emb.transpose(1, 2))


# --------------------------------------------------------------------------------------
# Seeded synthetic checkpoints (shapes = the reference constructors'; make_golden.py
# asserts key-for-key equality with the imported reference's state_dict()).
# --------------------------------------------------------------------------------------
def _affine_entries(prefix, cin, cout, context, key_style="tdnn"):
    _, _, tot = context_span(context)
    if key_style == "tdnn":
        return [(prefix + ".affine.weight", (cout, cin, tot), ("w", cin * len(context))),
                (prefix + ".affine.bias", (cout,), ("b", 0))]
    return [(prefix + ".weight", (cout, cin, tot), ("w", cin)), (prefix + ".bias", (cout,), ("b", 0))]


def _bn_entries(prefix, c, affine=True):
    out = []
    if affine:
        out += [(prefix + ".weight", (c,), ("gamma", 0)), (prefix + ".bias", (c,), ("beta", 0))]
    out += [(prefix + ".running_mean", (c,), ("rmean", 0)), (prefix + ".running_var", (c,), ("rvar", 0)),
            (prefix + ".num_batches_tracked", (), ("nbt", 0))]
    return out


def xvector_spec(inputs_dim):
    """Keys/shapes of Xvector(inputs_dim, N, training=False).state_dict() (xvector.py:26-33)."""
    spec = []
    dims = [(inputs_dim, 512), (512, 512), (512, 512), (512, 512), (512, 1500)]
    for (name, ctx), (cin, cout) in zip(XVECTOR_LAYERS, dims):
        spec += _affine_entries(name, cin, cout, ctx) + _bn_entries(name + ".batchnorm", cout)
    spec += _affine_entries("tdnn6", 3000, 512, [0]) + _bn_entries("tdnn6.batchnorm", 512)
    spec += _affine_entries("tdnn7", 512, 512, [0]) + _bn_entries("tdnn7.batchnorm", 512)
    return spec


def ecapa_spec(inputs_dim, channels=1024, embd_dim=192, mfa_conv=1536, hidden=128,
               fc2_bn_affine=False, scale=8, se_bottleneck=128, fc1=False):
    """Keys/shapes of ECAPA_TDNN(inputs_dim, N, training=False, <canonical params>).state_dict()
    (ecapa_tdnn_xvector.py:259-332)."""
    spec = _affine_entries("layer1", inputs_dim, channels, [-2, -1, 0, 1, 2]) + \
        _bn_entries("layer1.batchnorm", channels)
    width = channels // scale
    for li, d in (("layer2", 2), ("layer3", 3), ("layer4", 4)):
        spec += _affine_entries(li + ".conv_relu_bn1", channels, channels, [0]) + \
            _bn_entries(li + ".conv_relu_bn1.batchnorm", channels)
        for i in range(scale - 1):
            p = "{}.res2net_block.blocks.{}".format(li, i)
            spec += _affine_entries(p, width, width, [-d, 0, d]) + _bn_entries(p + ".batchnorm", width)
        spec += _affine_entries(li + ".conv_relu_bn2", channels, channels, [0]) + \
            _bn_entries(li + ".conv_relu_bn2.batchnorm", channels)
        spec += _affine_entries(li + ".se.se.1", channels, se_bottleneck, [0], "conv") + \
            _affine_entries(li + ".se.se.3", se_bottleneck, channels, [0], "conv")
    spec += _affine_entries("mfa", 3 * channels, mfa_conv, [0]) + _bn_entries("mfa.batchnorm", mfa_conv)
    spec += _affine_entries("stats.attention.0", 3 * mfa_conv, hidden, [0], "conv") + \
        _bn_entries("stats.attention.2", hidden) + \
        _affine_entries("stats.attention.4", hidden, mfa_conv, [0], "conv")
    spec += _bn_entries("bn_stats", 2 * mfa_conv)
    if fc1:       # ReluBatchNormTdnnLayer(2 * mfa_conv, embd_dim) with the constructor's default fc params (BN affine=True)
        spec += _affine_entries("fc1", 2 * mfa_conv, embd_dim, [0]) + _bn_entries("fc1.batchnorm", embd_dim)
    spec += _affine_entries("fc2", embd_dim if fc1 else 2 * mfa_conv, embd_dim, [0]) + \
        _bn_entries("fc2.batchnorm", embd_dim, affine=fc2_bn_affine)
    return spec


def make_state_dict(spec, seed):
    """Seeded synthetic checkpoint: He-scaled weights on *all* stored taps (masked taps get
    garbage on purpose), randomised BN statistics so that a folded or
    mis-ordered BN/ReLU epilogue cannot hide."""
    rng = np.random.RandomState(seed)
    sd = OrderedDict()
    for key, shape, (kind, fan_in) in spec:
        if kind == "w":
            v = rng.standard_normal(shape).astype(np.float32) * np.float32(math.sqrt(2.0 / fan_in))
        elif kind == "b":
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif kind == "gamma":
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif kind == "beta":
            v = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif kind == "rmean":
            v = rng.uniform(0.1, 0.6, shape).astype(np.float32)
        elif kind == "rvar":
            v = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif kind == "nbt":
            sd[key] = torch.tensor(1000, dtype=torch.long)
            continue
        elif kind == "lde_mu":
            v = (0.6 * rng.standard_normal(shape)).astype(np.float32)
        elif kind == "lde_s":
            v = rng.uniform(0.05, 0.15, shape).astype(np.float32)      # beta = s^2 ~ 0.01: soft assignments over C ~ 1e2..1e3 dims
        elif kind == "temp_fixed":
            sd[key] = torch.tensor([[[[max(1, (i // 2) * 5)]] for i in range(shape[1])]])
            continue
        else:
            raise ValueError(kind)
        sd[key] = torch.from_numpy(v)
    return sd


def synthetic_feats(num_utts, num_frames, feat_dim, seed):
    """N(0,1) frames, (num_utts, T, F) float32 (CMN-like zero mean)."""
    rng = np.random.RandomState(seed)
    return rng.standard_normal((num_utts, num_frames, feat_dim)).astype(np.float32)
