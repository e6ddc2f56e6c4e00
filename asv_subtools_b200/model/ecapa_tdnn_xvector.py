# -*- coding:utf-8 -*-
"""ECAPA-TDNN (SE-Res2Block + attentive statistics pooling) blueprint for the B200 path -- drop-in
for pytorch/model/ecapa_tdnn_xvector.py.

Same constructor signature, creation string and state_dict keys as the reference
(`layerN.conv_relu_bn1.affine.weight`, `layerN.res2net_block.blocks.i.affine.weight` (128,128,2d+1
with the masked taps), `layerN.se.se.1/3.weight`, `mfa.*`, `stats.attention.0/2/4.*`,
`bn_stats.*`, `fc2.*`; ecapa_tdnn_xvector.py:201-357) and the same `extract_embedding()` semantics
(:403-426).  Every contraction runs on the tcgen05 layer kernel; what the reference does with
chunk/cat/expand copies is expressed through channel-slice views instead:

  * Res2Net (:61-75): block i reads chunk i+1 of the 1024-channel tensor and, for i>=1, the previous
    block's output as a SECOND A source accumulated into the same TMEM accumulator
    (W.(sp + spx[i+1]) = W.sp + W.spx[i+1]) -- no add kernel, no chunk/cat copies;
  * SE (:97-111) = plane_mean -> two M=B GEMMs (ReLU, sigmoid epilogues) -> se_apply, which also
    writes the block output straight into its slot of the (B,T,3072) MFA input and the running sum
    x+x1(+x2) that feeds the next block (:405-409);
  * attentive pooling (:173-188): the (B,4608,T) concat is never built -- the time-constant
    [mean,std] part of the first conv becomes a per-utterance bias (a (B,3072)x(3072,128) GEMM),
    the softmax over T and the weighted moments are one streaming online-softmax pass;
  * bn_stats (:412) is folded into fc2's weights at build time.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from asv_subtools_b200 import ops  # noqa: E402
from asv_subtools_b200.nnet import ReluBatchNormTdnnLayer, TopVirtualNnet  # noqa: E402
from asv_subtools_b200.nnet.components import fold_batchnorm  # noqa: E402


def _merge(defaults, given):
    """The subset of utils.assign_params_dict (utils.py:319-356) the blueprint needs: recursive
    override of known keys."""
    out = dict(defaults)
    for k, v in (given or {}).items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


class Res2NetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, scale=8, kernel_size=3, dilation=1, bn_params={}):
        super().__init__()
        assert in_channels % scale == 0 and out_channels % scale == 0 and scale > 1
        width = in_channels // scale
        half = kernel_size // 2
        self.context = [i for i in range(-half * dilation, half * dilation + 1, dilation)]
        self.blocks = nn.ModuleList([ReluBatchNormTdnnLayer(width, out_channels // scale, self.context, **bn_params)
                                     for _ in range(scale - 1)])
        self.scale = scale
        self.width = width


class SE_Connect(nn.Module):
    def __init__(self, channels, bottleneck=128):
        super().__init__()
        self.se = nn.Sequential(nn.AdaptiveAvgPool1d(1), nn.Conv1d(channels, bottleneck, kernel_size=1), nn.ReLU(),
                                nn.Conv1d(bottleneck, channels, kernel_size=1), nn.Sigmoid())


class SE_Res2Block(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, dilation=1, scale=8, bn_params={}):
        super().__init__()
        if in_channels != out_channels:
            raise NotImplementedError("B200 SE_Res2Block implements in_channels == out_channels (no shortcut conv)")
        width = int(math.floor(in_channels / scale))
        self.conv_relu_bn1 = ReluBatchNormTdnnLayer(in_channels, width * scale, **bn_params)
        self.res2net_block = Res2NetBlock(width * scale, width * scale, scale=scale, kernel_size=kernel_size,
                                          dilation=dilation, bn_params=bn_params)
        self.conv_relu_bn2 = ReluBatchNormTdnnLayer(in_channels, width * scale, **bn_params)
        self.se = SE_Connect(out_channels)
        self.shortcut = None


class AttentiveStatsPool(nn.Module):
    def __init__(self, in_dim, bottleneck_dim=128, time_attention=False, bn={}):
        super().__init__()
        if not time_attention:
            raise NotImplementedError("B200 AttentiveStatsPool implements time_attention=True (the c1024 recipe)")
        self.in_dim, self.time_attention = in_dim, time_attention
        self.attention = nn.Sequential(nn.Conv1d(in_dim * 3, bottleneck_dim, kernel_size=1), nn.ReLU(),
                                       nn.BatchNorm1d(bottleneck_dim, **bn), nn.Tanh(),
                                       nn.Conv1d(bottleneck_dim, in_dim, kernel_size=1), nn.Softmax(dim=2))


class ECAPA_TDNN(TopVirtualNnet):
    def init(self, inputs_dim, num_targets, aug_dropout=0., tail_dropout=0., training=True,
             extracted_embedding="near", mixup=False, mixup_alpha=1.0, pooling="ecpa-attentive", pooling_params={},
             ecapa_params={}, fc1=False, fc1_params={}, fc2_params={},
             margin_loss=True, margin_loss_params={}, use_step=False, step_params={}, transfer_from="softmax_loss"):
        default_ecapa = {"channels": 1024, "embd_dim": 192, "mfa_conv": 1536,
                         "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}}
        default_pool = {"hidden_size": 128, "time_attention": True, "stddev": True}
        default_fc = {"nonlinearity": "relu", "nonlinearity_params": {"inplace": True}, "bn-relu": False, "bn": True,
                      "bn_params": {"momentum": 0.5, "affine": True, "track_running_stats": True}}
        if pooling != "ecpa-attentive":
            raise NotImplementedError("B200 ECAPA implements pooling='ecpa-attentive' (the reference default)")
        ecapa_params = _merge(default_ecapa, ecapa_params)
        pooling_params = _merge(default_pool, pooling_params)
        fc1_params = _merge(default_fc, fc1_params)
        fc2_params = _merge(default_fc, fc2_params)
        self.inputs_dim = inputs_dim
        self.use_step, self.step_params = use_step, step_params
        self.extracted_embedding = extracted_embedding
        self.embd_dim = ecapa_params["embd_dim"]
        channels, mfa_conv = ecapa_params["channels"], ecapa_params["mfa_conv"]
        self.layer1 = ReluBatchNormTdnnLayer(inputs_dim, channels, [-2, -1, 0, 1, 2], **ecapa_params)
        self.layer2 = SE_Res2Block(channels, channels, kernel_size=3, dilation=2, scale=8, bn_params=ecapa_params)
        self.layer3 = SE_Res2Block(channels, channels, kernel_size=3, dilation=3, scale=8, bn_params=ecapa_params)
        self.layer4 = SE_Res2Block(channels, channels, kernel_size=3, dilation=4, scale=8, bn_params=ecapa_params)
        self.mfa = ReluBatchNormTdnnLayer(channels * 3, mfa_conv, **ecapa_params)
        self.stats = AttentiveStatsPool(mfa_conv, pooling_params["hidden_size"], pooling_params["time_attention"])
        self.bn_stats = nn.BatchNorm1d(mfa_conv * 2, **ecapa_params["bn_params"])
        self.fc1 = ReluBatchNormTdnnLayer(mfa_conv * 2, self.embd_dim, **fc1_params) if fc1 else None      # :286-287
        self.fc2 = ReluBatchNormTdnnLayer(self.embd_dim if fc1 else mfa_conv * 2, self.embd_dim, **fc2_params)   # :326-333
        self.transform_keys = ["layer1", "layer2", "layer3", "layer4", "stats", "mfa", "bn_stats", "fc1", "fc2", "loss"]
        if margin_loss and transfer_from == "softmax_loss":
            self.rename_transform_keys = {"loss.affine.weight": "loss.weight"}

    def build_extractor(self):
        if self.extracted_embedding == "far":
            assert self.fc1 is not None, "extracted_embedding='far' needs fc1 (ecapa_tdnn_xvector.py:415-416)"
        elif self.extracted_embedding not in ("near", "near_affine"):
            raise TypeError("Expected far or near position, but got {}".format(self.extracted_embedding))
        dev = self.device_for_extraction()
        if os.environ.get("XVB_ECAPA_NATIVE", "1") == "0" or self.layer1.affine.output_dim != 1024:
            return EcapaExtractor(self, dev)       # op-by-op twin; also the path for other channel counts
        return NativeEcapaExtractor(self, dev)


class _Layer:
    """Device-side packed parameters of one TDNN / 1x1-conv layer."""

    def __init__(self, weight, bias, context, bn=None, relu=False, device=None, scale_shift=None):
        w = weight.detach().float().to(device).contiguous()
        self.context = list(context)
        self.w = ops.pack_tdnn_weight(w, self.context)
        self.cout = w.shape[0]
        # one-tap layers keep the (N, K) fp32 matrix too: the segment-level ones run on CUDA cores (ops.small_affine)
        self.w_f32 = w[:, :, 0].contiguous() if w.shape[2] == 1 and w.shape[1] % 4 == 0 else None
        self.bias = bias.detach().float().to(device).contiguous() if bias is not None else None
        self.relu = relu
        scale, shift = scale_shift if scale_shift is not None else fold_batchnorm(bn)
        self.scale = torch.from_numpy(scale).to(device) if scale is not None else None
        self.shift = torch.from_numpy(shift).to(device) if shift is not None else None

    def run(self, x, **kw):
        ops.tdnn_affine_ex(x, self.w, self.cout, self.context, bias=self.bias, bn_scale=self.scale, bn_shift=self.shift,
                           relu=self.relu, **kw)
        _mark("gemm K={}x{} N={}".format(len(self.context), x.channels, self.cout))

    def run_rows(self, x, sigmoid=False):
        """Segment-level form: x (B, K) fp32 -> (B, N) fp32 on CUDA cores (same kernel as the native extractor)."""
        y = ops.small_affine(x, self.w_f32, self.bias, self.scale, self.shift, relu=self.relu, sigmoid=sigmoid)
        _mark("rows K={} N={}".format(x.shape[1], self.cout))
        return y


_PROFILE = None  # list of (label, cuda event) when profiling (tools/bench_ecapa.py --profile)
SMALL_ROWS = os.environ.get("XVB_ECAPA_SMALL", "1") != "0"   # segment-level layers on CUDA cores (csrc/ecapa.cu small_affine)


def _mark(label):
    if _PROFILE is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _PROFILE.append((label, ev))


def _named_layers(m):
    """(name, weight (Cout,Cin,tot) ndarray, bias, context, scale, shift, relu) for xvb_ecapa_set_layer: the state_dict
    tensors as stored, eval BatchNorm folded; the attention conv split into its x / [mean|std] columns (:179) and
    bn_stats folded into fc2 (W' = W diag(s), b' = W t + b)."""
    f = lambda t: t.detach().float().cpu().numpy()  # noqa: E731

    def tdnn(name, layer):
        scale, shift = fold_batchnorm(layer.batchnorm)
        return (name, f(layer.affine.weight), f(layer.affine.bias) if layer.affine.bias is not None else None,
                list(layer.affine.context), scale, shift, layer.relu)

    out = [tdnn("layer1", m.layer1)]
    for li, blk in zip((2, 3, 4), (m.layer2, m.layer3, m.layer4)):
        p = "layer{}.".format(li)
        out.append(tdnn(p + "bn1", blk.conv_relu_bn1))
        out += [tdnn(p + "res{}".format(i), b) for i, b in enumerate(blk.res2net_block.blocks)]
        out.append(tdnn(p + "bn2", blk.conv_relu_bn2))
        out.append((p + "se1", f(blk.se.se[1].weight), f(blk.se.se[1].bias), [0], None, None, True))
        out.append((p + "se2", f(blk.se.se[3].weight), f(blk.se.se[3].bias), [0], None, None, False))
    out.append(tdnn("mfa", m.mfa))
    att, c = m.stats.attention, m.stats.in_dim
    w0 = f(att[0].weight)
    s, t = fold_batchnorm(att[2])
    out.append(("att_x", np.ascontiguousarray(w0[:, :c]), None, [0], s, t, True))
    out.append(("att_gs", np.ascontiguousarray(w0[:, c:]), f(att[0].bias), [0], None, None, False))
    out.append(("att2", f(att[4].weight), f(att[4].bias), [0], None, None, False))
    out += _segment_layers(m)
    return out


def _segment_layers(m):
    """[fc1 ->] [fc2] of ECAPA_TDNN.extract_embedding (:412-422) as (name, w, b, [0], scale, shift, relu) records: "far" =
    fc1.affine alone, "near_affine" = [fc1 full ->] fc2.affine, "near" = [fc1 full ->] fc2 full; bn_stats (eval BatchNorm on
    the pooled statistics, :412) is folded into whichever layer reads them: W' = W diag(s), b' = W t + b, in float64."""
    pos = m.extracted_embedding
    chain = ([("fc1", m.fc1, pos != "far")] if m.fc1 is not None else []) + \
            ([("fc2", m.fc2, pos == "near")] if pos != "far" else [])
    s, t = fold_batchnorm(m.bn_stats)
    out = []
    for i, (name, layer, full) in enumerate(chain):
        if full:
            w, b, scale, shift, relu = layer.export()
        else:
            w, b, scale, shift, relu = layer.affine.dense_weight(), layer.affine.bias.detach().float(), None, None, False
        w = w.double().cpu().numpy()[:, :, 0]
        b = b.double().cpu().numpy()
        if i == 0:
            b = w @ t.astype(np.float64) + b
            w = w * s.astype(np.float64)[None, :]
        out.append((name, w.astype(np.float32)[:, :, None], b.astype(np.float32), [0], scale, shift, relu))
    return out


class NativeEcapaExtractor:
    """xvb_ecapa_t: packed weights, workspace and the whole launch sequence in the C library."""

    def __init__(self, m=None, device=None, path=None):
        import ctypes as C
        from asv_subtools_b200._lib import check, int_array, lib
        self._C, self._lib, self._check = C, lib, check
        self._h = C.c_void_p()
        if path is not None:
            check(lib.xvb_ecapa_load(C.byref(self._h), str(path).encode()), "xvb_ecapa_load")
        else:
            check(lib.xvb_ecapa_create(C.byref(self._h), m.inputs_dim, m.layer1.affine.output_dim, m.stats.in_dim,
                                       m.stats.attention[0].out_channels, m.embd_dim), "xvb_ecapa_create")
            for name, w, b, ctx, scale, shift, relu in _named_layers(m):
                w = np.ascontiguousarray(w, dtype=np.float32)
                w3 = w.reshape(w.shape[0], w.shape[1], -1)
                arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (b, scale, shift)]
                ptr = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
                flags = (1 if relu else 0) | (2 if scale is not None else 0)
                check(lib.xvb_ecapa_set_layer(self._h, name.encode(), w3.shape[0], w3.shape[1], int_array(ctx), len(ctx),
                                              w3.ctypes.data_as(C.c_void_p), ptr[0], ptr[1], ptr[2], flags), "xvb_ecapa_set_layer")
            check(lib.xvb_ecapa_finalize(self._h), "xvb_ecapa_finalize")
        self.feat_dim = lib.xvb_ecapa_feat_dim(self._h)
        self.embed_dim = lib.xvb_ecapa_embed_dim(self._h)

    @classmethod
    def load(cls, path):
        return cls(path=path)

    def save(self, path):
        self._check(self._lib.xvb_ecapa_save(self._h, str(path).encode()), "xvb_ecapa_save")

    @property
    def last_launches(self):
        return self._lib.xvb_ecapa_last_launches(self._h)

    def extract(self, feats):
        if not (isinstance(feats, torch.Tensor) and feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()):
            raise TypeError("feats must be a contiguous CUDA float32 tensor")
        if feats.shape[2] != self.feat_dim:
            raise ValueError("expected feature dim {}, got {}".format(self.feat_dim, feats.shape[2]))
        B, T, _ = feats.shape
        emb = torch.empty(B, self.embed_dim, dtype=torch.float32, device=feats.device)
        C = self._C
        self._check(self._lib.xvb_ecapa_extract(self._h, C.c_void_p(feats.data_ptr()), B, T, C.c_void_p(emb.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "xvb_ecapa_extract")
        return emb

    def extract_shard(self, feats, batch=128, out=None):
        """feats (N,T,F) fp32 CUDA -> (N,D): the whole shard in `batch`-utterance batches, one C call."""
        if not (isinstance(feats, torch.Tensor) and feats.is_cuda and feats.dtype == torch.float32 and feats.is_contiguous()):
            raise TypeError("feats must be a contiguous CUDA float32 tensor")
        n, t, f = feats.shape
        if f != self.feat_dim:
            raise ValueError("expected feature dim {}, got {}".format(self.feat_dim, f))
        emb = out if out is not None else torch.empty(n, self.embed_dim, dtype=torch.float32, device=feats.device)
        C = self._C
        self._check(self._lib.xvb_ecapa_extract_shard(self._h, C.c_void_p(feats.data_ptr()), n, t, int(batch),
                                                      C.c_void_p(emb.data_ptr()),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "xvb_ecapa_extract_shard")
        return emb

    def set_gather(self, pointers, ntables, row0, ld):
        """Replicated-table form of the shard calls (parallel.PeerTable.attach)."""
        self._check(self._lib.xvb_ecapa_set_gather(self._h, pointers, int(ntables), int(row0), int(ld)), "xvb_ecapa_set_gather")

    def extract_shard_host(self, feats_ptr, n, t, emb_ptr, batch=128):
        """Pinned host feats (n,t,F) in, host embeddings (n,D) out; copies overlap the stack."""
        C = self._C
        self._check(self._lib.xvb_ecapa_extract_shard_host(self._h, C.c_void_p(feats_ptr), int(n), int(t), int(batch),
                                                           C.c_void_p(emb_ptr),
                                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    "xvb_ecapa_extract_shard_host")

    def extract_host(self, feats_np):
        """feats (B,T,F) float32 host array -> (B,D) float32 host array (H2D + D2H + one sync inside the call)."""
        feats_np = np.ascontiguousarray(feats_np, dtype=np.float32)
        b, t, f = feats_np.shape
        if f != self.feat_dim:
            raise ValueError("expected feature dim {}, got {}".format(self.feat_dim, f))
        emb = np.empty((b, self.embed_dim), dtype=np.float32)
        C = self._C
        self._check(self._lib.xvb_ecapa_extract_host(self._h, feats_np.ctypes.data_as(C.c_void_p), b, t,
                                                     emb.ctypes.data_as(C.c_void_p),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "xvb_ecapa_extract_host")
        return emb

    def close(self):
        h, self._h = self._h, None
        if h:
            self._lib.xvb_ecapa_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EcapaExtractor:
    """Packed weights on one device + the launch sequence of ECAPA_TDNN.extract_embedding (:403-426), driven from
    Python op by op: the A/B and profiling twin of NativeEcapaExtractor (XVB_ECAPA_NATIVE=0, tools/bench_ecapa.py
    --profile)."""

    def __init__(self, m, device):
        self.device = device
        self.feat_dim = m.inputs_dim
        self.ldf = (m.inputs_dim + 7) // 8 * 8
        self.last_launches = 0

        def tdnn(layer):
            return _Layer(layer.affine.weight, layer.affine.bias, layer.affine.context, bn=layer.batchnorm,
                          relu=layer.relu, device=device)

        self.layer1 = tdnn(m.layer1)
        self.blocks = []
        self.chain = os.environ.get("XVB_ECAPA_RES2NET", "chain") != "gemm"   # one persistent kernel per Res2Net block
        for blk in (m.layer2, m.layer3, m.layer4):
            res = [tdnn(b) for b in blk.res2net_block.blocks]
            self.blocks.append({
                "bn1": tdnn(blk.conv_relu_bn1),
                "res": res,
                "res_w_hi": torch.cat([r.w.hi for r in res], dim=0).contiguous(),
                "res_w_lo": torch.cat([r.w.lo for r in res], dim=0).contiguous(),
                "res_bias": torch.cat([r.bias for r in res]).contiguous(),
                "res_scale": torch.cat([r.scale for r in res]).contiguous(),
                "res_shift": torch.cat([r.shift for r in res]).contiguous(),
                "dilation": blk.res2net_block.context[-1],
                "nscale": blk.res2net_block.scale,
                "width": blk.res2net_block.width,
                "bn2": tdnn(blk.conv_relu_bn2),
                "se1": _Layer(blk.se.se[1].weight, blk.se.se[1].bias, [0], relu=True, device=device),
                "se2": _Layer(blk.se.se[3].weight, blk.se.se[3].bias, [0], device=device),
            })
        self.channels = m.layer1.affine.output_dim
        self.mfa = tdnn(m.mfa)
        att = m.stats.attention
        c = m.stats.in_dim
        w0 = att[0].weight.detach().float()
        self.att_x = _Layer(w0[:, :c].contiguous(), None, [0], bn=att[2], relu=True, device=device)
        self.att_gs = _Layer(w0[:, c:].contiguous(), att[0].bias, [0], device=device)  # [mean | std] columns
        self.att2 = _Layer(att[4].weight, att[4].bias, [0], device=device)
        self.mfa_dim = c
        # segment level: [fc1 ->] [fc2], bn_stats folded into the first (same records as the native extractor gets)
        self.segment = [_Layer(torch.from_numpy(w), torch.from_numpy(b), ctx, relu=relu, device=device, scale_shift=(scale, shift))
                        for _, w, b, ctx, scale, shift, relu in _segment_layers(m)]
        self.embed_dim = m.embd_dim

    def extract(self, feats):
        """feats (B,T,F) fp32 CUDA -> (B, embd_dim) fp32 CUDA (asynchronous on the current stream)."""
        if feats.shape[2] != self.feat_dim:
            raise ValueError("expected feature dim {}, got {}".format(self.feat_dim, feats.shape[2]))
        B, T, _ = feats.shape
        dev, C = feats.device, self.channels
        P = ops.SplitPlanes
        _mark("start")
        xin = ops.split_f32(feats, ld=self.ldf)
        _mark("split")
        X = P.empty((B, T, C), dev)
        self.layer1.run(xin, y=X)
        H, R, Z = P.empty((B, T, C), dev), P.empty((B, T, C), dev), P.empty((B, T, C), dev)
        N = P.empty((B, T, C), dev)
        CAT = P.empty((B, T, 3 * C), dev)
        gate = torch.empty(B, 1, C, dtype=torch.float32, device=dev)
        s1 = P.empty((B, 1, self.blocks[0]["se1"].cout), dev)
        cur = X
        for li, blk in enumerate(self.blocks):
            w = blk["width"]
            blk["bn1"].run(cur, y=H)
            if self.chain and w == 128:
                ops.res2net_block(H, blk["res_w_hi"], blk["res_w_lo"], blk["res_bias"], blk["res_scale"], blk["res_shift"],
                                  blk["dilation"], blk["nscale"], R)
                _mark("res2net chain kernel")
            else:
                ops.copy_planes(H.slice(0, w), R.slice(0, w))   # chunk 0 passes through (ecapa_tdnn_xvector.py:63-64)
                _mark("chunk0 copy")
                for i, layer in enumerate(blk["res"]):
                    layer.run(H.slice(w * (i + 1), w * (i + 2)), x2=R.slice(w * i, w * (i + 1)) if i >= 1 else None,
                              y=R.slice(w * (i + 1), w * (i + 2)))
            blk["bn2"].run(R, y=Z)
            zmean, zm = ops.plane_mean(Z)
            _mark("plane_mean")
            if SMALL_ROWS:
                gate = blk["se2"].run_rows(blk["se1"].run_rows(zmean), sigmoid=True).view(B, 1, C)
            else:
                blk["se1"].run(zm, y=s1)
                blk["se2"].run(s1, sigmoid=True, y_f32=gate)
            last = li + 1 == len(self.blocks)
            ops.se_apply(Z, cur, gate.view(B, C), CAT.slice(C * li, C * (li + 1)), None if last else N)
            _mark("se_apply")
            cur = N
        D = self.mfa_dim
        M = P.empty((B, T, D), dev)
        MF = torch.empty(B, T, D, dtype=torch.float32, device=dev)
        self.mfa.run(CAT, y=M, y_f32=MF)
        gstat, gp = ops.stats_pool_ex(MF, 1e-5, 1, planes=True)      # global mean | sqrt(var_unbiased + 1e-5)
        _mark("stats_pool(global)")
        if SMALL_ROWS:
            ub = self.att_gs.run_rows(gstat)
        else:
            ub = torch.empty(B, 1, self.att_gs.cout, dtype=torch.float32, device=dev)
            self.att_gs.run(gp, y_f32=ub)
        A1 = P.empty((B, T, self.att_x.cout), dev)
        self.att_x.run(M, utt_bias=ub.view(B, -1), tanh=True, y=A1)
        LOG = torch.empty(B, T, D, dtype=torch.float32, device=dev)
        self.att2.run(A1, y_f32=LOG)
        pstat, pp = ops.attn_stats_pool(LOG, MF, 1e-5, planes=True)
        _mark("attn_stats_pool")
        if SMALL_ROWS or len(self.segment) > 1:
            x = pstat
            for layer in self.segment:
                x = layer.run_rows(x)
            return x
        emb = torch.empty(B, 1, self.embed_dim, dtype=torch.float32, device=dev)
        self.segment[0].run(pp, y_f32=emb)
        return emb.view(B, self.embed_dim)

    def close(self):
        pass


# Test.
if __name__ == "__main__":
    print(ECAPA_TDNN(80, 10, training=False))
