"""Parameter containers mirroring pytorch/libs/nnet/components.py (TdnnAffine :20-165,
ReluBatchNormTdnnLayer :434-461).  They hold parameters under the reference's state_dict keys;
the forward arithmetic lives in csrc/tdnn_gemm.cu (tcgen05) and is driven by the owning model."""
import numpy as np
import torch


class TdnnAffine(torch.nn.Module):
    """y = splice(w * x, context) + b.  Same constructor contract as the reference
    (components.py:30-97): `weight` is (output_dim, input_dim // groups, tot_context) *including* the taps
    that `context` skips; stride=1, pad=True, no weight/feature normalisation on the B200 path.  `groups` > 1
    (the multi-head attention poolings, pooling.py:262-298) keeps the reference's parameter shape; `dense_weight()`
    expands it to the block-diagonal (output_dim, input_dim, tot_context) form the GEMM kernel packs."""

    def __init__(self, input_dim, output_dim, context=[0], bias=True, pad=True, stride=1, groups=1,
                 norm_w=False, norm_f=False):
        super().__init__()
        for i in range(len(context) - 1):
            if context[i] >= context[i + 1]:
                raise ValueError("Context tuple {} is invalid, such as the order.".format(context))
        if stride != 1 or not pad or norm_f:
            raise NotImplementedError("B200 TdnnAffine supports stride=1, pad=True, norm_f=False only")
        self.norm_w = bool(norm_w)
        if input_dim % groups != 0 or output_dim % groups != 0:
            raise ValueError("input_dim {} and output_dim {} must be divisible by groups {}".format(input_dim, output_dim, groups))
        self.input_dim, self.output_dim, self.context, self.groups = input_dim, output_dim, list(context), groups
        self.bool_bias = bias
        self.left_context = context[0] if context[0] < 0 else 0
        self.right_context = context[-1] if context[-1] > 0 else 0
        self.tot_context = self.right_context - self.left_context + 1
        self.weight = torch.nn.Parameter(torch.empty(output_dim, input_dim // groups, self.tot_context))
        self.bias = torch.nn.Parameter(torch.empty(output_dim)) if bias else None
        torch.nn.init.normal_(self.weight, 0.0, 0.01)  # components.py:99-104
        if self.bias is not None:
            torch.nn.init.constant_(self.bias, 0.0)

    def extra_repr(self):
        return "{}, {}, context={}, bias={}, groups={}".format(self.input_dim, self.output_dim, self.context, self.bool_bias,
                                                               self.groups)

    def dense_weight(self):
        """(output_dim, input_dim, tot_context) fp32: the stored weight, or its block-diagonal expansion for groups > 1
        (group g maps input channels [g*Cin/G, (g+1)*Cin/G) to output channels [g*Cout/G, (g+1)*Cout/G), conv1d's rule)."""
        w = self.weight.detach().float()
        if self.norm_w:      # F.normalize(filters, dim=1) at forward time (components.py:139-140): unit L2 per (output, tap)
            w = torch.nn.functional.normalize(w, dim=1)
        if self.groups == 1:
            return w
        g, ci, co = self.groups, self.input_dim // self.groups, self.output_dim // self.groups
        dense = torch.zeros(self.output_dim, self.input_dim, self.tot_context, dtype=torch.float32, device=w.device)
        for k in range(g):
            dense[k * co:(k + 1) * co, k * ci:(k + 1) * ci] = w[k * co:(k + 1) * co]
        return dense


class ReluBatchNormTdnnLayer(torch.nn.Module):
    """affine -> ReLU -> BatchNorm1d (eval), the reference's default "relu-bn" order
    (components.py:347-431).  Options accepted like the reference's **options; only the ones the
    extraction path depends on are interpreted: nonlinearity ('relu' or ''), bn, bn_params.affine."""

    def __init__(self, input_dim, output_dim, context=[0], affine_type="tdnn", **options):
        super().__init__()
        if affine_type != "tdnn":
            raise NotImplementedError("only affine_type='tdnn' is supported")
        if options.get("ln_replace", False):
            raise NotImplementedError("the LayerNorm variant is not on the B200 path")
        self.bn_relu = bool(options.get("bn-relu", False))     # affine -> BN -> ReLU instead of affine -> ReLU -> BN (:386-396)
        nonlin = options.get("nonlinearity", "relu")
        if nonlin not in ("relu", "", None, False):
            raise NotImplementedError("nonlinearity {!r} is not on the B200 path".format(nonlin))
        self.relu = nonlin == "relu"
        self.affine = TdnnAffine(input_dim, output_dim, context=context, bias=options.get("bias", True))
        self.batchnorm = None
        if options.get("bn", True):
            bn_params = {"momentum": 0.1, "affine": True, "track_running_stats": True}
            bn_params.update(options.get("bn_params", {}))
            self.batchnorm = torch.nn.BatchNorm1d(output_dim, **bn_params)

    def folded_bn(self):
        """eval-mode BatchNorm as (scale, shift) float32 arrays: y = x*scale + shift with
        scale = gamma / sqrt(running_var + eps), shift = beta - running_mean*scale."""
        return fold_batchnorm(self.batchnorm)

    def export(self):
        """(weight (Cout, Cin, tot) fp32 tensor, bias | None, scale | None, shift | None, relu) in the kernel's epilogue
        order +bias -> ReLU -> BN.  The "bn-relu" order (BN straight after the affine, then ReLU) has no epilogue of its own:
        eval-mode BN after an affine IS an affine, so it is folded into weight and bias (W' = s W, b' = s b + t) in float64."""
        w = self.affine.dense_weight()
        b = self.affine.bias.detach().float() if self.affine.bias is not None else None
        scale, shift = self.folded_bn()
        if not self.bn_relu or scale is None:
            return w, b, scale, shift, self.relu
        s64, t64 = torch.from_numpy(scale.astype(np.float64)), torch.from_numpy(shift.astype(np.float64))
        w2 = (w.double().cpu() * s64.view(-1, 1, 1)).float().to(w.device)
        b2 = ((b.double().cpu() if b is not None else 0.0) * s64 + t64).float().to(w.device)
        return w2, b2, None, None, self.relu


class FTdnnBlock(torch.nn.Module):
    """Factorised TDNN block (components.py:168-212): factor (in -> bottleneck, context [-c,0], no bias) ->
    affine (bottleneck -> out, context [0,c]) -> ReLU -> BatchNorm -> + bypass_scale * input.  Parameter
    containers only; the semi-orthogonal constraint step is training-side."""

    def __init__(self, input_dim, output_dim, bottleneck_dim, context_size=0, bypass_scale=0.66, pad=True):
        super().__init__()
        if bypass_scale != 0 and input_dim != output_dim:
            raise ValueError("bypass needs input_dim == output_dim")
        self.input_dim, self.output_dim, self.bottleneck_dim = input_dim, output_dim, bottleneck_dim
        self.context_size, self.bypass_scale = context_size, bypass_scale
        c1, c2 = ([-context_size, 0], [0, context_size]) if context_size > 0 else ([0], [0])
        self.factor = TdnnAffine(input_dim, bottleneck_dim, c1, pad=pad, bias=False)
        self.affine = TdnnAffine(bottleneck_dim, output_dim, c2, pad=pad, bias=True)
        self.bn = torch.nn.BatchNorm1d(output_dim, momentum=0.1, affine=True, track_running_stats=True)


def fold_batchnorm(bn):
    if bn is None:
        return None, None
    var = bn.running_var.detach().double().cpu().numpy()
    mean = bn.running_mean.detach().double().cpu().numpy()
    scale = 1.0 / np.sqrt(var + bn.eps)
    if bn.weight is not None:
        scale = scale * bn.weight.detach().double().cpu().numpy()
    shift = -mean * scale
    if bn.bias is not None:
        shift = shift + bn.bias.detach().double().cpu().numpy()
    return scale.astype(np.float32), shift.astype(np.float32)
