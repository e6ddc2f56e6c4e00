#!/usr/bin/env python
"""Can the tcgen05 layer kernel run with FEWER than three MMAs per algorithmic MAC and still hold the north star's
1e-4 (max|d| / max|ref|) on the embedding?  (VERDICT r1 "next" #9.)  CPU emulation, no GPU needed:

    python tools/split_error_table.py  > profiles/r04_split_error_table.md

Every GEMM of the x-vector stack (BASELINE configs[1]: 80-d, T = 200, "far") is evaluated in float64 on operands
ROUNDED the way a given operand-splitting scheme would round them (products and accumulation exact -- TMEM accumulates
in fp32, whose own error is far below every scheme's operand error), everything else (bias, ReLU, BatchNorm, pooling) in
float64 too, and the embedding is compared with the unrounded float64 forward.  Schemes:

  bf16x3        x = hi + lo (bf16 planes), w likewise; hi*hi + lo*hi + hi*lo           3 MMAs  (shipped)
  fp16x3        the same with fp16 planes                                                3 MMAs
  fp16 x22/w11  (x_hi + x_lo) * w_hi        activations to 22 bits, weights to 11        2 MMAs
  fp16 x11/w22  x_hi * (w_hi + w_lo)        activations to 11 bits, weights to 22        2 MMAs
  bf16 x8/w16   x_hi * (w_hi + w_lo)        drop lo*hi where the input is post-BN         2 MMAs
  tf32 x1       tf32(x) * tf32(w)           one kind::tf32 MMA = the time of 2 bf16 MMAs  "2"
  bf16 x1       x_hi * w_hi                                                              1 MMA
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nnet as onn  # noqa: E402


def rnd(x, dtype):
    return x.to(torch.float32).to(dtype).to(torch.float64)


def tf32(x):
    """Round-to-nearest-even to 10 explicit mantissa bits (the tensor core's tf32 input format)."""
    i = x.to(torch.float32).view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32).to(torch.float64)


def planes(x, dtype):
    hi = rnd(x, dtype)
    return hi, rnd(x - hi, dtype)


def operands(x, w, scheme):
    """-> list of (x_part, w_part) products a scheme issues."""
    if scheme == "exact":
        return [(x, w)]
    if scheme in ("bf16x3", "fp16x3"):
        dt = torch.bfloat16 if scheme == "bf16x3" else torch.float16
        (xh, xl), (wh, wl) = planes(x, dt), planes(w, dt)
        return [(xh, wh), (xl, wh), (xh, wl)]
    if scheme == "fp16 x22/w11":
        (xh, xl), (wh, _) = planes(x, torch.float16), planes(w, torch.float16)
        return [(xh, wh), (xl, wh)]
    if scheme == "fp16 x11/w22":
        (xh, _), (wh, wl) = planes(x, torch.float16), planes(w, torch.float16)
        return [(xh, wh), (xh, wl)]
    if scheme == "bf16 x8/w16":
        (xh, _), (wh, wl) = planes(x, torch.bfloat16), planes(w, torch.bfloat16)
        return [(xh, wh), (xh, wl)]
    if scheme == "tf32 x1":
        return [(tf32(x), tf32(w))]
    if scheme == "bf16 x1":
        return [(rnd(x, torch.bfloat16), rnd(w, torch.bfloat16))]
    raise ValueError(scheme)


def affine(x, w, b, context, scheme):
    """TdnnAffine on (B, C, T) float64 with the scheme's operand rounding; only the taps in `context`."""
    left, right, _ = onn.context_span(context)
    xp = torch.nn.functional.pad(x, (-left, right))
    t = x.shape[2]
    cols = torch.cat([xp[:, :, (c - left):(c - left) + t] for c in context], dim=1)            # (B, ntaps*C, T)
    wk = torch.cat([w[:, :, c - left] for c in context], dim=1)                                # (Cout, ntaps*C)
    y = 0
    for xa, wa in operands(cols, wk, scheme):
        y = y + torch.einsum("bkt,nk->bnt", xa, wa)
    return y + b.view(1, -1, 1)


def forward(sd, x, scheme, per_layer=None):
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    errs = {}
    for name, ctx in onn.XVECTOR_LAYERS:
        sch = scheme if per_layer is None else per_layer.get(name, "bf16x3")
        y = affine(x, sd[name + ".affine.weight"], sd[name + ".affine.bias"], ctx, sch)
        x = onn.batchnorm_eval(torch.relu(y), sd, name + ".batchnorm")
        errs[name] = x
    x = onn.statistics_pooling(x)
    sch = scheme if per_layer is None else per_layer.get("tdnn6", "bf16x3")
    return affine(x, sd["tdnn6.affine.weight"], sd["tdnn6.affine.bias"], [0], sch), errs


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = onn.make_state_dict(onn.xvector_spec(80), 102)                # the bench's checkpoint
    x = torch.from_numpy(onn.synthetic_feats(8, 200, 80, 1024)).transpose(1, 2).double()
    with torch.no_grad():
        ref, ref_l = forward(sd, x, "exact")
        print("# Operand-splitting schemes for the TDNN GEMM: error of the x-vector embedding (CPU emulation)\n")
        print("`tools/split_error_table.py`: 8 utterances x 200 frames x 80-d, the bench's seeded checkpoint, `far` embedding; float64")
        print("everywhere except the GEMM operands, which are rounded as each scheme rounds them.  rel = max|d| / max|ref| per tensor")
        print("(the tests' definition); budget on the embedding: 1e-4 (north star), measured on the GPU for the shipped scheme: 2.0e-5.\n")
        print("| scheme | MMAs per MAC | tdnn1 out | tdnn3 out | tdnn5 out | embedding | min cosine | verdict |")
        print("|---|---|---|---|---|---|---|---|")
        rows = [("bf16x3", 3), ("fp16x3", 3), ("fp16 x22/w11", 2), ("fp16 x11/w22", 2), ("bf16 x8/w16", 2), ("tf32 x1", "1 (= 2 bf16)"),
                ("bf16 x1", 1)]
        for scheme, n in rows:
            emb, lay = forward(sd, x, scheme)
            e = rel(emb, ref)
            cos = float(torch.nn.functional.cosine_similarity(emb.squeeze(2), ref.squeeze(2), dim=1).min())
            print("| {} | {} | {:.1e} | {:.1e} | {:.1e} | **{:.1e}** | 1 - {:.1e} | {} |".format(
                scheme, n, rel(lay["tdnn1"], ref_l["tdnn1"]), rel(lay["tdnn3"], ref_l["tdnn3"]), rel(lay["tdnn5"], ref_l["tdnn5"]), e,
                1 - cos, "holds 1e-4" if e < 1e-4 else "FAILS 1e-4"))
        print("\nMixed: the cheapest 2-MMA scheme on a subset of layers, bf16x3 elsewhere (does any single layer afford it?)\n")
        print("| 2-MMA scheme used on | embedding rel | verdict |")
        print("|---|---|---|")
        for scheme in ("fp16 x22/w11", "fp16 x11/w22"):
            for names in (["tdnn1"], ["tdnn4"], ["tdnn5"], ["tdnn6"], ["tdnn2", "tdnn3"], ["tdnn4", "tdnn5"]):
                emb, _ = forward(sd, x, None, per_layer={n: scheme for n in names})
                e = rel(emb, ref)
                print("| {} on {} | {:.1e} | {} |".format(scheme, "+".join(names), e, "holds 1e-4" if e < 1e-4 else "FAILS 1e-4"))
        print("\nReading: an fp16 (11-bit) or bf16 (8-bit) single plane on EITHER operand puts ~2^-12 (resp. 2^-9) of relative rounding on")
        print("every product of that GEMM; through six layers that lands the embedding at 1e-4 .. 1e-3, at or over the whole budget, where")
        print("bf16x3 sits at ~1e-5.  A 2-MMA variant would buy at most 1.5x on the tensor-bound layers and would spend the")
        print("entire tolerance (or more) to do it; kind::tf32 (one MMA at half rate) is no better than fp16 x11.  Closed: three MMAs stay.")


if __name__ == "__main__":
    main()
