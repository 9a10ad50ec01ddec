#!/usr/bin/env python
"""Per-GEMM sensitivity of the ViT-B/16 features to dropped cross terms of the f16f8 operand
encoding (numerics study on the CPU; imports the oracle, so a study tool, not product code).

Every dense contraction of the encoder is x_H w_H + x_L w_C + x_C w_L (csrc/split.cuh).  Per GEMM
kind (in_proj, out_proj, c_fc, c_proj, proj) this script selects

    full   all three terms                                   2   bf16-pass equivalents
    wh     drop x_C w_L: the weights travel as fp16 only     1.5
    xh     drop x_L w_C: the activations travel as fp16 only 1.5
    h      fp16 x fp16 only                                  1
    b3     split-bf16 x3 (what out_proj runs today)          3
    mx4    both cross terms as MXFP4 (e2m1 elements, one power-of-two scale per 32 values along K:
           tcgen05 kind::mxf4, 4x the fp16 rate)             1.5   -- a study of the NEXT operand
           mode, not something the kernels implement

and prints the rel-L2 / max error of the final features against the plain fp32 oracle.

    python scripts/numerics_passes.py [--frames 4] [--outlier 20]
"""
from __future__ import annotations

import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from anomalyclip_b200 import synthetic as syn  # noqa: E402
from oracle import anomalyclip_oracle as oracle  # noqa: E402

_real_linear = F.linear
SX, TX, TW = 4, 7, 4


def _e4m3(v):
    return v.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float64)


def _bf16(v):
    return v.to(torch.bfloat16).to(torch.float64)


_E2M1 = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6], dtype=torch.float64)


def _mxfp4(v, block=32):
    """[rows, K] -> MXFP4 along K: per (row, 32 values) a power-of-two scale chosen so that the
    block maximum fits (<= 6), elements rounded to the nearest e2m1 value."""
    r, k = v.shape
    pad = (-k) % block
    if pad:
        v = torch.cat((v, v.new_zeros(r, pad)), 1)
    b = v.reshape(r, -1, block)
    s = torch.exp2(torch.ceil(torch.log2(b.abs().amax(-1, keepdim=True).clamp_min(1e-300) / 6.0)))
    q = b / s
    idx = (q.abs().unsqueeze(-1) - _E2M1).abs().argmin(-1)
    return (torch.sign(q) * _E2M1[idx] * s).reshape(r, -1)[:, :k]


def kind_of(w: torch.Tensor, width: int) -> str:
    n, k = w.shape
    if (n, k) == (3 * width, width):
        return "in_proj"
    if (n, k) == (width, width):
        return "out_proj"
    if (n, k) == (4 * width, width):
        return "c_fc"
    if (n, k) == (width, 4 * width):
        return "c_proj"
    return "other"


def emulate(x, w, mode):
    x64, w64 = x.double(), w.double()
    if mode == "b3":
        xh, wh = _bf16(x), _bf16(w)
        xl, wl = _bf16((x64 - xh).float()), _bf16((w64 - wh).float())
        return xh @ wh.T + xl @ wh.T + xh @ wl.T
    sw = 15 - math.ceil(math.log2(float(w64.abs().max())))
    xs, ws = x64 * 2.0 ** SX, w64 * 2.0 ** sw
    xh, wh = xs.float().half().double(), ws.float().half().double()
    acc = xh @ wh.T
    if mode == "mx4":
        lead = xs.shape[:-1]
        x2, xh2 = xs.reshape(-1, xs.shape[-1]), xh.reshape(-1, xs.shape[-1])
        acc2 = _mxfp4(x2 - xh2) @ _mxfp4(ws).T + _mxfp4(x2) @ _mxfp4(ws - wh).T
        return (acc + acc2.reshape(*lead, -1)) * 2.0 ** -(SX + sw)
    if mode in ("full", "wh"):
        xl = _e4m3((xs - xh) * 2.0 ** TX)
        wc = _e4m3(w64 * 2.0 ** (sw - TX))
        acc = acc + xl @ wc.T
    if mode in ("full", "xh"):
        wl = _e4m3((ws - wh) * 2.0 ** TW)
        xc = _e4m3(x64 * 2.0 ** (SX - TW))
        acc = acc + xc @ wl.T
    return acc * 2.0 ** -(SX + sw)


def make_linear(modes: dict, width: int):
    def lin(x, w, b=None):
        mode = modes.get(kind_of(w, width), "fp32")
        if mode == "fp32":
            return _real_linear(x, w, b)
        y = emulate(x, w, mode).float()
        return y if b is None else y + b
    return lin


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--outlier", type=float, default=0.0)
    args = ap.parse_args()
    torch.manual_seed(0)
    w = syn.make_vit_weights(layers=args.layers)
    width = w["conv1.weight"].shape[0]
    if args.outlier > 0:
        for k, v in w.items():
            if k.endswith("ln_1.weight") or k.endswith("ln_2.weight"):
                v[::97] *= args.outlier
            if k.endswith("c_fc.weight") or k.endswith("in_proj_weight"):
                v[::131, ::53] *= args.outlier
    frames = syn.normalise_frames(syn.make_frames_u8(args.frames, seed=0))
    kinds = ("in_proj", "out_proj", "c_fc", "c_proj")
    cases = {"all full (2.0)": {k: "full" for k in kinds},
             "today: out_proj b3, rest full": {"in_proj": "full", "out_proj": "b3", "c_fc": "full", "c_proj": "full"},
             "all wh (1.5)": {k: "wh" for k in kinds},
             "all xh (1.5)": {k: "xh" for k in kinds},
             "all h (1.0)": {k: "h" for k in kinds}}
    att_h = {"in_proj": "h", "out_proj": "h"}
    cases["mixed = mode 5: attention side h, MLP full (1.64)"] = {**att_h, "c_fc": "full", "c_proj": "full"}
    cases["mode 6: mode 5 with c_proj wh (1.48)"] = {**att_h, "c_fc": "full", "c_proj": "wh"}
    cases["NEXT: attention side h, MLP mx4 (1.33)"] = {**att_h, "c_fc": "mx4", "c_proj": "mx4"}
    cases["NEXT: every GEMM mx4 (1.5)"] = {k: "mx4" for k in kinds}
    for k in kinds:
        for m in ("wh", "xh", "h"):
            cases[f"only {k} -> {m}"] = {**{q: "full" for q in kinds}, k: m}
    with torch.no_grad():
        ref = oracle.vit_forward(w, frames).double()
        for name, modes in cases.items():
            F.linear = make_linear(modes, width)
            try:
                out = oracle.vit_forward(w, frames).double()
            finally:
                F.linear = _real_linear
            rel = ((out - ref).norm() / ref.norm()).item()
            mx = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"{name:34s} rel-L2 {rel:.3e}   max-err/max|ref| {mx:.3e}", flush=True)


if __name__ == "__main__":
    main()
