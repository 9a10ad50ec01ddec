#!/usr/bin/env python
"""CPU emulation of operand encodings for the dense contractions of the ViT (numerics study).

Question: can the fp32-faithful GEMM be issued in fewer tensor-pipe cycles than the three bf16
passes (hi*hi + lo*hi + hi*lo)?  Candidate "f16f8": every fp32 operand value travels as

    h  = fp16(v * 2^s)                     main plane      (kind::f16 MMA, 1 pass at the bf16 rate)
    l  = e4m3((v * 2^s - h) * 2^t)         residual plane  (kind::f8f6f4 MMA, 2x the bf16 rate)
    c  = e4m3(v * 2^u)                     coarse copy     (multiplies the OTHER operand's residual)

    acc = sum x_h w_h + sum x_l w_c + sum x_c w_l          (one fp32 accumulator)
    y   = acc * 2^-(sx + sw)        with   uw = sw - tx,  ux = sx - tw

i.e. 1 + 2 * 1/2 = 2 pass-equivalents instead of 3.  This script pushes ViT-B/16 (synthetic
weights, anomalyclip_b200.synthetic) through the oracle with `F.linear` replaced by an emulation of
each encoding and prints the error of the final features against the plain fp32 oracle.  It imports
the oracle, so it is a study tool, not product code.

    python scripts/numerics_f16f8.py [--frames 4] [--layers 12]
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


def _e4m3(v: torch.Tensor) -> torch.Tensor:
    return v.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float64)


def _bf16(v):
    return v.to(torch.bfloat16).to(torch.float64)


def make_linear(mode: str, sx=4, tx=7, tw=4):
    def lin(x, w, b=None):
        x64, w64 = x.double(), w.double()
        if mode == "fp32":
            return _real_linear(x, w, b)
        if mode == "bf16x1":
            acc = _bf16(x) @ _bf16(w).T
        elif mode == "fp16x1":
            acc = x.half().double() @ w.half().double().T
        elif mode == "bf16x3":
            xh, wh = _bf16(x), _bf16(w)
            xl, wl = _bf16((x64 - xh).float()), _bf16((w64 - wh).float())
            acc = xh @ wh.T + xl @ wh.T + xh @ wl.T
        elif mode == "f16f8":
            # per-tensor weight scale chosen at pack time: max|w| * 2^sw in (2^14, 2^15]
            sw = 15 - math.ceil(math.log2(float(w64.abs().max())))
            xs, ws = x64 * 2.0 ** sx, w64 * 2.0 ** sw
            xh, wh = xs.float().half().double(), ws.float().half().double()
            xl, wl = _e4m3((xs - xh) * 2.0 ** tx), _e4m3((ws - wh) * 2.0 ** tw)
            xc, wc = _e4m3(x64 * 2.0 ** (sx - tw)), _e4m3(w64 * 2.0 ** (sw - tx))
            acc = (xh @ wh.T + xl @ wc.T + xc @ wl.T) * 2.0 ** -(sx + sw)
        else:
            raise ValueError(mode)
        y = acc.float()
        return y if b is None else y + b
    return lin


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--outlier", type=float, default=0.0,
                    help="scale a few residual-stream channels / weights by this factor (robustness)")
    args = ap.parse_args()
    torch.manual_seed(0)
    w = syn.make_vit_weights(layers=args.layers)
    if args.outlier > 0:
        for k, v in w.items():
            if k.endswith("ln_1.weight") or k.endswith("ln_2.weight"):
                v[::97] *= args.outlier
            if k.endswith("c_fc.weight") or k.endswith("in_proj_weight"):
                v[::131, ::53] *= args.outlier
    frames = syn.normalise_frames(syn.make_frames_u8(args.frames, seed=0))
    with torch.no_grad():
        ref = oracle.vit_forward(w, frames).double()
        for mode in ("bf16x1", "fp16x1", "bf16x3", "f16f8"):
            F.linear = make_linear(mode)
            try:
                out = oracle.vit_forward(w, frames).double()
            finally:
                F.linear = _real_linear
            rel = ((out - ref).norm() / ref.norm()).item()
            mx = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"{mode:8s} rel-L2 {rel:.3e}   max-err/max|ref| {mx:.3e}")


if __name__ == "__main__":
    main()
