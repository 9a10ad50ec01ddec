"""CPU oracle for the AnomalyCLIP inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module, and only as the checker or the timed CPU baseline.  The
product path (`anomalyclip_b200/`) never imports it and has no CPU fallback.

It is a functional fp32 restatement (plain torch CPU ops on explicit weight dictionaries, no
nn.Module) of what the reference computes with `test_mode=True`.  Each function cites the
reference lines it follows (paths relative to /root/reference).

Pinning status
  * ViT encoder, SelectorModel (test branch), ClassificationHead, test_step post-processing:
    PINNED -- `tests/golden/make_golden.py` runs the reference's own modules (imported from
    /root/reference in the build container) on seeded inputs and stores their outputs under
    `tests/golden/`; `tests/test_oracle.py` checks this file against those vectors.
  * TemporalModel's transformer: **PARITY UNPINNED**.  The arithmetic lives in the third-party
    package `axial_attention` (lucidrains/axial-attention; listed unpinned in the reference's
    requirements.txt:30, latest release 0.6.1; not vendored, not installable offline, and no
    reference test holds a vector for it).  `axial_image_transformer()` below restates the
    published 0.6.1 algorithm (AxialImageTransformer with reversible=True,
    axial_pos_emb_shape=(n, l)) and is anchored on the reference's call site
    `src/models/components/temporal_model.py:32-39,64` and state_dict naming.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# CLIP ViT image encoder                      src/models/components/clip/model.py:174-290
# --------------------------------------------------------------------------------------------
def quick_gelu(x: Tensor) -> Tensor:
    """clip/model.py:183-185 -- x * sigmoid(1.702 x) (not erf GELU)."""
    return x * torch.sigmoid(1.702 * x)


def layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """clip/model.py:174-180 -- nn.LayerNorm evaluated in fp32."""
    return F.layer_norm(x.float(), (x.shape[-1],), g, b, eps)


def multihead_self_attention(x: Tensor, in_w: Tensor, in_b: Tensor, out_w: Tensor, out_b: Tensor,
                             heads: int) -> Tensor:
    """nn.MultiheadAttention(x, x, x, need_weights=False, attn_mask=None) as used at
    clip/model.py:191,206-212.  x is (B, L, D) here (the reference runs it as (L, B, D))."""
    B, L, D = x.shape
    hd = D // heads
    qkv = F.linear(x, in_w, in_b)  # packed q, k, v rows of in_proj_weight
    q, k, v = qkv.split(D, dim=-1)
    q = q.reshape(B, L, heads, hd).transpose(1, 2) * (hd ** -0.5)
    k = k.reshape(B, L, heads, hd).transpose(1, 2)
    v = v.reshape(B, L, heads, hd).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, L, D)
    return F.linear(o, out_w, out_b)


def residual_attention_block(x: Tensor, w: Weights, prefix: str, heads: int) -> Tensor:
    """clip/model.py:214-217."""
    h = layer_norm(x, w[prefix + "ln_1.weight"], w[prefix + "ln_1.bias"])
    x = x + multihead_self_attention(h, w[prefix + "attn.in_proj_weight"],
                                     w[prefix + "attn.in_proj_bias"],
                                     w[prefix + "attn.out_proj.weight"],
                                     w[prefix + "attn.out_proj.bias"], heads)
    h = layer_norm(x, w[prefix + "ln_2.weight"], w[prefix + "ln_2.bias"])
    h = quick_gelu(F.linear(h, w[prefix + "mlp.c_fc.weight"], w[prefix + "mlp.c_fc.bias"]))
    return x + F.linear(h, w[prefix + "mlp.c_proj.weight"], w[prefix + "mlp.c_proj.bias"])


def vit_num_layers(w: Weights) -> int:
    n = 0
    while f"transformer.resblocks.{n}.ln_1.weight" in w:
        n += 1
    return n


def vit_forward(w: Weights, frames: Tensor, heads: Optional[int] = None,
                return_tokens: bool = False) -> Tensor:
    """VisionTransformer.forward, clip/model.py:266-290.  frames: (B, 3, R, R) fp32, already
    normalised.  Returns (B, output_dim)."""
    width = w["conv1.weight"].shape[0]
    patch = w["conv1.weight"].shape[-1]
    heads = heads if heads is not None else width // 64  # clip/model.py:487 (vision_heads)
    x = F.conv2d(frames, w["conv1.weight"], None, stride=patch)       # :267
    x = x.reshape(x.shape[0], width, -1).permute(0, 2, 1)              # :268-269
    cls = w["class_embedding"].expand(x.shape[0], 1, width)           # :270-277
    x = torch.cat([cls, x], dim=1) + w["positional_embedding"]        # :278
    x = layer_norm(x, w["ln_pre.weight"], w["ln_pre.bias"])           # :279
    for i in range(vit_num_layers(w)):                                 # :281-283
        x = residual_attention_block(x, w, f"transformer.resblocks.{i}.", heads)
    if return_tokens:
        return x
    x = layer_norm(x[:, 0, :], w["ln_post.weight"], w["ln_post.bias"])  # :285
    return x @ w["proj"]                                                # :287-288


# --------------------------------------------------------------------------------------------
# SelectorModel, test branch                   src/models/components/selector_model.py:32-69
# --------------------------------------------------------------------------------------------
def selector_forward(image_features: Tensor, text_features: Tensor, ncentroid: Tensor,
                     normal_id: int, bn_mean: Tensor, bn_var: Tensor,
                     bn_eps: float = 1e-5) -> Tensor:
    x = image_features.reshape(-1, image_features.shape[-1])                    # :40-42
    t = torch.cat((text_features[:normal_id], text_features[normal_id + 1:]))   # :44-50
    t = t - ncentroid                                                           # :53
    x = x - ncentroid                                                           # :54
    t = t / t.norm(dim=-1, keepdim=True)                                        # :57-59
    logits = x @ t.T                                                            # :62
    # BatchNorm1d(C-1, affine=False) in eval mode (:30,:65): running statistics
    return (logits - bn_mean) / torch.sqrt(bn_var + bn_eps)


# --------------------------------------------------------------------------------------------
# axial_attention 0.6.1 restatement (PARITY UNPINNED, see module docstring)
# --------------------------------------------------------------------------------------------
def _axial_self_attention(x: Tensor, w: Weights, prefix: str, heads: int) -> Tensor:
    """PreNorm(LayerNorm) + SelfAttention over sequences x: (b, t, d).
    to_q / to_kv have no bias, to_out has one; scale = dim_heads ** -0.5."""
    x = F.layer_norm(x, (x.shape[-1],), w[prefix + "norm.weight"], w[prefix + "norm.bias"])
    q = F.linear(x, w[prefix + "fn.to_q.weight"])
    k, v = F.linear(x, w[prefix + "fn.to_kv.weight"]).chunk(2, dim=-1)
    b, t, dh = q.shape
    e = dh // heads

    def merge(z):
        return z.reshape(b, -1, heads, e).transpose(1, 2).reshape(b * heads, -1, e)

    q, k, v = merge(q), merge(k), merge(v)
    dots = torch.einsum("bie,bje->bij", q, k) * (e ** -0.5)
    dots = dots.softmax(dim=-1)
    out = torch.einsum("bij,bje->bie", dots, v)
    out = out.reshape(b, heads, -1, e).transpose(1, 2).reshape(b, -1, dh)
    return F.linear(out, w[prefix + "fn.to_out.weight"], w[prefix + "fn.to_out.bias"])


def _permute_to_from(x: Tensor, permutation, fn) -> Tensor:
    """PermuteToFrom: move the attended axis and the channel axis last, fold the rest."""
    inv = [permutation.index(i) for i in range(len(permutation))]
    axial = x.permute(*permutation).contiguous()
    shape = axial.shape
    axial = fn(axial.reshape(-1, shape[-2], shape[-1]))
    return axial.reshape(*shape).permute(*inv).contiguous()


def _chan_layer_norm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """ChanLayerNorm over dim 1: (x - mean) / (std + eps) * g + b, biased variance.
    NOTE eps is added to the standard deviation, not to the variance."""
    std = torch.var(x, dim=1, unbiased=False, keepdim=True).sqrt()
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) / (std + eps) * g + b


def _conv_feed_forward(x: Tensor, w: Weights, prefix: str) -> Tensor:
    """Sequential(ChanLayerNorm, Conv2d(E,4E,3,pad 1), LeakyReLU(0.01), Conv2d(4E,E,3,pad 1))."""
    h = _chan_layer_norm(x, w[prefix + "0.g"], w[prefix + "0.b"])
    h = F.conv2d(h, w[prefix + "1.weight"], w[prefix + "1.bias"], padding=1)
    h = F.leaky_relu(h, 0.01)
    return F.conv2d(h, w[prefix + "3.weight"], w[prefix + "3.bias"], padding=1)


def axial_image_transformer(x: Tensor, w: Weights, prefix: str, depth: int, heads: int) -> Tensor:
    """AxialImageTransformer(dim=E, depth, heads, dim_heads=None, reversible=True,
    axial_pos_emb_shape=(n, l)).forward on x: (S, E, n, l).

    Layer list per depth d: blocks[2d]   = (f, g) = (attention along n, attention along l)
                            blocks[2d+1] = (f, g) = (conv feed-forward, conv feed-forward)
    ReversibleSequence: x = cat(x, x); per block y1 = x1 + f(x2); y2 = x2 + g(y1);
    result = mean of the two halves.  calculate_permutations(2, emb_dim=1) gives
    [0,3,2,1] (attend over axis n, sequences of length n) then [0,2,3,1] (attend over axis l)."""
    x = x + w[prefix + "pos_emb.param_0"] + w[prefix + "pos_emb.param_1"]
    x1, x2 = x, x
    perms = ([0, 3, 2, 1], [0, 2, 3, 1])
    for d in range(depth):
        pa = f"{prefix}layers.blocks.{2 * d}."
        y1 = x1 + _permute_to_from(
            x2, perms[0], lambda z: _axial_self_attention(z, w, pa + "f.net.fn.", heads))
        y2 = x2 + _permute_to_from(
            y1, perms[1], lambda z: _axial_self_attention(z, w, pa + "g.net.fn.", heads))
        x1, x2 = y1, y2
        pc = f"{prefix}layers.blocks.{2 * d + 1}."
        y1 = x1 + _conv_feed_forward(x2, w, pc + "f.net.")
        y2 = x2 + _conv_feed_forward(y1, w, pc + "g.net.")
        x1, x2 = y1, y2
    return (x1 + x2) / 2  # torch.stack(x.chunk(2, dim=1)).mean(dim=0)


# --------------------------------------------------------------------------------------------
# TemporalModel + ClassificationHead           temporal_model.py:42-77, classification_head.py
# --------------------------------------------------------------------------------------------
def classification_head(x: Tensor, w: Weights, prefix: str) -> Tensor:
    """classification_head.py:11-15 -- sigmoid(Linear(LayerNorm(x)))."""
    h = F.layer_norm(x, (x.shape[-1],), w[prefix + "layer_norm.weight"],
                     w[prefix + "layer_norm.bias"])
    return torch.sigmoid(F.linear(h, w[prefix + "linear.weight"], w[prefix + "linear.bias"]))


def temporal_forward(features: Tensor, w: Weights, segment_size: int, num_segments: int,
                     seg_length: int, depth: int, heads: int,
                     prefix: str = "temporal_model.") -> Tensor:
    """TemporalModel.forward(features, segment_size, test_mode=True).  features: (N, in_dim),
    rows ordered "(b n s l)"; returns scores (N, 1)."""
    n, s, l = num_segments, segment_size, seg_length
    x = F.linear(features, w[prefix + "projection.weight"], w[prefix + "projection.bias"])  # :43
    E = x.shape[-1]
    b = x.shape[0] // (n * s * l)
    x = x.reshape(b, n, s, l, E).permute(0, 2, 1, 3, 4).reshape(b * s, n, l, E)  # :46-53
    x = x.permute(0, 3, 1, 2)                                                    # :62  b d n l
    x = axial_image_transformer(x, w, prefix + "axial_attn.", depth, heads)      # :64
    x = x.permute(0, 2, 3, 1)                                                    # :67  b n l d
    x = x.reshape(b, s, n, l, E).permute(0, 2, 1, 3, 4).reshape(b * n * s * l, E)  # :70-71
    return classification_head(x, w, prefix + "classifier.")                     # :75


# --------------------------------------------------------------------------------------------
# AnomalyCLIP.forward(test_mode=True)           src/models/components/anomaly_clip.py:115-154
# --------------------------------------------------------------------------------------------
def anomaly_clip_forward(w: Weights, x: Tensor, ncentroid: Tensor, text_features: Tensor, *,
                         segment_size: int, normal_id: int, num_segments: int, seg_length: int,
                         depth: int, heads: int, concat_features: bool, stride: int = 1,
                         ncrops: int = 1, load_from_features: bool = True
                         ) -> Tuple[Tensor, Tensor]:
    """Returns (similarity (N*stride, C-1), scores (N*stride,)).  `text_features` (C, 512) is
    injected: the reference recomputes this per-checkpoint constant on every call (:136)."""
    if not load_from_features:                                                  # :118-131
        b, t, c, h, wd = x.shape
        sd = {k[len("image_encoder."):]: v for k, v in w.items() if k.startswith("image_encoder.")}
        feats = vit_forward(sd, x.reshape(-1, c, h, wd))
        x = feats.reshape(b, ncrops, -1, feats.shape[-1])
    b, nc, t, d = x.shape                                                        # :132
    x = x.reshape(-1, t, d)                                                      # :134
    similarity = selector_forward(x, text_features, ncentroid, normal_id,        # :138
                                  w["selector_model.bn_layer.running_mean"],
                                  w["selector_model.bn_layer.running_var"])
    x = (x - ncentroid).reshape(-1, d)                                           # :143, :224
    feats = torch.cat((similarity, x), dim=-1) if concat_features else x         # :227-231
    scores = temporal_forward(feats, w, segment_size, num_segments, seg_length, depth, heads)
    similarity = similarity.repeat_interleave(stride, dim=0)                     # :149
    scores = scores.repeat_interleave(stride, dim=0).reshape(-1)                 # :150-152
    return similarity, scores


def test_step_postprocess(similarity: Tensor, scores: Tensor, num_labels: Optional[int] = None
                          ) -> Tuple[Tensor, Tensor]:
    """anomaly_clip_module.py:473-483 -- class_probs = softmax(similarity) * score; trim pad."""
    class_probs = torch.softmax(similarity, dim=1) * scores.unsqueeze(1)
    if num_labels is not None:
        class_probs, scores = class_probs[:num_labels], scores[:num_labels]
    return class_probs, scores


test_step_postprocess.__test__ = False  # not a pytest test


# --------------------------------------------------------------------------------------------
# Test-mode frame/feature index plan            feature_dataset.py:17-27,252-259,352-366
#                                               video_dataset.py:237-244,331-343
# --------------------------------------------------------------------------------------------
def padded_length(num_frames: int, num_segments: int, seg_length: int, stride: int = 1) -> int:
    """Test mode pads the frame count up to a multiple of num_segments*seg_length*stride."""
    unit = num_segments * seg_length * stride
    return int(math.ceil(num_frames / unit) * unit)


def test_mode_frame_indices(num_frames: int, num_segments: int, seg_length: int, stride: int = 1):
    """The reference's own loops (feature_dataset.py:252-259,359-366): list of source-frame
    indices in append order, and segment_size (:373)."""
    end_frame = padded_length(num_frames, num_segments, seg_length, stride)
    n_starts = int(end_frame / (seg_length * stride))
    start_indices = [k * (seg_length * stride) for k in range(n_starts)]
    out = []
    for start_index in start_indices:
        for i in range(seg_length):
            out.append((int(start_index) + i * stride) % num_frames)
    return out, len(start_indices) // num_segments


test_mode_frame_indices.__test__ = False


def frame_labels(num_frames: int, start_frame: int, label: int, normal_id: int, intervals):
    """feature_dataset.py:336-349."""
    labels = []
    for i in range(num_frames):
        lab = normal_id
        for s, e in zip(intervals[::2], intervals[1::2]):
            if int(s) <= i + start_frame <= int(e):
                lab = label
        labels.append(lab)
    return labels


def resize_center_crop_u8(frame_hwc, plan):
    """Integer restatement of Pillow's two-pass 8-bit resample (horizontal, then vertical;
    libImaging/Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc) driven by a
    `ResizeCropPlan`; what GroupScale(224, BICUBIC) + GroupCenterCrop(224) produce
    (src/utils/augmentations.py:25-29).  frame_hwc: uint8 numpy (H, W, 3) -> uint8 (3, size, size).
    Pinned against Pillow itself in tests/test_data_cpu.py."""
    import numpy as np
    half = 1 << (22 - 1)
    src = frame_hwc[plan.row0:plan.row0 + plan.rows].astype(np.int64)
    tmp = np.zeros((plan.rows, plan.size, 3), dtype=np.int64)
    for x in range(plan.size):
        x0, n = plan.hbounds[x]
        acc = half + (src[:, x0:x0 + n, :] * plan.hcoeffs[x, :n].astype(np.int64)[None, :, None]).sum(1)
        tmp[:, x, :] = np.clip(acc >> 22, 0, 255)
    out = np.zeros((plan.size, plan.size, 3), dtype=np.int64)
    for y in range(plan.size):
        y0, n = plan.vbounds[y]
        acc = half + (tmp[y0:y0 + n] * plan.vcoeffs[y, :n].astype(np.int64)[:, None, None]).sum(0)
        out[y] = np.clip(acc >> 22, 0, 255)
    return out.astype(np.uint8).transpose(2, 0, 1)
