"""Seeded synthetic weights / inputs with the reference's state_dict names and shapes.

There is no network in the build or GPU environment, so neither the OpenAI ViT-B/16 weights
(clip/clip.py:38) nor the authors' checkpoints are available: tests, the smoke run and the
benchmark all use these random-init tensors (CLIP-style init, clip/model.py:352-381, plus
non-trivial biases / LayerNorm gains so that every term of every kernel is exercised).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch

Weights = Dict[str, torch.Tensor]


@dataclass(frozen=True)
class PathConfig:
    """The keys of configs/model/anomaly_clip_*.yaml + configs/data/*.yaml the hot path reads."""
    name: str
    num_classes: int
    normal_id: int
    emb_size: int
    depth: int
    concat_features: bool
    heads: int = 8
    num_segments: int = 32
    seg_length: int = 16
    stride: int = 1
    ncrops: int = 1
    feature_dim: int = 512

    @property
    def in_dim(self) -> int:  # anomaly_clip.py:91-93
        return self.feature_dim + (self.num_classes - 1) * int(self.concat_features)

    @property
    def unit(self) -> int:  # frames of one sub-video
        return self.num_segments * self.seg_length


PRESETS = {
    "ucfcrime": PathConfig("ucfcrime", 14, 7, 256, 1, False),
    "shanghaitech": PathConfig("shanghaitech", 18, 8, 256, 2, True),
    "xdviolence": PathConfig("xdviolence", 7, 4, 128, 1, False),
}


def make_vit_weights(width: int = 768, layers: int = 12, patch: int = 16, resolution: int = 224,
                     output_dim: int = 512, seed: int = 1234, prefix: str = "") -> Weights:
    """VisionTransformer state_dict (clip/model.py:233-264) with CLIP's init scales."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    scale = width ** -0.5
    tokens = (resolution // patch) ** 2 + 1
    w: Weights = {
        "conv1.weight": rn(width, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
        "class_embedding": rn(width, std=scale),
        "positional_embedding": rn(tokens, width, std=scale),
        "ln_pre.weight": 1.0 + rn(width, std=0.1),
        "ln_pre.bias": rn(width, std=0.05),
        "ln_post.weight": 1.0 + rn(width, std=0.1),
        "ln_post.bias": rn(width, std=0.05),
        "proj": rn(width, output_dim, std=scale),
    }
    proj_std = scale * (2 * layers) ** -0.5
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        w[p + "ln_1.weight"] = 1.0 + rn(width, std=0.1)
        w[p + "ln_1.bias"] = rn(width, std=0.05)
        w[p + "attn.in_proj_weight"] = rn(3 * width, width, std=scale)
        w[p + "attn.in_proj_bias"] = rn(3 * width, std=0.05)
        w[p + "attn.out_proj.weight"] = rn(width, width, std=proj_std)
        w[p + "attn.out_proj.bias"] = rn(width, std=0.02)
        w[p + "ln_2.weight"] = 1.0 + rn(width, std=0.1)
        w[p + "ln_2.bias"] = rn(width, std=0.05)
        w[p + "mlp.c_fc.weight"] = rn(4 * width, width, std=fc_std)
        w[p + "mlp.c_fc.bias"] = rn(4 * width, std=0.05)
        w[p + "mlp.c_proj.weight"] = rn(width, 4 * width, std=proj_std)
        w[p + "mlp.c_proj.bias"] = rn(width, std=0.02)
    return {prefix + k: v for k, v in w.items()}


def make_temporal_weights(in_dim: int, emb: int, depth: int, heads: int, n: int, l: int,
                          num_classes: int, seed: int = 4321) -> Weights:
    """selector_model.* and temporal_model.* entries of the checkpoint."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    E = emb
    w: Weights = {
        "selector_model.logit_scale": torch.tensor(4.6052),
        "selector_model.bn_layer.running_mean": rn(num_classes - 1),
        "selector_model.bn_layer.running_var": 0.5 + 1.5 * torch.rand(num_classes - 1, generator=g),
        "selector_model.bn_layer.num_batches_tracked": torch.tensor(100),
        "temporal_model.projection.weight": rn(E, in_dim, std=in_dim ** -0.5),
        "temporal_model.projection.bias": rn(E, std=0.05),
        "temporal_model.classifier.layer_norm.weight": 1.0 + rn(E, std=0.1),
        "temporal_model.classifier.layer_norm.bias": rn(E, std=0.05),
        "temporal_model.classifier.linear.weight": rn(1, E, std=E ** -0.5),
        "temporal_model.classifier.linear.bias": rn(1, std=0.1),
        "temporal_model.axial_attn.pos_emb.param_0": rn(1, E, n, 1, std=0.5),
        "temporal_model.axial_attn.pos_emb.param_1": rn(1, E, 1, l, std=0.5),
    }
    for d in range(depth):
        pa = f"temporal_model.axial_attn.layers.blocks.{2 * d}."
        for fg in ("f", "g"):
            q = pa + fg + ".net.fn."
            w[q + "norm.weight"] = 1.0 + rn(E, std=0.1)
            w[q + "norm.bias"] = rn(E, std=0.05)
            w[q + "fn.to_q.weight"] = rn(E, E, std=E ** -0.5)
            w[q + "fn.to_kv.weight"] = rn(2 * E, E, std=E ** -0.5)
            w[q + "fn.to_out.weight"] = rn(E, E, std=E ** -0.5)
            w[q + "fn.to_out.bias"] = rn(E, std=0.02)
        pc = f"temporal_model.axial_attn.layers.blocks.{2 * d + 1}."
        for fg in ("f", "g"):
            q = pc + fg + ".net."
            w[q + "0.g"] = 1.0 + rn(1, E, 1, 1, std=0.1)
            w[q + "0.b"] = rn(1, E, 1, 1, std=0.05)
            w[q + "1.weight"] = rn(4 * E, E, 3, 3, std=(9 * E) ** -0.5)
            w[q + "1.bias"] = rn(4 * E, std=0.05)
            w[q + "3.weight"] = rn(E, 4 * E, 3, 3, std=(36 * E) ** -0.5)
            w[q + "3.bias"] = rn(E, std=0.02)
    return w


def make_state_dict(cfg: PathConfig, with_vit: bool = True, seed: int = 1234,
                    vit_layers: int = 12) -> Weights:
    """`net.`-less state_dict of the hot path for a dataset preset."""
    w = make_temporal_weights(cfg.in_dim, cfg.emb_size, cfg.depth, cfg.heads, cfg.num_segments,
                              cfg.seg_length, cfg.num_classes, seed=seed + 1)
    if with_vit:
        w.update(make_vit_weights(seed=seed, prefix="image_encoder.", layers=vit_layers,
                                  output_dim=cfg.feature_dim))
    return w


def make_text_features(cfg: PathConfig, seed: int = 99) -> torch.Tensor:
    """Stand-in for TextEncoder(PromptLearner()) (anomaly_clip.py:217-221): a (C, 512) constant."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(cfg.num_classes, cfg.feature_dim, generator=g) * 0.4


def make_ncentroid(cfg: PathConfig, seed: int = 98) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(cfg.feature_dim, generator=g) * 0.1


def make_features(cfg: PathConfig, sub_videos: int, seed: int = 0) -> torch.Tensor:
    """(1, ncrops, sub_videos*unit, 512) pre-extracted feature rows (load_from_features=True)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, cfg.ncrops, sub_videos * cfg.unit, cfg.feature_dim, generator=g) * 0.5


def make_frames_u8(num_frames: int, resolution: int = 224, seed: int = 0) -> torch.Tensor:
    """(num_frames, 3, R, R) uint8 frames (already resized / centre-cropped)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (num_frames, 3, resolution, resolution), generator=g,
                         dtype=torch.uint8)


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # src/utils/augmentations.py:21-34
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def normalise_frames(frames_u8: torch.Tensor) -> torch.Tensor:
    """ToTensor (/255) + Normalize(CLIP mean/std): what the reference dataset hands the model."""
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (frames_u8.to(torch.float32) / 255.0 - mean) / std
