"""Generate the golden vectors under tests/golden/ from the REFERENCE's own modules.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):

    python tests/golden/make_golden.py

What is pinned
  vit_small.npz   reference `VisionTransformer` (src/models/components/clip/model.py:233-290),
                  a 2-layer / width-128 / 2-head instance on 32x32 frames, seeded weights.
  selector.npz    reference `SelectorModel.forward(test_mode=True)` (selector_model.py:32-69).
  head.npz        reference `ClassificationHead.forward` (classification_head.py:11-15).
  text.npz        reference `TextEncoder.forward` (text_encoder.py:14-25) on a small reference CLIP
                  text transformer with PromptLearner-style prompts (coop.py:82-90).
  temporal.npz    reference `TemporalModel.forward(test_mode=True)` (temporal_model.py:42-77) run
                  over a STAND-IN for the un-vendored `axial_attention` package: the stand-in is
                  this script's nn.Module restatement of axial-attention 0.6.1, so this vector pins
                  the reference's wiring (projection, einops regrouping, classifier, state_dict
                  names) but NOT the third-party arithmetic (parity unpinned, see oracle header).

The reference has no golden vectors of its own (its tests are template boilerplate), hence this.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def _load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def _randomise(module: nn.Module, gen: torch.Generator, scale: float = 1.0) -> None:
    """Give every parameter a non-trivial value (default inits leave biases 0 and LN at 1/0)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() == 1 and ("ln" in name or "norm" in name) and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))
            else:
                fan_in = p.shape[1] if p.dim() > 1 else p.shape[0]
                p.copy_(scale * torch.randn(p.shape, generator=gen) / (fan_in ** 0.5))


# ------------------------------------------------------------------------------------------
def make_vit():
    clip_model = _load_by_path("ref_clip_model", REF / "src/models/components/clip/model.py")
    torch.manual_seed(1234)
    vit = clip_model.VisionTransformer(input_resolution=32, patch_size=16, width=128, layers=2,
                                       heads=2, output_dim=64).eval()
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for name, p in vit.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))
            elif "ln_" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
    frames = torch.randn(3, 3, 32, 32, generator=gen)
    with torch.no_grad():
        out = vit(frames)
    np.savez(OUT / "vit_small.npz", frames=frames.numpy(), out=out.numpy(),
             **{"w." + k: v for k, v in _np(vit.state_dict()).items()})
    print("vit_small", tuple(out.shape), float(out.abs().mean()))


def make_selector():
    sys.path.insert(0, str(REF))
    from src.models.components.selector_model import SelectorModel

    gen = torch.Generator().manual_seed(11)
    C, D, normal_id = 14, 64, 7
    sel = SelectorModel(classnames=[f"c{i:02d}" for i in range(C)], normal_id=normal_id,
                        logit_scale=nn.Parameter(torch.tensor(4.6)), num_segments=4, seg_length=2,
                        select_idx_dropout_topk=0.7, select_idx_dropout_bottomk=0.7, num_topk=3,
                        num_bottomk=3).eval()
    with torch.no_grad():
        sel.bn_layer.running_mean.copy_(torch.randn(C - 1, generator=gen))
        sel.bn_layer.running_var.copy_(0.5 + 1.5 * torch.rand(C - 1, generator=gen))
    feats = torch.randn(3, 16, D, generator=gen)
    text = torch.randn(C, D, generator=gen)
    ncentroid = 0.1 * torch.randn(D, generator=gen)
    with torch.no_grad():
        out = sel(feats, text, None, ncentroid, True)
    np.savez(OUT / "selector.npz", feats=feats.numpy(), text=text.numpy(),
             ncentroid=ncentroid.numpy(), normal_id=np.int64(normal_id),
             bn_mean=sel.bn_layer.running_mean.numpy(), bn_var=sel.bn_layer.running_var.numpy(),
             out=out.numpy())
    print("selector", tuple(out.shape))


def make_head():
    sys.path.insert(0, str(REF))
    from src.models.components.classification_head import ClassificationHead

    gen = torch.Generator().manual_seed(13)
    head = ClassificationHead(32, 1).eval()
    _randomise(head, gen)
    x = 2.0 * torch.randn(40, 32, generator=gen)
    with torch.no_grad():
        out = head(x)
    np.savez(OUT / "head.npz", x=x.numpy(), out=out.numpy(),
             **{"w." + k: v for k, v in _np(head.state_dict()).items()})
    print("head", tuple(out.shape))


# ------------------------------------------------------------------------------------------
# Stand-in for `axial_attention` (lucidrains/axial-attention 0.6.1): module/parameter naming
# follows that package so the reference's state_dict keys come out as in a real checkpoint.
def _stand_in_axial_attention() -> types.ModuleType:
    class SelfAttention(nn.Module):
        def __init__(self, dim, heads, dim_heads=None):
            super().__init__()
            self.dim_heads = (dim // heads) if dim_heads is None else dim_heads
            dim_hidden = self.dim_heads * heads
            self.heads = heads
            self.to_q = nn.Linear(dim, dim_hidden, bias=False)
            self.to_kv = nn.Linear(dim, 2 * dim_hidden, bias=False)
            self.to_out = nn.Linear(dim_hidden, dim)

        def forward(self, x):
            q, (k, v) = self.to_q(x), self.to_kv(x).chunk(2, dim=-1)
            b, t, d, h, e = *q.shape, self.heads, self.dim_heads

            def merge_heads(z):
                return z.reshape(b, -1, h, e).transpose(1, 2).reshape(b * h, -1, e)

            q, k, v = map(merge_heads, (q, k, v))
            dots = torch.einsum("bie,bje->bij", q, k) * (e ** -0.5)
            dots = dots.softmax(dim=-1)
            out = torch.einsum("bij,bje->bie", dots, v)
            out = out.reshape(b, h, -1, e).transpose(1, 2).reshape(b, -1, d)
            return self.to_out(out)

    class PermuteToFrom(nn.Module):
        def __init__(self, permutation, fn):
            super().__init__()
            self.fn = fn
            self.permutation = permutation
            self.inv_permutation = [permutation.index(i) for i in range(len(permutation))]

        def forward(self, x):
            axial = x.permute(*self.permutation).contiguous()
            shape = axial.shape
            *_, t, d = shape
            axial = self.fn(axial.reshape(-1, t, d))
            return axial.reshape(*shape).permute(*self.inv_permutation).contiguous()

    class PreNorm(nn.Module):
        def __init__(self, dim, fn):
            super().__init__()
            self.fn = fn
            self.norm = nn.LayerNorm(dim)

        def forward(self, x):
            return self.fn(self.norm(x))

    class ChanLayerNorm(nn.Module):
        def __init__(self, dim, eps=1e-5):
            super().__init__()
            self.eps = eps
            self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
            self.b = nn.Parameter(torch.zeros(1, dim, 1, 1))

        def forward(self, x):
            std = torch.var(x, dim=1, unbiased=False, keepdim=True).sqrt()
            mean = torch.mean(x, dim=1, keepdim=True)
            return (x - mean) / (std + self.eps) * self.g + self.b

    class AxialPositionalEmbedding(nn.Module):
        def __init__(self, dim, shape, emb_dim_index=1):
            super().__init__()
            total = len(shape) + 2
            ax = [i for i in range(1, total) if i != emb_dim_index]
            self.num_axials = len(shape)
            for i, (axial_dim, axial_dim_index) in enumerate(zip(shape, ax)):
                s = [1] * total
                s[emb_dim_index] = dim
                s[axial_dim_index] = axial_dim
                setattr(self, f"param_{i}", nn.Parameter(torch.randn(*s)))

        def forward(self, x):
            for i in range(self.num_axials):
                x = x + getattr(self, f"param_{i}")
            return x

    class Deterministic(nn.Module):  # RNG bookkeeping wrapper of the reversible net (inference: id)
        def __init__(self, net):
            super().__init__()
            self.net = net

        def forward(self, *a, **k):
            return self.net(*a, **k)

    class ReversibleBlock(nn.Module):
        def __init__(self, f, g):
            super().__init__()
            self.f, self.g = Deterministic(f), Deterministic(g)

        def forward(self, x):
            x1, x2 = torch.chunk(x, 2, dim=1)
            y1 = x1 + self.f(x2)
            y2 = x2 + self.g(y1)
            return torch.cat([y1, y2], dim=1)

    class ReversibleSequence(nn.Module):
        def __init__(self, blocks):
            super().__init__()
            self.blocks = nn.ModuleList([ReversibleBlock(f, g) for f, g in blocks])

        def forward(self, x):
            x = torch.cat((x, x), dim=1)
            for blk in self.blocks:
                x = blk(x)
            return torch.stack(x.chunk(2, dim=1)).mean(dim=0)

    class AxialImageTransformer(nn.Module):
        def __init__(self, dim, depth, heads=8, dim_heads=None, dim_index=1, reversible=True,
                     axial_pos_emb_shape=None):
            super().__init__()
            assert reversible and dim_index == 1
            permutations = [[0, 3, 2, 1], [0, 2, 3, 1]]  # calculate_permutations(2, 1)

            def conv_ff():
                return nn.Sequential(ChanLayerNorm(dim), nn.Conv2d(dim, dim * 4, 3, padding=1),
                                     nn.LeakyReLU(inplace=True),
                                     nn.Conv2d(dim * 4, dim, 3, padding=1))

            self.pos_emb = (AxialPositionalEmbedding(dim, axial_pos_emb_shape, dim_index)
                            if axial_pos_emb_shape is not None else nn.Identity())
            layers = []
            for _ in range(depth):
                attn = [PermuteToFrom(p, PreNorm(dim, SelfAttention(dim, heads, dim_heads)))
                        for p in permutations]
                layers.append(attn)
                layers.append([conv_ff(), conv_ff()])
            self.layers = ReversibleSequence(layers)

        def forward(self, x):
            return self.layers(self.pos_emb(x))

    mod = types.ModuleType("axial_attention")
    mod.AxialImageTransformer = AxialImageTransformer
    return mod


def make_temporal():
    sys.path.insert(0, str(REF))
    sys.modules["axial_attention"] = _stand_in_axial_attention()
    from src.models.components.temporal_model import TemporalModel

    torch.manual_seed(99)
    n, l, s, b, E, in_dim, depth, heads = 4, 2, 3, 2, 16, 21, 2, 2
    tm = TemporalModel(input_size=in_dim, emb_size=E, output_size=1, heads=heads, dim_heads=None,
                       depth=depth, num_segments=n, seg_length=l).eval()
    gen = torch.Generator().manual_seed(17)
    _randomise(tm, gen)
    with torch.no_grad():  # ChanLayerNorm gains / pos-emb are 4-D: give them values too
        for name, p in tm.named_parameters():
            if name.endswith(".0.g"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            elif name.endswith(".0.b"):
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))
            elif "pos_emb" in name:
                p.copy_(torch.randn(p.shape, generator=gen))
    x = torch.randn(b * n * s * l, in_dim, generator=gen)
    with torch.no_grad():
        out = tm(x, s, True)
    np.savez(OUT / "temporal.npz", x=x.numpy(), out=out.numpy(),
             cfg=np.array([n, l, s, b, E, in_dim, depth, heads], dtype=np.int64),
             **{"w.temporal_model." + k: v for k, v in _np(tm.state_dict()).items()})
    print("temporal", tuple(out.shape), sorted(tm.state_dict().keys())[:6])


def make_text():
    """Reference TextEncoder (text_encoder.py:5-25) over a small reference CLIP text transformer
    (clip/model.py:293-431), fed with PromptLearner-style prompts [SOS][ctx][class .][EOT][pad]."""
    sys.path.insert(0, str(REF))
    clip_model_mod = _load_by_path("ref_clip_model_t", REF / "src/models/components/clip/model.py")
    from src.models.components.text_encoder import TextEncoder

    torch.manual_seed(5)
    n_cls, n_ctx, ctx_len, vocab, width = 5, 4, 20, 64, 64
    clip_model = clip_model_mod.CLIP(embed_dim=32, image_resolution=32, vision_layers=1, vision_width=64,
                                     vision_patch_size=16, context_length=ctx_len, vocab_size=vocab,
                                     transformer_width=width, transformer_heads=2,
                                     transformer_layers=2).float().eval()
    gen = torch.Generator().manual_seed(23)
    with torch.no_grad():
        for name, p_ in clip_model.transformer.named_parameters():
            if name.endswith("bias"):
                p_.copy_(0.1 * torch.randn(p_.shape, generator=gen))
        clip_model.ln_final.weight.copy_(1 + 0.2 * torch.randn(width, generator=gen))
        clip_model.ln_final.bias.copy_(0.1 * torch.randn(width, generator=gen))
    enc = TextEncoder(clip_model).eval()
    tokens = torch.zeros(n_cls, ctx_len, dtype=torch.long)
    for i in range(n_cls):   # SOS, ctx placeholders, name tokens, '.', EOT (= highest id), padding 0
        seq = [vocab - 2] + [7] * n_ctx + [10 + j for j in range(1 + i % 3)] + [9, vocab - 1]
        tokens[i, : len(seq)] = torch.tensor(seq)
    with torch.no_grad():
        emb = clip_model.token_embedding(tokens)
        ctx = 0.02 * torch.randn(n_cls, n_ctx, width, generator=gen)
        prompts = torch.cat([emb[:, :1], ctx, emb[:, 1 + n_ctx:]], dim=1)   # coop.py:82-90
        out = enc(prompts, tokens)
    sd = {"text_encoder." + k: v for k, v in enc.state_dict().items()}
    sd["token_embedding.weight"] = clip_model.token_embedding.weight
    sd["prompt_learner.ctx"] = ctx
    sd["prompt_learner.token_prefix"] = emb[:, :1]
    sd["prompt_learner.token_suffix"] = emb[:, 1 + n_ctx:]
    np.savez(OUT / "text.npz", out=out.numpy(), tokens=tokens.numpy(),
             cfg=np.array([n_cls, n_ctx, ctx_len, vocab, width, 2, 2, 32], dtype=np.int64),
             **{"w." + k: v for k, v in _np(sd).items()})
    print("text", tuple(out.shape), sorted(sd)[:4])


if __name__ == "__main__":
    make_text()
    make_vit()
    make_selector()
    make_head()
    make_temporal()
