"""aclip_vit_forward against (a) the reference VisionTransformer's own golden output and
(b) the CPU oracle on seeded ViT-B/16 weights."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import anomalyclip_oracle as oracle
from tests.parity import assert_parity
from tests.util_weights import make_frames_u8, make_vit_weights, normalise_frames

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def _encoder(sd, heads=None, micro_batch=256, passes=3):
    from anomalyclip_b200.engine import PackedVit, VitEncoder
    return VitEncoder(PackedVit(sd, torch.device("cuda"), heads=heads, passes=passes), micro_batch, passes)


def test_small_vit_matches_reference_golden():
    z = np.load(GOLD / "vit_small.npz")
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    enc = _encoder(sd, heads=2)
    out = enc(torch.from_numpy(z["frames"]).cuda())
    assert_parity(out, torch.from_numpy(z["out"]), "small ViT vs reference module")


@pytest.fixture(scope="module")
def vitb16():
    sd = make_vit_weights()
    return sd, _encoder(sd)


def test_vit_b16_fp32_frames_match_oracle(vitb16):
    sd, enc = vitb16
    torch.manual_seed(5)
    frames = torch.randn(3, 3, 224, 224)
    ref = oracle.vit_forward(sd, frames)
    out = enc(frames.cuda())
    assert_parity(out, ref, "ViT-B/16 features")


def test_vit_b16_uint8_frames_normalised_on_gpu(vitb16):
    sd, enc = vitb16
    u8 = make_frames_u8(2, seed=3)
    ref = oracle.vit_forward(sd, normalise_frames(u8))
    out = enc(u8.cuda())
    assert_parity(out, ref, "ViT-B/16 features from uint8 frames")


def test_micro_batching_does_not_change_results(vitb16):
    sd, enc = vitb16
    from anomalyclip_b200.engine import VitEncoder
    frames = make_frames_u8(5, seed=9).cuda()
    a = enc(frames)
    b = VitEncoder(enc.packed, micro_batch=2)(frames)
    assert torch.equal(a, b)


def test_bf16_single_pass_mode_is_less_accurate_but_close(vitb16):
    sd, enc = vitb16
    from anomalyclip_b200.engine import VitEncoder
    torch.manual_seed(6)
    frames = torch.randn(2, 3, 224, 224)
    ref = oracle.vit_forward(sd, frames)
    out = VitEncoder(enc.packed, passes=1)(frames.cuda())
    from tests.parity import rel_l2
    err = rel_l2(out, ref)
    print(f"passes=1 rel-L2 {err:.3e}")
    assert 1e-4 < err < 5e-2  # the plain-bf16 mode misses the 1e-3 bar: this is why passes=3 is the default


# ---- passes = 2: f16f8 operands (fp16 main product + e4m3 cross terms), two pass-equivalents
def test_vit_b16_f16f8_mode_matches_oracle(vitb16):
    sd, enc3 = vitb16
    enc2 = _encoder(sd, passes=2)
    torch.manual_seed(5)
    frames = torch.randn(3, 3, 224, 224)
    ref = oracle.vit_forward(sd, frames)
    out = enc2(frames.cuda())
    assert_parity(out, ref, "ViT-B/16 features, f16f8 GEMM operands")
    from tests.parity import rel_l2
    e2, e3 = rel_l2(out, ref), rel_l2(enc3(frames.cuda()), ref)
    print(f"rel-L2 vs oracle: f16f8 {e2:.3e}, bf16x3 {e3:.3e}")
    assert e2 < 1e-4   # two orders of magnitude inside the 1e-3 bar, like the three-pass mode
    u8 = make_frames_u8(5, seed=3)
    assert_parity(enc2(u8.cuda()), oracle.vit_forward(sd, normalise_frames(u8)),
                  "ViT-B/16 features from uint8 frames, f16f8")
    from anomalyclip_b200.engine import VitEncoder
    assert torch.equal(enc2(u8.cuda()), VitEncoder(enc2.packed, micro_batch=2, passes=2)(u8.cuda()))


def test_small_wide_vit_f16f8_mode():
    sd = make_vit_weights(width=256, layers=2, patch=16, resolution=64, output_dim=256, seed=7)
    torch.manual_seed(8)
    frames = torch.randn(9, 3, 64, 64)
    ref = oracle.vit_forward(sd, frames)
    assert_parity(_encoder(sd, passes=2)(frames.cuda()), ref, "width-256 ViT, f16f8")


def test_f16f8_mode_rejects_unsupported_shapes_and_mixed_packing():
    from anomalyclip_b200._lib import AclipError
    from anomalyclip_b200.engine import PackedVit, VitEncoder
    sd = make_vit_weights(width=128, layers=1, patch=16, resolution=32, output_dim=256, seed=7)
    with pytest.raises(AclipError):   # width must be a multiple of 256 for the CTA-pair kernel
        _encoder(sd, heads=2, passes=2)(torch.zeros(1, 3, 32, 32, device="cuda"))
    with pytest.raises(AclipError):   # bf16-packed weights cannot be run with passes=2
        VitEncoder(PackedVit(sd, torch.device("cuda"), heads=2), passes=2)


def test_bad_inputs_raise():
    from anomalyclip_b200._lib import AclipError
    sd = make_vit_weights(layers=1)
    enc = _encoder(sd)
    with pytest.raises(ValueError):
        enc(torch.zeros(1, 3, 32, 32, device="cuda"))
    with pytest.raises(AclipError):
        enc(torch.zeros(1, 3, 224, 224))
    assert enc(torch.zeros(0, 3, 224, 224, device="cuda")).shape == (0, 512)
