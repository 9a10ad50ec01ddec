"""GPU-side frame ingest (aclip_resize_crop_u8) against Pillow / torchvision, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pil_pipeline(frame_hwc: np.ndarray) -> np.ndarray:
    import torchvision.transforms as T
    from PIL import Image
    tf = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])
    return np.asarray(tf(Image.fromarray(frame_hwc))).transpose(2, 0, 1)


@pytest.mark.parametrize("h,w", [(240, 320), (480, 856), (120, 160), (224, 224), (400, 225)])
def test_resize_crop_bit_exact_with_pillow(h, w):
    from anomalyclip_b200.data import GpuFrameIngest
    rng = np.random.default_rng(h + w)
    frames = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    ingest = GpuFrameIngest(h, w, torch.device("cuda"))
    out = ingest(torch.from_numpy(frames).cuda()).cpu().numpy()
    ref = np.stack([_pil_pipeline(f) for f in frames])
    assert out.shape == (3, 3, 224, 224)
    assert np.array_equal(out, ref)


def test_ingest_feeds_the_encoder_like_the_reference_dataset():
    """decoded frames -> GPU resize/crop -> uint8 encoder path == ToTensor+Normalize of the PIL
    pipeline through the oracle ViT (2 layers)."""
    from anomalyclip_b200.data import GpuFrameIngest
    from anomalyclip_b200.engine import PackedVit, VitEncoder
    from oracle import anomalyclip_oracle as oracle
    from tests.parity import assert_parity
    from tests.util_weights import make_vit_weights, normalise_frames
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2, 240, 320, 3), dtype=np.uint8)
    sd = make_vit_weights(layers=2)
    ref_u8 = torch.from_numpy(np.stack([_pil_pipeline(f) for f in frames]))
    ref = oracle.vit_forward(sd, normalise_frames(ref_u8))
    dev = torch.device("cuda")
    out = VitEncoder(PackedVit(sd, dev))(GpuFrameIngest(240, 320, dev)(torch.from_numpy(frames).to(dev)))
    assert_parity(out, ref, "features from GPU-ingested frames")
