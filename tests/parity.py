"""Shared parity bar: north_star's 1e-3 relative fp32 tolerance, argmax bit-exact."""
import torch

RTOL = 1e-3


def rel_l2(out: torch.Tensor, ref: torch.Tensor) -> float:
    out, ref = out.detach().double().cpu(), ref.detach().double().cpu()
    return ((out - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def assert_parity(out: torch.Tensor, ref: torch.Tensor, what: str, rtol: float = RTOL) -> float:
    """rel-L2 <= rtol and max |err| <= rtol * max |ref| (so no single element is off either)."""
    out, ref = out.detach().double().cpu(), ref.detach().double().cpu()
    assert out.shape == ref.shape, f"{what}: shape {tuple(out.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(out).all(), f"{what}: non-finite values"
    err = rel_l2(out, ref)
    mx = (out - ref).abs().max().item() / ref.abs().max().clamp_min(1e-30).item()
    print(f"{what}: rel-L2 {err:.3e}  max-err/max-ref {mx:.3e}")
    assert err <= rtol, f"{what}: rel-L2 {err:.3e} > {rtol}"
    assert mx <= rtol, f"{what}: max error {mx:.3e} > {rtol}"
    return err
