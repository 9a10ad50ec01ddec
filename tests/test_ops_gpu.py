"""Parity of the building-block kernels (LayerNorm, ViT attention, axial attention) through the
C ABI against plain fp32/fp64 torch restatements of the same operator."""
import pytest
import torch

from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from anomalyclip_b200 import ops as _ops
    return _ops


def _unsplit(s):
    return s[0].float() + s[1].float()


@pytest.mark.parametrize("rows,D", [(1, 768), (197 * 3, 768), (1000, 256), (513, 128), (9, 64)])
@pytest.mark.parametrize("chan", [False, True])
def test_layernorm(ops, rows, D, chan):
    torch.manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda") * 3 + 0.5
    g = 1 + 0.2 * torch.randn(D, device="cuda")
    b = 0.1 * torch.randn(D, device="cuda")
    xd = x.double()
    mean = xd.mean(-1, keepdim=True)
    var = xd.var(-1, unbiased=False, keepdim=True)
    if chan:
        ref = (xd - mean) / (var.sqrt() + 1e-5) * g.double() + b.double()
    else:
        ref = (xd - mean) / (var + 1e-5).sqrt() * g.double() + b.double()
    out, s = ops.layernorm(x, g, b, chan_mode=chan, want_f32=True, want_split=True)
    assert_parity(out, ref, "layernorm fp32", rtol=1e-5)
    assert_parity(_unsplit(s), ref, "layernorm split", rtol=3e-5)


@pytest.mark.parametrize("B,L,heads", [(1, 197, 12), (3, 197, 12), (2, 5, 2), (2, 64, 1), (1, 224, 3),
                                        (4, 17, 2), (2, 130, 4), (40, 197, 12)])
@pytest.mark.parametrize("kernel", [2])
def test_vit_attention(ops, B, L, heads, kernel):
    torch.manual_seed(B * 1000 + L)
    W = heads * 64
    qkv = torch.randn(B * L, 3 * W, device="cuda") * 1.5
    s = ops.split(qkv)
    out = _unsplit(ops.vit_attention(s, B, L, heads, kernel=kernel))
    x = _unsplit(s).double().reshape(B, L, 3, heads, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(B * L, W)
    assert_parity(out, ref, f"vit attention kernel={kernel} B={B} L={L} h={heads}", rtol=1e-4)


def _f16f8_reference_planes(y, ops):
    """torch restatement of the activation encoding (csrc/split.cuh) of fp32 values y."""
    e_main, e_res, e_coarse = ops.ACT_EXP
    ym = y.float() * 2.0 ** e_main
    h = ym.clamp(-65504, 65504).half()
    l = ((ym - h.float()) * 2.0 ** e_res).clamp(-448, 448).to(torch.float8_e4m3fn)
    c = (y.float() * 2.0 ** e_coarse).clamp(-448, 448).to(torch.float8_e4m3fn)
    return h, l, c


@pytest.mark.parametrize("rows,D", [(1, 768), (197 * 3, 768), (1000, 256)])
def test_layernorm_f16f8_output(ops, rows, D):
    """The f16f8 output planes are exactly the encoding of the kernel's own fp32 output."""
    torch.manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda") * 3 + 0.5
    g = 1 + 0.2 * torch.randn(D, device="cuda")
    b = 0.1 * torch.randn(D, device="cuda")
    out, enc = ops.layernorm(x, g, b, want_f32=True, want_split=True, out_enc=1)
    h, l, c = enc.planes()
    eh, el, ec = _f16f8_reference_planes(out, ops)
    assert torch.equal(h, eh)
    assert torch.equal(l.view(torch.uint8), el.view(torch.uint8))
    assert torch.equal(c.view(torch.uint8), ec.view(torch.uint8))
    assert_parity(enc.decode(), out, "layernorm f16f8 decode", rtol=3e-5)


@pytest.mark.parametrize("B,L,heads", [(1, 197, 12), (3, 197, 4), (2, 5, 2), (2, 130, 4), (40, 197, 12)])
def test_vit_attention_f16f8_output(ops, B, L, heads):
    torch.manual_seed(B * 1000 + L)
    W = heads * 64
    s = ops.split(torch.randn(B * L, 3 * W, device="cuda") * 1.5)
    ref = _unsplit(ops.vit_attention(s, B, L, heads))          # bf16 hi/lo output of the same kernel
    enc = ops.vit_attention(s, B, L, heads, out_enc=1)
    assert_parity(enc.decode(), ref, f"vit attention f16f8 B={B} L={L} h={heads}", rtol=3e-5)
    h, l, c = enc.planes()
    # coarse plane = e4m3 of the value itself (within one e4m3 ulp of the decoded value)
    assert_parity(c.float(), ref.clamp(-448, 448), "coarse plane", rtol=7e-2)


@pytest.mark.parametrize("E,heads", [(256, 8), (128, 8), (64, 2)])
@pytest.mark.parametrize("axis", [0, 1])
def test_axial_attention(ops, E, heads, axis):
    torch.manual_seed(E + axis)
    S, n, l = 3, 32, 16
    qkv = torch.randn(S * n * l, 3 * E, device="cuda")
    out = _unsplit(ops.axial_attention(qkv, S, n, l, heads, axis))
    e = E // heads
    x = qkv.double().reshape(S, n, l, 3, heads, e)
    if axis == 0:   # sequences along n: batch (S, l, heads)
        q, k, v = (x[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))   # S l h n e
    else:           # sequences along l: batch (S, n, heads)
        q, k, v = (x[:, :, :, i].permute(0, 1, 3, 2, 4) for i in range(3))   # S n h l e
    p = torch.softmax(q @ k.transpose(-1, -2) * e ** -0.5, dim=-1)
    o = p @ v
    o = o.permute(0, 3, 1, 2, 4) if axis == 0 else o.permute(0, 1, 3, 2, 4)   # S n l h e
    ref = o.reshape(S * n * l, E)
    assert_parity(out, ref, f"axial attention E={E} axis={axis}", rtol=3e-5)


def test_errors_are_reported_not_thrown_across_the_abi(ops):
    from anomalyclip_b200._lib import AclipError
    x = torch.randn(4, 770, device="cuda")
    with pytest.raises(AclipError):
        ops.layernorm(x, torch.ones(770, device="cuda"), torch.zeros(770, device="cuda"))
