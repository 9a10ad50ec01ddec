"""The f16mx operand encoding (csrc/mx.cuh: fp16 plane + two MXFP4 planes with UE8M0 block scales)
and the GEMM that multiplies it (aclip_gemm passes = 7)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from anomalyclip_b200 import ops as _ops
    return _ops


def _emulate_planes(x, e_main):
    """torch restatement of mx_pack32: H, L4, C4 as fp64 values in units of 2^e_main."""
    grid = torch.tensor([0, .5, 1, 1.5, 2, 3, 4, 6], dtype=torch.float64, device=x.device)
    xs = (x.float() * 2.0 ** e_main)
    h = xs.half().double()
    def mx(v):
        r, k = v.shape
        b = v.reshape(r, k // 32, 32)
        amax = b.abs().amax(-1, keepdim=True)
        t = (amax.float() * torch.tensor(1.0 / 6.0, dtype=torch.float32, device=x.device))
        bits = t.view(torch.int32)
        e = (bits >> 23) + ((bits & 0x7fffff) != 0).int()
        s = torch.exp2(e.double() - 127.0)
        q = (b / s).float().double()
        a = q.abs().clamp(max=6.0)
        # round to nearest, ties to even mantissa (cvt.rn): candidates sorted, pick nearest; ties -> even index
        d = (a.unsqueeze(-1) - grid).abs()
        idx = d.argmin(-1)
        lo = (idx - 1).clamp(min=0)
        tie_lo = (d.gather(-1, lo.unsqueeze(-1)).squeeze(-1) == d.gather(-1, idx.unsqueeze(-1)).squeeze(-1)) & (lo != idx)
        idx = torch.where(tie_lo & (lo % 2 == 0), lo, idx)
        hi = (idx + 1).clamp(max=7)
        tie_hi = (d.gather(-1, hi.unsqueeze(-1)).squeeze(-1) == d.gather(-1, idx.unsqueeze(-1)).squeeze(-1)) & (hi != idx)
        idx = torch.where(tie_hi & (hi % 2 == 0), hi, idx)
        return (torch.sign(q) * grid[idx] * s).reshape(r, k)
    return h, mx((xs.double() - h).float().double()), mx(xs.double())


@pytest.mark.parametrize("rows,cols", [(128, 64), (300, 768), (197, 3072), (1, 100)])
def test_encode_f16mx_planes(ops, rows, cols):
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda") * torch.logspace(-2, 1, cols, device="cuda")
    enc = ops.encode_f16mx(x)
    ld = enc.ld
    xp = torch.zeros(rows, ld, device="cuda")
    xp[:, :cols] = x
    h, l4, c4 = enc.planes()
    eh, el4, ec4 = _emulate_planes(xp, enc.exp)
    assert torch.equal(h.double(), eh)
    # scale bytes and elements follow the restatement (ties of the e2m1 rounding may differ by one grid step)
    for got, want, name in ((l4, el4, "L4"), (c4, ec4, "C4")):
        bad = (got != want).double().mean().item()
        assert bad < 2e-3, f"{name}: {bad:.2e} of the elements differ from the restatement"
    # H + L4 reconstructs the value to a quarter of an fp16 ulp of its 32-block's residual range
    rec = enc.decode()[:, :cols]
    rel = ((rec - x.double()).norm() / x.double().norm()).item()
    assert rel < 1.2e-4, rel
    assert ((c4 * 2.0 ** -enc.exp)[:, :cols] - x.double()).norm() / x.double().norm() < 0.2
    w = ops.encode_f16mx(x * 1e-3, weight=True)
    assert ((w.decode()[:, :cols] - (x * 1e-3).double()).norm() / (x * 1e-3).double().norm()).item() < 1.2e-4


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _effective(enc):
    """What the three MMA products of the GEMM see: (H, L4, C4) as fp64 in real units."""
    h, l4, c4 = enc.planes()
    s = 2.0 ** -enc.exp
    return h.double() * s, l4 * s, c4 * s


@pytest.mark.parametrize("M,N,K", [(256, 192, 64), (256, 384, 256), (300, 192, 768), (1000, 768, 3072),
                                   (197 * 8, 3072, 768), (129, 576, 128)])
def test_gemm_f16mx(ops, M, N, K):
    """x_H w_H + x_L4 w_C4 + x_C4 w_L4 on the CTA-pair kernel: against the same three products in
    fp64 from the decoded planes (tight: checks the operand/scale-factor plumbing), and against the
    fp64 product of the inputs (the accuracy the mode buys: ~5e-5, fp16 alone ~3e-4)."""
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda") * torch.logspace(-1, 1, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    ea, ew = ops.encode_f16mx(a), ops.encode_f16mx(w, weight=True)
    out = ops.gemm(ea, ew, passes=7)
    torch.cuda.synchronize()
    ah, al, ac = _effective(ea)
    wh, wl, wc = _effective(ew)
    exact_planes = ah @ wh.T + al @ wc.T + ac @ wl.T
    e1 = _rel(out, exact_planes)
    e2 = _rel(out, a.double() @ w.double().T)
    eh = _rel(ah @ wh.T, a.double() @ w.double().T)
    print(f"f16mx gemm M={M} N={N} K={K}: vs decoded planes {e1:.2e}, vs fp64 {e2:.2e} (fp16 planes alone {eh:.2e})")
    assert e1 < 3e-6
    assert e2 < 1e-4 and e2 < 0.4 * eh


def test_gemm_f16mx_epilogues(ops):
    """bias + QuickGELU into an f16mx output (c_fc -> c_proj's operand), and bias + fp32 residual in
    place (c_proj), ragged M."""
    torch.manual_seed(5)
    M, N, K = 777, 768, 384
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    ea, ew = ops.encode_f16mx(a), ops.encode_f16mx(w, weight=True)
    z = a.double() @ w.double().T + bias.double()
    gelu = z * torch.sigmoid(1.702 * z)
    enc = ops.gemm(ea, ew, bias=bias, act=ops.ACT_QUICKGELU, passes=7, want_split=True, out_enc=3)
    assert isinstance(enc, ops.F16MX) and enc.rows == M and enc.ld == N
    assert _rel(enc.decode(), gelu) < 1.5e-4
    # the encoded output is what the packer would make of the fp32 result of the same GEMM
    f32 = ops.gemm(ea, ew, bias=bias, act=ops.ACT_QUICKGELU, passes=7)
    ref = ops.encode_f16mx(f32)
    h1, l1, c1 = enc.planes()
    h2, l2, c2 = ref.planes()
    assert torch.equal(h1, h2) and torch.equal(l1, l2) and torch.equal(c1, c2)
    out = res.clone()
    ops.gemm(ea, ew, bias=bias, residual=out, out_f32=out, passes=7)
    assert _rel(out, z + res.double()) < 1e-4
    # and it chains: c_fc -> c_proj on the encoded hidden
    w2 = torch.randn(192, N, device="cuda") * 0.05
    y = ops.gemm(enc, ops.encode_f16mx(w2, weight=True), passes=7)
    assert _rel(y, gelu @ w2.double().T) < 1.5e-4


def test_layernorm_f16mx_output(ops):
    """LayerNorm straight into the f16mx encoding = the packer applied to the fp32 LayerNorm output."""
    torch.manual_seed(9)
    rows, D = 333, 768
    x = torch.randn(rows, D, device="cuda") * 3 + 0.5
    g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    enc = ops.layernorm(x, g, b, want_f32=False, want_split=True, out_enc=3)
    # against the packer applied to the fp32 LayerNorm output: the two kernels sum a row in different
    # orders, so an fp16 rounding may flip here and there, no more
    ref = ops.encode_f16mx(ops.layernorm(x, g, b))
    h, l4, c4 = enc.planes()
    rh, rl4, rc4 = ref.planes()
    assert (h != rh).double().mean().item() < 2e-3
    assert (c4 != rc4).double().mean().item() < 2e-3
    assert _rel(enc.decode(), ref.decode()) < 2e-5
    want = torch.nn.functional.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-5)
    assert _rel(enc.decode(), want) < 1.5e-4


def test_vit_b16_mx_mode():
    """passes=7: mode 5 with the MLP pair of every block on f16mx operands (LayerNorm and the c_fc
    epilogue write the encoding, c_fc and c_proj multiply it)."""
    from anomalyclip_b200.engine import PackedVit, VitEncoder
    from oracle import anomalyclip_oracle as oracle
    from tests.util_weights import make_frames_u8, make_vit_weights, normalise_frames
    from tests.parity import assert_parity
    sd = make_vit_weights()
    packed = PackedVit(sd, torch.device("cuda"), passes=7)
    enc7 = VitEncoder(packed, 256, 7)
    torch.manual_seed(5)
    frames = torch.randn(3, 3, 224, 224)
    ref = oracle.vit_forward(sd, frames)
    e7 = assert_parity(enc7(frames.cuda()), ref, "ViT-B/16 features, MXFP4 cross-term mode", rtol=3e-4)
    e5 = assert_parity(VitEncoder(packed, 256, 5)(frames.cuda()), ref, "ViT-B/16 features, mixed mode", rtol=3e-4)
    u8 = make_frames_u8(5, seed=3)
    e7u = assert_parity(enc7(u8.cuda()), oracle.vit_forward(sd, normalise_frames(u8)),
                        "ViT-B/16 features from uint8 frames, MXFP4 cross-term mode", rtol=3e-4)
    print(f"mode 7 rel-L2 vs oracle: {e7:.3e} (fp32 frames), {e7u:.3e} (uint8 frames); mode 5: {e5:.3e}")
    # micro-batching (ragged last micro-batch: 5 = 2 + 2 + 1 frames) does not change a bit
    assert torch.equal(enc7(u8.cuda()), VitEncoder(packed, micro_batch=2, passes=7)(u8.cuda()))
