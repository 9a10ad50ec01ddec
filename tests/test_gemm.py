"""Parity of the tcgen05 GEMM (C ABI: aclip_gemm / aclip_split_f32) against torch fp64/fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _unsplit(s):
    return s[0].float() + s[1].float()


@pytest.fixture(scope="module")
def ops():
    from anomalyclip_b200 import ops as _ops
    return _ops


def test_split_roundtrip(ops):
    torch.manual_seed(0)
    x = torch.randn(301, 530, device="cuda") * 3
    s = ops.split(x, ld_out=576)
    assert s.shape == (2, 301, 576)
    back = _unsplit(s)
    assert torch.all(back[:, 530:] == 0)
    # hi+lo carries 16 significand bits
    assert _rel(back[:, :530], x) < 2e-5
    assert torch.equal(s[0, :, :530], x.to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 768), (1000, 768, 768),
                                   (197 * 8, 2304, 768), (555, 128, 256), (129, 384, 3072),
                                   (64, 512, 768), (4096, 3072, 768)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_plain(ops, M, N, K, passes):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    out = ops.gemm(ops.split(a), ops.split(w), passes=passes)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    err = _rel(out, ref)
    print(f"gemm M={M} N={N} K={K} passes={passes}: rel err {err:.3e}")
    assert err < (3e-5 if passes == 3 else 1e-2)
    if passes == 1:  # must equal the bf16-rounded product, not just be "close"
        ref1 = a.to(torch.bfloat16).double() @ w.to(torch.bfloat16).double().T
        assert _rel(out, ref1) < 1e-5


def test_gemm_epilogue_bias_act_residual(ops):
    torch.manual_seed(1)
    M, N, K = 777, 768, 512
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    pre = a.double() @ w.double().T + bias.double()
    for act, fn in ((ops.ACT_NONE, lambda x: x),
                    (ops.ACT_QUICKGELU, lambda x: x * torch.sigmoid(1.702 * x)),
                    (ops.ACT_LEAKYRELU, lambda x: torch.nn.functional.leaky_relu(x, 0.01))):
        out = ops.gemm(ops.split(a), ops.split(w), bias=bias, act=act, residual=res)
        ref = fn(pre) + res.double()
        err = _rel(out, ref)
        print(f"epilogue act={act}: rel err {err:.3e}")
        assert err < 3e-5
    # in-place residual update x += a @ w^T + b
    x = res.clone()
    ops.gemm(ops.split(a), ops.split(w), bias=bias, residual=x, out_f32=x)
    assert _rel(x, pre + res.double()) < 3e-5


def test_gemm_split_output_and_rowmap(ops):
    torch.manual_seed(2)
    frames, g = 5, 196
    M, N, K = frames * g, 768, 768
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    pos = torch.randn(g, N, device="cuda")
    ref = (a.double() @ w.double().T).reshape(frames, g, N) + pos.double()
    # split output
    s = ops.gemm(ops.split(a), ops.split(w), want_split=True)
    assert _rel(_unsplit(s), ref.reshape(M, N) - pos.double().repeat(frames, 1)) < 3e-5
    # periodic residual (positional table) + row remap 196 -> 197 with offset 1
    out = torch.zeros(frames * 197, N, device="cuda")
    ops.gemm(ops.split(a), ops.split(w), residual=pos, res_mod=g, out_f32=out,
             row_map=(g, 197, 1))
    out = out.reshape(frames, 197, N)
    assert torch.all(out[:, 0] == 0)
    assert _rel(out[:, 1:], ref) < 3e-5


@pytest.mark.parametrize("S,Cin,Cout", [(1, 64, 256), (3, 256, 1024), (2, 1024, 256), (2, 128, 128)])
def test_conv3x3_implicit_gemm(ops, S, Cin, Cout):
    torch.manual_seed(3)
    H, W = 32, 16
    x = torch.randn(S, Cin, H, W, device="cuda")
    wt = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02
    b = torch.randn(Cout, device="cuda")
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(S * H * W, Cout)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().reshape(S * H * W, Cin)
    a = ops.split(x_nhwc)  # [2, S*H*W, Cin] == [2, S, H, W, Cin]
    wk = wt.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin)  # k = tap * Cin + c
    out = ops.gemm(a, ops.split(wk), bias=b, conv=(S, H, W, Cin))
    err = _rel(out, ref)
    print(f"conv3x3 S={S} {Cin}->{Cout}: rel err {err:.3e}")
    assert err < 3e-5


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 256, 768), (300, 512, 768), (1000, 768, 768),
                                   (197 * 8, 2304, 768), (129, 768, 3072), (4096, 3072, 768),
                                   (50432, 768, 768)])
@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_cta_pair_kernel(ops, M, N, K, passes):
    """The cta_group::2 kernel (256x256 tiles over two SMs) against fp64 and against the
    single-CTA kernel (identical accumulation order -> identical bits)."""
    torch.manual_seed(M + N + K + 1)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    sa, sw = ops.split(a), ops.split(w)
    out2 = ops.gemm(sa, sw, bias=bias, residual=res, act=ops.ACT_QUICKGELU, passes=passes, kernel=2)
    out1 = ops.gemm(sa, sw, bias=bias, residual=res, act=ops.ACT_QUICKGELU, passes=passes, kernel=1)
    torch.cuda.synchronize()
    pre = a.double() @ w.double().T + bias.double()
    ref = pre * torch.sigmoid(1.702 * pre) + res.double()
    err = _rel(out2, ref)
    print(f"pair gemm M={M} N={N} K={K} passes={passes}: rel err {err:.3e}")
    assert err < (3e-5 if passes == 3 else 1e-2)
    assert torch.equal(out1, out2)


def test_gemm_cta_pair_split_output_rowmap_and_conv(ops):
    torch.manual_seed(7)
    frames, g = 6, 196
    M, N, K = frames * g, 768, 768
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    pos = torch.randn(g, N, device="cuda")
    ref = (a.double() @ w.double().T).reshape(frames, g, N) + pos.double()
    out = torch.zeros(frames * 197, N, device="cuda")
    ops.gemm(ops.split(a), ops.split(w), residual=pos, res_mod=g, out_f32=out, row_map=(g, 197, 1),
             kernel=2)
    out = out.reshape(frames, 197, N)
    assert torch.all(out[:, 0] == 0)
    assert _rel(out[:, 1:], ref) < 3e-5
    s = ops.gemm(ops.split(a), ops.split(w), want_split=True, kernel=2)
    assert _rel(_unsplit(s), a.double() @ w.double().T) < 3e-5
    # implicit-GEMM conv on the pair kernel
    S, Cin, Cout, H, W = 3, 256, 1024, 32, 16
    x = torch.randn(S, Cin, H, W, device="cuda")
    wt = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02
    b = torch.randn(Cout, device="cuda")
    cref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1)
    cref = cref.permute(0, 2, 3, 1).reshape(S * H * W, Cout)
    xs = ops.split(x.permute(0, 2, 3, 1).contiguous().reshape(S * H * W, Cin))
    wk = ops.split(wt.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin))
    c2 = ops.gemm(xs, wk, bias=b, conv=(S, H, W, Cin), kernel=2)
    c1 = ops.gemm(xs, wk, bias=b, conv=(S, H, W, Cin), kernel=1)
    assert _rel(c2, cref) < 3e-5
    assert torch.equal(c1, c2)


def test_gemm_random_shapes(ops):
    """Ragged M, every legal N granule, K not a multiple of the 64-wide K block."""
    import random
    rnd = random.Random(1)
    for _ in range(25):
        M = rnd.choice([1, 2, 31, 127, 128, 129, 255, 257, 511, 700, 1023])
        N = 32 * rnd.randint(1, 20)
        K = 8 * rnd.randint(1, 90)
        torch.manual_seed(M * 7 + N * 3 + K)
        a = torch.randn(M, K, device="cuda")
        w = torch.randn(N, K, device="cuda") * 0.05
        bias = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")
        out = ops.gemm(ops.split(a), ops.split(w), bias=bias, residual=res)
        s = ops.gemm(ops.split(a), ops.split(w), bias=bias, want_split=True)
        ref = a.double() @ w.double().T + bias.double()
        assert _rel(out, ref + res.double()) < 3e-5, (M, N, K)
        assert _rel(_unsplit(s), ref) < 3e-5, (M, N, K)


# ---------------------------------------------------------------------------------------------
# f16f8 operands (passes = 2): fp16 main product + e4m3 cross terms, two pass-equivalents
# ---------------------------------------------------------------------------------------------
def _emulate_f16f8(x, e_main, e_res, e_coarse):
    """torch restatement of csrc/split.cuh:f16f8_pack2 -> (H fp16, L e4m3, C e4m3)."""
    xm = x.float() * 2.0 ** e_main
    h = xm.clamp(-65504, 65504).half()
    l = ((xm - h.float()) * 2.0 ** e_res).clamp(-448, 448).to(torch.float8_e4m3fn)
    c = (x.float() * 2.0 ** e_coarse).clamp(-448, 448).to(torch.float8_e4m3fn)
    return h, l, c


def test_encode_f16f8_planes(ops):
    torch.manual_seed(3)
    x = torch.randn(131, 200, device="cuda") * 2
    x[0, :4] = torch.tensor([5000.0, -7000.0, 1e-6, 0.0], device="cuda")  # saturation / tiny / zero
    t = ops.encode_f16f8(x, ld_out=208)
    h, l, c = t.planes()
    eh, el, ec = _emulate_f16f8(x, *ops.ACT_EXP)
    assert torch.equal(h[:, :200], eh) and torch.all(h[:, 200:] == 0)
    assert torch.equal(l[:, :200].view(torch.uint8), el.view(torch.uint8))
    assert torch.equal(c[:, :200].view(torch.uint8), ec.view(torch.uint8))
    ok = x.abs() < 4000
    assert _rel(t.decode()[:, :200][ok], x[ok]) < 3e-5
    w = torch.randn(64, 256, device="cuda") * 0.03
    tw = ops.encode_f16f8(w, weight=True)
    assert 2 ** 14 < float(w.abs().max()) * 2.0 ** tw.exp <= 2 ** 15
    assert _rel(tw.decode(e_res=ops.WGT_EXP_RES), w) < 3e-5


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 256, 768), (1000, 768, 768),
                                   (197 * 8, 2304, 768), (129, 768, 3072), (4096, 3072, 768),
                                   (50432, 768, 768)])
def test_gemm_f16f8(ops, M, N, K):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    out = ops.gemm(ops.encode_f16f8(a), ops.encode_f16f8(w, weight=True), passes=2)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T
    err = _rel(out, ref)
    out3 = ops.gemm(ops.split(a), ops.split(w), passes=3)
    print(f"gemm f16f8 M={M} N={N} K={K}: rel err {err:.3e} (bf16x3: {_rel(out3, ref):.3e})")
    assert err < 3e-5


def test_gemm_f16f8_epilogue_and_encoded_output(ops):
    torch.manual_seed(5)
    M, N, K = 777, 768, 512
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.05
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    ea, ew = ops.encode_f16f8(a), ops.encode_f16f8(w, weight=True)
    pre = a.double() @ w.double().T + bias.double()
    out = ops.gemm(ea, ew, bias=bias, act=ops.ACT_QUICKGELU, residual=res, passes=2)
    assert _rel(out, pre * torch.sigmoid(1.702 * pre) + res.double()) < 3e-5
    # hidden activations leave the epilogue already encoded for the next GEMM (c_fc -> c_proj)
    enc = ops.gemm(ea, ew, bias=bias, act=ops.ACT_QUICKGELU, passes=2, want_split=True, out_enc=1)
    f32 = ops.gemm(ea, ew, bias=bias, act=ops.ACT_QUICKGELU, passes=2)
    h, l, c = enc.planes()
    eh, el, ec = _emulate_f16f8(f32, *ops.ACT_EXP)
    assert torch.equal(h, eh)
    assert torch.equal(l.view(torch.uint8), el.view(torch.uint8))
    assert torch.equal(c.view(torch.uint8), ec.view(torch.uint8))
    # bf16 hi/lo output from f16f8 operands (in_proj feeds the attention kernel)
    s = ops.gemm(ea, ew, bias=bias, passes=2, want_split=True)
    assert _rel(_unsplit(s), pre) < 3e-5


@pytest.mark.parametrize("S,Cin,Cout", [(1, 64, 256), (3, 256, 1024), (2, 1024, 256), (9, 256, 256)])
def test_conv3x3_implicit_gemm_f16f8(ops, S, Cin, Cout):
    """The 3x3 "same" convolution over an f16f8 NHWC grid (5-D TMA maps for the fp16 plane and for
    the two e4m3 planes, hardware zero fill at the borders), LeakyReLU epilogue, f16f8 output."""
    torch.manual_seed(3)
    H, W = 32, 16
    x = torch.randn(S, Cin, H, W, device="cuda")
    wt = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02
    b = torch.randn(Cout, device="cuda")
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), b.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(S * H * W, Cout)
    a = ops.encode_f16f8(x.permute(0, 2, 3, 1).contiguous().reshape(S * H * W, Cin))
    wk = ops.encode_f16f8(wt.permute(0, 2, 3, 1).contiguous().reshape(Cout, 9 * Cin), weight=True)
    out = ops.gemm(a, wk, bias=b, conv=(S, H, W, Cin), passes=2)
    err = _rel(out, ref)
    print(f"conv3x3 f16f8 S={S} {Cin}->{Cout}: rel err {err:.3e}")
    assert err < 3e-5
    enc = ops.gemm(a, wk, bias=b, act=ops.ACT_LEAKYRELU, conv=(S, H, W, Cin), passes=2,
                   want_split=True, out_enc=1)
    assert _rel(enc.decode(), torch.nn.functional.leaky_relu(ref, 0.01)) < 3e-5


@pytest.mark.parametrize("M,N,K", [(512, 256, 9216), (64, 32, 64), (130, 96, 200), (1000, 768, 768),
                                   (515, 256, 512)])
def test_small_tile_is_bit_identical_to_the_128_row_tile(ops, M, N, K):
    """The 64 x 32 tile of the small problems (tile=2) against the 128-row tiles (tile=1): the
    accumulation order along K is per element, so every output bit must agree -- split-bf16 x3 and
    fp16 one-pass operands, bias / activation / residual / row remap / encoded outputs, ragged M."""
    torch.manual_seed(M + 7 * N + K)
    Kp = (K + 7) // 8 * 8
    a = torch.randn(M, Kp, device="cuda")
    w = torch.randn(N, Kp, device="cuda") * 0.05
    a[:, K:] = 0
    w[:, K:] = 0
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    sa, sw = ops.split(a), ops.split(w)
    big = ops.gemm(sa, sw, bias=bias, act=ops.ACT_QUICKGELU, residual=res, kernel=1, tile=1)
    small = ops.gemm(sa, sw, bias=bias, act=ops.ACT_QUICKGELU, residual=res, kernel=1, tile=2)
    assert torch.equal(big, small)
    z = a.double() @ w.double().T + bias.double()
    assert _rel(small, z * torch.sigmoid(1.702 * z) + res.double()) < 5e-5
    sb = ops.gemm(sa, sw, want_split=True, kernel=1, tile=1)
    ss = ops.gemm(sa, sw, want_split=True, kernel=1, tile=2)
    assert torch.equal(sb, ss)
    # row remap (groups of 16 rows -> pitch 17, offset 1), in place over a prefilled output
    if M % 16 == 0:
        o1 = torch.full((M // 16 * 17, N), 7.0, device="cuda")
        o2 = o1.clone()
        ops.gemm(sa, sw, out_f32=o1, row_map=(16, 17, 1), kernel=1, tile=1)
        ops.gemm(sa, sw, out_f32=o2, row_map=(16, 17, 1), kernel=1, tile=2)
        assert torch.equal(o1, o2) and torch.all(o2[::17] == 7.0)
    # fp16 one-pass operands, fp32 and fp16-plane outputs
    ea, ew = ops.encode_f16(a), ops.encode_f16f8(w, weight=True)
    f1 = ops.gemm(ea, ew, bias=bias, residual=res, passes=4, kernel=1, tile=1)
    f2 = ops.gemm(ea, ew, bias=bias, residual=res, passes=4, kernel=1, tile=2)
    assert torch.equal(f1, f2)
    h1 = ops.gemm(ea, ew, bias=bias, act=ops.ACT_LEAKYRELU, passes=4, want_split=True, out_enc=2, kernel=1, tile=1)
    h2 = ops.gemm(ea, ew, bias=bias, act=ops.ACT_LEAKYRELU, passes=4, want_split=True, out_enc=2, kernel=1, tile=2)
    assert torch.equal(h1, h2)


@pytest.mark.parametrize("S,Cin,Cout", [(1, 1024, 256), (1, 256, 1024), (3, 128, 96)])
def test_small_tile_conv3x3_is_bit_identical(ops, S, Cin, Cout):
    torch.manual_seed(5)
    H, W = 32, 16
    x = torch.randn(S * H * W, Cin, device="cuda")
    wk = torch.randn(Cout, 9 * Cin, device="cuda") * 0.02
    b = torch.randn(Cout, device="cuda")
    res = torch.randn(S * H * W, Cout, device="cuda")
    sa, sw = ops.split(x), ops.split(wk)
    c1 = ops.gemm(sa, sw, bias=b, residual=res, conv=(S, H, W, Cin), kernel=1, tile=1)
    c2 = ops.gemm(sa, sw, bias=b, residual=res, conv=(S, H, W, Cin), kernel=1, tile=2)
    assert torch.equal(c1, c2)
    ref = torch.nn.functional.conv2d(x.reshape(S, H, W, Cin).permute(0, 3, 1, 2).double(),
                                     wk.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2).double(), b.double(),
                                     padding=1).permute(0, 2, 3, 1).reshape(S * H * W, Cout) + res.double()
    assert _rel(c2, ref) < 3e-5
    ea, ew = ops.encode_f16(x), ops.encode_f16f8(wk, weight=True)
    f1 = ops.gemm(ea, ew, bias=b, residual=res, conv=(S, H, W, Cin), passes=4, kernel=1, tile=1)
    f2 = ops.gemm(ea, ew, bias=b, residual=res, conv=(S, H, W, Cin), passes=4, kernel=1, tile=2)
    assert torch.equal(f1, f2)
    # and the automatic choice (small problems pick the small tile) agrees as well
    assert torch.equal(ops.gemm(ea, ew, bias=b, residual=res, conv=(S, H, W, Cin), passes=4), f1)
