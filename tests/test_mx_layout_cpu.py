"""The f16mx tensor layout (anomalyclip_b200/csrc/mx.cuh) as the host-side view decodes it: plane
offsets, nibble order of the packed e2m1 planes and the 32 x 16-byte scale chunks that tcgen05.cp
copies into TMEM.  No GPU needed: the bytes are written here by a plain restatement of the header."""
import numpy as np
import torch

from anomalyclip_b200 import ops

E2M1 = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def test_f16mx_view_follows_the_documented_layout():
    rows, ld = 300, 192                      # 3 row blocks of 128 (the last one partial), 3 K atoms of 64
    t = ops.F16MX(rows, ld, torch.device("cpu"), exp=4)
    P = rows * ld
    blocks, atoms = (rows + 127) // 128, ld // 64
    assert t.buf.numel() == 3 * P + atoms * blocks * 512 and t.row_blocks == blocks
    rng = np.random.default_rng(0)
    buf = t.buf.numpy()
    h = rng.standard_normal((rows, ld)).astype(np.float16)
    buf[: 2 * P] = h.view(np.uint8).reshape(-1)
    codes = {name: rng.integers(0, 16, (rows, ld), dtype=np.uint8) for name in ("L4", "C4")}
    sfb = {name: rng.integers(100, 140, (rows, ld // 32), dtype=np.uint8) for name in ("L4", "C4")}
    for plane, name in enumerate(("L4", "C4")):
        c = codes[name]
        packed = (c[:, 0::2] | (c[:, 1::2] << 4)).astype(np.uint8)           # element k -> nibble k & 1 of byte k // 2
        buf[2 * P + plane * (P // 2): 2 * P + (plane + 1) * (P // 2)] = packed.reshape(-1)
        for m in range(rows):
            for kb in range(ld // 32):                                        # 32-value block kb of row m
                off = 3 * P + ((kb >> 1) * blocks + (m >> 7)) * 512 + (m & 31) * 16 + ((m >> 5) & 3) * 4 + (kb & 1)
                buf[off + 2 * plane] = sfb[name][m, kb]                       # bytes 0..1: L4 scales, 2..3: C4 scales
    got_h, got_l, got_c = t.planes()
    assert np.array_equal(got_h.numpy(), h)
    val = np.array(E2M1 + [-v for v in E2M1])
    for got, name in ((got_l, "L4"), (got_c, "C4")):
        want = val[codes[name]] * np.exp2(np.repeat(sfb[name].astype(np.float64), 32, axis=1) - 127.0)
        assert np.array_equal(got.numpy(), want), name
    # decode() = (H + L4) / 2^exp
    want = (h.astype(np.float64) + val[codes["L4"]] * np.exp2(np.repeat(sfb["L4"].astype(np.float64), 32, axis=1) - 127.0)) / 16.0
    assert np.array_equal(t.decode().numpy(), want)


def test_f16mx_needs_rows_of_64():
    import pytest
    with pytest.raises(ValueError):
        ops.F16MX(4, 96, torch.device("cpu"))
