"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host logic
(partitioning, row gather under gloo, EOT recovery, state_dict names, no CPU fallback)."""
import os
import re
import socket
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from anomalyclip_b200 import _lib
    header = (ROOT / "include" / "aclip_b200.h").read_text()
    declared = set(re.findall(r"\b(aclip_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.aclip_version() >= 100


def test_invalid_arguments_return_status_not_crash():
    from anomalyclip_b200 import _lib
    lib = _lib.load()
    assert lib.aclip_gemm(None, None) == -1
    assert b"NULL" in lib.aclip_last_error()
    assert lib.aclip_vit_forward(None, None, 0, 0, 0, None, None, None, None, 0, 3, None) == -1
    assert lib.aclip_temporal_forward(None, None, 0, 1, None, None, None, None, 0, 3, None) == -1
    assert lib.aclip_vit_workspace_bytes(None, 4) == 0
    assert lib.aclip_encode_f16f8(None, 4, 16, 16, None, 16, 64, 4, 7, 0, None) == -1
    # a passes = 2 GEMM without its accumulator scale / with N % 256 != 0 is refused before any launch
    g = _lib.GemmArgs()
    g.a, g.w, g.M, g.N, g.K, g.lda, g.ldw = 1024, 2048, 128, 256, 64, 64, 64
    g.a_plane_stride, g.w_plane_stride, g.passes, g.out_f32, g.ldc = 128 * 64, 256 * 64, 2, 4096, 256
    import ctypes as C
    assert lib.aclip_gemm(C.byref(g), None) == -1 and b"out_scale" in lib.aclip_last_error()
    g.out_scale, g.N = 1.0, 128
    assert lib.aclip_gemm(C.byref(g), None) == -1 and b"CTA-pair" in lib.aclip_last_error()


def test_no_cpu_fallback():
    from anomalyclip_b200._lib import AclipError
    from anomalyclip_b200.engine import PackedVit
    from tests.util_weights import make_vit_weights
    with pytest.raises(AclipError, match="no CPU"):
        PackedVit(make_vit_weights(layers=1), torch.device("cpu"))


def test_product_code_never_imports_the_oracle():
    for path in (ROOT / "anomalyclip_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    for path in (ROOT / "src").rglob("*.py"):
        assert "oracle" not in path.read_text(), path


def test_partition_is_balanced_and_contiguous():
    from anomalyclip_b200.distributed import partition
    assert partition(4, 8) == [(0, 1), (1, 1), (2, 1), (3, 1), (4, 0), (4, 0), (4, 0), (4, 0)]
    assert partition(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    for units in range(0, 40):
        for world in (1, 2, 3, 8):
            blocks = partition(units, world)
            assert sum(c for _, c in blocks) == units
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, units, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anomalyclip_b200.distributed import run_sharded
    unit_rows, width = 4, 3

    def compute(start, count):  # stands in for the GPU path: row r of unit u is [u, r, rank]
        u = torch.arange(start, start + count).repeat_interleave(unit_rows)
        r = torch.arange(unit_rows).repeat(count)
        return torch.stack((u, r, torch.full_like(u, rank)), dim=1).float()

    rows = run_sharded(units, unit_rows, compute)
    q.put((rank, rows.tolist()))  # plain lists: no shared-memory handles to outlive the worker
    dist.destroy_process_group()


@pytest.mark.parametrize("units", [4, 5, 1])
def test_sharded_run_gathers_all_rows_in_unit_order_gloo(units):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {r: torch.tensor(v).reshape(-1, 3) for r, v in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(got[0], got[1])
    rows = got[0]
    assert rows.shape == (units * 4, 3)
    assert rows[:, 0].tolist() == [float(u) for u in range(units) for _ in range(4)]
    assert rows[:, 1].tolist() == [float(r) for _ in range(units) for r in range(4)]
    first = (units + 1) // 2  # rank 0 owns the first ceil(units/2) units
    assert rows[: first * 4, 2].eq(0).all() and rows[first * 4:, 2].eq(1).all()


def _mode_worker(rank, world, port, modes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from anomalyclip_b200.distributed import agree_on_mode
    q.put((rank, agree_on_mode(modes[rank], torch.device("cpu"))))
    dist.destroy_process_group()


@pytest.mark.parametrize("modes,agreed", [((4, 4), 4), ((4, 5), 5), ((5, 2), 2), ((2, 4), 2), ((7, 5), 5), ((7, 7), 7)])
def test_ranks_agree_on_the_most_conservative_operand_mode_gloo(modes, agreed):
    """Ranks calibrate an "auto" encoder on their own frames; everybody then runs the most conservative
    of the selected modes (4 fastest, then 7, 5, 2), so a sharded video is scored in ONE mode."""
    from anomalyclip_b200.distributed import agree_on_mode
    assert agree_on_mode(5, torch.device("cpu")) == 5          # no process group: unchanged
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mode_worker, args=(r, world, port, modes, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: agreed, 1: agreed}


def test_operand_mode_bookkeeping_needs_no_gpu():
    """The host-side mode table: which `passes` values share the fp16-based weight pack, what "auto"
    tries and in which order, how the image encoder's mode maps to the temporal stage's."""
    from anomalyclip_b200 import engine
    assert set(engine.FP16_PACKED_MODES) == {2, 4, 5, 6, 7, "auto"}
    assert engine.AUTO_CANDIDATES == (5,)         # 7, 6 and 4 can flip a class index at a reference tie: opt-in only
    scorer = engine.TemporalScorer.__new__(engine.TemporalScorer)
    for given, mapped in ((3, 3), (2, 2), (4, 4), (5, "auto"), (6, "auto"), (7, "auto"), ("auto", "auto")):
        engine.TemporalScorer.__init__(scorer, packed=None, passes=given)
        assert scorer.passes == mapped and (scorer.mode is None) == (mapped == "auto")


def test_eot_positions_without_tokenizer():
    from anomalyclip_b200.models import eot_positions
    torch.manual_seed(0)
    n_cls, n_ctx, dim, ctx_len = 3, 8, 16, 77
    emb = torch.randn(50, dim)
    name_lens = [1, 3, 2]
    tokens = torch.zeros(n_cls, ctx_len, dtype=torch.long)
    for i, nl in enumerate(name_lens):  # SOS X*8 name... '.' EOT 0 0 0 ...
        seq = [48] + [5] * n_ctx + list(range(10, 10 + nl)) + [7, 49]
        tokens[i, : len(seq)] = torch.tensor(seq)
    suffix = emb[tokens][:, 1 + n_ctx:, :]
    assert eot_positions(suffix, emb[0], n_ctx).tolist() == tokens.argmax(-1).tolist()


def test_mirror_state_dict_has_the_reference_key_names():
    from anomalyclip_b200.models import AnomalyCLIP
    from tests.util_weights import PRESETS, make_state_dict
    cfg = PRESETS["shanghaitech"]
    net = AnomalyCLIP(arch="ViT-B/16", classnames=[f"c{i}" for i in range(cfg.num_classes)],
                      emb_size=cfg.emb_size, depth=cfg.depth, heads=cfg.heads, dim_heads=None,
                      num_segments=32, seg_length=16, concat_features=True, normal_id=cfg.normal_id,
                      stride=1, load_from_features=True, ncrops=1, n_ctx=8)
    keys = set(net.state_dict())
    synth = make_state_dict(cfg, with_vit=True)
    assert set(synth) <= keys, sorted(set(synth) - keys)[:5]
    for k, v in synth.items():
        assert tuple(net.state_dict()[k].shape) == tuple(v.shape), k
    for k in ("prompt_learner.ctx", "prompt_learner.token_prefix", "prompt_learner.token_suffix",
              "token_embedding.weight", "text_encoder.positional_embedding",
              "text_encoder.text_projection", "text_encoder.ln_final.weight",
              "text_encoder.transformer.resblocks.11.attn.in_proj_weight",
              "image_encoder.transformer.resblocks.0.mlp.c_fc.weight",
              "temporal_model.axial_attn.layers.blocks.3.g.net.3.weight",
              "temporal_model.axial_attn.layers.blocks.0.f.net.fn.fn.to_kv.weight"):
        assert k in keys, k
    assert net.state_dict()["prompt_learner.token_suffix"].shape == (cfg.num_classes, 68, 512)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 1, 512, 512), None, torch.zeros(512), 1, False)


def test_target_paths_of_the_reference_configs_resolve():
    import importlib
    for target in ("src.models.anomaly_clip_module.AnomalyCLIPModule",
                   "src.models.components.anomaly_clip.AnomalyCLIP",
                   "src.models.components.selector_model.SelectorModel",
                   "src.models.components.temporal_model.TemporalModel",
                   "src.models.components.classification_head.ClassificationHead"):
        mod, name = target.rsplit(".", 1)
        assert hasattr(importlib.import_module(mod), name), target


def test_text_tower_matches_reference_text_encoder():
    """The stock-PyTorch text tower of the mirror (evaluated once per checkpoint) against the
    reference TextEncoder's own output, with the EOT positions recovered without the tokenizer."""
    import numpy as np
    from anomalyclip_b200.models import PromptLearner, TextEncoder, eot_positions
    z = np.load(ROOT / "tests" / "golden" / "text.npz")
    w = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    n_cls, n_ctx, ctx_len, vocab, width, heads, layers, embed = (int(v) for v in z["cfg"])
    enc = TextEncoder(width, layers, heads, ctx_len, embed)
    missing, unexpected = enc.load_state_dict(
        {k[len("text_encoder."):]: v for k, v in w.items() if k.startswith("text_encoder.")}, strict=True)
    pl = PromptLearner(n_cls, n_ctx, width, ctx_len, shared_context=False)
    pl.load_state_dict({k[len("prompt_learner."):]: v for k, v in w.items() if k.startswith("prompt_learner.")})
    eot = eot_positions(pl.token_suffix, w["token_embedding.weight"][0], n_ctx)
    assert eot.tolist() == torch.from_numpy(z["tokens"]).argmax(-1).tolist()
    with torch.no_grad():
        out = enc(pl(), eot)
    torch.testing.assert_close(out, torch.from_numpy(z["out"]), rtol=1e-4, atol=1e-5)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs) emits exactly one JSON line with
    the contract's keys; everything else goes to stderr."""
    import json
    import subprocess
    import sys
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_text_features_are_computed_once_from_checkpoint_buffers():
    """AnomalyCLIP.get_text_features: prompts from prompt_learner buffers, EOT positions recovered
    from token_suffix, result cached until the text-side parameters change (the reference
    recomputes it on every forward, anomaly_clip.py:136)."""
    from anomalyclip_b200.models import AnomalyCLIP
    torch.manual_seed(0)
    C = 7
    net = AnomalyCLIP(arch="ViT-B/16", classnames=[f"c{i}" for i in range(C)], emb_size=128, depth=1,
                      heads=8, dim_heads=None, num_segments=32, seg_length=16, concat_features=False,
                      normal_id=4, stride=1, load_from_features=True, ncrops=1, n_ctx=8).eval()
    emb = net.token_embedding.weight.detach()
    tokens = torch.zeros(C, 77, dtype=torch.long)
    for i in range(C):   # SOS X*8 <name tokens> . EOT pad...
        seq = [49406] + [343] * 8 + [1000 + i + j for j in range(1 + i % 3)] + [269, 49407]
        tokens[i, : len(seq)] = torch.tensor(seq)
    full = emb[tokens]
    with torch.no_grad():
        net.prompt_learner.token_prefix.copy_(full[:, :1])
        net.prompt_learner.token_suffix.copy_(full[:, 9:])
    a = net.get_text_features()
    assert a.shape == (C, 512) and torch.isfinite(a).all()
    with torch.no_grad():
        ref = net.text_encoder(net.prompt_learner(), tokens.argmax(-1))
    torch.testing.assert_close(a, ref)
    assert net.get_text_features() is a                      # cached
    with torch.no_grad():
        net.prompt_learner.ctx.add_(0.01)
    b = net.get_text_features()
    assert b is not a and not torch.allclose(a, b)           # recomputed after the context changed


def test_f16f8_encoding_identity_on_the_host():
    """The exponent rules of the f16f8 operand encoding (csrc/split.cuh, ops.weight_exponent),
    restated in torch: x_H w_H + x_L w_C + x_C w_L = 2^(ex + ew) x w to ~1e-5, the weight's main
    plane stays finite, and the coarse / residual planes stay inside e4m3's range."""
    from anomalyclip_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(96, 256) * 1.5
    for scale in (0.02, 1.0, 37.0):
        w = torch.randn(64, 256) * scale
        ew = ops.weight_exponent(w)
        assert 2 ** 14 < float(w.abs().max()) * 2.0 ** ew <= 2 ** 15
        ex, rx, cx = ops.ACT_EXP
        rw, cw = ops.WGT_EXP_RES, ew - rx
        assert cx == ex - rw            # both cross terms carry 2^(ex + ew), like the main product

        def planes(v, e, r, c):
            vm = v.double() * 2.0 ** e
            h = vm.float().clamp(-65504, 65504).half().double()
            res, coarse = (vm - h) * 2.0 ** r, v.double() * 2.0 ** c
            assert res.abs().max() <= 448 and coarse.abs().max() <= 448
            return h, res.float().to(torch.float8_e4m3fn).double(), coarse.float().to(torch.float8_e4m3fn).double()

        xh, xl, xc = planes(x, ex, rx, cx)
        wh, wl, wc = planes(w, ew, rw, cw)
        acc = (xh @ wh.T + xl @ wc.T + xc @ wl.T) * 2.0 ** -(ex + ew)
        ref = x.double() @ w.double().T
        err = ((acc - ref).norm() / ref.norm()).item()
        assert err < 3e-5, err
        # without the cross terms the same product is only fp16-accurate
        err16 = ((xh @ wh.T * 2.0 ** -(ex + ew) - ref).norm() / ref.norm()).item()
        assert err16 > 10 * err


def test_f16f8_emulated_through_a_small_vit_stays_fp32_faithful():
    """End-to-end numerics of the f16f8 GEMM operands without a GPU: every linear layer of a small
    ViT (oracle arithmetic) is replaced by the emulated encoding (scripts/numerics_f16f8.py); the
    features stay within 1e-4 of the fp32 oracle (bar 1e-3) while plain fp16 / bf16 operands do not
    get anywhere near that."""
    import importlib.util
    import torch.nn.functional as F
    from anomalyclip_b200 import synthetic as syn
    from oracle import anomalyclip_oracle as oracle
    spec = importlib.util.spec_from_file_location("numerics_f16f8", ROOT / "scripts" / "numerics_f16f8.py")
    study = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(study)
    w = syn.make_vit_weights(width=256, layers=3, patch=16, resolution=64, output_dim=256, seed=5)
    torch.manual_seed(1)
    frames = torch.randn(4, 3, 64, 64)
    real = F.linear
    errs = {}
    with torch.no_grad():
        ref = oracle.vit_forward(w, frames).double()
        for mode in ("bf16x1", "fp16x1", "bf16x3", "f16f8"):
            F.linear = study.make_linear(mode)
            try:
                out = oracle.vit_forward(w, frames).double()
            finally:
                F.linear = real
            errs[mode] = ((out - ref).norm() / ref.norm()).item()
    assert errs["f16f8"] < 1e-4 and errs["bf16x3"] < 1e-4, errs
    assert errs["fp16x1"] > 5 * errs["f16f8"] and errs["bf16x1"] > 50 * errs["f16f8"], errs


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/aclip_b200.h must compile as C99 on its own."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    res = subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror",
                          str(ROOT / "include" / "aclip_b200.h")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_training_side_targets_resolve_and_the_schedule_is_the_reference_curve():
    """`loss._target_` and `scheduler._target_` of configs/model/*.yaml must import (Hydra builds
    them for an evaluation run too); the schedule is checked against the reference's own class when
    /root/reference is present."""
    import importlib
    import importlib.util
    import math
    import pickle
    Loss = getattr(importlib.import_module("src.models.components.loss"), "ComputeLoss")
    Sched = getattr(importlib.import_module("src.models.components.scheduler"), "WarmupCosineAnnealingLR")
    loss = Loss(normal_id=7, num_topk=3, lambda_dir_abn=1.0, lambda_dir_nor=1.0, lambda_topk_abn=1.0,
                lambda_bottomk_abn=1.0, lambda_topk_nor=1.0, lambda_smooth=8e-4, lambda_sparse=8e-3,
                frames_per_segment=16, num_segments=32)
    assert pickle.loads(pickle.dumps(loss)).lambda_smooth == 8e-4        # hyper_parameters round-trip
    with pytest.raises(NotImplementedError):
        loss(None)

    def curve(cls):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([{"params": [p], "lr": 0.1}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 0.02}])
        sched = cls(opt, total_epoch=50, warmup_epochs=5)
        out = []
        for _ in range(60):
            out.append([g["lr"] for g in opt.param_groups])
            opt.step()
            sched.step()
        return out
    mine = curve(Sched)
    assert mine[0] == [0.0, 0.0] and abs(mine[5][0] - 0.1) < 1e-12 and abs(mine[55][0]) < 1e-12
    assert abs(mine[27][0] - 0.1 * (1 + math.cos(math.pi * 22 / 45)) / 2) < 1e-12
    ref_file = Path("/root/reference/src/models/components/scheduler.py")
    if ref_file.exists():
        spec = importlib.util.spec_from_file_location("_ref_scheduler", ref_file)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        want = curve(ref.WarmupCosineAnnealingLR)
        assert all(abs(a - b) < 1e-12 for x, y in zip(mine, want) for a, b in zip(x, y))
