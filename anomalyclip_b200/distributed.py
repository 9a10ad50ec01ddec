"""Multi-GPU plumbing of the path: sub-videos shard across ranks with no data-path collective;
the only exchange is ONE all-gather of the per-frame result rows [score | class_probs] at the end
(SURVEY 8e).  The reference has no working multi-GPU inference (its test_step is
`@rank_zero_only`, anomaly_clip_module.py:458), so this is new surface, kept minimal.

One process per GPU, `torch.distributed` with the NCCL backend on the B200 box (NVLink 5 /
NVSwitch); the same code runs under gloo on CPU for the world_size-2 tests.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(num_units: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced (start, count) blocks of sub-videos, one per rank; the first
    `num_units % world_size` ranks take one extra unit."""
    if num_units < 0 or world_size < 1:
        raise ValueError("partition: num_units >= 0 and world_size >= 1 required")
    base, extra = divmod(num_units, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def gather_rows(local: torch.Tensor, counts: Sequence[int], group: Optional[dist.ProcessGroup] = None
                ) -> torch.Tensor:
    """All-gather of row blocks of (possibly) different heights: `local` is this rank's
    [counts[rank], width] block; returns the concatenation [sum(counts), width] on every rank.
    A single collective: blocks are padded to the tallest one."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(counts) != world or local.shape[0] != counts[rank]:
        raise ValueError("gather_rows: counts do not describe the local block")
    tallest = max(counts)
    if tallest == 0:
        return local
    width = local.shape[1:]
    send = local
    if local.shape[0] != tallest:
        send = local.new_zeros((tallest, *width))
        send[: local.shape[0]] = local
    send = send.contiguous()
    recv = send.new_empty((world * tallest, *width))
    dist.all_gather_into_tensor(recv, send, group=group) if _has_flat_gather(send) else \
        _gather_list(recv, send, world, tallest, group)
    if all(c == tallest for c in counts):
        return recv
    return torch.cat([recv[r * tallest: r * tallest + c] for r, c in enumerate(counts)], dim=0)


def _has_flat_gather(t: torch.Tensor) -> bool:
    return t.is_cuda  # NCCL: one flat all-gather; gloo (CPU tests): list form


def _gather_list(recv, send, world, tallest, group) -> None:
    parts = [recv[r * tallest: (r + 1) * tallest] for r in range(world)]
    dist.all_gather(parts, send, group=group)


def run_sharded(num_units: int, unit_rows: int, compute: Callable[[int, int], torch.Tensor],
                group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Run `compute(first_unit, unit_count) -> [unit_count*unit_rows, width]` on this rank's block
    of sub-videos and return the gathered rows of ALL units in unit order."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    blocks = partition(num_units, world)
    start, cnt = blocks[rank]
    local = compute(start, cnt)
    if local.shape[0] != cnt * unit_rows:
        raise ValueError("run_sharded: compute returned the wrong number of rows")
    return gather_rows(local, [c * unit_rows for _, c in blocks], group)


def sharded_mean(local_sum: torch.Tensor, local_count: int,
                 group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Mean over all ranks of row sums accumulated per rank: ONE all-reduce of (sum[D], count).
    Used for `ncentroid` (mean feature of the normal training videos,
    src/models/anomaly_clip_module.py:419-445) when the normal set is sharded over the GPUs."""
    packed = torch.cat((local_sum.to(torch.float64).reshape(-1),
                        torch.tensor([float(local_count)], dtype=torch.float64,
                                     device=local_sum.device)))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return (packed[:-1] / packed[-1].clamp_min(1.0)).to(torch.float32).reshape(local_sum.shape)


_MODE_SPEED = {4: 3, 7: 2, 5: 1, 2: 0}     # operand modes of the image encoder, fastest = highest


def agree_on_mode(mode: int, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> int:
    """Ranks calibrate an "auto" encoder on their own frames; everybody then runs the most
    conservative of the chosen modes (one tiny all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or mode not in _MODE_SPEED:
        return mode
    t = torch.tensor([_MODE_SPEED[mode]], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return {v: k for k, v in _MODE_SPEED.items()}[int(t.item())]


class PeerRowGather:
    """Fused all-gather of the per-frame result rows over NVLink peer memory.

    Every rank owns a symmetric buffer (torch.distributed._symmetric_memory: the allocation is
    mapped into every peer of the group) holding two parity copies of the gathered rows
    [world * rows_per_rank, width] and a flag word per rank.  `TemporalScorer(..., peer=self)`
    makes the head kernel store each row it computes into ALL ranks' buffers and raise the flags
    from its last CTA; `wait()` enqueues a one-warp kernel that holds the STREAM until every rank's
    flag arrived, and returns the gathered rows -- no NCCL call, no host synchronisation.

    Double buffering by call parity plus stream-ordered consumption keeps a fast rank from
    overwriting rows a slow rank is still reading: a rank can be at most one call ahead, because
    its next-but-one call needs this rank's flag of the call in between."""

    def __init__(self, rows_per_rank: int, width: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, extra_floats: int = 0) -> None:
        """extra_floats: additional symmetric floats behind the flag words (`extra()`), for callers
        that exchange a second payload through the same mapping (the feature rows of a
        frame-sharded encoder, `PeerFeatureGather`)."""
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._lib = _lib
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerRowGather supports up to 8 ranks (one NVSwitch domain)")
        self.rows_per_rank, self.width, self.device = rows_per_rank, width, device
        self.block = self.world * rows_per_rank * width             # floats per parity copy
        self.buf = symm_mem.empty(2 * self.block + 64 + extra_floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        self._marks_host = torch.zeros(self.world, dtype=torch.int32).pin_memory()
        self._marks_event: Optional[torch.cuda.Event] = None
        self.handle = symm_mem.rendezvous(self.buf, group.group_name)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                                          # all buffers zeroed and mapped

    def descriptor(self, n_rows: int, width: int, partial: bool = False):
        """Descriptor of the next call (advances the epoch).  partial: this rank may contribute
        fewer than rows_per_rank rows (uneven unit counts); the tail of its block is then stale."""
        if width != self.width or n_rows > self.rows_per_rank or \
                (n_rows != self.rows_per_rank and not partial):
            raise ValueError("PeerRowGather: shape differs from the one it was built for")
        self.epoch += 1
        g = self._lib.PeerGather()
        g.world, g.rank, g.rows_per_rank, g.width = self.world, self.rank, self.rows_per_rank, self.width
        parity = self.epoch & 1
        for r in range(self.world):
            base = int(self.handle.buffer_ptrs[r])
            g.rows[r] = base + 4 * parity * self.block
            g.flags[r] = base + 4 * 2 * self.block
        g.epoch = self.epoch
        g.counter = self.counter.data_ptr()
        return g

    def _marks(self) -> torch.Tensor:
        return self.buf[2 * self.block + self.world: 2 * self.block + 2 * self.world].view(torch.int32)

    def _raise_if_marked(self, marks: List[int]) -> None:
        late = [r for r, v in enumerate(marks) if v != 0]
        if late:
            self._marks().zero_()          # a later epoch starts clean
            raise self._lib.AclipError(
                f"PeerRowGather: the rows of rank(s) {late} did not arrive within the wait kernel's "
                "bound (~5 s); the gathered buffer of that call was stale.  A peer stalled or died; "
                "fall back to distributed.gather_rows (one NCCL all-gather)")

    def wait(self) -> torch.Tensor:
        """Gathered rows [world * rows_per_rank, width] of the latest call (a view of the local
        symmetric buffer, valid until the call after next).

        The wait kernel gives up after a bounded time instead of hanging the device and marks the
        late ranks; those marks are copied to pinned host memory behind it, and the NEXT wait()
        (or check()) raises `AclipError` if any was set -- a timeout is never silent, and the
        steady state costs no host synchronisation."""
        if self._marks_event is not None and self._marks_event.query():
            self._marks_event = None
            self._raise_if_marked(self._marks_host.tolist())
        lib = self._lib.load()
        stream = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            self._lib.check(lib.aclip_peer_wait(self.buf.data_ptr() + 4 * 2 * self.block, self.world,
                                                self.epoch, stream.cuda_stream))
            if self._marks_event is None:      # previous copy consumed: take the next snapshot
                self._marks_host.copy_(self._marks(), non_blocking=True)
                self._marks_event = torch.cuda.Event()
                self._marks_event.record(stream)
        parity = self.epoch & 1
        return self.buf[parity * self.block:(parity + 1) * self.block].view(
            self.world * self.rows_per_rank, self.width)

    def signal(self) -> None:
        """Take part in the next exchange without contributing rows (a rank with no unit in this
        call): raises this rank's flag of the new epoch on every peer."""
        g = self.descriptor(0, self.width, partial=True)
        import ctypes as C
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.load().aclip_peer_signal(
                C.addressof(g), torch.cuda.current_stream(self.device).cuda_stream))

    def timed_out(self) -> List[int]:
        """Ranks whose rows did not arrive within the wait kernel's bound (synchronises)."""
        return [r for r, v in enumerate(self._marks().tolist()) if v != 0]

    def check(self) -> None:
        """Synchronise and raise if any wait so far timed out (call before trusting results that
        were consumed on the device only)."""
        self._marks_event = None
        self._raise_if_marked(self._marks().tolist())

    def extra(self) -> torch.Tensor:
        """The caller-defined symmetric floats behind the flag words (see `extra_floats`)."""
        return self.buf[2 * self.block + 64:]

    def extra_ptr(self, rank: int) -> int:
        """Peer-mapped address of rank `rank`'s `extra()` region."""
        return int(self.handle.buffer_ptrs[rank]) + 4 * (2 * self.block + 64)


class FrameShardedScorer:
    """One video (or batch of 512-frame units) scored by ALL ranks of one NVSwitch domain with the
    frames sharded over the ranks -- BASELINE configs[3] / SURVEY 8e option (1): 2 048 frames over
    8 GPUs are 256 frames per GPU, half a temporal unit, so the image encoder is sharded by FRAME
    (frames are independent, anomaly_clip.py:119-123) and the temporal stage by 512-frame unit
    (temporal_model.py:46-53).

    Per call: every rank encodes its contiguous block of `frames_per_rank` frames; the epilogue of
    the encoder's output projection stores the feature rows into every rank's gathered buffer over
    NVLink peer memory (aclip_vit_forward_ex); after a stream-side flag wait each rank runs
    selector + temporal + head on ITS block of units (`partition`), whose head kernel stores the
    result rows [score | class_probs] into every rank's buffer again (PeerRowGather).  Two fused
    exchanges, no NCCL call and no host synchronisation on the path; every rank ends with all rows.
    """

    def __init__(self, net, total_frames: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None) -> None:
        group = group if group is not None else dist.group.WORLD
        self.net, self.device = net, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.unit = net.num_segments * net.seg_length
        if total_frames % self.world != 0 or total_frames % self.unit != 0:
            raise ValueError(f"FrameShardedScorer: {total_frames} frames must divide into {self.world} "
                             f"equal rank blocks and into {self.unit}-frame units")
        self.total_frames = total_frames
        self.frames_per_rank = total_frames // self.world
        self.blocks = partition(total_frames // self.unit, self.world)       # units per rank
        self.width = len(net.classnames)                                      # [score | class_probs]
        dim = net.image_encoder.output_dim
        self.features = PeerRowGather(self.frames_per_rank, dim, device, group)
        most = max(c for _, c in self.blocks)
        self.rows = PeerRowGather(most * self.unit, self.width, device, group)
        self._local = torch.empty((self.frames_per_rank, dim), dtype=torch.float32, device=device)
        self._group = group

    def _agree_on_mode(self, encoder, local_frames: torch.Tensor) -> None:
        """An "auto" encoder calibrates on the frames it sees; ranks see different frames, so the
        decision is made collectively: the one-pass mode only if EVERY rank's calibration accepts
        it (one tiny all-reduce, once per encoder)."""
        if encoder.mode is not None:
            return
        encoder.calibrate(local_frames.contiguous())
        encoder.mode = agree_on_mode(encoder.mode, self.device, self._group)
        encoder.calibration["mode"] = encoder.mode
        encoder.calibration["agreed_over_ranks"] = self.world

    def frame_block(self) -> Tuple[int, int]:
        """(first frame, count) of the frames this rank encodes."""
        return self.rank * self.frames_per_rank, self.frames_per_rank

    @torch.no_grad()
    def __call__(self, local_frames: torch.Tensor, ncentroid: torch.Tensor) -> torch.Tensor:
        """local_frames: this rank's (frames_per_rank, 3, R, R) block (uint8 or fp32 normalised) of
        the `total_frames` frames, units laid out back to back (segment_size 1 per unit).
        Returns rows [total_frames, 1 + (C-1)] = [score | class_probs] of ALL frames."""
        net = self.net
        if local_frames.shape[0] != self.frames_per_rank:
            raise ValueError("FrameShardedScorer: wrong number of local frames")
        encoder = net.image_encoder.encoder()
        self._agree_on_mode(encoder, local_frames)
        encoder(local_frames, out=self._local, peer=self.features)
        feats = self.features.wait()                     # [total_frames, dim], all ranks' rows
        start, count = self.blocks[self.rank]
        scorer = net.scorer()
        if count:
            text = net.get_text_features()
            if text.device != self.device:
                text = net._text_features = text.to(self.device)
            scorer.packed.set_directions(text, ncentroid)
        if scorer.mode is None:      # "auto": ranks with units calibrate, everybody runs the agreed mode
            if count:
                scorer.calibrate(feats[start * self.unit:(start + count) * self.unit].contiguous(), 1)
            scorer.mode = agree_on_mode(scorer.mode if scorer.mode is not None else 4, self.device, self._group)
        if count:
            scorer(feats[start * self.unit:(start + count) * self.unit], 1, peer=self.rows)
        else:
            self.rows.signal()
        rows = self.rows.wait()
        per = self.rows.rows_per_rank
        if all(c * self.unit == per for _, c in self.blocks):
            return rows
        return torch.cat([rows[r * per: r * per + c * self.unit]
                          for r, (_, c) in enumerate(self.blocks) if c], dim=0)

    def check(self) -> None:
        self.features.check()
        self.rows.check()
