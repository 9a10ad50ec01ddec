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
                 group: Optional[dist.ProcessGroup] = None) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._lib = _lib
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerRowGather supports up to 8 ranks (one NVSwitch domain)")
        self.rows_per_rank, self.width, self.device = rows_per_rank, width, device
        self.block = self.world * rows_per_rank * width             # floats per parity copy
        self.buf = symm_mem.empty(2 * self.block + 64, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group.group_name)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)                                          # all buffers zeroed and mapped

    def descriptor(self, n_rows: int, width: int):
        if n_rows != self.rows_per_rank or width != self.width:
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

    def wait(self) -> torch.Tensor:
        """Gathered rows [world * rows_per_rank, width] of the latest call (a view of the local
        symmetric buffer, valid until the call after next)."""
        lib = self._lib.load()
        self._lib.check(lib.aclip_peer_wait(self.buf.data_ptr() + 4 * 2 * self.block, self.world,
                                            self.epoch, torch.cuda.current_stream(self.device).cuda_stream))
        parity = self.epoch & 1
        return self.buf[parity * self.block:(parity + 1) * self.block].view(
            self.world * self.rows_per_rank, self.width)

    def timed_out(self) -> List[int]:
        """Ranks whose rows did not arrive within the wait kernel's bound (synchronises)."""
        marks = self.buf[2 * self.block + self.world: 2 * self.block + 2 * self.world]
        return [r for r, v in enumerate(marks.view(torch.int32).tolist()) if v != 0]
