"""Test-mode frame-batch assembly: what the reference's datasets hand to
`test_step`, without the per-frame Python loops.

Reference: src/data/components/feature_dataset.py:17-27,243-259,306-376 and
video_dataset.py:228-244,291-351 (same index plan), collated with batch_size_test = 1
(src/data/anomaly_clip_datamodule.py:175-183).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def padded_length(num_frames: int, num_segments: int, seg_length: int, stride: int = 1) -> int:
    """round_to_nearest(num_frames, n*l*stride): feature_dataset.py:17-27,252-255."""
    unit = num_segments * seg_length * stride
    return int(math.ceil(num_frames / unit) * unit)


def test_mode_indices(num_frames: int, num_segments: int, seg_length: int, stride: int = 1
                      ) -> Tuple[np.ndarray, int]:
    """Source-frame index of every sampled row, in the order the reference appends them
    (feature_dataset.py:256-259,359-366): segment starts every l*stride frames up to the padded
    length, l frames per segment `stride` apart, wrapped around with `% num_frames`.
    Returns (indices [padded/stride], segment_size)."""
    if num_frames <= 0:
        raise ValueError("a video needs at least one frame")
    end = padded_length(num_frames, num_segments, seg_length, stride)
    starts = np.arange(end // (seg_length * stride), dtype=np.int64) * (seg_length * stride)
    idx = (starts[:, None] + np.arange(seg_length, dtype=np.int64)[None, :] * stride) % num_frames
    return idx.reshape(-1), len(starts) // num_segments


test_mode_indices.__test__ = False  # not a pytest test


def frame_labels(num_frames: int, start_frame: int, label: int, normal_id: int,
                 intervals: Sequence[int]) -> np.ndarray:
    """Per-frame class label from the temporal annotation [s0, e0, s1, e1, ...] (inclusive
    bounds on the absolute frame id): feature_dataset.py:336-349."""
    ids = np.arange(num_frames, dtype=np.int64) + start_frame
    out = np.full(num_frames, normal_id, dtype=np.int64)
    for s, e in zip(intervals[::2], intervals[1::2]):
        out[(ids >= int(s)) & (ids <= int(e))] = label
    return out


@dataclass
class VideoRecord:
    path: str
    start_frame: int
    end_frame: int
    label: int

    @property
    def num_frames(self) -> int:
        return self.end_frame - self.start_frame + 1


class FeatureVideoDataset(torch.utils.data.Dataset):
    """Pre-extracted features (`.npy`, [frames*ncrops, 512]) in test mode
    (data.load_from_features=True).  `__getitem__` returns the reference's 5-tuple
    (features (ncrops, T, 512), labels [frames], label, segment_size, path)."""

    def __init__(self, records: List[VideoRecord], num_segments: int, seg_length: int, stride: int = 1,
                 ncrops: int = 1, normal_id: int = 0,
                 annotations: Optional[Dict[str, Sequence[int]]] = None) -> None:
        self.records, self.annotations = records, annotations or {}
        self.num_segments, self.seg_length, self.stride = num_segments, seg_length, stride
        self.ncrops, self.normal_id = ncrops, normal_id

    @classmethod
    def from_annotation_file(cls, annotation_file: str, root: str, **kw) -> "FeatureVideoDataset":
        recs = []
        for line in Path(annotation_file).read_text().splitlines():
            p = line.strip().split()
            if len(p) >= 4:
                recs.append(VideoRecord(str(Path(root) / p[0]), int(p[1]), int(p[2]), int(p[3])))
        return cls(recs, **kw)

    def __len__(self) -> int:
        return len(self.records)

    def assemble(self, feats: torch.Tensor, rec: VideoRecord):
        frames = feats.shape[0] // self.ncrops
        labels = frame_labels(frames, rec.start_frame, rec.label, self.normal_id,
                              self.annotations.get(Path(rec.path).stem, ()))
        idx, segment_size = test_mode_indices(frames, self.num_segments, self.seg_length, self.stride)
        x = feats.reshape(frames, self.ncrops, feats.shape[-1])
        x = x.index_select(0, torch.from_numpy(idx)).permute(1, 0, 2).contiguous()  # (ncrops, T, D)
        return x, labels, rec.label, segment_size, rec.path

    def __getitem__(self, i: int):
        rec = self.records[i]
        feats = torch.from_numpy(np.load(rec.path, allow_pickle=True)).to(torch.float32)
        return self.assemble(feats, rec)


def gather_test_frames(frames: torch.Tensor, num_segments: int, seg_length: int, stride: int = 1
                       ) -> Tuple[torch.Tensor, int]:
    """Raw-frame variant (video_dataset.py:331-346) for frames already decoded to a
    (F, 3, H, W) tensor (uint8 or normalised fp32, host or device): the padded, wrapped-around
    batch (T, 3, H, W) and the segment_size.  Runs as one index_select on the tensor's device."""
    idx, segment_size = test_mode_indices(frames.shape[0], num_segments, seg_length, stride)
    return frames.index_select(0, torch.from_numpy(idx).to(frames.device)), segment_size


class DevicePrefetcher:
    """Wraps an iterable of test-mode batches (the reference's 5-tuples, first element on the
    host, ideally pinned) and keeps the NEXT batch's host->device copy in flight on a side stream
    while the current one is being computed, so the copy of 150 KB/frame of uint8 pixels (or
    2 KB/row of features) hides behind the encoder.  Yields the same tuples with element 0 on the
    device.  Two device staging buffers are reused in turn (no allocation per batch): the tensor
    handed out for batch i is overwritten when batch i+2 is staged, i.e. it is valid until the
    consumer asks for batch i+1's successor -- consume each batch before advancing twice."""

    def __init__(self, batches, device: torch.device) -> None:
        self.batches, self.device = batches, device
        self.copy_stream = torch.cuda.Stream(device=device)
        self._buf = [None, None]
        self._consumed = [None, None]   # compute-stream events: buffer k may be overwritten

    def _stage(self, batch, k: int):
        src = batch[0]
        buf = self._buf[k]
        if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
            buf = self._buf[k] = torch.empty(src.shape, dtype=src.dtype, device=self.device)
        if self._consumed[k] is not None:
            self.copy_stream.wait_event(self._consumed[k])
        with torch.cuda.stream(self.copy_stream):
            buf.copy_(src, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(self.copy_stream)
        return (buf, *batch[1:]), ready

    def __iter__(self):
        it = iter(self.batches)
        compute = torch.cuda.current_stream(self.device)
        k = 0
        try:
            nxt = self._stage(next(it), k)
        except StopIteration:
            return
        while nxt is not None:
            cur, ready = nxt
            cur_k = k
            k ^= 1
            try:
                nxt = self._stage(next(it), k)
            except StopIteration:
                nxt = None
            compute.wait_event(ready)
            yield cur
            done = torch.cuda.Event()      # everything the consumer enqueued on batch cur_k
            done.record(compute)
            self._consumed[cur_k] = done
