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
        return cls(read_annotation_file(annotation_file, root), **kw)

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
        # the reference's annotation rows name the video without the suffix (feature_dataset.py:74)
        if not rec.path.endswith(".npy"):
            rec = VideoRecord(rec.path + ".npy", rec.start_frame, rec.end_frame, rec.label)
        feats = torch.from_numpy(np.load(rec.path, allow_pickle=True)).to(torch.float32)
        return self.assemble(feats, rec)


def read_annotation_file(annotation_file: str, root: str) -> List[VideoRecord]:
    """One `VIDEO_PATH START_FRAME END_FRAME LABEL` row per video (video_dataset.py:195-199)."""
    recs = []
    for line in Path(annotation_file).read_text().splitlines():
        p = line.strip().split()
        if len(p) >= 4:
            recs.append(VideoRecord(str(Path(root) / p[0]), int(p[1]), int(p[2]), int(p[3])))
    return recs


def read_temporal_annotations(path: Optional[str]) -> Dict[str, List[int]]:
    """`NAME LABEL START END [START END ...]` rows -> {stem: [start, end, ...]}
    (video_dataset.py:201-211, feature_dataset.py the same)."""
    out: Dict[str, List[int]] = {}
    if path:
        for line in Path(path).read_text().splitlines():
            p = line.strip().split()
            if len(p) >= 2:
                out[Path(p[0]).stem] = [int(v) for v in p[2:]]
    return out


class FrameVideoDataset(torch.utils.data.Dataset):
    """Raw-frame videos in test mode (data.load_from_features=False): the reference's
    `VideoFrameDataset(test_mode=True)` (video_dataset.py:237-244, 293-351) with the test transform of
    src/utils/augmentations.py:21-34.  Same constructor keywords; `__getitem__` returns the same
    5-tuple (frames (T, 3, S, S), labels [frames], label, segment_size, path).

    `output` selects how far the reference's transform runs on the host:
      "uint8"       PIL bicubic resize of the shorter edge to `input_size` + centre crop (exactly
                    GroupScale + GroupCenterCrop), frames stay uint8: ToTensor + Normalize happen on
                    the GPU inside the patchify kernel (4x less host->device traffic).  Default.
      "normalised"  the reference's full transform: fp32, /255, mean/std normalised.
      "raw"         decoded frames (T, H, W, 3) uint8 untouched, for `GpuFrameIngest` (resize + crop
                    on the GPU as well); all frames of a video must have one size.
    Frames that the wrap-around padding repeats are decoded once."""

    MEAN = (0.48145466, 0.4578275, 0.40821073)
    STD = (0.26862954, 0.26130258, 0.27577711)

    def __init__(self, root_path: str, annotationfile_path: str, normal_id: int, num_segments: int = 32,
                 frames_per_segment: int = 16, imagefile_template: str = "{:06d}.jpg", transform=None,
                 test_mode: bool = True, val_mode: bool = False, ncrops: int = 1,
                 temporal_annotation_file: Optional[str] = None, labels_file: Optional[str] = None,
                 stride: int = 1, spatialannotationdir_path: Optional[str] = None,
                 input_size: int = 224, output: str = "uint8") -> None:
        if not test_mode or val_mode:
            raise NotImplementedError("FrameVideoDataset: only test_mode=True (the inference path) exists; "
                                      "the random training sampler is out of scope")
        if output not in ("uint8", "normalised", "raw"):
            raise ValueError(f"FrameVideoDataset: unknown output '{output}'")
        if ncrops < 1:
            raise ValueError("FrameVideoDataset: ncrops must be >= 1")
        self.root_path, self.annotationfile_path = root_path, annotationfile_path
        self.normal_id, self.num_segments, self.frames_per_segment = normal_id, num_segments, frames_per_segment
        self.imagefile_template, self.transform, self.stride, self.ncrops = imagefile_template, transform, stride, ncrops
        self.input_size, self.output = input_size, output
        self.labels_file = labels_file
        self.video_list = read_annotation_file(annotationfile_path, root_path)
        self.annotations = read_temporal_annotations(temporal_annotation_file)

    def __len__(self) -> int:
        return len(self.video_list)

    def _load(self, directory: str, frame: int) -> np.ndarray:
        from PIL import Image
        img = Image.open(str(Path(directory) / self.imagefile_template.format(frame))).convert("RGB")
        if self.output != "raw":
            w, h = img.size
            s = self.input_size
            ow, oh = (s, int(s * h / w)) if w <= h else (int(s * w / h), s)   # torchvision Resize(int)
            img = img.resize((ow, oh), Image.BICUBIC)
            top, left = int(round((oh - s) / 2.0)), int(round((ow - s) / 2.0))  # torchvision CenterCrop
            img = img.crop((left, top, left + s, top + s))
        return np.asarray(img)

    def __getitem__(self, i: int):
        rec = self.video_list[i]
        # ncrops only shortens the label vector here (video_dataset.py:323); the frames are not cropped
        labels = frame_labels(rec.num_frames // self.ncrops, rec.start_frame, rec.label, self.normal_id,
                              self.annotations.get(Path(rec.path).stem, ()))
        idx, segment_size = test_mode_indices(rec.num_frames, self.num_segments,
                                              self.frames_per_segment, self.stride)
        uniq, inverse = np.unique(idx, return_inverse=True)
        decoded = np.stack([self._load(rec.path, int(f) + rec.start_frame) for f in uniq])  # (U, H, W, 3)
        frames = torch.from_numpy(decoded)
        if self.output != "raw":
            frames = frames.permute(0, 3, 1, 2)                                       # (U, 3, S, S)
            if self.output == "normalised":
                mean = torch.tensor(self.MEAN).view(1, 3, 1, 1)
                std = torch.tensor(self.STD).view(1, 3, 1, 1)
                frames = (frames.to(torch.float32).div(255.0) - mean) / std
        frames = frames.index_select(0, torch.from_numpy(inverse.astype(np.int64))).contiguous()
        if self.transform is not None:
            frames = self.transform(frames)
        return frames, labels, rec.label, segment_size, rec.path


def save_features(path: str, feats: torch.Tensor, ncrops: int = 1) -> str:
    """Write encoder outputs in the on-disk format `FeatureVideoDataset` (and the reference's
    feature_dataset.py:74,326-349) reads: `<path>.npy`, fp32 `[frames * ncrops, D]`, frame-major with
    the crops of a frame adjacent.  `feats` is (frames, D), (frames, ncrops, D) or (ncrops, frames, D)
    as returned by the dataset (crop-major); returns the file name."""
    f = feats.detach().to(device="cpu", dtype=torch.float32)
    if f.dim() == 3:
        if f.shape[0] == ncrops and f.shape[1] != ncrops:      # (ncrops, frames, D) -> frame-major
            f = f.permute(1, 0, 2)
        if f.shape[1] != ncrops:
            raise ValueError(f"save_features: {tuple(feats.shape)} does not hold {ncrops} crops per frame")
        f = f.reshape(-1, f.shape[-1])
    elif f.dim() != 2 or f.shape[0] % ncrops != 0:
        raise ValueError(f"save_features: expected (frames*ncrops, D) rows, got {tuple(feats.shape)}")
    name = path if path.endswith(".npy") else path + ".npy"
    np.save(name, f.contiguous().numpy())
    return name


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


# --------------------------------------------------------------------------------------------
# GPU-side frame ingest: PIL-exact bicubic resize + centre crop plan
# --------------------------------------------------------------------------------------------
_PRECISION_BITS = 32 - 8 - 2   # Pillow, libImaging/Resample.c (8 bits per channel)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_resample_plan(in_size: int, out_size: int, first: int = 0, count: Optional[int] = None):
    """Pillow's `precompute_coeffs` + `normalize_coeffs_8bpc` for the bicubic filter over the full
    input extent: for output samples [first, first+count) the first input index, the tap count
    and the fixed-point (22-bit) taps.  Returns (bounds int32 [count, 2], coeffs int32 [count, ksize])."""
    count = out_size - first if count is None else count
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((count, 2), dtype=np.int32)
    coeffs = np.zeros((count, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for i in range(count):
        xx = first + i
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            coeffs[i, x] = int(-0.5 + w * (1 << _PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << _PRECISION_BITS))
        bounds[i] = (xmin, xmax)
    return bounds, coeffs


@dataclass
class ResizeCropPlan:
    """Everything `aclip_resize_crop_u8` needs for frames of one source size: the reference's
    GroupScale(size, BICUBIC) + GroupCenterCrop(size) (src/utils/augmentations.py:25-29) restricted
    to the pixels the crop keeps."""
    in_h: int
    in_w: int
    size: int
    row0: int                 # first source row the vertical pass reads
    rows: int                 # number of source rows it reads (height of the intermediate image)
    hbounds: np.ndarray
    hcoeffs: np.ndarray
    vbounds: np.ndarray       # relative to row0
    vcoeffs: np.ndarray


def resize_crop_plan(in_h: int, in_w: int, size: int = 224) -> ResizeCropPlan:
    # torchvision.transforms.Resize(int): the smaller edge becomes `size`
    if in_w <= in_h:
        out_w, out_h = size, int(size * in_h / in_w)
    else:
        out_h, out_w = size, int(size * in_w / in_h)
    # torchvision center_crop
    top = int(round((out_h - size) / 2.0))
    left = int(round((out_w - size) / 2.0))
    hb, hc = pil_resample_plan(in_w, out_w, left, size)
    vb, vc = pil_resample_plan(in_h, out_h, top, size)
    row0 = int(vb[:, 0].min())
    rows = int((vb[:, 0] + vb[:, 1]).max()) - row0
    vb = vb.copy()
    vb[:, 0] -= row0
    return ResizeCropPlan(in_h, in_w, size, row0, rows, hb, hc, vb, vc)


class GpuFrameIngest:
    """Decoded uint8 frames (F, H, W, 3) on the device -> (F, 3, size, size) uint8, bit-exact with
    the reference's PIL resize + centre crop; feed the result to the image encoder (which
    normalises uint8 frames on the fly).  One plan per source resolution."""

    def __init__(self, in_h: int, in_w: int, device: torch.device, size: int = 224) -> None:
        from . import _lib
        self._lib = _lib
        if device.type != "cuda":
            raise _lib.AclipError("GpuFrameIngest runs only on a CUDA device (no CPU fallback)")
        self.plan, self.device = resize_crop_plan(in_h, in_w, size), device
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self._hb, self._hc = up(self.plan.hbounds), up(self.plan.hcoeffs)
        self._vb, self._vc = up(self.plan.vbounds), up(self.plan.vcoeffs)
        self._tmp = None

    def __call__(self, frames_hwc: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        p = self.plan
        if not frames_hwc.is_cuda or frames_hwc.dtype != torch.uint8 or \
                tuple(frames_hwc.shape[1:]) != (p.in_h, p.in_w, 3):
            raise ValueError(f"GpuFrameIngest: expected CUDA uint8 (F,{p.in_h},{p.in_w},3) frames")
        frames_hwc = frames_hwc.contiguous()
        n = frames_hwc.shape[0]
        if out is None:
            out = torch.empty((n, 3, p.size, p.size), dtype=torch.uint8, device=self.device)
        if n == 0:
            return out
        need = n * p.rows * p.size * 3
        if self._tmp is None or self._tmp.numel() < need:
            self._tmp = torch.empty(need, dtype=torch.uint8, device=self.device)
        lib = self._lib.load()
        self._lib.check(lib.aclip_resize_crop_u8(
            frames_hwc.data_ptr(), n, p.in_h, p.in_w, p.row0, p.rows, p.size, self._hb.data_ptr(),
            self._hc.data_ptr(), p.hcoeffs.shape[1], self._vb.data_ptr(), self._vc.data_ptr(),
            p.vcoeffs.shape[1], self._tmp.data_ptr(), out.data_ptr(),
            torch.cuda.current_stream(self.device).cuda_stream))
        return out
