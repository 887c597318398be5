"""Synthetic MODIS/VIIRS-like tiles with plume-like masks.

The reference ships no data (all scripts read absolute paths on the author's machines,
``src/config/filepaths.py:7-9``) and there is no network, so training / benchmarking inputs are generated:
per-band N(0,1) noise plus an aerosol-like field made of 1-3 anisotropic Gaussian blobs; the mask is the
field thresholded so that roughly 5-15 % of the pixels are plume (the reference's plume-size acceptance
window is 100-2000 px, ``src/features/plume_identifier_gaussian_profile.py:38-39``).  The field leaks
into the first bands with different gains so the segmentation is learnable.

torch is used here as plumbing (RNG, device memory); nothing in this file is on the timed path.
"""
from __future__ import annotations

from collections import deque
from typing import Iterable, Iterator, Tuple

import torch


def synthetic_batch(n: int, h: int, w: int, channels: int, seed: int, device="cpu",
                    dtype=torch.bfloat16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (tiles [n,h,w,channels] `dtype` NHWC, masks [n,h,w] uint8), deterministic in `seed`.
    Generation happens on CPU so that CPU oracle and GPU runs see bit-identical inputs."""
    g = torch.Generator().manual_seed(int(seed))
    yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing="ij")
    field = torch.zeros(n, h, w)
    nblob = torch.randint(1, 4, (n,), generator=g)
    for i in range(n):
        for _ in range(int(nblob[i])):
            cy = torch.rand(1, generator=g).item() * h
            cx = torch.rand(1, generator=g).item() * w
            # anisotropic: long axis 3-8x the short one, random orientation (plumes are elongated)
            s_short = (0.03 + 0.05 * torch.rand(1, generator=g).item()) * min(h, w)
            s_long = s_short * (3.0 + 5.0 * torch.rand(1, generator=g).item())
            th = torch.rand(1, generator=g).item() * 3.14159265
            ct, st = torch.cos(torch.tensor(th)), torch.sin(torch.tensor(th))
            u = (xx - cx) * ct + (yy - cy) * st
            v = -(xx - cx) * st + (yy - cy) * ct
            amp = 0.6 + 0.8 * torch.rand(1, generator=g).item()
            field[i] += amp * torch.exp(-0.5 * ((u / s_long) ** 2 + (v / s_short) ** 2))
    mask = (field > 0.35).to(torch.uint8)
    x = torch.randn(n, h, w, channels, generator=g)
    gains = torch.linspace(2.0, 0.25, steps=min(channels, 4))
    for b in range(gains.numel()):
        x[..., b] += gains[b] * field
    x = x.to(dtype)
    if str(device) != "cpu":
        x, mask = x.to(device), mask.to(device)
    return x, mask


def synthetic_scene(h: int, w: int, channels: int, seed: int, dtype=torch.bfloat16) -> torch.Tensor:
    """One large scene [h, w, channels] (CPU) for tiled inference; built tile-row by tile-row so that a
    4096 x 4096 swath does not need a giant meshgrid temp."""
    g = torch.Generator().manual_seed(int(seed))
    out = torch.empty(h, w, channels, dtype=dtype)
    nblob = max(1, (h * w) // (256 * 256) // 2)
    cy = torch.rand(nblob, generator=g) * h
    cx = torch.rand(nblob, generator=g) * w
    ss = 4.0 + 10.0 * torch.rand(nblob, generator=g)
    sl = ss * (3.0 + 5.0 * torch.rand(nblob, generator=g))
    th = torch.rand(nblob, generator=g) * 3.14159265
    amp = 0.6 + 0.8 * torch.rand(nblob, generator=g)
    gains = torch.linspace(2.0, 0.25, steps=min(channels, 4))
    rows = 256
    xs = torch.arange(w, dtype=torch.float32)
    for y0 in range(0, h, rows):
        y1 = min(y0 + rows, h)
        ys = torch.arange(y0, y1, dtype=torch.float32)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        field = torch.zeros(y1 - y0, w)
        near = ((cy + 4 * sl) >= y0) & ((cy - 4 * sl) <= y1)
        for k in torch.nonzero(near).flatten().tolist():
            ct, st = torch.cos(th[k]), torch.sin(th[k])
            u = (xx - cx[k]) * ct + (yy - cy[k]) * st
            v = -(xx - cx[k]) * st + (yy - cy[k]) * ct
            field += amp[k] * torch.exp(-0.5 * ((u / sl[k]) ** 2 + (v / ss[k]) ** 2))
        blk = torch.randn(y1 - y0, w, channels, generator=g)
        for b in range(gains.numel()):
            blk[..., b] += gains[b] * field
        out[y0:y1] = blk.to(dtype)
    return out


class DevicePrefetcher:
    """Iterates device copies of an iterable of pinned host batches ``(tiles, masks)``, keeping the
    host->device copy of the next batches in flight on a copy stream while the caller trains on the current
    one.  Every batch is copied exactly once, when it is `depth` batches ahead of use; the yielded tensors
    are valid until the next item is requested.  On a CPU device the host batches pass through unchanged."""

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], device, depth: int = 2):
        self.batches, self.device, self.depth = batches, torch.device(device), max(1, int(depth))

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        if self.device.type != "cuda":
            yield from self.batches
            return
        copy_stream = torch.cuda.Stream(device=self.device)
        src = iter(self.batches)
        slots, queue = [], deque()

        def issue(slot, hb) -> None:
            x, t = hb
            if slot["x"] is None or slot["x"].shape != x.shape or slot["x"].dtype != x.dtype:
                slot["x"] = torch.empty(x.shape, dtype=x.dtype, device=self.device)
                slot["t"] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])  # the consumer's kernels on this slot have run
            with torch.cuda.stream(copy_stream):
                slot["x"].copy_(x, non_blocking=True)
                slot["t"].copy_(t, non_blocking=True)
                slot["ready"] = torch.cuda.Event()
                slot["ready"].record(copy_stream)
            queue.append(slot)

        for _ in range(self.depth):
            hb = next(src, None)
            if hb is None:
                break
            slots.append({"x": None, "t": None, "free": None, "ready": None})
            issue(slots[-1], hb)
        while queue:
            slot = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(slot["ready"])
            yield slot["x"], slot["t"]
            slot["free"] = torch.cuda.Event()
            slot["free"].record(torch.cuda.current_stream(self.device))
            hb = next(src, None)
            if hb is not None:
                issue(slot, hb)
