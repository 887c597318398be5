"""Measurement of the fire -> pixel geolocation, one JSON line: a 1200 x 1200 MAIAC-like
lat/lon grid (float64) and 2048 fires, all located in one call.  Device time by CUDA events (20 calls after 3
warm-ups).  Algorithmic bytes: the two float64 grids are read once per pass (2 x 23 MB) plus the fire arrays;
the roofline entry reports that against the HBM copy peak.  cpu_baseline: the plain-C
oracle (the reference's algorithm) on a bounded sample of the same fires, one core."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200 import fires  # noqa: E402
from tests.grids import sinusoidal_grid  # noqa: E402

H = W = 1200
N = 2048
lat, lon = sinusoidal_grid(H, W, -5.0, 110.0)
rng = np.random.default_rng(11)
idx = rng.integers(0, H * W, N)
flat = lat.ravel()[idx] + rng.normal(0, 0.004, N)
flon = lon.ravel()[idx] + rng.normal(0, 0.004, N)

loc = fires.FireLocator(lat, lon, "cuda:0")
fl = torch.as_tensor(flat).cuda()
fo = torch.as_tensor(flon).cuda()
out = torch.empty(N, 2, dtype=torch.int32, device="cuda:0")
for _ in range(3):
    loc.ops.locate_fires(loc.lats, loc.lons, fl, fo, fires.HALF_BOX_DEG, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    loc.ops.locate_fires(loc.lats, loc.lons, fl, fo, fires.HALF_BOX_DEG, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps

t0 = time.perf_counter()
for _ in range(5):
    rows, cols = loc.locate(flat, flon)
e2e_s = (time.perf_counter() - t0) / 5

from oracle import c_ref, fire_ref  # noqa: E402
c_ref.nearest_pixels(flat[:1], flon[:1], lat[:8, :8], lon[:8, :8])      # build / load outside the timed region
S = 256
t0 = time.perf_counter()
ref = c_ref.nearest_pixels(flat[:S], flon[:S], lat, lon)               # plain-C oracle, one core
cpu_s = time.perf_counter() - t0
ok = bool(np.array_equal(ref, out.cpu().numpy()[:S].astype(np.int64)))
ok = ok and bool(np.array_equal(fire_ref.nearest_pixel_ref(flat[:16], flon[:16], lat, lon), ref[:16]))

peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
peak = float(peaks.get("hbm_gbps", peaks.get("hbm_gbs", 6541.0)))
streamed = 2 * (2 * H * W * 8) + 48 * N
line = {
    "metric": "fires_located_per_sec", "value": N / (ms * 1e-3), "unit": "fires/s", "n_gpus": 1, "ms_per_call": ms,
    "dtype": "f64 haversine, u64/u32 index reductions", "data": "synthetic",
    "config": {"workload": f"{H}x{W} float64 lat/lon grid, {N} fires, one call (2 passes over the grids)"},
    "e2e": {"value": N / e2e_s, "unit": "fires/s", "h2d_bytes_per_step": 16 * N, "d2h_bytes_per_step": 8 * N,
            "note": "FireLocator.locate from host fire arrays, edge filter included, grid resident on the device"},
    "roofline": {"bound": "hbm", "achieved": streamed / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": streamed / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                 "note": "bound by the per-chunk fire culling and the float64 haversine of the candidates, not by HBM"},
    "cpu_baseline": {"value": S / cpu_s, "unit": "fires/s", "cores": 1, "kind": "port",
                     "sample": f"plain-C oracle (per fire one pass over the image, as the reference's per-fire masks), first {S} fires ({cpu_s:.2f} s)"},
    "parity_on_sample": ok,
}
print(json.dumps(line))
