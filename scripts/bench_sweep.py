"""Measurement of the threshold sweep, one JSON line: a 1200 x 1200 AOD grid (the MAIAC
tile size), 64 fire clusters and the reference's three sweeps (25 thresholds each: steps 0.02 / 0.03 / 0.04 up to
0.5 / 0.75 / 1.0, plume_identifier_gaussian_profile.py:34-35, 489-495) = what the reference does per timestamp.
Device time by CUDA events around masks + labelling + extents of all three sweeps (10 repetitions after 2 warm-ups).
Algorithmic bytes per sweep: the image read once, per threshold the mask written and read (1 B) and the label
and size planes written (4 B each).  cpu_baseline: the plain-C oracle on all three sweeps, one core."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200 import sweep  # noqa: E402
from tests.sweep_data import synthetic_aod  # noqa: E402

H = W = 1200
aod, _ = synthetic_aod(H, W, 5)
rng = np.random.default_rng(3)
rows = rng.integers(16, H - 16, 64)
cols = rng.integers(16, W - 16, 64)
sweeps = [np.abs(np.arange(0, tmax, step) - tmax) for step, tmax in [(0.02, 0.5), (0.03, 0.75), (0.04, 1)]]

sw = sweep.ThresholdSweep("cuda:0")
a = torch.from_numpy(aod).cuda()
thr = [torch.tensor(t).cuda() for t in sweeps]
rc = torch.tensor(np.stack([rows, cols], 1), dtype=torch.int32).cuda()
bufs = []
for t in sweeps:
    n = len(t)
    bufs.append((torch.empty(n, H, W, dtype=torch.uint8, device="cuda"), torch.empty(n, H, W, dtype=torch.int32, device="cuda"),
                 torch.empty(n, H, W, dtype=torch.int32, device="cuda"), torch.empty(n, 64, dtype=torch.int32, device="cuda")))


def run():
    for t, (m, lab, sz, ext) in zip(thr, bufs):
        sw.ops.threshold_masks(a, t, m)
        sw.ops.label_components(m, lab, sz)
        sw.ops.fire_extents(lab, sz, rc, sweep.P_ID_WIN_SIZE, ext)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n_thr = sum(len(t) for t in sweeps)
alg_bytes = 3 * H * W * 4 + n_thr * H * W * (1 + 1 + 4 + 4)

t0 = time.perf_counter()
for _ in range(3):
    res = [sw.extents(aod, t, rows, cols) for t in sweeps]
e2e_s = (time.perf_counter() - t0) / 3

from oracle import c_ref, sweep_ref  # noqa: E402
c_ref.label8(np.zeros((4, 4), dtype=np.uint8))                         # build / load outside the timed region
t0 = time.perf_counter()
ref = [c_ref.plume_extents(c_ref.threshold_masks(aod, t), rows, cols) for t in sweeps]   # plain-C oracle, one core
cpu_s = time.perf_counter() - t0
ok = all(bool(np.array_equal(a, b)) for a, b in zip(ref, res))
small = sweep_ref.find_plume_extents_ref(sweep_ref.threshold_masks_ref(aod[:300, :300], sweeps[0][:5]),
                                         np.clip(rows[:8], 16, 283), np.clip(cols[:8], 16, 283))
ok = ok and bool(np.array_equal(small, c_ref.plume_extents(c_ref.threshold_masks(aod[:300, :300], sweeps[0][:5]),
                                                             np.clip(rows[:8], 16, 283), np.clip(cols[:8], 16, 283))))

peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
peak = float(peaks.get("hbm_gbps", peaks.get("hbm_gbs", 6541.0)))
line = {
    "metric": "threshold_sweep_timestamps_per_sec", "value": 1e3 / ms, "unit": "timestamps/s", "n_gpus": 1,
    "ms_per_timestamp": ms, "dtype": "u8 masks, int32 union-find", "data": "synthetic",
    "config": {"workload": f"{H}x{W} AOD, 3 sweeps x 25 thresholds = {n_thr} masks + labelled planes, 64 fires"},
    "e2e": {"value": 1.0 / e2e_s, "unit": "timestamps/s", "h2d_bytes_per_step": 3 * H * W * 4, "d2h_bytes_per_step": n_thr * 64 * 4,
            "note": "ThresholdSweep.extents from a host image per sweep, extents copied back"},
    "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                 "note": "union-find merge and flatten are latency / atomic bound, not streaming"},
    "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "timestamps/s", "cores": 1, "kind": "port",
                     "sample": f"plain-C oracle (two-pass union-find labelling), all three sweeps of the timestamp ({cpu_s:.2f} s)"},
    "parity_on_sample": ok,
}
print(json.dumps(line))
