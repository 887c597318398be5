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


def run_dense():                                   # round-1 path: byte masks, int32 label / size planes
    for t, (m, lab, sz, ext) in zip(thr, bufs):
        sw.ops.threshold_masks(a, t, m)
        sw.ops.label_components(m, lab, sz)
        sw.ops.fire_extents(lab, sz, rc, sweep.P_ID_WIN_SIZE, ext)


thr_all = torch.cat(thr)
ws = torch.empty(sw.ops.sweep_workspace_bytes(H, W, thr_all.numel()), dtype=torch.uint8, device="cuda")
ext_all = torch.empty(thr_all.numel(), 64, dtype=torch.int32, device="cuda")


def run_bits_per_sweep():                          # bit planes, one call per sweep as the reference loops
    for t, (_, _, _, ext) in zip(thr, bufs):
        sw.ops.sweep_extents(a, t, rc, sweep.P_ID_WIN_SIZE, ws, ext)


def run():                                         # bit planes, the three sweeps of the timestamp in one call
    sw.ops.sweep_extents(a, thr_all, rc, sweep.P_ID_WIN_SIZE, ws, ext_all)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_dense = timed(run_dense)
ms_per_sweep = timed(run_bits_per_sweep)
ms = timed(run)
agree = bool(torch.equal(ext_all, torch.cat([b[3] for b in bufs])))          # one call == three calls
run_dense()
agree = agree and bool(torch.equal(ext_all, torch.cat([b[3] for b in bufs])))  # bit planes == dense planes
n_thr = sum(len(t) for t in sweeps)
# round-1 definition (dense formulation): image read once per sweep, per threshold the mask written and read (1 B)
# and the label and size planes written (4 B each) -- kept so that the fraction is comparable across rounds
alg_bytes = 3 * H * W * 4 + n_thr * H * W * (1 + 1 + 4 + 4)
# bit-plane formulation: image read once, per threshold the bit plane written once and read by init / merge / flatten
bits_bytes = H * W * 4 + n_thr * H * ((W + 31) // 32) * 4 * 4

t0 = time.perf_counter()
for _ in range(3):
    res_all = sw.extents(aod, np.concatenate(sweeps), rows, cols)
e2e_s = (time.perf_counter() - t0) / 3
res = np.split(res_all, np.cumsum([len(t) for t in sweeps])[:-1])

# the reference's own call sequence, one sweep at a time (identify(), gaussian_profile.py:489-503), host image in
t0 = time.perf_counter()
for _ in range(3):
    named = []
    for t in sweeps:
        md = sweep.generate_mask_dict(aod, t)
        named.append(sweep.find_plume_extents(md, rows, cols))
        sweep.find_threshold_index(named[-1])
named_s = (time.perf_counter() - t0) / 3
agree = agree and all(bool(np.array_equal(a, b)) for a, b in zip(named, res))

from oracle import c_ref, sweep_ref  # noqa: E402
c_ref.label8(np.zeros((4, 4), dtype=np.uint8))                         # build / load outside the timed region
t0 = time.perf_counter()
ref = [c_ref.plume_extents(c_ref.threshold_masks(aod, t), rows, cols) for t in sweeps]   # plain-C oracle, one core
cpu_s = time.perf_counter() - t0
ok = agree and all(bool(np.array_equal(a, b)) for a, b in zip(ref, res))
small = sweep_ref.find_plume_extents_ref(sweep_ref.threshold_masks_ref(aod[:300, :300], sweeps[0][:5]),
                                         np.clip(rows[:8], 16, 283), np.clip(cols[:8], 16, 283))
ok = ok and bool(np.array_equal(small, c_ref.plume_extents(c_ref.threshold_masks(aod[:300, :300], sweeps[0][:5]),
                                                             np.clip(rows[:8], 16, 283), np.clip(cols[:8], 16, 283))))

# the step before the sweep in the reference's loop: interpolate_aod_nearest (gaussian_profile.py:613)
from tests.sweep_data import synthetic_null_aod  # noqa: E402
from scipy import interpolate  # noqa: E402
nul = synthetic_null_aod(H, W, 9)
nul_dev = torch.from_numpy(nul).cuda()
fill_ws = torch.empty(sw.ops.fill_nearest_workspace_bytes(H, W), dtype=torch.uint8, device="cuda")
fill_out = torch.empty_like(nul_dev)
ms_fill = timed(lambda: sw.ops.fill_nearest(nul_dev, -999, fill_ws, fill_out))
one = torch.full((H, W), -999.0, dtype=torch.float64, device="cuda")
one[7, 11] = 0.5                                                      # worst case: every pixel walks all rows
ms_fill_worst = timed(lambda: sw.ops.fill_nearest(one, -999, fill_ws, fill_out))
sw.ops.fill_nearest(nul_dev, -999, fill_ws, fill_out)
t0 = time.perf_counter()
good = nul != -999
xx, yy = np.meshgrid(np.arange(W), np.arange(H))
scipy_filled = interpolate.NearestNDInterpolator(np.vstack((xx[good], yy[good])).T, nul[good])(np.ravel(xx), np.ravel(yy)).reshape(H, W)
fill_cpu_s = time.perf_counter() - t0
fill_agree = float((fill_out.cpu().numpy() == scipy_filled).mean())          # < 1 only by distance ties

# the whole data-parallel front half of a timestamp through ThresholdSweep.timestamp: host float64 image with nulls in
ts_aod = np.round(aod.astype(np.float64) * 1000) * 0.001
ts_aod[nul == -999] = -999
fr = np.concatenate([rows + d for d in (0, 0, 1)])                  # 64 clusters of three touching fire pixels each
fc = np.concatenate([cols + d for d in (0, 1, 1)])
keep = (fr > 20) & (fr < H - 21) & (fc > 20) & (fc < W - 21)
sw.timestamp(ts_aod, fr[keep], fc[keep])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ts_out = sw.timestamp(ts_aod, fr[keep], fc[keep])
torch.cuda.synchronize()
timestamp_s = (time.perf_counter() - t0) / 3
n_plume_masks = sum(m is not None for s_ in ts_out["sweeps"] for m in s_["plume_masks"])

peaks = {}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))
peak = float(peaks.get("hbm_gbps", peaks.get("hbm_gbs", 6541.0)))
line = {
    "metric": "threshold_sweep_timestamps_per_sec", "value": 1e3 / ms, "unit": "timestamps/s", "n_gpus": 1,
    "ms_per_timestamp": ms, "dtype": "u32 bit planes (32 px / word), int32 union-find over runs", "data": "synthetic",
    "config": {"workload": f"{H}x{W} AOD, 3 sweeps x 25 thresholds = {n_thr} masks, components, 64 fires; one call"},
    "ms_per_timestamp_one_call_per_sweep": ms_per_sweep, "ms_per_timestamp_dense_planes": ms_dense,
    "e2e": {"value": 1.0 / e2e_s, "unit": "timestamps/s", "h2d_bytes_per_step": H * W * 4, "d2h_bytes_per_step": n_thr * 64 * 4,
            "note": "ThresholdSweep.extents from a host image (75 thresholds in one call), extents copied back",
            "reference_named_calls_ms": named_s * 1e3,
            "reference_named_calls": "generate_mask_dict + find_plume_extents + find_threshold_index per sweep, three sweeps, host image in"},
    "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                 "bytes_definition": "dense formulation of round 1 (image + 10 B per pixel and threshold), for comparison across rounds",
                 "bit_plane_bytes": bits_bytes, "bit_plane_gbps": bits_bytes / (ms * 1e-3) / 1e9},
    "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "timestamps/s", "cores": 1, "kind": "port",
                     "sample": f"plain-C oracle (two-pass union-find labelling), all three sweeps of the timestamp ({cpu_s:.2f} s)"},
    "parity_on_sample": ok,
    "timestamp_front_half": {"ms": timestamp_s * 1e3, "fire_clusters": int(len(ts_out["fire_rows"])), "plume_masks_returned": int(n_plume_masks),
                             "what": "ThresholdSweep.timestamp from a host float64 image with nulls: fill, fire clustering, 75 thresholds, "
                                     "threshold index, plume masks unpacked on the host (main :611-613 + identify :478-499)"},
    "nearest_fill": {"ms": ms_fill, "ms_single_valid_pixel": ms_fill_worst, "null_fraction": float((~good).mean()), "dtype": "f64",
                     "gbps": (H * W * 8 * 2 + H * W * 4 * 2) / (ms_fill * 1e-3) / 1e9,
                     "bytes_definition": "image read + written (8 B each) and the row-offset plane written + read (4 B each)",
                     "cpu_scipy_s": fill_cpu_s, "cpu_cores": 1, "agreement_with_scipy": fill_agree},
}
print(json.dumps(line))
