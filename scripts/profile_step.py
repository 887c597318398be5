"""One training step of BASELINE.json configs[1] between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py [--batch 32] [--tile 256]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.data import synthetic_batch  # noqa: E402
from kcl_ltss_bioatm_b200.spec import UNetSpec  # noqa: E402
from kcl_ltss_bioatm_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--tile", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
spec = UNetSpec()
tr = Trainer(spec, device="cuda:0")
x, t = synthetic_batch(a.batch, a.tile, a.tile, spec.in_channels, seed=1)
x, t = x.cuda(), t.cuda()
for _ in range(a.warmup):
    tr.step(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    tr.step(x, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", tr.model.loss_out.tolist())
