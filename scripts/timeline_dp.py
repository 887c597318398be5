"""Diagnostic (torchrun, N ranks): kernel timeline of one data-parallel training step on rank 0 via torch.profiler, with the
NCCL all-reduce kernels marked and the time during which ONLY collective kernels run (= exposed communication).
  torchrun --nproc-per-node N scripts/timeline_dp.py out.txt [default|wide] [graph|eager]
Not a measurement: numbers under a profiler are never bench values."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from kcl_ltss_bioatm_b200.data import synthetic_batch  # noqa: E402
from kcl_ltss_bioatm_b200.spec import UNetSpec  # noqa: E402
from kcl_ltss_bioatm_b200.trainer import Trainer, init_distributed  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_dp.txt"
which = sys.argv[2] if len(sys.argv) > 2 else "default"
mode = sys.argv[3] if len(sys.argv) > 3 else "graph"
rank, world, local, pg = init_distributed("cuda")
dev = torch.device("cuda", local)
spec = UNetSpec.wide() if which == "wide" else UNetSpec()
tr = Trainer(spec, device=dev, process_group=pg, seed=0)
n, hw = (1, 1024) if which == "wide" else (32, 256)
x, t = synthetic_batch(n, hw, hw, spec.in_channels, seed=1 + rank)
x, t = x.to(dev), t.to(dev)
step = tr.step_graphed if mode == "graph" else tr.step
for _ in range(5):
    step(x, t)
torch.cuda.synchronize()
if pg is not None:
    torch.distributed.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, t)
    torch.cuda.synchronize()
if rank == 0:
    tmp = out + ".json"
    prof.export_chrome_trace(tmp)
    ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") == "kernel"]
    os.remove(tmp)
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    is_nccl = lambda e: "nccl" in e["name"].lower()  # noqa: E731
    # sweep line: time covered by compute kernels, by NCCL kernels, by NCCL only
    pts = []
    for e in ev:
        pts.append((e["ts"], 1, is_nccl(e)))
        pts.append((e["ts"] + e["dur"], -1, is_nccl(e)))
    pts.sort()
    comp = coll = 0
    only_coll = both = 0.0
    last = pts[0][0]
    for ts, d, nc in pts:
        if coll > 0 and comp == 0:
            only_coll += ts - last
        if coll > 0 and comp > 0:
            both += ts - last
        last = ts
        if nc:
            coll += d
        else:
            comp += d
    with open(out, "w") as f:
        f.write(f"# {which} spec, {world} ranks, {mode}: rank 0 kernel timeline of one step (torch.profiler)\n")
        f.write("start_us  dur_us  end_us  stream  kernel\n")
        for e in ev:
            if is_nccl(e) or e["dur"] > 150 or "adam" in e["name"]:
                f.write(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {e['ts'] - t0 + e['dur']:9.1f} {e['args'].get('stream')}  "
                        f"{e['name'][:80]}\n")
        span = ev[-1]["ts"] + ev[-1]["dur"] - t0
        nc = [e for e in ev if is_nccl(e)]
        f.write(f"# span {span:.1f} us; {len(nc)} NCCL kernels, {sum(e['dur'] for e in nc):.1f} us in total; "
                f"{both:.1f} us overlapped with compute kernels, {only_coll:.1f} us with NO compute kernel running "
                f"(exposed communication)\n")
    print(open(out).read()[-1500:])
tr.release_graphs()
if pg is not None:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
