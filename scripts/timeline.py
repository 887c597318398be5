"""Diagnostic: kernel timeline (start, duration, stream) of one training step via torch.profiler (CUPTI).
Shows which kernels actually overlap.  Not a measurement: numbers under a profiler are never bench values."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec
from kcl_ltss_bioatm_b200.trainer import Trainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.txt"
mode = sys.argv[2] if len(sys.argv) > 2 else "graph"
spec = UNetSpec()
tr = Trainer(spec, device="cuda:0")
x, t = synthetic_batch(32, 256, 256, spec.in_channels, seed=1)
x, t = x.cuda(), t.cuda()
step = tr.step_graphed if mode == 'graph' else tr.step
for _ in range(5):
    step(x, t)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, t)
    torch.cuda.synchronize()
tmp = out + ".json"
prof.export_chrome_trace(tmp)
ev = [e for e in json.load(open(tmp))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
with open(out, "w") as f:
    f.write("start_us  dur_us  end_us  stream  kernel\n")
    for e in ev:
        f.write(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {e['ts'] - t0 + e['dur']:9.1f} {e['args'].get('stream')}  {e['name'][:70]}\n")
    f.write(f"span {ev[-1]['ts'] + ev[-1]['dur'] - t0:.1f} us, sum of durations {sum(e['dur'] for e in ev):.1f} us\n")
os.remove(tmp)
print(open(out).read()[-200:])
