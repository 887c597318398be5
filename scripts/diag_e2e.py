"""Where does the end-to-end leg lose time against the device-resident leg?  Variants of the loop, wall clock per step."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.data import DevicePrefetcher, synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec
from kcl_ltss_bioatm_b200.trainer import LossLog, Trainer

dev = torch.device("cuda:0")
spec = UNetSpec()
tr = Trainer(spec, device=dev)
host = []
for i in range(4):
    x, t = synthetic_batch(32, 256, 256, spec.in_channels, seed=i)
    host.append((x.pin_memory(), t.pin_memory()))
devb = [(x.to(dev), t.to(dev)) for x, t in host]
for i in range(5):
    tr.step_graphed(*devb[i % 4])
torch.cuda.synchronize()


def wall(fn, n):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def resident(n):
    for i in range(n):
        tr.step_graphed(*devb[i % 4])


def resident_losslog(n):
    ring = LossLog()
    for i in range(n):
        ring.push(tr.step_graphed(*devb[i % 4]))
    ring.flush()


def resident_sync(n):
    for i in range(n):
        tr.step_graphed(*devb[i % 4])
        torch.cuda.current_stream().synchronize()


def prefetch_losslog(n):
    ring = LossLog()
    for x, t in DevicePrefetcher((host[i % 4] for i in range(n)), dev, depth=2):
        ring.push(tr.step_graphed(x, t))
    ring.flush()


def prefetch_only(n):
    for x, t in DevicePrefetcher((host[i % 4] for i in range(n)), dev, depth=2):
        tr.step_graphed(x, t)


def h2d_alone(n):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n):
            devb[0][0].copy_(host[i % 4][0], non_blocking=True)
    s.synchronize()


for name, fn in (("resident", resident), ("resident+losslog", resident_losslog), ("resident+sync", resident_sync),
                 ("prefetch", prefetch_only), ("prefetch+losslog", prefetch_losslog), ("h2d alone", h2d_alone)):
    fn(3)
    print(f"{name:20s} n=20: {wall(fn, 20):7.3f} ms/step   n=100: {wall(fn, 100):7.3f} ms/step", flush=True)
