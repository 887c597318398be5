import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec
from kcl_ltss_bioatm_b200.trainer import Trainer
only = len(sys.argv) > 1 and sys.argv[1] == 'graph-only'
for (b, hw) in ([(32, 256)] if only else [(32, 256), (1, 256)]):
    spec = UNetSpec()
    tr = Trainer(spec, device="cuda:0")
    x, t = synthetic_batch(b, hw, hw, spec.in_channels, seed=1)
    x, t = x.cuda(), t.cuda()
    for mode in (("graph",) if only else ("eager", "graph")):
        fn = tr.step if mode == "eager" else tr.step_graphed
        for _ in range(5):
            fn(x, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        n = 40
        for _ in range(n):
            fn(x, t)
        e1.record()
        host = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        print(f"B={b} {hw}^2 {mode}: {e0.elapsed_time(e1) / n:.3f} ms/step (host enqueue {host:.3f} ms/step), loss {tr.model.loss_out[0].item():.4f}", flush=True)
    del tr
