"""Role-level cycle counters of the conv3 kernel for a few layer shapes (diagnostics)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.ops import CudaOps
from kcl_ltss_bioatm_b200.lib import ptr
cu = CudaOps()
prof = torch.zeros(148, 8, dtype=torch.int64, device="cuda")
names = ["prod_total", "prod_wait_slot", "mma_total", "mma_wait_acc", "mma_wait_A", "mma_wait_B", "epi0_total", "epi0_wait_acc"]
for (n, h, w, cin, cout, stats) in [(32, 256, 256, 64, 64, True), (32, 256, 256, 64, 64, False), (32, 256, 256, 128, 64, True),
                                    (32, 128, 128, 128, 128, True), (32, 64, 64, 256, 256, True), (32, 32, 32, 1024, 512, True)]:
    x = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(cout, 3, 3, cin, device="cuda") / (9 * cin) ** 0.5).to(torch.bfloat16)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    ss, sq = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    for rep in range(2):
        cu.lib.plume_debug_set_prof(ptr(prof) if rep == 1 else None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cu.conv3x3_fwd(x, wt, None, bias, 0, y, ss if stats else None, sq if stats else None)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * 9 * cin * cout * n * h * w
    p = prof.float().mean(0).tolist()
    tiles = (w // 8) * (h // 16) * n * max(1, cout // 256) / 148.0
    print(f"N{n} {h}x{w} {cin}->{cout} stats={stats}: {ms*1e3:.1f} us, {fl/ms/1e9:.0f} TFLOP/s, tiles/SM {tiles:.1f}")
    print("   " + ", ".join(f"{k}={v/1e3:.0f}k" for k, v in zip(names, p)) + f"  | cycles/tile {p[2]/tiles:.0f}")
cu.lib.plume_debug_set_prof(None)
