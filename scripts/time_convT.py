"""Times the transposed-conv forward / dgrad / wgrad launches alone on the network's shapes (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.ops import CudaOps
ops = CudaOps()
for n, h, w, cin, cout in [(32, 128, 128, 128, 64), (32, 64, 64, 256, 128), (32, 32, 32, 512, 256), (32, 16, 16, 1024, 512)]:
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, device="cuda", dtype=torch.bfloat16)
    wf = (torch.randn(4, cout, cin, device="cuda") / cin ** 0.5).bfloat16()
    wd = wf.permute(2, 0, 1).contiguous()
    bias = torch.zeros(cout, device="cuda")
    dx = torch.empty_like(x)
    def t(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    f = t(lambda: ops.convT_fwd(x, wf, bias, cat[..., cout:]))
    d = t(lambda: ops.convT_dgrad(cat[..., cout:], wd, dx))
    print(f"convT {n}x{h}x{w} {cin}->{cout}: fwd {f:7.1f} us  dgrad {d:7.1f} us", flush=True)
