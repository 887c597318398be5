"""Times the BatchNorm streaming kernels alone on the network's tensor shapes (CUDA events, 30 launches)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.ops import CudaOps
ops = CudaOps()
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30 * 1e3
tot = [0.0, 0.0, 0.0]
for n, h, w, c in [(32, 256, 256, 64), (32, 128, 128, 128), (32, 64, 64, 256), (32, 32, 32, 512), (32, 16, 16, 1024)]:
    y = torch.randn(n, h, w, c, device="cuda").bfloat16(); da = torch.randn_like(y); out = torch.empty_like(y)
    sc = torch.ones(c, device="cuda"); sh = torch.zeros(c, device="cuda"); mu = torch.zeros(c, device="cuda"); inv = torch.ones(c, device="cuda")
    sg = torch.zeros(c, device="cuda"); sgx = torch.zeros(c, device="cuda"); sdy = torch.zeros(c, device="cuda")
    mb = y.numel() * 2 / 1e6
    a = t(lambda: ops.scale_shift_act(y, sc, sh, 1, out))
    r = t(lambda: ops.bn_bwd_reduce(da, y, sc, sh, mu, inv, 1, sg, sgx))
    p = t(lambda: ops.bn_bwd_apply(da, y, sc, sh, mu, inv, 1, sg, sgx, out, sdy))
    tot = [tot[0] + a, tot[1] + r, tot[2] + p]
    print(f"{n}x{h}x{w}x{c} ({mb:6.1f} MB): act {a:6.1f} us ({2*mb/a/1e3:5.2f} TB/s)  reduce {r:6.1f} us ({2*mb/r/1e3:5.2f})  apply {p:6.1f} us ({3*mb/p/1e3:5.2f})", flush=True)
print(f"sum over the five shapes: act {tot[0]:.1f} reduce {tot[1]:.1f} apply {tot[2]:.1f} us")
