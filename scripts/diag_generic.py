"""Diagnostics: generic forward-type GEMM (small images) in both precisions over a grid of shapes."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from kcl_ltss_bioatm_b200.ops import CudaOps

DEV, BF = "cuda:0", torch.bfloat16


def rnd(*s, seed=0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed))


def split(x):
    hi = x.to(BF)
    return torch.stack([hi, (x - hi.float()).to(BF)], dim=-2).contiguous()


def join(t):
    t = t.float().cpu()
    return t[..., 0, :] + t[..., 1, :]


def l2(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / b.float().norm()).item()


for prec in ("bf16", "bf16x3"):
    ops = CudaOps(precision=prec)
    P = ops.planes
    for (n, h, w, cin, cout) in [(5, 4, 4, 128, 64), (5, 4, 4, 64, 128), (5, 4, 4, 64, 64), (5, 4, 4, 128, 128),
                                 (8, 4, 4, 64, 128), (1, 4, 4, 64, 128), (2, 8, 8, 64, 128), (3, 8, 8, 64, 128),
                                 (5, 4, 4, 64, 256), (2, 4, 4, 64, 128), (16, 4, 4, 64, 128), (9, 4, 4, 64, 128)]:
        x = rnd(n, h, w, cin, seed=1)
        wt = rnd(cout, 3, 3, cin, seed=2) / (9 * cin) ** 0.5
        wf = torch.zeros(P * wt.numel(), dtype=BF, device=DEV)
        wd = torch.zeros(P * wt.numel(), dtype=BF, device=DEV)
        ops.pack_batch([("conv3x3", wt.to(DEV), wf, wd)])
        if P == 2:
            xs = split(x).to(DEV)
            xq = join(xs)
            y = torch.full((n, h, w, 2, cout), float("nan"), dtype=BF, device=DEV)
        else:
            xs = x.to(BF).to(DEV)
            xq = xs.float().cpu()
            y = torch.full((n, h, w, cout), float("nan"), dtype=BF, device=DEV)
        wq = (wf.float().cpu()[:wt.numel()] + (wf.float().cpu()[wt.numel():] if P == 2 else 0)).view(cout, 3, 3, cin)
        ops.conv3x3_fwd(xs, wf, None, None, 0, y)
        torch.cuda.synchronize()
        ref = F.conv2d(xq.permute(0, 3, 1, 2).double(), wq.permute(0, 3, 1, 2).double(), padding=1).permute(0, 2, 3, 1)
        got = join(y) if P == 2 else y.float().cpu()
        print(f"{prec:7s} fwd n={n} {h}x{w} {cin}->{cout}: rel L2 {l2(got, ref):.3e}", flush=True)
