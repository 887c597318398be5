"""Times plume_conv3x3_wgrad alone (CUDA events, 20 launches after 3 warm-ups) on the network's layer shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kcl_ltss_bioatm_b200.ops import CudaOps

ops = CudaOps()
shapes = [(32, 256, 256, 64, 64), (32, 256, 256, 128, 64), (32, 128, 128, 128, 128), (32, 128, 128, 256, 128),
          (32, 64, 64, 256, 256), (32, 64, 64, 512, 256), (32, 32, 32, 512, 512), (32, 16, 16, 1024, 1024)]
if len(sys.argv) > 1:
    shapes = [shapes[int(a)] for a in sys.argv[1:]]
for n, h, w, cin, cout in shapes:
    x = torch.randn(n, h, w, cin, device="cuda").bfloat16()
    dy = torch.randn(n, h, w, cout, device="cuda").bfloat16()
    dw = torch.zeros(cout, 3, 3, cin, device="cuda")
    for _ in range(3):
        ops.conv3x3_wgrad(x, dy, dw, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv3x3_wgrad(x, dy, dw, True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * 9 * cin * cout * n * h * w
    print(f"wgrad {n}x{h}x{w} {cin}->{cout}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
