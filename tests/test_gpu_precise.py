"""GPU parity of the bf16x3 high-precision mode (north_star's "tf32 mode": forward logits within 1e-3 of the fp32
oracle) -- operators and the whole network, through the C ABI's *_x3 entry points.

Storage: a value is hi + lo (two bf16), activations are [N, H, W, 2, C]; the GEMMs run three bf16 MMA passes.
Because storage rounding is 2^-17 instead of 2^-9, this mode is also where the hand-written backward SCHEDULE is
pinned tightly on the GPU: every parameter gradient against fp32 autograd of the oracle network, per tensor.
"""
import pytest
import torch
import torch.nn.functional as F

from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec
from oracle.unet_ref import UNetRef, plume_loss

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def split(x: torch.Tensor) -> torch.Tensor:
    """fp32 [..., C] -> bf16 [..., 2, C] (hi, lo)."""
    hi = x.to(BF)
    lo = (x - hi.float()).to(BF)
    return torch.stack([hi, lo], dim=-2).contiguous()


def join(t: torch.Tensor) -> torch.Tensor:
    t = t.detach().float().cpu()
    return t[..., 0, :] + t[..., 1, :]


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def l2rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.fixture(scope="module")
def ops():
    from kcl_ltss_bioatm_b200.ops import CudaOps

    return CudaOps(precision="bf16x3")


def pack(ops, kind, w):
    """fp32 master weights -> (w_fwd, w_dgrad) in the hi|lo operand format."""
    wf = torch.zeros(2 * w.numel(), dtype=BF, device=DEV)
    wd = torch.zeros(2 * w.numel(), dtype=BF, device=DEV)
    ops.pack_batch([(kind, w.to(DEV).contiguous(), wf, wd)])
    return wf, wd


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 16, 64, 64), (2, 16, 16, 64, 128), (1, 32, 32, 128, 256),
                                            (3, 24, 40, 64, 64), (5, 4, 4, 128, 64), (1, 16, 16, 1024, 256)])
def test_conv3x3_x3_fwd_dgrad_wgrad(ops, n, h, w, cin, cout):
    x = rnd(n, h, w, cin, seed=1)
    wt = rnd(cout, 3, 3, cin, seed=2) / (9 * cin) ** 0.5
    scale, shift = 1 + 0.1 * rnd(cout, seed=3), 0.1 * rnd(cout, seed=4)
    xs, xq = split(x), join(split(x))                      # xq: what the device actually holds
    wf, wd = pack(ops, "conv3x3", wt)
    # forward with epilogue and statistics
    y = torch.full((n, h, w, 2, cout), float("nan"), dtype=BF, device=DEV)
    ss, sq = (torch.zeros(cout, dtype=torch.float64, device=DEV) for _ in range(2))
    ops.conv3x3_fwd(xs.to(DEV), wf, scale.to(DEV), shift.to(DEV), 1, y, ss, sq)
    ref = F.conv2d(nchw(xq).double(), wt.permute(0, 3, 1, 2).double(), padding=1)
    ref = (ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu().permute(0, 2, 3, 1).float()
    torch.cuda.synchronize()
    assert torch.isfinite(join(y)).all()
    assert l2rel(join(y), ref) < 1e-4, l2rel(join(y), ref)   # 4.5e-5 at K = 9 x 1024
    assert l2rel(ss, ref.double().sum(dim=(0, 1, 2))) < 1e-4
    assert l2rel(sq, (ref.double() ** 2).sum(dim=(0, 1, 2))) < 1e-4
    # dgrad
    dy = rnd(n, h, w, cout, seed=5)
    dys, dyq = split(dy), join(split(dy))
    dx = torch.full((n, h, w, 2, cin), float("nan"), dtype=BF, device=DEV)
    ops.conv3x3_dgrad(dys.to(DEV), wd, dx)
    xg = nchw(xq).double().requires_grad_(True)
    wg = wt.permute(0, 3, 1, 2).double().requires_grad_(True)
    F.conv2d(xg, wg, padding=1).backward(nchw(dyq).double())
    torch.cuda.synchronize()
    assert l2rel(join(dx), xg.grad.permute(0, 2, 3, 1).float()) < 1e-4
    # wgrad (fp32 output, accumulate on top of ones)
    dw = torch.ones(cout, 3, 3, cin, dtype=torch.float32, device=DEV)
    ops.conv3x3_wgrad(xs.to(DEV), dys.to(DEV), dw, True)
    torch.cuda.synchronize()
    assert l2rel(dw.cpu() - 1.0, wg.grad.permute(0, 2, 3, 1).float()) < 1e-4


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128), (3, 4, 4, 512, 256)])
def test_convT_x3_fwd_dgrad_wgrad(ops, n, h, w, cin, cout):
    x = rnd(n, h, w, cin, seed=11)
    wt = rnd(4, cout, cin, seed=12) / cin ** 0.5            # [ij][co][ci]
    bias = 0.1 * rnd(cout, seed=13)
    xs, xq = split(x), join(split(x))
    wf, wd = pack(ops, "convT", wt)
    cat = torch.zeros(n, 2 * h, 2 * w, 2, 2 * cout, dtype=BF, device=DEV)   # writes channels [cout, 2cout)
    ops.convT_fwd(xs.to(DEV), wf, bias.to(DEV), cat[..., cout:])
    wpt = wt.view(2, 2, cout, cin).permute(3, 2, 0, 1).double()             # ConvTranspose2d [Cin][Cout][i][j]
    xg = nchw(xq).double().requires_grad_(True)
    wg = wpt.clone().requires_grad_(True)
    out = F.conv_transpose2d(xg, wg, bias=bias.double(), stride=2)
    torch.cuda.synchronize()
    got = join(cat)
    assert torch.equal(got[..., :cout], torch.zeros_like(got[..., :cout]))  # the skip half is untouched
    assert l2rel(got[..., cout:], out.permute(0, 2, 3, 1).float()) < 3e-5
    du = rnd(n, 2 * h, 2 * w, cout, seed=14)
    dus, duq = split(du), join(split(du))
    gcat = torch.zeros(n, 2 * h, 2 * w, 2, 2 * cout, dtype=BF, device=DEV)
    gcat[..., cout:] = dus.to(DEV)
    out.backward(nchw(duq).double())
    dx = torch.full((n, h, w, 2, cin), float("nan"), dtype=BF, device=DEV)
    ops.convT_dgrad(gcat[..., cout:], wd, dx)
    dw = torch.zeros(4, cout, cin, dtype=torch.float32, device=DEV)
    ops.convT_wgrad(xs.to(DEV), gcat[..., cout:], dw, False)
    torch.cuda.synchronize()
    assert l2rel(join(dx), xg.grad.permute(0, 2, 3, 1).float()) < 3e-5
    assert l2rel(dw.cpu(), wg.grad.permute(2, 3, 1, 0).reshape(4, cout, cin).float()) < 3e-5


def test_bandwidth_kernels_x3_roundtrip_and_pool(ops):
    """scale/shift/ReLU, pool + argmax and their backward in the split format against fp32 torch."""
    n, h, w, c = 2, 8, 12, 64
    y = rnd(n, h, w, c, seed=21)
    scale, shift = 1 + 0.1 * rnd(c, seed=22), 0.1 * rnd(c, seed=23)
    ys, yq = split(y), join(split(y))
    a = torch.empty(n, h, w, 2, c, dtype=BF, device=DEV)
    ops.scale_shift_act(ys.to(DEV), scale.to(DEV), shift.to(DEV), 1, a)
    ref = (yq * scale + shift).relu()
    torch.cuda.synchronize()
    assert l2rel(join(a), ref) < 2e-5
    cat = torch.zeros(n, h, w, 2, 2 * c, dtype=BF, device=DEV)
    pooled = torch.empty(n, h // 2, w // 2, 2, c, dtype=BF, device=DEV)
    am = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=DEV)
    ops.scale_shift_act_pool(ys.to(DEV), scale.to(DEV), shift.to(DEV), 1, cat[..., :c], pooled, am)
    torch.cuda.synchronize()
    skip = join(cat)[..., :c]
    assert torch.equal(skip, join(a))                                        # same rounding as the unfused kernel
    pr, idx = F.max_pool2d(nchw(skip), 2, return_indices=True)
    assert torch.equal(join(pooled), pr.permute(0, 2, 3, 1))                 # pooling the stored values is exact
    # pad_channels: plain bf16 in, split out with an empty lo plane
    xin = rnd(n, h, w, 8, seed=24).to(BF)
    x0 = torch.full((n, h, w, 2, 64), float("nan"), dtype=BF, device=DEV)
    ops.pad_channels(xin.to(DEV), x0)
    torch.cuda.synchronize()
    assert torch.equal(x0.cpu()[..., 0, :8], xin) and x0.cpu()[..., 0, 8:].abs().max() == 0
    assert x0.cpu()[..., 1, :].abs().max() == 0


def _net_pair(spec, seed=0):
    from kcl_ltss_bioatm_b200.unet import UNetB200

    torch.manual_seed(seed)
    ref = UNetRef(spec).train()
    net = UNetB200(spec, device=DEV, seed=seed)
    return ref, net


@pytest.mark.parametrize("spec,n,hw", [(UNetSpec(base_filters=64, depth=2, precision="bf16x3"), 2, 64),
                                       (UNetSpec(precision="tf32"), 2, 128),
                                       (UNetSpec(precision="bf16x3"), 1, 256)])
def test_training_logits_within_1e_3_of_fp32_oracle(spec, n, hw):
    """north_star: forward logits within 1e-3 relative (tf32 mode).  Training-mode forward (batch statistics) at
    random init on the DEFAULT spec (23 layers), where the bf16 path's storage floor is 1.06e-2."""
    assert spec.precision == "bf16x3"
    ref, net = _net_pair(spec)
    x, t = synthetic_batch(n, hw, hw, spec.in_channels, seed=7)
    with torch.no_grad():
        z_ref = ref(nchw(x))[:, 0]
        loss_ref = float(plume_loss(z_ref, t, spec))
    z = net.forward(x.to(DEV), t.to(DEV))
    torch.cuda.synchronize()
    e2 = l2rel(z, z_ref)
    em = ((z.cpu() - z_ref).abs().max() / z_ref.abs().max()).item()
    print(f"bf16x3 train-mode logits vs fp32 oracle: rel L2 {e2:.3e}, max-norm {em:.3e}; "
          f"loss {net.loss_out[0].item():.6f} vs {loss_ref:.6f}")
    assert e2 <= 1e-3 and em <= 1e-3
    assert abs(net.loss_out[0].item() - loss_ref) <= 1e-4 * abs(loss_ref)


@pytest.mark.parametrize("spec,n,hw,bar", [
    (UNetSpec(norm="none", precision="bf16x3"), 2, 128, 1e-2),
    (UNetSpec(base_filters=64, depth=2, precision="bf16x3"), 4, 32, 2e-2),
    (UNetSpec(precision="bf16x3"), 2, 128, 8e-2)])
def test_every_gradient_matches_fp32_autograd(spec, n, hw, bar):
    """The whole hand-written backward schedule on the GPU against autograd of the fp32 oracle network, relative L2
    error per parameter tensor.  How tight the bar can be is set by the conditioning of the network, which the oracle
    itself shows (fp32 autograd vs fp64 autograd of the same network, measured on CPU):
      * without BatchNorm the gradients are well conditioned (fp32 vs fp64: 3.5e-6): measured 3.6e-3 worst tensor
        (the bottleneck, whose activations have shrunk to 1e-4 of the input without normalisation) and 2.7e-5 over
        all tensors together on the 23 layers of the default depth; bars 1e-2 / 1e-3 -- this pins every tap, the
        concat / pool / transposed-conv routing and the head (a wrong tap, a missing bucket or a mis-scaled term is
        O(0.1 - 1));
      * with BatchNorm at random init they are not (fp32 vs fp64: 8e-4 per tensor on the default spec, i.e. an
        amplification of ~1e4 of the unit roundoff), so a 16-bit storage format measures 7e-3 (depth 2) / 2.7e-2
        (default); the bars sit a factor 3 above that.  Conv biases in front of a BatchNorm have a mathematically
        zero gradient (both sides hold rounding noise) and are skipped there."""
    ref, net = _net_pair(spec)
    x, t = synthetic_batch(n, hw, hw, spec.in_channels, seed=3)
    plume_loss(ref(nchw(x))[:, 0], t, spec).backward()
    net.forward(x.to(DEV), t.to(DEV))
    net.backward()
    torch.cuda.synchronize()
    gd = net.grad_dict()
    worst_k, worst, num, den = None, 0.0, 0.0, 0.0
    for k, p in ref.named_parameters():
        if spec.norm == "batch" and (k.endswith("conv1.bias") or k.endswith("conv2.bias")):
            continue
        e = l2rel(gd[k], p.grad)
        num += float((gd[k].float().cpu() - p.grad).norm() ** 2)
        den += float(p.grad.norm() ** 2)
        if e > worst:
            worst_k, worst = k, e
    print(f"bf16x3 gradients vs fp32 autograd ({spec.norm}, depth {spec.depth}): worst per-tensor rel L2 {worst:.3e} "
          f"({worst_k}), all tensors together {(num / den) ** 0.5:.3e}")
    assert worst <= bar, (worst_k, worst)
    assert (num / den) ** 0.5 <= (1e-3 if spec.norm == "none" else bar / 2)


def test_eval_mask_and_tiled_scene_in_bf16x3():
    """Eval-mode forward (folded BatchNorm) and the tiled scene path in the split format."""
    from kcl_ltss_bioatm_b200.data import synthetic_scene
    from kcl_ltss_bioatm_b200.predict import ScenePredictor

    spec = UNetSpec(base_filters=64, depth=2, precision="bf16x3")
    ref, net = _net_pair(spec)
    for i in range(3):                                    # a few training steps so the running statistics move
        x, t = synthetic_batch(2, 64, 64, spec.in_channels, seed=100 + i)
        net.train_step(x.to(DEV), t.to(DEV))
    ref.load_state_dict(net.state_dict())
    ref.eval()
    x, _ = synthetic_batch(2, 64, 64, spec.in_channels, seed=55)
    with torch.no_grad():
        z_ref = ref(nchw(x))[:, 0]
    z = net.predict_logits(x.to(DEV))
    torch.cuda.synchronize()
    assert l2rel(z, z_ref) <= 1e-3, l2rel(z, z_ref)
    scene = synthetic_scene(200, 264, spec.in_channels, seed=5)
    mask = ScenePredictor(net, tile=64, margin=8, batch_tiles=8).predict_scene(scene.to(DEV))
    torch.cuda.synchronize()
    assert mask.dtype == torch.uint8 and tuple(mask.shape) == (200, 264)
