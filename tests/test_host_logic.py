"""CPU: the host-side schedule (kcl_ltss_bioatm_b200.unet.UNetB200) driven by the CPU operator oracle,
against autograd of the whole-network oracle (oracle/unet_ref.py).

This checks everything the GPU cannot tell us cheaply: buffer wiring, zero-copy concat, the hand-written
backward schedule, parameter layout / state_dict conversions, BatchNorm bookkeeping and Adam -- with the
same bf16 rounding points the CUDA kernels use.  Two arithmetic modes of the operator oracle are used:
  * fp32 buffers: the hand-written schedule must reproduce autograd to fp32 accuracy (1e-4);
  * bf16 buffers (what the GPU does): logits within 1e-2 of the logit range (the north-star bound);
    parameter gradients are compared by direction (cosine >= 0.95) because BatchNorm's backward
    subtracts batch means from bf16-stored gradients and the weight gradient is a strongly cancelling
    sum, which amplifies rounding noise to ~10-20 % of a tensor's norm on a random, tiny batch.
"""
import copy

import pytest
import torch

from kcl_ltss_bioatm_b200.spec import UNetSpec, build_layout, fwd_flops_per_tile, train_flops_per_tile
from kcl_ltss_bioatm_b200.unet import UNetB200
from oracle.ops_ref import RefOps
from oracle.unet_ref import UNetRef, make_optimizer, plume_loss

BF = torch.bfloat16


def make_batch(n, h, w, c, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, c, generator=g).to(BF)
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    t = torch.zeros(n, h, w, dtype=torch.uint8)
    for i in range(n):
        cy, cx = torch.randint(0, h, (1,), generator=g).item(), torch.randint(0, w, (1,), generator=g).item()
        t[i] = (((yy - cy) ** 2 + 2 * (xx - cx) ** 2) < (h * w) / 12).to(torch.uint8)
    return x, t


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def cosine(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def test_schedule_is_exact_in_fp32():
    """No bf16 anywhere: forward, loss and every parameter gradient equal autograd's."""
    spec = UNetSpec(base_filters=64, depth=2)
    torch.manual_seed(0)
    ref = UNetRef(spec)
    net = UNetB200(spec, ops=RefOps(torch.float32), device="cpu", seed=0)
    x, t = make_batch(2, 16, 16, spec.in_channels, seed=1)
    x = x.float()
    logits_ref = ref(x.permute(0, 3, 1, 2))[:, 0]
    loss_ref = plume_loss(logits_ref, t, spec)
    loss_ref.backward()
    logits = net.forward(x, t)
    net.backward()
    assert rel(logits, logits_ref) < 1e-4
    assert abs(net.loss_out[0].item() - loss_ref.item()) < 1e-5
    gd = net.grad_dict()
    for k, p in ref.named_parameters():
        if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
            assert gd[k].abs().max() < 1e-6  # cancelled exactly by BatchNorm
            continue
        assert rel(gd[k], p.grad) < 1e-4, k
    # one Adam step lands on the same weights
    opt = make_optimizer(ref, spec)
    opt.step()
    net.optimizer_step()
    sd, sd_ref = net.state_dict(), ref.state_dict()
    for k, p in ref.named_parameters():
        if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
            continue  # sign of ~1e-9 noise decides a +-lr step under Adam
        big = p.grad.abs() > 1e-5  # below that, Adam's eps makes the step size sensitive to 1e-9 noise
        assert (sd[k] - sd_ref[k])[big].abs().max() < 2e-5, k


@pytest.fixture(scope="module")
def small():
    spec = UNetSpec(base_filters=64, depth=2)
    torch.manual_seed(0)
    ref = UNetRef(spec)
    net = UNetB200(spec, ops=RefOps(), device="cpu", seed=0)
    return spec, ref, net


def test_same_seed_gives_identical_weights(small):
    spec, ref, net = small
    sd_ref, sd = ref.state_dict(), net.state_dict()
    assert list(sd_ref.keys()) == list(sd.keys())
    for k in sd_ref:
        assert sd_ref[k].shape == sd[k].shape and sd_ref[k].dtype == sd[k].dtype, k
        assert torch.equal(sd_ref[k], sd[k]), k


def test_param_count_default_spec_matches_survey():
    # 31 046 401 logical parameters for the default spec (31 043 521 weights+biases
    # ... counted with BatchNorm affine); check through the layout without allocating device memory
    spec = UNetSpec()
    torch.manual_seed(0)
    n = sum(p.numel() for p in UNetRef(spec).parameters())
    assert n == 31_046_401
    assert abs(fwd_flops_per_tile(spec, 256, 256)["total"] / 1e9 - 96.7) < 0.2
    assert abs(train_flops_per_tile(spec, 256, 256) / 1e9 - 289.5) < 1.0
    lay = build_layout(spec)
    assert lay.order[0] == "head.weight" and lay.order[-1].startswith("enc0.")
    assert all(s.offset % 64 == 0 for s in lay.slots.values())


def test_forward_backward_step_match_autograd(small):
    spec, ref, net = small
    ref = copy.deepcopy(ref)
    net.load_state_dict(ref.state_dict())
    x, t = make_batch(2, 16, 16, spec.in_channels, seed=1)
    # ---- training forward
    ref.train()
    xr = x.float().permute(0, 3, 1, 2)
    logits_ref = ref(xr)[:, 0]
    loss_ref = plume_loss(logits_ref, t, spec)
    net.train()
    logits = net.forward(x, t)
    assert rel(logits, logits_ref) < 1e-2
    assert abs(net.loss_out[0].item() - loss_ref.item()) < 1e-2 * abs(loss_ref.item())
    # ---- backward
    opt = make_optimizer(ref, spec)
    opt.zero_grad()
    loss_ref.backward()
    net.backward()
    gd = net.grad_dict()
    worst = 0.0
    for k, p in ref.named_parameters():
        if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
            continue  # exactly cancelled by BatchNorm: both are rounding noise around zero
        c = cosine(gd[k], p.grad)
        assert c > 0.95, (k, c)
    # ---- optimizer step, BatchNorm running statistics
    opt.step()
    net.optimizer_step()
    sd, sd_ref = net.state_dict(), ref.state_dict()
    for k in sd_ref:
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(sd_ref[k])
        elif "running" in k:
            assert rel(sd[k], sd_ref[k]) < 1e-2, k
    # Adam moves every weight by ~lr in the gradient's sign: where the gradient is well above the bf16
    # noise floor, nearly all signs must agree
    for k, p in ref.named_parameters():
        if "bias" in k and "bn" not in k and k != "head.bias":
            continue
        g = p.grad
        big = g.abs() > 0.5 * g.abs().max()
        agree = (torch.sign(gd[k][big]) == torch.sign(g[big])).float().mean().item()
        assert agree > 0.98, (k, agree)


def test_eval_forward_uses_running_stats_and_matches(small):
    spec, ref, net = small
    ref = copy.deepcopy(ref)
    # make the running statistics non-trivial
    x0, _ = make_batch(2, 16, 16, spec.in_channels, seed=5)
    ref.train()
    with torch.no_grad():
        ref(x0.float().permute(0, 3, 1, 2))
    net.load_state_dict(ref.state_dict())
    x, _ = make_batch(3, 16, 32, spec.in_channels, seed=2)
    ref.eval()
    with torch.no_grad():
        lr = ref(x.float().permute(0, 3, 1, 2))[:, 0]
    lg = net.predict_logits(x)
    assert rel(lg, lr) < 1e-2
    m = net.predict_mask(x)
    agree = (m == (torch.sigmoid(lr) >= spec.mask_threshold).to(torch.uint8)).float().mean().item()
    assert agree >= 0.99  # tiny random-weight logits sit near the threshold; the GPU test uses trained weights
    assert m.dtype == torch.uint8 and tuple(m.shape) == (3, 16, 32)


def test_state_dict_roundtrip_is_bit_exact(small):
    spec, ref, net = small
    sd = net.state_dict()
    other = UNetB200(spec, ops=RefOps(), device="cpu", seed=None)
    other.load_state_dict(sd)
    sd2 = other.state_dict()
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k
    ref2 = UNetRef(spec)
    ref2.load_state_dict(sd)  # strict: every oracle key present, shapes agree


def test_norm_none_variant_matches_autograd():
    spec = UNetSpec(base_filters=64, depth=1, norm="none")
    torch.manual_seed(3)
    ref = UNetRef(spec)
    net = UNetB200(spec, ops=RefOps(), device="cpu", seed=3)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, net.state_dict()[k]), k
    x, t = make_batch(2, 8, 8, spec.in_channels, seed=4)
    logits_ref = ref(x.float().permute(0, 3, 1, 2))[:, 0]
    loss_ref = plume_loss(logits_ref, t, spec)
    loss_ref.backward()
    logits = net.forward(x, t)
    assert rel(logits, logits_ref) < 1e-2
    net.backward()
    gd = net.grad_dict()
    for k, p in ref.named_parameters():
        assert cosine(gd[k], p.grad) > 0.97, k


def test_gradient_accumulation_equals_big_batch_without_bn():
    spec = UNetSpec(base_filters=64, depth=1, norm="none")
    net = UNetB200(spec, ops=RefOps(), device="cpu", seed=0)
    x, t = make_batch(4, 8, 8, spec.in_channels, seed=7)
    net.forward(x, t)
    net.backward()
    full = net.grads.clone()
    # BCE is a mean over pixels, so two half batches at loss_scale 1/2 reproduce its gradient; Dice is a
    # ratio of batch sums and is excluded here
    spec2 = UNetSpec(base_filters=64, depth=1, norm="none", dice_weight=0.0)
    a = UNetB200(spec2, ops=RefOps(), device="cpu", seed=0)
    a.forward(x, t)
    a.backward()
    full = a.grads.clone()
    a.forward(x[:2], t[:2])
    a.backward(loss_scale=0.5)
    a.forward(x[2:], t[2:])
    a.backward(accumulate=True, loss_scale=0.5)
    assert rel(a.grads, full) < 2e-2


def test_bad_geometry_is_rejected(small):
    spec, _, net = small
    x = torch.zeros(1, 10, 16, spec.in_channels, dtype=BF)
    with pytest.raises(ValueError):
        net.forward(x)
    with pytest.raises(ValueError):
        net.forward(torch.zeros(1, 16, 16, spec.in_channels + 8, dtype=BF))


def test_batchnorm_micro_batches_second_slice_equals_standalone_backward():
    """With BatchNorm, backward(accumulate=True) of slice 2 must add exactly the gradient a standalone backward of
    slice 2 produces: the BatchNorm-backward sums of a slice are per-backward scratch, never the (accumulating)
    dgamma / dbeta slots.  fp32 operator oracle, so the identity is exact to rounding."""
    spec = UNetSpec(base_filters=64, depth=2)
    x, t = make_batch(4, 16, 16, spec.in_channels, seed=11)
    x = x.float()

    def grads_of(slices):
        net = UNetB200(spec, ops=RefOps(torch.float32), device="cpu", seed=0)
        for i, sl in enumerate(slices):
            net.forward(x[sl], t[sl])
            net.backward(accumulate=i > 0, loss_scale=0.5)
        return net.grads.clone()

    g1, g2 = grads_of([slice(0, 2)]), grads_of([slice(2, 4)])
    both = grads_of([slice(0, 2), slice(2, 4)])
    assert rel(both, g1 + g2) < 1e-5
    # and against autograd of the oracle doing the same two half-batch passes
    torch.manual_seed(0)
    ref = UNetRef(spec).train()
    for sl in (slice(0, 2), slice(2, 4)):
        (0.5 * plume_loss(ref(x[sl].permute(0, 3, 1, 2))[:, 0], t[sl], spec)).backward()
    net = UNetB200(spec, ops=RefOps(torch.float32), device="cpu", seed=0)
    net.grads.copy_(both)
    gd = net.grad_dict()
    for k, p in ref.named_parameters():
        if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
            continue
        assert rel(gd[k], p.grad) < 1e-4, k


def test_tile_file_loader_visits_every_tile_and_validates(tmp_path):
    """src/models/train_model.py's file-backed loader: tiles are indexed globally across files (none skipped), every
    rank gets `per_rank` tiles, shapes are validated with the offending path in the message."""
    from src.models.train_model import TileFiles

    spec = UNetSpec(base_filters=64, depth=2)
    counts, base = [5, 2, 3], 0
    for i, n in enumerate(counts):
        xs = torch.zeros(n, 8, 8, spec.in_channels)
        for k in range(n):
            xs[k] = base + k          # the tile's global index
        torch.save({"x": xs, "mask": torch.zeros(n, 8, 8, dtype=torch.uint8)}, tmp_path / f"tiles_{i}.pt")
        base += n
    files = sorted(str(p) for p in tmp_path.glob("*.pt"))
    idx = TileFiles(files, spec)
    assert idx.total == 10
    seen = []
    for it in range(3):
        for rank in range(2):
            xb, mb = idx.batch(it, 2, rank, 2)
            assert xb.shape == (2, 8, 8, spec.in_channels) and xb.dtype == BF and mb.dtype == torch.uint8
            seen += [int(v) for v in xb[:, 0, 0, 0].float().tolist()]
    assert seen == [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1]          # consecutive, wrapping, nothing skipped
    torch.save({"x": torch.zeros(1, 8, 8, 3), "mask": torch.zeros(1, 8, 8, dtype=torch.uint8)}, tmp_path / "zz_bad.pt")
    with pytest.raises(ValueError, match="zz_bad.pt"):
        TileFiles(files + [str(tmp_path / "zz_bad.pt")], spec)
    torch.save({"x": torch.zeros(1, 6, 8, spec.in_channels), "mask": torch.zeros(1, 6, 8, dtype=torch.uint8)},
               tmp_path / "zz_odd.pt")
    with pytest.raises(ValueError, match="zz_odd.pt"):
        TileFiles([str(tmp_path / "zz_odd.pt")], spec)
