"""CPU: the per-operator oracle (oracle/ops_ref.py) against torch autograd / nn.functional definitions.

The reference repository pins nothing for this path (no model code, no tests: SURVEY.md sections 0, 4),
so the oracle's backward operators are pinned against autograd of its own forward definitions.
"""
import torch
import torch.nn.functional as F

from oracle.ops_ref import RefOps

rf = RefOps()
BF = torch.bfloat16


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def close(a, b, tol):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item() <= tol * (b.abs().max().item() + 1e-12)


def test_conv3x3_fwd_dgrad_wgrad_match_autograd():
    n, h, w, cin, cout = 2, 6, 10, 16, 24
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wm = rnd(cout, 3, 3, cin, seed=2) * 0.1
    wf = torch.empty(cout, 3, 3, cin, dtype=BF)
    wd = torch.empty(cin, 3, 3, cout, dtype=BF)
    rf.pack_conv3x3(wm, wf, wd)
    y = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x, wf, None, None, 0, y)
    xx = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ww = wf.float().permute(0, 3, 1, 2).requires_grad_(True)
    out = F.conv2d(xx, ww, padding=1)
    assert close(y, out.permute(0, 2, 3, 1), 1e-2)
    dy = rnd(n, h, w, cout, seed=3).to(BF)
    out.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.empty(n, h, w, cin, dtype=BF)
    rf.conv3x3_dgrad(dy, wd, dx)
    assert close(dx, xx.grad.permute(0, 2, 3, 1), 1e-2)
    dw = torch.zeros(cout, 3, 3, cin)
    rf.conv3x3_wgrad(x, dy, dw)
    assert close(dw, ww.grad.permute(0, 2, 3, 1), 1e-4)
    rf.conv3x3_wgrad(x, dy, dw, accumulate=True)
    assert close(dw, 2 * ww.grad.permute(0, 2, 3, 1), 1e-4)


def test_conv3x3_epilogue_and_stats():
    n, h, w, cin, cout = 1, 4, 4, 8, 8
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wf = (rnd(cout, 3, 3, cin, seed=2) * 0.2).to(BF)
    scale, shift = 1 + 0.1 * rnd(cout, seed=3), 0.1 * rnd(cout, seed=4)
    y = torch.empty(n, h, w, cout, dtype=BF)
    ss, sq = torch.zeros(cout), torch.zeros(cout)
    rf.conv3x3_fwd(x, wf, scale, shift, 1, y, ss, sq)
    assert (y.float() >= 0).all()
    assert close(ss, y.float().sum((0, 1, 2)), 1e-6) and close(sq, (y.float() ** 2).sum((0, 1, 2)), 1e-6)


def test_convT_matches_autograd():
    n, h, w, cin, cout = 2, 3, 5, 16, 8
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wm = rnd(4, cout, cin, seed=2) * 0.2
    bias = 0.1 * rnd(cout, seed=3)
    wf = torch.empty(4, cout, cin, dtype=BF)
    wd = torch.empty(cin, 4, cout, dtype=BF)
    rf.pack_convT(wm, wf, wd)
    cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=BF)
    rf.convT_fwd(x, wf, bias, cat[..., cout:])
    assert cat[..., :cout].abs().max() == 0
    # definition check: u[n,2h+i,2w+j,co] = sum_ci x[n,h,w,ci] w[ij][co][ci] + b
    u = torch.einsum("nhwc,ijoc->nhiwjo", x.float(), wf.float().view(2, 2, cout, cin)).reshape(n, 2 * h, 2 * w, cout) + bias
    assert close(cat[..., cout:], u, 1e-2)
    xx = x.float().requires_grad_(True)
    wv = wf.float().view(2, 2, cout, cin).requires_grad_(True)
    uu = torch.einsum("nhwc,ijoc->nhiwjo", xx, wv).reshape(n, 2 * h, 2 * w, cout)
    du = rnd(n, 2 * h, 2 * w, cout, seed=4).to(BF)
    uu.backward(du.float())
    dx = torch.empty(n, h, w, cin, dtype=BF)
    rf.convT_dgrad(du, wd, dx)
    assert close(dx, xx.grad, 1e-2)
    dw = torch.zeros(4, cout, cin)
    rf.convT_wgrad(x, du, dw)
    assert close(dw, wv.grad.reshape(4, cout, cin), 1e-4)


def test_pool_matches_torch_and_bwd_is_adjoint():
    n, h, w, c = 2, 6, 8, 16
    x = rnd(n, h, w, c, seed=1).to(BF)
    y = torch.empty(n, h // 2, w // 2, c, dtype=BF)
    am = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8)
    rf.maxpool_fwd(x, y, am)
    ref, idx = F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, return_indices=True)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    # argmax position agrees with torch's flat index
    ii = (idx // w) % 2
    jj = (idx % w) % 2
    assert torch.equal(am.long(), (ii * 2 + jj).permute(0, 2, 3, 1))
    dy = rnd(n, h // 2, w // 2, c, seed=2).to(BF)
    dskip = rnd(n, h, w, c, seed=3).to(BF)
    dx = torch.empty(n, h, w, c, dtype=BF)
    rf.maxpool_bwd(dy, am, None, dx)
    xx = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.max_pool2d(xx, 2).backward(dy.float().permute(0, 3, 1, 2))
    assert torch.equal(dx.float(), xx.grad.permute(0, 2, 3, 1))
    rf.maxpool_bwd(dy, am, dskip, dx)
    assert close(dx, xx.grad.permute(0, 2, 3, 1) + dskip.float(), 1e-2)


def test_pool_ties_pick_first():
    x = torch.zeros(1, 2, 2, 8, dtype=BF)
    y = torch.empty(1, 1, 1, 8, dtype=BF)
    am = torch.full((1, 1, 1, 8), 9, dtype=torch.uint8)
    rf.maxpool_fwd(x, y, am)
    assert (am == 0).all()


def test_bn_train_fwd_bwd_matches_autograd():
    n, h, w, c = 3, 4, 6, 16
    y = rnd(n, h, w, c, seed=1).to(BF)
    gamma, beta = 1 + 0.1 * rnd(c, seed=2), 0.1 * rnd(c, seed=3)
    cnt = n * h * w
    ss, sq = y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))
    scale, shift, mean, invstd = (torch.zeros(c) for _ in range(4))
    rm, rv = torch.zeros(c), torch.ones(c)
    rf.bn_finalize(ss, sq, cnt, gamma, beta, 1e-5, 0.1, rm, rv, scale, shift, mean, invstd)
    a = torch.empty(n, h, w, c, dtype=BF)
    rf.scale_shift_act(y, scale, shift, 1, a)
    bn = torch.nn.BatchNorm2d(c, eps=1e-5, momentum=0.1)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    yy = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    out = bn(yy).relu()
    assert close(a, out.permute(0, 2, 3, 1), 1e-2)
    assert close(rm, bn.running_mean, 1e-5) and close(rv, bn.running_var, 1e-5)
    da = rnd(n, h, w, c, seed=4).to(BF)
    out.backward(da.float().permute(0, 3, 1, 2))
    sg, sgx = torch.zeros(c), torch.zeros(c)
    rf.bn_bwd_reduce(da, y, scale, shift, mean, invstd, 1, sg, sgx)
    dy = torch.empty(n, h, w, c, dtype=BF)
    sdy = torch.zeros(c)
    rf.bn_bwd_apply(da, y, scale, shift, mean, invstd, 1, sg, sgx, dy, sdy)
    assert close(dy, yy.grad.permute(0, 2, 3, 1), 1e-2)
    assert close(sgx, bn.weight.grad, 1e-4) and close(sg, bn.bias.grad, 1e-4)


def test_bn_fold_eval_matches_module():
    c = 8
    bn = torch.nn.BatchNorm2d(c).eval()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * rnd(c, seed=1))
        bn.bias.copy_(0.1 * rnd(c, seed=2))
        bn.running_mean.copy_(rnd(c, seed=3))
        bn.running_var.copy_(rnd(c, seed=4).abs() + 0.5)
    cb = rnd(c, seed=5)
    scale, shift = torch.zeros(c), torch.zeros(c)
    rf.bn_fold_eval(bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, cb, bn.eps, scale, shift)
    z = rnd(2, c, 3, 3, seed=6)
    ref = bn(z + cb.view(1, -1, 1, 1))
    assert close(z * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), ref, 1e-5)


def test_head_loss_and_grads_match_autograd():
    n, h, w, c = 2, 5, 7, 16
    feat = rnd(n, h, w, c, seed=1).to(BF)
    wv, b = rnd(c, seed=2) * 0.3, torch.tensor([0.1])
    tgt = (rnd(n, h, w, seed=3) > 0.5).to(torch.uint8)
    lg, sums, loss = torch.zeros(n, h, w), torch.zeros(4), torch.zeros(3)
    rf.head_fwd(feat, wv, b, tgt, lg, sums)
    rf.head_loss(sums, n * h * w, 1.0, 1.0, 1.0, loss)
    ff = feat.float().requires_grad_(True)
    ww = wv.clone().requires_grad_(True)
    bb = b.clone().requires_grad_(True)
    z = ff @ ww + bb
    t = tgt.float()
    p = torch.sigmoid(z)
    ref = F.binary_cross_entropy_with_logits(z, t) + 1 - (2 * (p * t).sum() + 1) / (p.sum() + t.sum() + 1)
    assert abs(loss[0].item() - ref.item()) < 1e-5
    (0.5 * ref).backward()
    df, dw, db = torch.empty(n, h, w, c, dtype=BF), torch.zeros(c), torch.zeros(1)
    rf.head_bwd(feat, wv, lg, tgt, sums, 1.0, 1.0, 1.0, 0.5, df, dw, db)
    assert close(df, ff.grad, 1e-2) and close(dw, ww.grad, 1e-4) and close(db, bb.grad, 1e-4)


def test_adam_matches_torch_optim():
    p0, g = rnd(100, seed=1), rnd(100, seed=2)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    q, m, v = p0.clone(), torch.zeros(100), torch.zeros(100)
    for step in (1, 2, 3):
        p.grad = g.clone() * step
        opt.step()
        rf.adam(q, g * step, m, v, 1e-3, 0.9, 0.999, 1e-8, step)
    assert close(q, p.data, 1e-6)


def test_tiles_roundtrip_covers_scene_once():
    hs, ws, T, margin = 70, 90, 32, 4
    stride = T - 2 * margin
    ys = list(range(0, hs - 2 * margin, stride))
    xs = list(range(0, ws - 2 * margin, stride))
    yy = torch.tensor([y for y in ys for _ in xs], dtype=torch.int32)
    xx = torch.tensor([x for _ in ys for x in xs], dtype=torch.int32)
    scene = rnd(hs, ws, 8, seed=1).to(BF)
    tiles = torch.empty(yy.numel(), T, T, 16, dtype=BF)
    rf.extract_tiles(scene, yy, xx, T, tiles)
    # "logits" = channel 0 of each tile: stitching must reproduce the scene's channel 0 exactly
    logits = tiles[..., 0].float().contiguous()
    mask = torch.full((hs, ws), 7, dtype=torch.uint8)
    prob = torch.full((hs, ws), -1.0)
    rf.stitch_threshold(logits, yy, xx, T, margin, 0.0, mask, prob)
    assert (mask != 7).all()
    assert torch.equal(mask, (scene[..., 0].float() >= 0).to(torch.uint8))
    assert close(prob, torch.sigmoid(scene[..., 0].float()), 1e-6)
