"""GPU parity, operator by operator: every C-ABI entry point (through ctypes) against the CPU oracle
(oracle/ops_ref.py) on the same seeded inputs.

Tolerances (written per test): bf16 outputs of a GEMM are compared to the oracle's bf16 outputs within
2 bf16 ulps of the tensor's max (fp32 accumulation order differs); fp32 reductions within 1e-3 relative;
integer / index / copy work (argmax, masks, packing, padding, tiles) bit-exact.
"""
import pytest
import torch

from oracle.ops_ref import RefOps

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from kcl_ltss_bioatm_b200.ops import CudaOps

    return CudaOps(), RefOps()


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def relmax(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert torch.isfinite(got).all()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


CONV_SHAPES = [  # N, H, W, Cin, Cout
    (1, 16, 16, 64, 64), (2, 16, 16, 64, 128), (1, 32, 32, 128, 256), (2, 8, 8, 256, 512),
    (3, 24, 40, 64, 64),    # ragged: partial tiles along w and h
    (1, 4, 4, 64, 64),      # one GEMM tile spans several images (and is mostly out of bounds)
    (5, 4, 4, 128, 64),
    (1, 16, 16, 1024, 256), (2, 48, 24, 128, 128),
    (2, 32, 16, 128, 64),   # 144 KB weight slice resident beside three halo slots (half-tile staging), kb = 2
    (1, 40, 24, 64, 128),   # the same for a 128-wide tile, ragged
]


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_SHAPES)
def test_conv3x3_fwd_epilogue_stats(ops, n, h, w, cin, cout):
    cu, rf = ops
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wt = (rnd(cout, 3, 3, cin, seed=2) / (9 * cin) ** 0.5).to(BF)
    scale, shift = 1 + 0.1 * rnd(cout, seed=3), 0.1 * rnd(cout, seed=4)
    for relu, stats in ((0, False), (1, True)):
        y_ref = torch.empty(n, h, w, cout, dtype=BF)
        ss_r, sq_r = torch.zeros(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64)
        rf.conv3x3_fwd(x, wt, scale, shift, relu, y_ref, ss_r if stats else None, sq_r if stats else None)
        y = torch.full((n, h, w, cout), float("nan"), dtype=BF, device=DEV)
        ss, sq = (torch.zeros(cout, dtype=torch.float64, device=DEV) for _ in range(2))
        cu.conv3x3_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), shift.to(DEV), relu, y,
                       ss if stats else None, sq if stats else None)
        torch.cuda.synchronize()
        assert relmax(y, y_ref) < 1.6e-2  # 2 bf16 ulps at the top of the range
        if stats:
            assert relmax(ss, ss_r) < 5e-3 and relmax(sq, sq_r) < 5e-3


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_SHAPES[:6] + CONV_SHAPES[-2:])
def test_conv3x3_concat_slices_and_dgrad(ops, n, h, w, cin, cout):
    cu, rf = ops
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wm = rnd(cout, 3, 3, cin, seed=5) / (9 * cout) ** 0.5
    wf_r, wd_r = torch.empty(cout, 3, 3, cin, dtype=BF), torch.empty(cin, 3, 3, cout, dtype=BF)
    rf.pack_conv3x3(wm, wf_r, wd_r)
    wf = torch.empty(cout, 3, 3, cin, dtype=BF, device=DEV)
    wd = torch.empty(cin, 3, 3, cout, dtype=BF, device=DEV)
    cu.pack_conv3x3(wm.to(DEV), wf, wd)
    torch.cuda.synchronize()
    assert torch.equal(wf.cpu(), wf_r) and torch.equal(wd.cpu(), wd_r)  # packing is bit-exact
    # input read from, and output written to, channel slices of wider buffers
    xb = torch.zeros(n, h, w, cin + 64, dtype=BF, device=DEV)
    xb[..., 64:] = x.to(DEV)
    yb = torch.zeros(n, h, w, cout + 64, dtype=BF, device=DEV)
    cu.conv3x3_fwd(xb[..., 64:], wf, None, None, 0, yb[..., :cout])
    torch.cuda.synchronize()
    y_ref = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x, wf_r, None, None, 0, y_ref)
    assert relmax(yb[..., :cout], y_ref) < 1.6e-2
    assert yb[..., cout:].abs().max().item() == 0  # nothing leaks outside the slice
    dy = rnd(n, h, w, cout, seed=6).to(BF)
    dx_ref = torch.empty(n, h, w, cin, dtype=BF)
    rf.conv3x3_dgrad(dy, wd_r, dx_ref)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=BF, device=DEV)
    cu.conv3x3_dgrad(dy.to(DEV), wd, dx)
    torch.cuda.synchronize()
    assert relmax(dx, dx_ref) < 1.6e-2


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (1, 16, 16, 64, 64), (2, 16, 16, 64, 128), (2, 16, 16, 128, 64), (1, 32, 32, 128, 256),
    (4, 8, 8, 256, 128), (3, 24, 40, 64, 64), (5, 4, 4, 128, 64), (8, 64, 64, 64, 64), (2, 16, 16, 512, 256),
    (3, 24, 40, 256, 256), (2, 32, 32, 1024, 128)])   # CTA pairs (Cin >= 256): ragged tiles, many channel slices
def test_conv3x3_wgrad_splitk(ops, n, h, w, cin, cout):
    _wgrad_splitk(ops, n, h, w, cin, cout)


@pytest.mark.parametrize("n,h,w,cin,cout", [(4, 8, 8, 256, 128), (3, 24, 40, 256, 256), (2, 32, 32, 1024, 128)])
def test_conv3x3_wgrad_cta_pairs(n, h, w, cin, cout):
    """The opt-in CTA-pair weight gradient (PLUME_WGRAD3_PAIR=1, read once per process: run in a child process)."""
    import os
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from tests.test_gpu_ops import _wgrad_splitk\n"
            "from kcl_ltss_bioatm_b200.ops import CudaOps\n"
            "from oracle.ops_ref import RefOps\n"
            "_wgrad_splitk((CudaOps(), RefOps()), %d, %d, %d, %d, %d)\nprint('pair wgrad ok')\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), n, h, w, cin, cout))
    env = dict(os.environ, PLUME_WGRAD3_PAIR="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "pair wgrad ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _wgrad_splitk(ops, n, h, w, cin, cout):
    cu, rf = ops
    x, dy = rnd(n, h, w, cin, seed=1).to(BF), rnd(n, h, w, cout, seed=2).to(BF)
    dw_ref = torch.zeros(cout, 3, 3, cin)
    rf.conv3x3_wgrad(x, dy, dw_ref)
    dw = torch.full((cout, 3, 3, cin), float("nan"), device=DEV)
    cu.conv3x3_wgrad(x.to(DEV), dy.to(DEV), dw)
    torch.cuda.synchronize()
    assert relmax(dw, dw_ref) < 1e-4  # fp32 accumulate of exact bf16 products: only summation order differs
    cu.conv3x3_wgrad(x.to(DEV), dy.to(DEV), dw, accumulate=True)
    torch.cuda.synchronize()
    assert relmax(dw, 2 * dw_ref) < 1e-4


def test_conv3x3_linearity_and_zero_at_full_size(ops):
    """Size-independent properties at BASELINE configs[1] layer size (32 x 256 x 256, 64 -> 64), where the
    CPU oracle would take too long: conv(0) == 0 exactly, conv(2x) == 2 conv(x) exactly (powers of two
    commute with bf16 rounding), and a checksum against the oracle on one image."""
    cu, rf = ops
    n, h, w, c = 32, 256, 256, 64
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(n, h, w, c, generator=g, device=DEV).to(BF)
    wt = (torch.randn(c, 3, 3, c, generator=g, device=DEV) / 24).to(BF)
    y0 = torch.full((n, h, w, c), float("nan"), dtype=BF, device=DEV)
    cu.conv3x3_fwd(torch.zeros_like(x), wt, None, None, 0, y0)
    assert y0.abs().max().item() == 0
    y1, y2 = torch.empty_like(y0), torch.empty_like(y0)
    cu.conv3x3_fwd(x, wt, None, None, 0, y1)
    cu.conv3x3_fwd(x * 2, wt, None, None, 0, y2)
    torch.cuda.synchronize()
    assert torch.equal(y2, y1 * 2)
    y_ref = torch.empty(1, h, w, c, dtype=BF)
    rf.conv3x3_fwd(x[7:8].cpu(), wt.cpu(), None, None, 0, y_ref)
    assert relmax(y1[7:8], y_ref) < 1.6e-2


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 8, 8, 128, 64), (2, 8, 8, 256, 128), (2, 4, 4, 512, 256),
                                            (3, 12, 20, 128, 64), (1, 16, 16, 1024, 512)])
def test_convT2x2_fwd_dgrad_wgrad(ops, n, h, w, cin, cout):
    cu, rf = ops
    x = rnd(n, h, w, cin, seed=1).to(BF)
    wm, bias = rnd(4, cout, cin, seed=2) / cin ** 0.5, 0.1 * rnd(cout, seed=3)
    wf_r, wd_r = torch.empty(4, cout, cin, dtype=BF), torch.empty(cin, 4, cout, dtype=BF)
    rf.pack_convT(wm, wf_r, wd_r)
    wf, wd = torch.empty(4, cout, cin, dtype=BF, device=DEV), torch.empty(cin, 4, cout, dtype=BF, device=DEV)
    cu.pack_convT(wm.to(DEV), wf, wd)
    torch.cuda.synchronize()
    assert torch.equal(wf.cpu(), wf_r) and torch.equal(wd.cpu(), wd_r)
    cat_r = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=BF)
    rf.convT_fwd(x, wf_r, bias, cat_r[..., cout:])
    cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=BF, device=DEV)
    cu.convT_fwd(x.to(DEV), wf, bias.to(DEV), cat[..., cout:])
    torch.cuda.synchronize()
    assert relmax(cat, cat_r) < 1.6e-2 and cat[..., :cout].abs().max().item() == 0
    du = rnd(n, 2 * h, 2 * w, 2 * cout, seed=4).to(BF)
    dud = du.to(DEV)
    dx_ref = torch.empty(n, h, w, cin, dtype=BF)
    rf.convT_dgrad(du[..., cout:], wd_r, dx_ref)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=BF, device=DEV)
    cu.convT_dgrad(dud[..., cout:], wd, dx)
    dw_ref = torch.zeros(4, cout, cin)
    rf.convT_wgrad(x, du[..., cout:], dw_ref)
    dw = torch.full((4, cout, cin), float("nan"), device=DEV)
    cu.convT_wgrad(x.to(DEV), dud[..., cout:], dw)
    torch.cuda.synchronize()
    assert relmax(dx, dx_ref) < 1.6e-2
    assert relmax(dw, dw_ref) < 1e-4


@pytest.mark.parametrize("n,h,w,c,ld", [(2, 16, 16, 64, 64), (3, 8, 12, 128, 256), (1, 4, 4, 1024, 1024),
                                        (2, 6, 10, 192, 192), (1, 2, 2, 4096, 4096)])
def test_bn_pool_relu_kernels(ops, n, h, w, c, ld):
    cu, rf = ops
    ybuf = rnd(n, h, w, ld, seed=1).to(BF)
    y, yd = ybuf[..., ld - c:], ybuf.to(DEV)[..., ld - c:]
    scale, shift = 1 + 0.2 * rnd(c, seed=2), 0.3 * rnd(c, seed=3)
    a_ref = torch.empty(n, h, w, c, dtype=BF)
    rf.scale_shift_act(y, scale, shift, 1, a_ref)
    a = torch.empty(n, h, w, c, dtype=BF, device=DEV)
    cu.scale_shift_act(yd, scale.to(DEV), shift.to(DEV), 1, a)
    torch.cuda.synchronize()
    assert relmax(a, a_ref) < 4e-3  # fma vs mul+add: at most 1 bf16 ulp apart
    skip_r = torch.zeros(n, h, w, 2 * c, dtype=BF)
    pooled_r = torch.empty(n, h // 2, w // 2, c, dtype=BF)
    am_r = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8)
    rf.scale_shift_act_pool(y, scale, shift, 1, skip_r[..., :c], pooled_r, am_r)
    skip = torch.zeros(n, h, w, 2 * c, dtype=BF, device=DEV)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=BF, device=DEV)
    am = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=DEV)
    cu.scale_shift_act_pool(yd, scale.to(DEV), shift.to(DEV), 1, skip[..., :c], pooled, am)
    torch.cuda.synchronize()
    assert relmax(skip, skip_r) < 4e-3 and relmax(pooled, pooled_r) < 4e-3
    # plain pool on identical inputs: values and argmax indices bit-exact (first maximum wins ties)
    p2, am2 = torch.empty_like(pooled), torch.empty_like(am)
    cu.maxpool_fwd(skip[..., :c], p2, am2)
    p2_r, am2_r = torch.empty_like(pooled_r), torch.empty_like(am_r)
    rf.maxpool_fwd(skip.cpu()[..., :c], p2_r, am2_r)
    torch.cuda.synchronize()
    assert torch.equal(p2.cpu(), p2_r) and torch.equal(am2.cpu(), am2_r)
    assert torch.equal(pooled, p2) and torch.equal(am, am2)  # fused and plain pool agree exactly
    dyp = rnd(n, h // 2, w // 2, c, seed=5).to(BF)
    dcat = rnd(n, h, w, 2 * c, seed=6).to(BF)
    dx_r = torch.empty(n, h, w, c, dtype=BF)
    dx = torch.empty(n, h, w, c, dtype=BF, device=DEV)
    rf.maxpool_bwd(dyp, am2_r, None, dx_r)
    cu.maxpool_bwd(dyp.to(DEV), am2, None, dx)
    torch.cuda.synchronize()
    assert torch.equal(dx.cpu(), dx_r)
    rf.maxpool_bwd(dyp, am2_r, dcat[..., :c], dx_r)
    cu.maxpool_bwd(dyp.to(DEV), am2, dcat.to(DEV)[..., :c], dx)
    torch.cuda.synchronize()
    assert relmax(dx, dx_r) < 4e-3
    # BatchNorm finalize / fold / backward
    cnt = n * h * w
    ss, sq = y.double().sum((0, 1, 2)), (y.double() ** 2).sum((0, 1, 2))
    gamma, beta = 1 + 0.1 * rnd(c, seed=7), 0.1 * rnd(c, seed=8)
    outs_r = [torch.zeros(c) for _ in range(4)]
    rm_r, rv_r = torch.zeros(c), torch.ones(c)
    rf.bn_finalize(ss, sq, cnt, gamma, beta, 1e-5, 0.1, rm_r, rv_r, *outs_r)
    outs = [torch.zeros(c, device=DEV) for _ in range(4)]
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    cu.bn_finalize(ss.to(DEV), sq.to(DEV), cnt, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, rm, rv, *outs)
    torch.cuda.synchronize()
    for o, o_r in zip(outs + [rm, rv], outs_r + [rm_r, rv_r]):
        assert relmax(o, o_r) < 1e-4
    fs_r, fh_r = torch.zeros(c), torch.zeros(c)
    rf.bn_fold_eval(gamma, beta, rm_r, rv_r, shift, 1e-5, fs_r, fh_r)
    fs, fh = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    cu.bn_fold_eval(gamma.to(DEV), beta.to(DEV), rm, rv, shift.to(DEV), 1e-5, fs, fh)
    torch.cuda.synchronize()
    assert relmax(fs, fs_r) < 1e-4 and relmax(fh, fh_r) < 1e-4
    da = rnd(n, h, w, c, seed=9).to(BF)
    dad = da.to(DEV)
    coef_d = [t.to(DEV) for t in outs_r]
    sg_r, sgx_r = torch.zeros(c), torch.zeros(c)
    rf.bn_bwd_reduce(da, y, *outs_r, 1, sg_r, sgx_r)
    sg, sgx = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    cu.bn_bwd_reduce(dad, yd, *coef_d, 1, sg, sgx)
    torch.cuda.synchronize()
    assert relmax(sg, sg_r) < 2e-3 and relmax(sgx, sgx_r) < 2e-3
    dy_r, sdy_r = torch.empty(n, h, w, c, dtype=BF), torch.zeros(c)
    rf.bn_bwd_apply(da, y, *outs_r, 1, sg_r, sgx_r, dy_r, sdy_r)
    dyo, sdy = torch.empty(n, h, w, c, dtype=BF, device=DEV), torch.zeros(c, device=DEV)
    # parameter-gradient hand-over: overwrite, then accumulate on top of what is there
    dgam, dbet = torch.full((c,), 3.0, device=DEV), torch.full((c,), -2.0, device=DEV)
    cu.bn_bwd_apply(dad, yd, *coef_d, 1, sg_r.to(DEV), sgx_r.to(DEV), dyo, sdy, dgam, dbet, False)
    torch.cuda.synchronize()
    assert torch.equal(dgam.cpu(), sgx_r) and torch.equal(dbet.cpu(), sg_r)
    cu.bn_bwd_apply(dad, yd, *coef_d, 1, sg_r.to(DEV), sgx_r.to(DEV), dyo, None, dgam, dbet, True)
    torch.cuda.synchronize()
    assert torch.equal(dgam.cpu(), 2 * sgx_r) and torch.equal(dbet.cpu(), 2 * sg_r)
    assert relmax(dyo, dy_r) < 4e-3
    assert (sdy.cpu() - sdy_r).abs().max().item() < 2e-2 * dy_r.float().abs().max().item() * cnt ** 0.5
    rdy_r, rs_r = torch.empty(n, h, w, c, dtype=BF), torch.zeros(c)
    rf.relu_bwd(da, a_ref, rdy_r, rs_r)
    rdy, rs = torch.empty(n, h, w, c, dtype=BF, device=DEV), torch.zeros(c, device=DEV)
    cu.relu_bwd(dad, a_ref.to(DEV), rdy, rs)
    cs_r, cs = torch.zeros(c), torch.zeros(c, device=DEV)
    rf.channel_sum(y, cs_r)
    cu.channel_sum(yd, cs)
    torch.cuda.synchronize()
    assert torch.equal(rdy.cpu(), rdy_r) and relmax(rs, rs_r) < 1e-3 and relmax(cs, cs_r) < 1e-3


def test_pad_channels_and_empty_inputs(ops):
    cu, rf = ops
    xin = rnd(2, 8, 8, 8, seed=1).to(BF)
    out = torch.full((2, 8, 8, 64), float("nan"), dtype=BF, device=DEV)
    cu.pad_channels(xin.to(DEV), out)
    out_r = torch.empty(2, 8, 8, 64, dtype=BF)
    rf.pad_channels(xin, out_r)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), out_r)
    # zero-pixel launches are no-ops, not errors
    lib = cu.lib
    from kcl_ltss_bioatm_b200.lib import ptr

    assert lib.plume_pad_channels(ptr(xin.to(DEV)), 8, ptr(out), 64, 0, None) == 0
    # misuse is reported, not swallowed
    assert lib.plume_pad_channels(ptr(out), 64, ptr(out), 8, 10, None) != 0
    assert b"multiples of 8" in lib.plume_last_error() or b"Cs <= Cd" in lib.plume_last_error()
    z = torch.zeros(1, 8, 8, 72, dtype=BF, device=DEV)
    with pytest.raises(Exception):
        cu.conv3x3_fwd(z, torch.zeros(64 * 9 * 72, dtype=BF, device=DEV), None, None, 0,
                       torch.zeros(1, 8, 8, 64, dtype=BF, device=DEV))  # Cin not a multiple of 64


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 8, 24, 128), (3, 4, 4, 64), (32, 64, 64, 64)])
def test_head_loss_fwd_bwd(ops, n, h, w, c):
    cu, rf = ops
    feat = rnd(n, h, w, c, seed=1).to(BF)
    wv, b = rnd(c, seed=2) / c ** 0.5, torch.tensor([0.05])
    tgt = (rnd(n, h, w, seed=3) > 0.8).to(torch.uint8)
    lg_r, sums_r, loss_r = torch.zeros(n, h, w), torch.zeros(4), torch.zeros(3)
    rf.head_fwd(feat, wv, b, tgt, lg_r, sums_r)
    rf.head_loss(sums_r, n * h * w, 1.0, 1.0, 1.0, loss_r)
    lg, sums, loss = (torch.zeros(n, h, w, device=DEV), torch.zeros(4, device=DEV), torch.zeros(3, device=DEV))
    cu.head_fwd(feat.to(DEV), wv.to(DEV), b.to(DEV), tgt.to(DEV), lg, sums)
    cu.head_loss(sums, n * h * w, 1.0, 1.0, 1.0, loss)
    torch.cuda.synchronize()
    assert relmax(lg, lg_r) < 1e-5 and relmax(sums, sums_r) < 1e-4 and relmax(loss, loss_r) < 1e-4
    df_r, dw_r, db_r = torch.empty(n, h, w, c, dtype=BF), torch.zeros(c), torch.zeros(1)
    rf.head_bwd(feat, wv, lg_r, tgt, sums_r, 1.0, 1.0, 1.0, 0.5, df_r, dw_r, db_r)
    df, dw, db = (torch.empty(n, h, w, c, dtype=BF, device=DEV), torch.zeros(c, device=DEV),
                  torch.zeros(1, device=DEV))
    cu.head_bwd(feat.to(DEV), wv.to(DEV), lg_r.to(DEV), tgt.to(DEV), sums_r.to(DEV), 1.0, 1.0, 1.0, 0.5, df, dw, db)
    torch.cuda.synchronize()
    assert relmax(df, df_r) < 8e-3 and relmax(dw, dw_r) < 1e-3 and relmax(db, db_r) < 1e-3


@pytest.mark.parametrize("n", [1 << 20, 12345, 3])
def test_adam_three_steps(ops, n):
    cu, rf = ops
    p, g = rnd(n, seed=1), rnd(n, seed=2)
    m, v = torch.zeros(n), torch.zeros(n)
    pd, gd, md, vd = p.to(DEV), g.to(DEV), m.to(DEV), v.to(DEV)
    for step in (1, 2, 3):
        rf.adam(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step, 0.5)
        cu.adam(pd, gd, md, vd, 1e-3, 0.9, 0.999, 1e-8, step, 0.5)
    torch.cuda.synchronize()
    assert relmax(pd, p) < 1e-6 and relmax(md, m) < 1e-6 and relmax(vd, v) < 1e-6


def test_tiles_extract_stitch_bit_exact(ops):
    cu, rf = ops
    from kcl_ltss_bioatm_b200.predict import tile_grid

    hs, ws, cs, T, margin = 100, 130, 8, 64, 8
    scene = rnd(hs, ws, cs, seed=1).to(BF)
    ys_l, xs_l = tile_grid(hs, ws, T, margin)
    ys, xs = torch.tensor(ys_l, dtype=torch.int32), torch.tensor(xs_l, dtype=torch.int32)
    k = ys.numel()
    tiles_r = torch.empty(k, T, T, 64, dtype=BF)
    rf.extract_tiles(scene, ys, xs, T, tiles_r)
    tiles = torch.full((k, T, T, 64), float("nan"), dtype=BF, device=DEV)
    cu.extract_tiles(scene.to(DEV), ys.to(DEV), xs.to(DEV), T, tiles)
    torch.cuda.synchronize()
    assert torch.equal(tiles.cpu(), tiles_r)
    logits = rnd(k, T, T, seed=2)
    mask_r, prob_r = torch.full((hs, ws), 7, dtype=torch.uint8), torch.zeros(hs, ws)
    rf.stitch_threshold(logits, ys, xs, T, margin, 0.0, mask_r, prob_r)
    mask, prob = torch.full((hs, ws), 7, dtype=torch.uint8, device=DEV), torch.zeros(hs, ws, device=DEV)
    cu.stitch_threshold(logits.to(DEV), ys.to(DEV), xs.to(DEV), T, margin, 0.0, mask, prob)
    torch.cuda.synchronize()
    assert (mask_r != 7).all()                     # every scene pixel owned by exactly one tile
    assert torch.equal(mask.cpu(), mask_r) and relmax(prob, prob_r) < 1e-5


def test_pack_batch_matches_per_layer_packing(ops):
    """One table-driven launch for several layers == the per-layer packers (bit-exact), incl. odd sizes,
    a layer without a dgrad layout and a transposed-conv layer."""
    cu, rf = ops
    jobs_r, jobs_c = [], []
    for i, (kind, cout, cin, need_wd) in enumerate([("conv3x3", 64, 64, False), ("conv3x3", 128, 64, True),
                                                    ("convT", 64, 128, True), ("conv3x3", 72, 40, True),
                                                    ("convT", 40, 24, True)]):
        if kind == "conv3x3":
            w = rnd(cout, 3, 3, cin, seed=20 + i)
            wf, wd = torch.empty(cout, 3, 3, cin, dtype=BF), torch.empty(cin, 3, 3, cout, dtype=BF)
        else:
            w = rnd(4, cout, cin, seed=20 + i)
            wf, wd = torch.empty(4, cout, cin, dtype=BF), torch.empty(cin, 4, cout, dtype=BF)
        jobs_r.append((kind, w, wf, wd if need_wd else None))
        jobs_c.append((kind, w.to(DEV), torch.zeros_like(wf, device=DEV),
                       torch.zeros_like(wd, device=DEV) if need_wd else None))
    rf.pack_batch(jobs_r)
    cu.pack_batch(jobs_c)
    cu.pack_batch(jobs_c)  # cached table path
    torch.cuda.synchronize()
    for (_, _, wf_r, wd_r), (_, _, wf, wd) in zip(jobs_r, jobs_c):
        assert torch.equal(wf.cpu(), wf_r)
        if wd_r is not None:
            assert torch.equal(wd.cpu(), wd_r)


@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (1, 8, 8), (1, 20, 12)])
def test_conv3x3_fewer_input_channels_than_weights(ops, n, h, w):
    """The 8-band input against first-layer weights zero-padded to 64 input channels (ldx < Cin): TMA
    zero-fills the missing channels in the forward and the weight-gradient kernels (halo and generic paths)."""
    cu, rf = ops
    cx, kcin, cout = 8, 64, 64
    x = rnd(n, h, w, cx, seed=3).to(BF)
    wm = torch.zeros(cout, 3, 3, kcin)
    wm[..., :cx] = rnd(cout, 3, 3, cx, seed=4) / (9 * cx) ** 0.5
    wf_r = wm.to(BF)
    y_ref = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x, wf_r, None, None, 1, y_ref)
    y = torch.empty(n, h, w, cout, dtype=BF, device=DEV)
    cu.conv3x3_fwd(x.to(DEV), wf_r.to(DEV), None, None, 1, y)
    torch.cuda.synchronize()
    assert relmax(y, y_ref) < 1.6e-2
    # padded copy gives the bit-identical result
    xp = torch.zeros(n, h, w, kcin, dtype=BF, device=DEV)
    xp[..., :cx] = x.to(DEV)
    y2 = torch.empty_like(y)
    cu.conv3x3_fwd(xp, wf_r.to(DEV), None, None, 1, y2)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    dy = rnd(n, h, w, cout, seed=6).to(BF)
    dw_ref = torch.empty(cout, 3, 3, kcin)
    rf.conv3x3_wgrad(x, dy, dw_ref)
    dw = torch.full((cout, 3, 3, kcin), 7.0, device=DEV)
    cu.conv3x3_wgrad(x.to(DEV), dy.to(DEV), dw)
    torch.cuda.synchronize()
    assert relmax(dw, dw_ref) < 1e-3
    assert dw[..., cx:].abs().max().item() == 0


@pytest.mark.parametrize("cin,cout,mode", [(64, 64, "<64,0> resident weights"), (128, 64, "<64,1> weight triples"),
                                           (128, 128, "<128,1> triples, half-tile staging"),
                                           (256, 256, "<256,2> single weight tiles")])
def test_conv3_repeated_launches_are_bit_identical(ops, cin, cout, mode):
    """200 launches of igemm_conv3_kernel on the same inputs give bit-identical outputs in every mode.  The kernel's
    relay-thread protocol (mbarrier observed by a relay, counter published through shared memory, MMAs issued behind
    a weak load of that counter) has no summation-order freedom (each issuer owns a TMEM tile, the epilogue adds
    them in a fixed order), so any run-to-run difference would be a race on operand visibility."""
    cu, rf = ops
    n, h, w = 3, 48, 40      # more tiles than SMs -> several tiles per CTA, ragged edges along both axes
    x = rnd(n, h, w, cin, seed=11).to(BF).to(DEV)
    wt = (rnd(cout, 3, 3, cin, seed=12) / (9 * cin) ** 0.5).to(BF).to(DEV)
    scale, shift = (1 + 0.1 * rnd(cout, seed=13)).to(DEV), (0.1 * rnd(cout, seed=14)).to(DEV)
    first = torch.empty(n, h, w, cout, dtype=BF, device=DEV)
    cu.conv3x3_fwd(x, wt, scale, shift, 1, first)
    y_ref = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x.cpu(), wt.cpu(), scale.cpu(), shift.cpu(), 1, y_ref)
    torch.cuda.synchronize()
    assert relmax(first, y_ref) < 1.6e-2
    y = torch.empty_like(first)
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    for i in range(200):
        y.fill_(float("nan"))
        cu.conv3x3_fwd(x, wt, scale, shift, 1, y)
        bad += (y.view(torch.int16) != first.view(torch.int16)).sum()
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, f"{mode}: {int(bad.item())} differing elements over 200 launches"


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (4, 48, 40, 256, 512), (2, 32, 32, 128, 256), (2, 64, 24, 512, 256),       # 256-wide ring of single weight tiles
    (4, 48, 40, 64, 64), (2, 40, 24, 128, 64),                                 # 64-wide, weights resident (kb = 1, 2)
    (2, 32, 32, 64, 128), (2, 48, 24, 128, 128),                               # 128-wide resident, full / half staging
    (2, 32, 32, 256, 128), (2, 32, 16, 256, 64)])                              # rings of weight triples
def test_conv3_cta_pairs_match_oracle_and_repeat_bit_identically(ops, n, h, w, cin, cout):
    """Tiles run as CTA pairs (tcgen05.mma.cta_group::2) when the M tiles pair up: forward with epilogue and
    BatchNorm statistics against the oracle, dgrad against the oracle, and 100 repeated launches bit-identical (both
    CTAs' TMA loads complete on the leader's barriers; a visibility race would show as run-to-run differences)."""
    cu, rf = ops
    x = rnd(n, h, w, cin, seed=21).to(BF)
    wt = (rnd(cout, 3, 3, cin, seed=22) / (9 * cin) ** 0.5).to(BF)
    scale, shift = 1 + 0.1 * rnd(cout, seed=23), 0.1 * rnd(cout, seed=24)
    y_ref = torch.empty(n, h, w, cout, dtype=BF)
    ss_r, sq_r = torch.zeros(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64)
    rf.conv3x3_fwd(x, wt, scale, shift, 1, y_ref, ss_r, sq_r)
    xd, wd_, scd, shd = x.to(DEV), wt.to(DEV), scale.to(DEV), shift.to(DEV)
    first = torch.full((n, h, w, cout), float("nan"), dtype=BF, device=DEV)
    ss, sq = (torch.zeros(cout, dtype=torch.float64, device=DEV) for _ in range(2))
    cu.conv3x3_fwd(xd, wd_, scd, shd, 1, first, ss, sq)
    torch.cuda.synchronize()
    assert relmax(first, y_ref) < 1.6e-2
    assert relmax(ss, ss_r) < 2e-3 and relmax(sq, sq_r) < 2e-3
    y = torch.empty_like(first)
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    for _ in range(100):
        y.fill_(float("nan"))
        cu.conv3x3_fwd(xd, wd_, scd, shd, 1, y)
        bad += (y.view(torch.int16) != first.view(torch.int16)).sum()
    torch.cuda.synchronize()
    assert int(bad.item()) == 0, f"{int(bad.item())} differing elements over 100 launches"
    # dgrad: cout -> cin through the rotated weights (cin >= 256 also runs as pairs)
    wf_r, wd_r = torch.empty(cout, 3, 3, cin, dtype=BF), torch.empty(cin, 3, 3, cout, dtype=BF)
    rf.pack_conv3x3(wt.float(), wf_r, wd_r)
    dy = rnd(n, h, w, cout, seed=25).to(BF)
    dx_ref = torch.empty(n, h, w, cin, dtype=BF)
    rf.conv3x3_dgrad(dy, wd_r, dx_ref)
    dx = torch.full((n, h, w, cin), float("nan"), dtype=BF, device=DEV)
    cu.conv3x3_dgrad(dy.to(DEV), wd_r.to(DEV), dx)
    torch.cuda.synchronize()
    assert relmax(dx, dx_ref) < 1.6e-2


def test_wgrad3_repeated_launches_agree_to_rounding(ops):
    """The split-K weight gradient adds its partial sums with fp32 atomics (order not fixed): repeated launches must
    agree to fp32 rounding of the sum, nothing more."""
    cu, _ = ops
    n, h, w, cin, cout = 4, 32, 32, 128, 128
    x, dy = rnd(n, h, w, cin, seed=1).to(BF).to(DEV), rnd(n, h, w, cout, seed=2).to(BF).to(DEV)
    first = torch.empty(cout, 3, 3, cin, device=DEV)
    cu.conv3x3_wgrad(x, dy, first)
    worst = torch.zeros((), device=DEV)
    dw = torch.empty_like(first)
    for _ in range(50):
        cu.conv3x3_wgrad(x, dy, dw)
        worst = torch.maximum(worst, (dw - first).abs().max())
    torch.cuda.synchronize()
    assert worst.item() <= 2e-5 * first.abs().max().item()


@pytest.mark.parametrize("ratio", [10.0, 100.0])
def test_batchnorm_statistics_survive_large_mean_over_std(ops, ratio):
    """A conv output population with |mean| / std = 10 and 100 (a bias-dominated channel): the batch statistics
    come from per-tile fp32 partial sums accumulated in fp64, and bn_finalize forms E[y^2] - E[y]^2 in fp64, so
    the variance keeps its digits (in fp32 the cross-tile sums alone lose ~4 % of it at ratio 100)."""
    cu, rf = ops
    n, h, w, cin, cout = 8, 64, 64, 64, 64
    x = rnd(n, h, w, cin, seed=21).to(BF)
    wt = (rnd(cout, 3, 3, cin, seed=22) / (9 * cin) ** 0.5).to(BF)
    y0 = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x, wt, None, None, 0, y0)
    std0 = y0.float().std().item()
    shift = torch.full((cout,), ratio * std0)                # conv bias: mean = ratio x std
    y_ref = torch.empty(n, h, w, cout, dtype=BF)
    rf.conv3x3_fwd(x, wt, None, shift, 0, y_ref)
    y = torch.empty(n, h, w, cout, dtype=BF, device=DEV)
    ss, sq = (torch.zeros(cout, dtype=torch.float64, device=DEV) for _ in range(2))
    cu.conv3x3_fwd(x.to(DEV), wt.to(DEV), None, shift.to(DEV), 0, y, ss, sq)
    outs = [torch.zeros(cout, device=DEV) for _ in range(4)]
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    cu.bn_finalize(ss, sq, n * h * w, None, None, 1e-5, 0.1, rm, rv, *outs)
    torch.cuda.synchronize()
    yd = y.double().cpu().reshape(-1, cout)                 # exact statistics of what the kernel stored
    mean_true, var_true = yd.mean(0), yd.var(0, unbiased=False)
    mean, invstd = outs[2].double().cpu(), outs[3].double().cpu()
    assert ((mean - mean_true).abs() / mean_true.abs()).max().item() < 1e-6
    assert ((1 / invstd ** 2 - 1e-5 - var_true).abs() / var_true).max().item() < 2e-3
    assert relmax(y, y_ref) < 1.6e-2


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 24, 64), (3, 8, 8, 256), (1, 64, 64, 128)])
def test_pool_and_head_backward_fused_with_bn_reduction(ops, n, h, w, c):
    """plume_maxpool2x2_bwd_bn / plume_head_bwd_bn: the gradient they store is bit-identical to the plain kernels',
    and the BatchNorm-backward sums they accumulate equal a separate bn_bwd_reduce over that stored gradient."""
    cu, _ = ops
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(n, h // 2, w // 2, c, generator=g).to(BF).to(DEV)
    am = torch.randint(0, 4, (n, h // 2, w // 2, c), generator=g, dtype=torch.uint8).to(DEV)
    dskip = torch.randn(n, h, w, c, generator=g).to(BF).to(DEV)
    y = torch.randn(n, h, w, c, generator=g).to(BF).to(DEV)
    scale, shift = (1 + 0.1 * rnd(c, seed=3)).to(DEV), (0.1 * rnd(c, seed=4)).to(DEV)
    mean, invstd = (0.1 * rnd(c, seed=5)).to(DEV), (1 + 0.1 * rnd(c, seed=6)).abs().to(DEV)
    dx0, dx1 = (torch.empty(n, h, w, c, dtype=BF, device=DEV) for _ in range(2))
    sg0, sx0, sg1, sx1 = (torch.zeros(c, device=DEV) for _ in range(4))
    cu.maxpool_bwd(dy, am, dskip, dx0)
    cu.bn_bwd_reduce(dx0, y, scale, shift, mean, invstd, 1, sg0, sx0)
    cu.maxpool_bwd(dy, am, dskip, dx1, bn=(y, scale, shift, mean, invstd, 1, sg1, sx1))
    torch.cuda.synchronize()
    assert torch.equal(dx0, dx1)
    assert relmax(sg1, sg0) < 1e-4 and relmax(sx1, sx0) < 1e-4
    if c <= 256:   # the head takes C = 8 * 2^k <= 256 feature channels
        feat = torch.randn(n, h, w, c, generator=g).to(BF).to(DEV)
        wh, logits = (0.1 * rnd(c, seed=7)).to(DEV), rnd(n, h, w, seed=8).to(DEV)
        target = (torch.rand(n, h, w, generator=g) > 0.7).to(torch.uint8).to(DEV)
        sums = torch.tensor([1.0, 2.0, 30.0, 40.0], device=DEV)
        outs = []
        for fused in (False, True):
            df = torch.empty(n, h, w, c, dtype=BF, device=DEV)
            dw, db, sg, sx = torch.zeros(c, device=DEV), torch.zeros(1, device=DEV), torch.zeros(c, device=DEV), \
                torch.zeros(c, device=DEV)
            if fused:
                cu.head_bwd(feat, wh, logits, target, sums, 1.0, 1.0, 1.0, 0.5, df, dw, db,
                            bn=(y, scale, shift, mean, invstd, 1, sg, sx))
            else:
                cu.head_bwd(feat, wh, logits, target, sums, 1.0, 1.0, 1.0, 0.5, df, dw, db)
                cu.bn_bwd_reduce(df, y, scale, shift, mean, invstd, 1, sg, sx)
            torch.cuda.synchronize()
            outs.append((df, dw, db, sg, sx))
        assert torch.equal(outs[0][0], outs[1][0])
        for a, b in zip(outs[0][1:], outs[1][1:]):
            assert relmax(b, a) < 1e-4
