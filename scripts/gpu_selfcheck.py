"""Verbose per-operator parity run on a GPU box: CUDA kernels (through the C ABI) vs the CPU oracle.

Usage: python scripts/gpu_selfcheck.py [group ...]     groups: conv wgrad convT bw head adam tiles
Each group should be run in its own process (a kernel trap poisons the CUDA context):
    for g in conv wgrad convT bw head adam tiles; do timeout 300 python scripts/gpu_selfcheck.py $g; done
This is a diagnostic tool (it prints error structure); the pass/fail gate is tests/test_gpu_*.py.
"""
from __future__ import annotations

import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from kcl_ltss_bioatm_b200.ops import CudaOps  # noqa: E402
from oracle.ops_ref import RefOps  # noqa: E402

DEV = "cuda:0"
FAILS = []


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def report(name, got, ref, tol=2e-2, detail=True):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    mx = err.max().item()
    rel = mx / denom
    bad = (err > tol * denom).float().mean().item()
    ok = rel <= tol and torch.isfinite(got).all().item()
    print(f"[{'ok' if ok else 'FAIL'}] {name}: max_abs={mx:.4g} rel_to_max={rel:.4g} frac_bad={bad:.4g} "
          f"ref_absmax={denom:.4g} got_absmax={got.abs().max().item():.4g}", flush=True)
    if not ok:
        FAILS.append(name)
        if detail and got.dim() == 4:
            e = err / denom
            print("   err by channel chunk of 8:", [round(v, 3) for v in
                  e.mean(dim=(0, 1, 2)).view(-1, 8).mean(1)[:16].tolist()])
            print("   err by w (first 20):", [round(v, 3) for v in e.mean(dim=(0, 1, 3))[:20].tolist()])
            print("   err by h (first 20):", [round(v, 3) for v in e.mean(dim=(0, 2, 3))[:20].tolist()])
            print("   err by n:", [round(v, 3) for v in e.mean(dim=(1, 2, 3))[:8].tolist()])
            print("   got[0,0,0,:8]", got[0, 0, 0, :8].tolist())
            print("   ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
            print("   got[0,1,1,:8]", got[0, 1, 1, :8].tolist())
            print("   ref[0,1,1,:8]", ref[0, 1, 1, :8].tolist())
    return ok


def sync(ops, what):
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"[FAIL] {what}: CUDA error after launch: {e}; debug_word={ops.lib.plume_debug_word():#x}", flush=True)
        raise


def group_conv(cu, rf):
    shapes = [  # N, H, W, Cin, Cout
        (1, 16, 16, 64, 64),
        (2, 16, 16, 64, 128),
        (1, 32, 32, 128, 256),
        (2, 8, 8, 256, 512),
        (3, 24, 40, 64, 64),     # ragged: partial tiles in w and h
        (1, 4, 4, 64, 64),       # tile spans several images
        (5, 4, 4, 128, 64),
    ]
    for (n, h, w, cin, cout) in shapes:
        x = rnd(n, h, w, cin, seed=1).to(torch.bfloat16)
        wt = (rnd(cout, 3, 3, cin, seed=2) * (1.0 / (9 * cin) ** 0.5)).to(torch.bfloat16)
        scale = (1 + 0.1 * rnd(cout, seed=3)).float()
        shift = (0.1 * rnd(cout, seed=4)).float()
        for relu, stats in ((0, False), (1, True)):
            y_ref = torch.empty(n, h, w, cout, dtype=torch.bfloat16)
            ss_r, sq_r = torch.zeros(cout, dtype=torch.float64), torch.zeros(cout, dtype=torch.float64)
            rf.conv3x3_fwd(x, wt, scale, shift, relu, y_ref, ss_r if stats else None, sq_r if stats else None)
            y = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            ss, sq = (torch.zeros(cout, dtype=torch.float64, device=DEV) for _ in range(2))
            cu.conv3x3_fwd(x.to(DEV), wt.to(DEV), scale.to(DEV), shift.to(DEV), relu, y,
                           ss if stats else None, sq if stats else None)
            sync(cu, "conv3x3_fwd")
            report(f"conv3x3_fwd N{n} {h}x{w} {cin}->{cout} relu={relu}", y, y_ref)
            if stats:
                report("   stats sum", ss, ss_r, tol=5e-3, detail=False)
                report("   stats sq", sq, sq_r, tol=5e-3, detail=False)
        # strided (concat-slice) input and output
        xb = torch.zeros(n, h, w, cin + 64, dtype=torch.bfloat16, device=DEV)
        xb[..., 64:] = x.to(DEV)
        yb = torch.zeros(n, h, w, cout + 64, dtype=torch.bfloat16, device=DEV)
        cu.conv3x3_fwd(xb[..., 64:], wt.to(DEV), None, None, 0, yb[..., :cout])
        sync(cu, "conv3x3_fwd strided")
        y_ref = torch.empty(n, h, w, cout, dtype=torch.bfloat16)
        rf.conv3x3_fwd(x, wt, None, None, 0, y_ref)
        report(f"conv3x3_fwd strided N{n} {h}x{w} {cin}->{cout}", yb[..., :cout], y_ref)
        assert yb[..., cout:].abs().max().item() == 0, "store leaked outside the channel slice"
        # dgrad through the packer
        wm = rnd(cout, 3, 3, cin, seed=5) * (1.0 / (9 * cout) ** 0.5)
        wf_r = torch.empty(cout, 3, 3, cin, dtype=torch.bfloat16)
        wd_r = torch.empty(cin, 3, 3, cout, dtype=torch.bfloat16)
        rf.pack_conv3x3(wm, wf_r, wd_r)
        wf = torch.empty(cout, 3, 3, cin, dtype=torch.bfloat16, device=DEV)
        wd = torch.empty(cin, 3, 3, cout, dtype=torch.bfloat16, device=DEV)
        cu.pack_conv3x3(wm.to(DEV), wf, wd)
        sync(cu, "pack_conv3x3")
        report("   pack fwd", wf, wf_r, tol=0, detail=False)
        report("   pack dgrad", wd, wd_r, tol=0, detail=False)
        dy = rnd(n, h, w, cout, seed=6).to(torch.bfloat16)
        dx_ref = torch.empty(n, h, w, cin, dtype=torch.bfloat16)
        rf.conv3x3_dgrad(dy, wd_r, dx_ref)
        # cross-check the oracle's dgrad definition against autograd
        xx = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        out = torch.nn.functional.conv2d(xx, wf_r.float().permute(0, 3, 1, 2), padding=1)
        out.backward(dy.float().permute(0, 3, 1, 2))
        report("   oracle dgrad vs autograd", dx_ref, xx.grad.permute(0, 2, 3, 1), tol=1e-2, detail=False)
        dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
        cu.conv3x3_dgrad(dy.to(DEV), wd, dx)
        sync(cu, "conv3x3_dgrad")
        report(f"conv3x3_dgrad N{n} {h}x{w} {cin}<-{cout}", dx, dx_ref)


def group_wgrad(cu, rf):
    shapes = [
        (1, 16, 16, 64, 64),
        (2, 16, 16, 64, 128),
        (2, 16, 16, 128, 64),
        (1, 32, 32, 128, 256),
        (4, 8, 8, 256, 128),
        (3, 24, 40, 64, 64),
        (5, 4, 4, 128, 64),
        (8, 64, 64, 64, 64),    # several K splits
    ]
    for (n, h, w, cin, cout) in shapes:
        x = rnd(n, h, w, cin, seed=1).to(torch.bfloat16)
        dy = rnd(n, h, w, cout, seed=2).to(torch.bfloat16)
        dw_ref = torch.zeros(cout, 3, 3, cin)
        rf.conv3x3_wgrad(x, dy, dw_ref)
        dw = torch.full((cout, 3, 3, cin), float("nan"), device=DEV)
        cu.conv3x3_wgrad(x.to(DEV), dy.to(DEV), dw)
        sync(cu, "conv3x3_wgrad")
        splits = cu.lib.plume_wgrad_splits(n, h, w, 9, cin, cout)
        ok = report(f"conv3x3_wgrad N{n} {h}x{w} {cin}x{cout} splits={splits}", dw, dw_ref, tol=1e-2, detail=False)
        if not ok:
            e = ((dw.cpu() - dw_ref).abs() / (dw_ref.abs().max() + 1e-9))
            print("   err by tap:", [round(v, 3) for v in e.mean(dim=(0, 3)).flatten().tolist()])
            print("   err by ci chunk8:", [round(v, 3) for v in e.mean(dim=(0, 1, 2)).view(-1, 8).mean(1)[:16].tolist()])
            print("   err by co chunk8:", [round(v, 3) for v in e.mean(dim=(1, 2, 3)).view(-1, 8).mean(1)[:16].tolist()])
            print("   got[0,1,1,:8]", dw[0, 1, 1, :8].tolist())
            print("   ref[0,1,1,:8]", dw_ref[0, 1, 1, :8].tolist())
        cu.conv3x3_wgrad(x.to(DEV), dy.to(DEV), dw, accumulate=True)
        sync(cu, "conv3x3_wgrad acc")
        report("   accumulate", dw, 2 * dw_ref, tol=1e-2, detail=False)


def group_convT(cu, rf):
    shapes = [(1, 8, 8, 128, 64), (2, 8, 8, 256, 128), (2, 4, 4, 512, 256), (3, 12, 20, 128, 64), (1, 16, 16, 1024, 512)]
    for (n, h, w, cin, cout) in shapes:
        x = rnd(n, h, w, cin, seed=1).to(torch.bfloat16)
        wm = rnd(4, cout, cin, seed=2) * (1.0 / cin ** 0.5)
        bias = 0.1 * rnd(cout, seed=3)
        wf_r = torch.empty(4, cout, cin, dtype=torch.bfloat16)
        wd_r = torch.empty(cin, 4, cout, dtype=torch.bfloat16)
        rf.pack_convT(wm, wf_r, wd_r)
        wf = torch.empty(4, cout, cin, dtype=torch.bfloat16, device=DEV)
        wd = torch.empty(cin, 4, cout, dtype=torch.bfloat16, device=DEV)
        cu.pack_convT(wm.to(DEV), wf, wd)
        sync(cu, "pack_convT")
        report("   packT fwd", wf, wf_r, tol=0, detail=False)
        report("   packT dgrad", wd, wd_r, tol=0, detail=False)
        # forward into the upper half of a concat buffer
        cat_r = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16)
        rf.convT_fwd(x, wf_r, bias, cat_r[..., cout:])
        cat = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device=DEV)
        cu.convT_fwd(x.to(DEV), wf, bias.to(DEV), cat[..., cout:])
        sync(cu, "convT_fwd")
        report(f"convT_fwd N{n} {h}x{w} {cin}->{cout}", cat, cat_r)
        du = rnd(n, 2 * h, 2 * w, 2 * cout, seed=4).to(torch.bfloat16)
        dx_ref = torch.empty(n, h, w, cin, dtype=torch.bfloat16)
        rf.convT_dgrad(du[..., cout:], wd_r, dx_ref)
        dx = torch.full((n, h, w, cin), float("nan"), dtype=torch.bfloat16, device=DEV)
        dud = du.to(DEV)
        cu.convT_dgrad(dud[..., cout:], wd, dx)
        sync(cu, "convT_dgrad")
        report(f"convT_dgrad N{n} {h}x{w}", dx, dx_ref)
        dw_ref = torch.zeros(4, cout, cin)
        rf.convT_wgrad(x, du[..., cout:], dw_ref)
        dw = torch.full((4, cout, cin), float("nan"), device=DEV)
        cu.convT_wgrad(x.to(DEV), dud[..., cout:], dw)
        sync(cu, "convT_wgrad")
        report(f"convT_wgrad N{n} {h}x{w}", dw, dw_ref, tol=1e-2, detail=False)


def group_bw(cu, rf):
    for (n, h, w, c, ld) in [(2, 16, 16, 64, 64), (3, 8, 12, 128, 256), (1, 4, 4, 1024, 1024), (2, 6, 10, 192, 192)]:
        ybuf = rnd(n, h, w, ld, seed=1).to(torch.bfloat16)
        y = ybuf[..., ld - c:]
        scale = (1 + 0.2 * rnd(c, seed=2)).float()
        shift = (0.3 * rnd(c, seed=3)).float()
        a_ref = torch.empty(n, h, w, c, dtype=torch.bfloat16)
        rf.scale_shift_act(y, scale, shift, 1, a_ref)
        yd = ybuf.to(DEV)[..., ld - c:]
        a = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
        cu.scale_shift_act(yd, scale.to(DEV), shift.to(DEV), 1, a)
        sync(cu, "scale_shift_act")
        report(f"scale_shift_act N{n} {h}x{w} C{c} ld{ld}", a, a_ref, tol=4e-3)
        # fused pool
        skip_r = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16)
        pooled_r = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16)
        am_r = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8)
        rf.scale_shift_act_pool(y, scale, shift, 1, skip_r[..., :c], pooled_r, am_r)
        skip = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device=DEV)
        pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=DEV)
        am = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=DEV)
        cu.scale_shift_act_pool(yd, scale.to(DEV), shift.to(DEV), 1, skip[..., :c], pooled, am)
        sync(cu, "scale_shift_act_pool")
        report("   pool skip", skip, skip_r, tol=4e-3)
        report("   pool pooled", pooled, pooled_r, tol=4e-3)
        # argmax may differ only where rounding makes a tie-break differ; compare on exact skip equality
        same = (skip.cpu() == skip_r).all().item()
        mism = (am.cpu() != am_r).float().mean().item()
        print(f"   argmax mismatch frac={mism:.5f} (skip bit-identical={same})", flush=True)
        if same and mism > 0:
            FAILS.append("argmax")
        # plain pool + backward
        pooled2 = torch.empty_like(pooled)
        am2 = torch.empty_like(am)
        cu.maxpool_fwd(skip[..., :c], pooled2, am2)
        sync(cu, "maxpool_fwd")
        p2_r = torch.empty_like(pooled_r)
        am2_r = torch.empty_like(am_r)
        rf.maxpool_fwd(skip.cpu()[..., :c], p2_r, am2_r)
        report("   maxpool_fwd", pooled2, p2_r, tol=0)
        print("   maxpool argmax equal:", (am2.cpu() == am2_r).all().item(), flush=True)
        if not (am2.cpu() == am2_r).all().item():
            FAILS.append("maxpool argmax")
        dyp = rnd(n, h // 2, w // 2, c, seed=5).to(torch.bfloat16)
        dcat = rnd(n, h, w, 2 * c, seed=6).to(torch.bfloat16)
        dx_r = torch.empty(n, h, w, c, dtype=torch.bfloat16)
        rf.maxpool_bwd(dyp, am2_r, dcat[..., :c], dx_r)
        dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
        cu.maxpool_bwd(dyp.to(DEV), am2, dcat.to(DEV)[..., :c], dx)
        sync(cu, "maxpool_bwd")
        report("   maxpool_bwd(+skip)", dx, dx_r, tol=4e-3)
        rf.maxpool_bwd(dyp, am2_r, None, dx_r)
        cu.maxpool_bwd(dyp.to(DEV), am2, None, dx)
        sync(cu, "maxpool_bwd noskip")
        report("   maxpool_bwd", dx, dx_r, tol=0)
        # BN finalize + backward
        cnt = n * h * w
        ss = y.double().sum(dim=(0, 1, 2))
        sq = (y.double() ** 2).sum(dim=(0, 1, 2))
        gamma = (1 + 0.1 * rnd(c, seed=7)).float()
        beta = (0.1 * rnd(c, seed=8)).float()
        outs_r = [torch.zeros(c) for _ in range(4)]
        rm_r, rv_r = torch.zeros(c), torch.ones(c)
        rf.bn_finalize(ss, sq, cnt, gamma, beta, 1e-5, 0.1, rm_r, rv_r, *outs_r)
        outs = [torch.zeros(c, device=DEV) for _ in range(4)]
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        cu.bn_finalize(ss.to(DEV), sq.to(DEV), cnt, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, rm, rv, *outs)
        sync(cu, "bn_finalize")
        for nm, o, o_r in zip(("scale", "shift", "mean", "invstd"), outs, outs_r):
            report(f"   bn_finalize {nm}", o, o_r, tol=1e-4, detail=False)
        report("   bn_finalize running_mean", rm, rm_r, tol=1e-4, detail=False)
        report("   bn_finalize running_var", rv, rv_r, tol=1e-4, detail=False)
        fs_r, fh_r = torch.zeros(c), torch.zeros(c)
        rf.bn_fold_eval(gamma, beta, rm_r, rv_r, shift, 1e-5, fs_r, fh_r)
        fs, fh = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        cu.bn_fold_eval(gamma.to(DEV), beta.to(DEV), rm, rv, shift.to(DEV), 1e-5, fs, fh)
        sync(cu, "bn_fold_eval")
        report("   bn_fold scale", fs, fs_r, tol=1e-4, detail=False)
        report("   bn_fold shift", fh, fh_r, tol=1e-4, detail=False)
        da = rnd(n, h, w, c, seed=9).to(torch.bfloat16)
        sc_r, sh_r, mu_r, is_r = outs_r
        sg_r, sgx_r = torch.zeros(c), torch.zeros(c)
        rf.bn_bwd_reduce(da, y, sc_r, sh_r, mu_r, is_r, 1, sg_r, sgx_r)
        sg, sgx = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
        dad = da.to(DEV)
        cu.bn_bwd_reduce(dad, yd, *[t.to(DEV) for t in outs_r], 1, sg, sgx)
        sync(cu, "bn_bwd_reduce")
        report("   bn_bwd_reduce sum_g", sg, sg_r, tol=2e-3, detail=False)
        report("   bn_bwd_reduce sum_gx", sgx, sgx_r, tol=2e-3, detail=False)
        dy_r = torch.empty(n, h, w, c, dtype=torch.bfloat16)
        sdy_r = torch.zeros(c)
        rf.bn_bwd_apply(da, y, sc_r, sh_r, mu_r, is_r, 1, sg_r, sgx_r, dy_r, sdy_r)
        dyo = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
        sdy = torch.zeros(c, device=DEV)
        cu.bn_bwd_apply(dad, yd, *[t.to(DEV) for t in outs_r], 1, sg_r.to(DEV), sgx_r.to(DEV), dyo, sdy)
        sync(cu, "bn_bwd_apply")
        report("   bn_bwd_apply dy", dyo, dy_r, tol=4e-3)
        report("   bn_bwd_apply sum_dy", sdy, sdy_r, tol=2e-2, detail=False)
        rdy_r = torch.empty(n, h, w, c, dtype=torch.bfloat16)
        rs_r = torch.zeros(c)
        rf.relu_bwd(da, a_ref, rdy_r, rs_r)
        rdy = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV)
        rs = torch.zeros(c, device=DEV)
        cu.relu_bwd(dad, a_ref.to(DEV), rdy, rs)
        sync(cu, "relu_bwd")
        report("   relu_bwd", rdy, rdy_r, tol=0)
        report("   relu_bwd sum", rs, rs_r, tol=1e-3, detail=False)
        cs_r = torch.zeros(c)
        rf.channel_sum(y, cs_r)
        cs = torch.zeros(c, device=DEV)
        cu.channel_sum(yd, cs)
        sync(cu, "channel_sum")
        report("   channel_sum", cs, cs_r, tol=1e-3, detail=False)
    xin = rnd(2, 8, 8, 8, seed=1).to(torch.bfloat16)
    out = torch.full((2, 8, 8, 64), float("nan"), dtype=torch.bfloat16, device=DEV)
    cu.pad_channels(xin.to(DEV), out)
    sync(cu, "pad_channels")
    out_r = torch.empty(2, 8, 8, 64, dtype=torch.bfloat16)
    rf.pad_channels(xin, out_r)
    report("pad_channels", out, out_r, tol=0)


def group_head(cu, rf):
    for (n, h, w, c) in [(2, 16, 16, 64), (1, 8, 24, 128), (3, 4, 4, 64)]:
        feat = rnd(n, h, w, c, seed=1).to(torch.bfloat16)
        wv = (rnd(c, seed=2) / c ** 0.5).float()
        b = torch.tensor([0.05])
        tgt = (rnd(n, h, w, seed=3) > 0.8).to(torch.uint8)
        lg_r, sums_r, loss_r = torch.zeros(n, h, w), torch.zeros(4), torch.zeros(3)
        rf.head_fwd(feat, wv, b, tgt, lg_r, sums_r)
        rf.head_loss(sums_r, n * h * w, 1.0, 1.0, 1.0, loss_r)
        lg, sums, loss = torch.zeros(n, h, w, device=DEV), torch.zeros(4, device=DEV), torch.zeros(3, device=DEV)
        cu.head_fwd(feat.to(DEV), wv.to(DEV), b.to(DEV), tgt.to(DEV), lg, sums)
        cu.head_loss(sums, n * h * w, 1.0, 1.0, 1.0, loss)
        sync(cu, "head_fwd")
        report(f"head_fwd logits N{n} {h}x{w} C{c}", lg, lg_r, tol=1e-4, detail=False)
        report("   sums", sums, sums_r, tol=1e-4, detail=False)
        report("   loss", loss, loss_r, tol=1e-4, detail=False)
        df_r, dw_r, db_r = torch.empty(n, h, w, c, dtype=torch.bfloat16), torch.zeros(c), torch.zeros(1)
        rf.head_bwd(feat, wv, lg_r, tgt, sums_r, 1.0, 1.0, 1.0, 0.5, df_r, dw_r, db_r)
        df, dw, db = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=DEV), torch.zeros(c, device=DEV), torch.zeros(1, device=DEV)
        cu.head_bwd(feat.to(DEV), wv.to(DEV), lg_r.to(DEV), tgt.to(DEV), sums_r.to(DEV), 1.0, 1.0, 1.0, 0.5, df, dw, db)
        sync(cu, "head_bwd")
        report("   head_bwd dfeat", df, df_r, tol=8e-3)
        report("   head_bwd dw", dw, dw_r, tol=1e-3, detail=False)
        report("   head_bwd db", db, db_r, tol=1e-3, detail=False)


def group_adam(cu, rf):
    for n in (1 << 20, 12345, 3):
        p, g = rnd(n, seed=1), rnd(n, seed=2)
        m, v = torch.zeros(n), torch.zeros(n)
        pd, gd, md, vd = p.to(DEV), g.to(DEV), m.to(DEV), v.to(DEV)
        for step in (1, 2, 3):
            rf.adam(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, step, 0.5)
            cu.adam(pd, gd, md, vd, 1e-3, 0.9, 0.999, 1e-8, step, 0.5)
        sync(cu, "adam")
        report(f"adam n={n} param", pd, p, tol=1e-6, detail=False)
        report("   m", md, m, tol=1e-6, detail=False)
        report("   v", vd, v, tol=1e-5, detail=False)


def group_tiles(cu, rf):
    hs, ws, cs, T, margin = 100, 130, 8, 64, 8
    scene = rnd(hs, ws, cs, seed=1).to(torch.bfloat16)
    stride = T - 2 * margin
    ys_l, xs_l = [], []
    for y0 in range(0, max(hs - 2 * margin, 1), stride):
        for x0 in range(0, max(ws - 2 * margin, 1), stride):
            ys_l.append(y0)
            xs_l.append(x0)
    ys, xs = torch.tensor(ys_l, dtype=torch.int32), torch.tensor(xs_l, dtype=torch.int32)
    k = ys.numel()
    tiles_r = torch.empty(k, T, T, 64, dtype=torch.bfloat16)
    rf.extract_tiles(scene, ys, xs, T, tiles_r)
    tiles = torch.full((k, T, T, 64), float("nan"), dtype=torch.bfloat16, device=DEV)
    cu.extract_tiles(scene.to(DEV), ys.to(DEV), xs.to(DEV), T, tiles)
    sync(cu, "extract_tiles")
    report("extract_tiles", tiles, tiles_r, tol=0)
    logits = rnd(k, T, T, seed=2)
    mask_r = torch.full((hs, ws), 7, dtype=torch.uint8)
    prob_r = torch.zeros(hs, ws)
    rf.stitch_threshold(logits, ys, xs, T, margin, 0.0, mask_r, prob_r)
    mask = torch.full((hs, ws), 7, dtype=torch.uint8, device=DEV)
    prob = torch.zeros(hs, ws, device=DEV)
    cu.stitch_threshold(logits.to(DEV), ys.to(DEV), xs.to(DEV), T, margin, 0.0, mask, prob)
    sync(cu, "stitch")
    print("   stitch coverage complete (oracle):", (mask_r != 7).all().item(), flush=True)
    report("stitch mask", mask, mask_r, tol=0, detail=False)
    report("stitch prob", prob, prob_r, tol=1e-5, detail=False)


GROUPS = {"conv": group_conv, "wgrad": group_wgrad, "convT": group_convT, "bw": group_bw,
          "head": group_head, "adam": group_adam, "tiles": group_tiles}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    cu, rf = CudaOps(), RefOps()
    print("device:", torch.cuda.get_device_name(0), "| lib:", cu.lib.plume_version().decode(), flush=True)
    for nm in names:
        t0 = time.time()
        print(f"===== group {nm}", flush=True)
        try:
            GROUPS[nm](cu, rf)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            FAILS.append(f"{nm}: exception {e}")
            break
        print(f"===== group {nm} done in {time.time() - t0:.1f}s", flush=True)
    print("FAILED:" if FAILS else "ALL OK", FAILS, flush=True)
    sys.exit(1 if FAILS else 0)
