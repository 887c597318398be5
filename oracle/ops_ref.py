"""TEST INFRASTRUCTURE -- per-operator CPU oracle for the C ABI in ``include/plume_b200.h``.

PARITY UNPINNED: the reference repository (gridl/kcl-ltss-bioatm) contains no model code at all
(``/root/reference/src/models/__init__.py`` is 0 bytes), so there are no golden
vectors or reference outputs to pin these restatements against.  Each method states the operator's
definition with ``torch.nn.functional`` in fp32 on CPU; whole-network behaviour is pinned against
``oracle/unet_ref.py`` (the frozen spec) by ``tests/test_host_logic.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  It mirrors ``kcl_ltss_bioatm_b200.ops.CudaOps`` method for method (same names,
argument meaning, in-place outputs, bf16 rounding points) so the same host code can be checked on CPU.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _nchw(t: torch.Tensor) -> torch.Tensor:
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 2, 3, 1)


class RefOps:
    """fp32 CPU statement of every operator; outputs are written in place like the CUDA kernels."""

    name = "ref"

    def __init__(self, act_dtype: torch.dtype = torch.bfloat16) -> None:
        # act_dtype=torch.float32 removes every bf16 rounding point: the host schedule can then be
        # checked against autograd to fp32 accuracy (tests/test_host_logic.py)
        self.act_dtype = act_dtype
        self.launches = 0

    # ---- 3x3 convolution (weights KRSC = [Cout][3][3][Cin]) ------------------------------------
    def conv3x3_fwd(self, x, w_fwd, scale, shift, relu, y, stat_sum=None, stat_sq=None):
        cout, cin = y.shape[-1], x.shape[-1]
        kcin = w_fwd.numel() // (9 * cout)  # the weights may carry zero-padded input channels beyond x's
        w = w_fwd.float().view(cout, 3, 3, kcin)[..., :cin].permute(0, 3, 1, 2)
        o = F.conv2d(_nchw(x), w, padding=1)
        if scale is not None:
            o = o * scale.view(1, -1, 1, 1)
        if shift is not None:
            o = o + shift.view(1, -1, 1, 1)
        if relu:
            o = o.relu()
        y.copy_(_nhwc(o).to(self.act_dtype))
        if stat_sum is not None:
            yr = y.to(stat_sum.dtype)      # the accumulators are fp64 on the device
            stat_sum += yr.sum(dim=(0, 1, 2))
            stat_sq += (yr * yr).sum(dim=(0, 1, 2))

    def conv3x3_dgrad(self, dy, w_dgrad, dx):
        # w_dgrad[ci][r][s][co] = w[co][2-r][2-s][ci]: a plain 3x3 correlation of dy with it is dgrad
        cin, cout = dx.shape[-1], dy.shape[-1]
        w = w_dgrad.float().view(cin, 3, 3, cout).permute(0, 3, 1, 2)
        dx.copy_(_nhwc(F.conv2d(_nchw(dy), w, padding=1)).to(self.act_dtype))

    def conv3x3_wgrad(self, x, dy, dw, accumulate=False):
        cout, cin = dy.shape[-1], x.shape[-1]
        g = torch.nn.grad.conv2d_weight(_nchw(x), (cout, cin, 3, 3), _nchw(dy), padding=1)
        g = g.permute(0, 2, 3, 1)
        kcin = dw.numel() // (9 * cout)
        if kcin != cin:  # dW of the weights' zero-padded input channels is zero
            g = F.pad(g, (0, kcin - cin))
        g = g.reshape(dw.shape)
        if accumulate:
            dw += g
        else:
            dw.copy_(g)

    # ---- transposed conv 2x2 / stride 2 (weights [ij][Cout][Cin]) -------------------------------
    def convT_fwd(self, x, w_fwd, bias, u):
        cout, cin = u.shape[-1], x.shape[-1]
        w = w_fwd.float().view(2, 2, cout, cin).permute(3, 2, 0, 1)  # [Cin][Cout][i][j]
        o = F.conv_transpose2d(_nchw(x), w, bias=bias, stride=2)
        u.copy_(_nhwc(o).to(self.act_dtype))

    def convT_dgrad(self, du, w_dgrad, dx):
        cin, cout = dx.shape[-1], du.shape[-1]
        w = w_dgrad.float().view(cin, 2, 2, cout).permute(0, 3, 1, 2)  # conv weight [Cin][Cout][i][j]
        dx.copy_(_nhwc(F.conv2d(_nchw(du), w, stride=2)).to(self.act_dtype))

    def convT_wgrad(self, x, du, dw, accumulate=False):
        cout, cin = du.shape[-1], x.shape[-1]
        xf, duf = x.float(), du.float()
        n, h, w, _ = x.shape
        d = duf.reshape(n, h, 2, w, 2, cout)
        g = torch.einsum("nhiwjo,nhwc->ijoc", d, xf).reshape(dw.shape)
        if accumulate:
            dw += g
        else:
            dw.copy_(g)

    # ---- packing ---------------------------------------------------------------------------------
    def pack_conv3x3(self, w, wf, wd):
        wb = w.to(self.act_dtype)
        if wf is not None:
            wf.view_as(w).copy_(wb)
        if wd is not None:
            cout, _, _, cin = w.shape
            wd.view(cin, 3, 3, cout).copy_(wb.flip(1, 2).permute(3, 1, 2, 0))

    def pack_convT(self, w, wf, wd):
        wb = w.to(self.act_dtype)
        if wf is not None:
            wf.view_as(w).copy_(wb)
        if wd is not None:
            _, cout, cin = w.shape
            wd.view(cin, 4, cout).copy_(wb.permute(2, 0, 1))

    def pack_batch(self, jobs):
        for kind, w, wf, wd in jobs:
            (self.pack_conv3x3 if kind == "conv3x3" else self.pack_convT)(w, wf, wd)

    # ---- bandwidth operators ----------------------------------------------------------------------
    def pad_channels(self, x, out):
        out.zero_()
        out[..., : x.shape[-1]] = x

    def bn_finalize(self, ssum, ssq, count, gamma, beta, eps, momentum, running_mean, running_var,
                    scale, shift, mean, invstd):
        m = ssum.double() / count
        var = (ssq.double() / count - m * m).clamp_min(0).float()
        m = m.float()
        istd = torch.rsqrt(var + eps)
        g = gamma if gamma is not None else torch.ones_like(m)
        b = beta if beta is not None else torch.zeros_like(m)
        scale.copy_(g * istd)
        shift.copy_(b - m * g * istd)
        if mean is not None:
            mean.copy_(m)
        if invstd is not None:
            invstd.copy_(istd)
        unbias = count / (count - 1) if count > 1 else 1.0
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * m)
        if running_var is not None:
            running_var.mul_(1 - momentum).add_(momentum * unbias * var)

    def bn_fold_eval(self, gamma, beta, running_mean, running_var, conv_bias, eps, scale, shift):
        g = gamma if gamma is not None else torch.ones_like(running_mean)
        b = beta if beta is not None else torch.zeros_like(running_mean)
        cb = conv_bias if conv_bias is not None else torch.zeros_like(running_mean)
        sc = g * torch.rsqrt(running_var + eps)
        scale.copy_(sc)
        shift.copy_((cb - running_mean) * sc + b)

    def scale_shift_act(self, y, scale, shift, relu, a):
        o = y.float() * scale + shift
        if relu:
            o = o.relu()
        a.copy_(o.to(self.act_dtype))

    @staticmethod
    def _pool(v: torch.Tensor):
        # v: [N,H,W,C] fp32 -> pooled, argmax (first maximum in (0,0),(0,1),(1,0),(1,1) order)
        n, h, w, c = v.shape
        win = v.reshape(n, h // 2, 2, w // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(n, h // 2, w // 2, c, 4)
        best = win[..., 0].clone()
        idx = torch.zeros_like(best, dtype=torch.uint8)
        for k in range(1, 4):
            better = win[..., k] > best
            best = torch.where(better, win[..., k], best)
            idx = torch.where(better, torch.full_like(idx, k), idx)
        return best, idx

    def scale_shift_act_pool(self, y, scale, shift, relu, skip, pooled, argmax):
        o = y.float() * scale + shift
        if relu:
            o = o.relu()
        ob = o.to(self.act_dtype)
        if skip is not None:
            skip.copy_(ob)
        best, idx = self._pool(ob.float())
        pooled.copy_(best.to(self.act_dtype))
        argmax.view_as(idx).copy_(idx)

    def maxpool_fwd(self, x, y, argmax):
        best, idx = self._pool(x.float())
        y.copy_(best.to(self.act_dtype))
        argmax.view_as(idx).copy_(idx)

    def maxpool_bwd(self, dy, argmax, dskip, dx, bn=None):
        n, h, w, c = dx.shape
        g = dy.float()
        idx = argmax.view(n, h // 2, w // 2, c).long()
        out = torch.zeros(n, h // 2, w // 2, c, 4)
        out.scatter_(-1, idx.unsqueeze(-1), g.unsqueeze(-1))
        out = out.reshape(n, h // 2, w // 2, c, 2, 2).permute(0, 1, 4, 2, 5, 3).reshape(n, h, w, c)
        if dskip is not None:
            out = out + dskip.float()
        dx.copy_(out.to(self.act_dtype))
        if bn is not None:   # fused BatchNorm-backward reduction of the stored gradient
            y, scale, shift, mean, invstd, relu, sum_g, sum_gx = bn
            self.bn_bwd_reduce(dx, y, scale, shift, mean, invstd, relu, sum_g, sum_gx)

    def bn_bwd_reduce(self, da, y, scale, shift, mean, invstd, relu, sum_g, sum_gx):
        yf = y.float()
        g = da.float()
        if relu:
            g = torch.where(yf * scale + shift > 0, g, torch.zeros_like(g))
        xhat = (yf - mean) * invstd
        sum_g += g.sum(dim=(0, 1, 2))
        sum_gx += (g * xhat).sum(dim=(0, 1, 2))

    def bn_bwd_apply(self, da, y, scale, shift, mean, invstd, relu, sum_g, sum_gx, dy, sum_dy,
                     dgamma=None, dbeta=None, accumulate=False):
        if dgamma is not None:
            if accumulate:
                dgamma += sum_gx
                dbeta += sum_g
            else:
                dgamma.copy_(sum_gx)
                dbeta.copy_(sum_g)
        yf = y.float()
        g = da.float()
        if relu:
            g = torch.where(yf * scale + shift > 0, g, torch.zeros_like(g))
        xhat = (yf - mean) * invstd
        cnt = yf.numel() // yf.shape[-1]
        o = scale * (g - sum_g / cnt - xhat * (sum_gx / cnt))
        ob = o.to(self.act_dtype)
        dy.copy_(ob)
        if sum_dy is not None:
            sum_dy += ob.float().sum(dim=(0, 1, 2))

    def relu_bwd(self, da, a, dy, sum_dy):
        g = torch.where(a.float() > 0, da.float(), torch.zeros_like(da, dtype=torch.float32))
        gb = g.to(self.act_dtype)
        dy.copy_(gb)
        if sum_dy is not None:
            sum_dy += gb.float().sum(dim=(0, 1, 2))

    def channel_sum(self, x, out):
        out += x.float().sum(dim=(0, 1, 2))

    def head_fwd(self, feat, w, b, target, logits, sums):
        z = feat.float() @ w + (b if b is not None else 0.0)
        logits.view_as(z).copy_(z)
        if target is not None:
            t = (target.view_as(z) != 0).float()
            p = torch.sigmoid(z)
            sums[0] += F.binary_cross_entropy_with_logits(z, t, reduction="sum")
            sums[1] += (p * t).sum()
            sums[2] += p.sum()
            sums[3] += t.sum()

    def head_loss(self, sums, pixels, bce_w, dice_w, eps, loss_out):
        bce = sums[0] / pixels
        dice = 1 - (2 * sums[1] + eps) / (sums[2] + sums[3] + eps)
        loss_out[0] = bce_w * bce + dice_w * dice
        loss_out[1] = bce
        loss_out[2] = dice

    def head_bwd(self, feat, w, logits, target, sums, bce_w, dice_w, eps, grad_scale, dfeat, dw, db, bn=None):
        n, h, wd, c = feat.shape
        z = logits.view(n, h, wd)
        t = (target.view(n, h, wd) != 0).float()
        p = torch.sigmoid(z)
        pixels = z.numel()
        S = sums[2] + sums[3] + eps
        I2 = 2 * sums[1] + eps
        ddice = -(2 * t * S - I2) / (S * S)
        dz = grad_scale * (bce_w * (p - t) / pixels + dice_w * ddice * p * (1 - p))
        dfeat.copy_((dz.unsqueeze(-1) * w).to(self.act_dtype))
        dw += (dz.unsqueeze(-1) * feat.float()).sum(dim=(0, 1, 2))
        db += dz.sum()
        if bn is not None:
            y, scale, shift, mean, invstd, relu, sum_g, sum_gx = bn
            self.bn_bwd_reduce(dfeat, y, scale, shift, mean, invstd, relu, sum_g, sum_gx)

    def cast_f32_bf16(self, src, dst):
        dst.copy_(src.to(torch.bfloat16))

    def cast_bf16_f32(self, src, dst):
        dst.copy_(src.float())

    def adam(self, param, grad, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
        g = grad * grad_scale
        m.mul_(beta1).add_((1 - beta1) * g)
        v.mul_(beta2).add_((1 - beta2) * g * g)
        bc1 = 1 - beta1 ** step
        bc2 = 1 - beta2 ** step
        param.sub_((lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + eps))

    def adam_dev(self, param, grad, m, v, coef):
        lr_t, b1, b2, omb1, omb2, eps, inv_bc2_sqrt, gs = [float(c) for c in coef.tolist()]
        g = grad * gs
        m.mul_(b1).add_(omb1 * g)
        v.mul_(b2).add_(omb2 * g * g)
        param.sub_(lr_t * m / (v.sqrt() * inv_bc2_sqrt + eps))

    # ---- tiled inference -----------------------------------------------------------------------------
    def extract_tiles(self, scene, ys, xs, tile, tiles):
        hs, ws, cs = scene.shape
        tiles[: ys.numel()].zero_()
        for k, (y0, x0) in enumerate(zip(ys.tolist(), xs.tolist())):
            y1, x1 = min(y0 + tile, hs), min(x0 + tile, ws)
            yy0, xx0 = max(y0, 0), max(x0, 0)
            if y1 > yy0 and x1 > xx0:
                tiles[k, yy0 - y0:y1 - y0, xx0 - x0:x1 - x0, :cs] = scene[yy0:y1, xx0:x1]

    def stitch_threshold(self, logits, ys, xs, tile, margin, logit_threshold, mask, prob=None):
        hs, ws = mask.shape
        lg = logits.view(-1, tile, tile)
        for k, (y0, x0) in enumerate(zip(ys.tolist(), xs.tolist())):
            ya = 0 if y0 <= 0 else margin
            yb = tile if y0 + tile >= hs else tile - margin
            xa = 0 if x0 <= 0 else margin
            xb = tile if x0 + tile >= ws else tile - margin
            sy0, sy1 = max(y0 + ya, 0), min(y0 + yb, hs)
            sx0, sx1 = max(x0 + xa, 0), min(x0 + xb, ws)
            if sy1 <= sy0 or sx1 <= sx0:
                continue
            z = lg[k, sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0]
            mask[sy0:sy1, sx0:sx1] = (z >= logit_threshold).to(torch.uint8)
            if prob is not None:
                prob[sy0:sy1, sx0:sx1] = torch.sigmoid(z)

    # ---- label geometry ----------------------------------------------------------------------------
    def rasterize_hulls(self, verts, offsets, bbox, masks, ys=None, xs=None):
        from oracle.hull_ref import rasterize_ref

        v, o = verts.cpu().numpy(), offsets.cpu().numpy()
        hulls = [(v[o[i]:o[i + 1], 0], v[o[i]:o[i + 1], 1]) for i in range(len(o) - 1)]
        if ys is None:
            h, w = masks.shape
            masks.copy_(torch.from_numpy(rasterize_ref(hulls, h, w)))
        else:
            _, h, w = masks.shape
            for k in range(masks.shape[0]):
                masks[k].copy_(torch.from_numpy(rasterize_ref(hulls, h, w, origin=(int(ys[k]), int(xs[k])))))

    def locate_fires(self, lats, lons, fire_lat, fire_lon, half_box_deg, out_rc):
        from oracle import fire_ref

        assert half_box_deg == fire_ref.HALF_BOX_DEG
        rc = fire_ref.nearest_pixel_ref(fire_lat.cpu().numpy(), fire_lon.cpu().numpy(), lats.cpu().numpy(),
                                        lons.cpu().numpy())
        out_rc.copy_(torch.from_numpy(rc).to(torch.int32))

    # ---- threshold sweep ---------------------------------------------------------------------------
    def threshold_masks(self, aod, thresholds, masks):
        from oracle.sweep_ref import threshold_masks_ref

        masks.copy_(torch.from_numpy(threshold_masks_ref(aod.cpu().numpy(), thresholds.cpu().numpy()).astype("uint8")))

    def label_components(self, masks, labels, sizes):
        import numpy as np

        from oracle.sweep_ref import label_ref

        for t in range(masks.shape[0]):
            lab = label_ref(masks[t].cpu().numpy().astype(bool))          # 0 background, else 1 + smallest index
            labels[t].copy_(torch.from_numpy((lab - 1).astype("int32")))
            cnt = np.bincount(lab.ravel(), minlength=lab.size + 1)[1:]
            cnt[lab.ravel() == 0] = cnt[lab.ravel() == 0] * 0
            sizes[t].copy_(torch.from_numpy(cnt.reshape(lab.shape).astype("int32")))

    def fire_extents(self, labels, sizes, fire_rc, win, extents):
        from oracle.sweep_ref import extract_label_ref

        for t in range(labels.shape[0]):
            lab = labels[t].cpu().numpy().astype("int64") + 1
            sz = sizes[t].cpu().numpy().ravel()
            for f, (r, c) in enumerate(fire_rc.cpu().numpy()):
                l = extract_label_ref(lab, int(r), int(c), win)
                extents[t, f] = 0 if l is None else int(sz[l - 1])


class RefOpsF64Accum(RefOps):
    """RefOps with the same bf16 rounding points but the 3x3 convolutions accumulated in float64, i.e. a different
    summation order.  Two implementations that differ only in accumulation order flip isolated bf16 roundings, and
    BatchNorm's backward amplifies those flips; the distance between RefOps and this class is therefore the NOISE
    FLOOR of any "same rounding points" comparison (tests/test_gpu_unet.py::_grad_parity measures it per tensor)."""

    name = "ref-f64-accum"

    def conv3x3_fwd(self, x, w_fwd, scale, shift, relu, y, stat_sum=None, stat_sq=None):
        cout, cin = y.shape[-1], x.shape[-1]
        kcin = w_fwd.numel() // (9 * cout)
        w = w_fwd.double().view(cout, 3, 3, kcin)[..., :cin].permute(0, 3, 1, 2)
        o = F.conv2d(_nchw(x).double(), w, padding=1).float()
        if scale is not None:
            o = o * scale.view(1, -1, 1, 1)
        if shift is not None:
            o = o + shift.view(1, -1, 1, 1)
        if relu:
            o = o.relu()
        y.copy_(_nhwc(o).to(self.act_dtype))
        if stat_sum is not None:
            yr = y.to(stat_sum.dtype)
            stat_sum += yr.sum(dim=(0, 1, 2))
            stat_sq += (yr * yr).sum(dim=(0, 1, 2))

    def conv3x3_dgrad(self, dy, w_dgrad, dx):
        cin, cout = dx.shape[-1], dy.shape[-1]
        w = w_dgrad.double().view(cin, 3, 3, cout).permute(0, 3, 1, 2)
        dx.copy_(_nhwc(F.conv2d(_nchw(dy).double(), w, padding=1).float()).to(self.act_dtype))
