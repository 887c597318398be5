"""Operator layer over the C ABI: one method per entry point of ``include/plume_b200.h``.

Tensors are torch CUDA tensors in NHWC; bf16 activations may be channel slices of a wider buffer
(``concat[..., :C]``) -- the pixel stride is read from ``stride(2)``.  Outputs are preallocated by the
caller (nothing is allocated per call; ``locate_fires`` reuses one growable scratch buffer), so a whole step is
a fixed sequence of launches on the current stream.

The same method names and argument meaning are implemented on CPU in ``oracle/ops_ref.py``; that class
is test infrastructure (the checker), never imported from here.
"""
from __future__ import annotations

import torch

from . import lib as _lib
from .lib import check, current_stream, ptr


def _act(t: torch.Tensor, name: str, planes: int = 1):
    """(pointer, pixel stride, N, H, W, C) of an NHWC bf16 activation view.  planes == 2 (bf16x3 mode): the tensor
    is [N, H, W, 2, C] -- a hi and a lo plane per pixel -- and the pixel stride covers both planes."""
    if t.dtype != torch.bfloat16 or t.dim() != (4 if planes == 1 else 5):
        raise TypeError(f"{name}: expected a {4 if planes == 1 else 5}-D bf16 NHWC tensor, got {t.dtype} "
                        f"{tuple(t.shape)}")
    if not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if planes == 1:
        n, h, w, c = t.shape
        sn, sh, sw, sc = t.stride()
        ld = sw if w > 1 else (sh if h > 1 else (sn if n > 1 else c))
        ok = (sc == 1 or c == 1) and ld >= c and ld % 8 == 0
    else:
        n, h, w, pl, c = t.shape
        sn, sh, sw, sp, sc = t.stride()
        ld = 2 * sp
        ok = pl == 2 and (sc == 1 or c == 1) and sp >= c and sp % 8 == 0
    ok = ok and (w == 1 or sw == ld) and (h == 1 or sh == w * ld) and (n == 1 or sn == h * w * ld)
    if not ok:
        raise ValueError(f"{name}: not a dense-pixel NHWC view: shape {tuple(t.shape)} strides {t.stride()}")
    return ptr(t), int(ld), int(n), int(h), int(w), int(c)


def _weight_cin(w_numel: int, cout: int, cin: int, ldx: int) -> int:
    """Input channels per tap of a 3x3 weight tensor.  It may exceed the activation's channel count when
    the activation is dense (ldx == cin): the C ABI then reads the missing channels as zero (ldx < Cin)."""
    kcin = w_numel // (9 * cout)
    if kcin * 9 * cout != w_numel or kcin < cin or (kcin > cin and ldx != cin):  # ldx: channels of one plane
        raise ValueError(f"3x3 weights with {w_numel} elements do not fit Cout={cout}, Cin={cin} (ld {ldx})")
    return kcin


def _f32(t, name: str, numel: int | None = None):
    if t is None:
        return ptr(None)
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise TypeError(f"{name}: expected a contiguous fp32 CUDA tensor")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{name}: needs {numel} elements, has {t.numel()}")
    return ptr(t)


def _f64(t, name: str, numel: int | None = None):
    if t is None:
        return ptr(None)
    if t.dtype != torch.float64 or not t.is_contiguous() or not t.is_cuda:
        raise TypeError(f"{name}: expected a contiguous fp64 CUDA tensor")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{name}: needs {numel} elements, has {t.numel()}")
    return ptr(t)


class CudaOps:
    """Launches the sm_100a kernels.  Fails loudly when the library or a GPU is missing."""

    name = "b200"

    def __init__(self, precision: str = "bf16") -> None:
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PlumeError("CUDA device required: the B200 path has no CPU fallback")
        if precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3'")
        self.precision = precision
        # bf16x3: activations are [N, H, W, 2, C] (hi / lo planes), the bf16 weight copies hold a hi and a lo
        # matrix, and every activation-typed entry point is its *_x3 variant (include/plume_b200.h)
        self.planes = 2 if precision == "bf16x3" else 1
        self._sfx = "_x3" if self.planes == 2 else ""
        self._ws: torch.Tensor | None = None
        self.launches = 0  # kernels of ours enqueued so far (bench.py reports the per-step count)

    def _fn(self, name: str):
        return getattr(self.lib, name + self._sfx)

    def set_deterministic(self, on: bool) -> None:
        """Library-wide switch (also PLUME_DETERMINISTIC=1): ordered partial-sum reductions instead of fp atomics."""
        self.lib.plume_set_deterministic(int(bool(on)))

    @property
    def is_deterministic(self) -> bool:
        return bool(self.lib.plume_get_deterministic())

    def _a(self, t: torch.Tensor, name: str):
        return _act(t, name, self.planes)

    # ------------------------------------------------------------------ helpers
    def _workspace(self, nbytes: int, device) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self._ws

    # ------------------------------------------------------------------ tensor-core GEMMs
    def conv3x3_fwd(self, x, w_fwd, scale, shift, relu, y, stat_sum=None, stat_sq=None):
        xp, ldx, n, h, w, cin = self._a(x, "x")
        yp, ldy, n2, h2, w2, cout = self._a(y, "y")
        assert (n, h, w) == (n2, h2, w2)
        kcin = _weight_cin(w_fwd.numel() // self.planes, cout, cin, ldx // self.planes)
        assert w_fwd.dtype == torch.bfloat16 and w_fwd.is_contiguous()
        check(self._fn("plume_conv3x3_fwd")(xp, ldx, ptr(w_fwd), _f32(scale, "scale", cout),
                                         _f32(shift, "shift", cout), int(bool(relu)), yp, ldy,
                                         _f64(stat_sum, "stat_sum", cout), _f64(stat_sq, "stat_sq", cout),
                                         n, h, w, kcin, cout, current_stream()), "plume_conv3x3_fwd")
        self.launches += 1

    def conv3x3_dgrad(self, dy, w_dgrad, dx):
        dyp, lddy, n, h, w, cout = self._a(dy, "dy")
        dxp, lddx, n2, h2, w2, cin = self._a(dx, "dx")
        assert (n, h, w) == (n2, h2, w2)
        assert w_dgrad.dtype == torch.bfloat16 and w_dgrad.numel() == self.planes * cout * 9 * cin
        check(self._fn("plume_conv3x3_dgrad")(dyp, lddy, ptr(w_dgrad), dxp, lddx, n, h, w, cin, cout,
                                           current_stream()), "plume_conv3x3_dgrad")
        self.launches += 1

    def conv3x3_wgrad(self, x, dy, dw, accumulate=False):
        xp, ldx, n, h, w, cin = self._a(x, "x")
        dyp, lddy, n2, h2, w2, cout = self._a(dy, "dy")
        assert (n, h, w) == (n2, h2, w2)
        kcin = _weight_cin(dw.numel(), cout, cin, ldx // self.planes)
        check(self._fn("plume_conv3x3_wgrad")(xp, ldx, dyp, lddy, _f32(dw, "dw", cout * 9 * kcin),
                                           int(bool(accumulate)), ptr(None), 0, n, h, w, kcin, cout,
                                           current_stream()), "plume_conv3x3_wgrad")
        self.launches += 1

    def convT_fwd(self, x, w_fwd, bias, u):
        xp, ldx, n, h, w, cin = self._a(x, "x")
        up, ldu, n2, h2, w2, cout = self._a(u, "u")
        assert (n2, h2, w2) == (n, 2 * h, 2 * w)
        assert w_fwd.dtype == torch.bfloat16 and w_fwd.numel() == self.planes * 4 * cout * cin
        check(self._fn("plume_convT2x2_concat_fwd")(xp, ldx, ptr(w_fwd), _f32(bias, "bias", cout), up, ldu,
                                                 n, h, w, cin, cout, current_stream()),
              "plume_convT2x2_concat_fwd")
        self.launches += 1

    def convT_dgrad(self, du, w_dgrad, dx):
        dup, lddu, n2, h2, w2, cout = self._a(du, "du")
        dxp, lddx, n, h, w, cin = self._a(dx, "dx")
        assert (n2, h2, w2) == (n, 2 * h, 2 * w)
        check(self._fn("plume_convT2x2_dgrad")(dup, lddu, ptr(w_dgrad), dxp, lddx, n, h, w, cin, cout,
                                            current_stream()), "plume_convT2x2_dgrad")
        self.launches += 1

    def convT_wgrad(self, x, du, dw, accumulate=False):
        xp, ldx, n, h, w, cin = self._a(x, "x")
        dup, lddu, n2, h2, w2, cout = self._a(du, "du")
        assert (n2, h2, w2) == (n, 2 * h, 2 * w)
        check(self._fn("plume_convT2x2_wgrad")(xp, ldx, dup, lddu, _f32(dw, "dw", 4 * cout * cin),
                                            int(bool(accumulate)), ptr(None), 0, n, h, w, cin, cout,
                                            current_stream()), "plume_convT2x2_wgrad")
        self.launches += 1

    # ------------------------------------------------------------------ packing
    def pack_conv3x3(self, w, wf, wd):
        cout, _, _, cin = w.shape
        check(self.lib.plume_pack_conv3x3(_f32(w, "w"), ptr(wf), ptr(wd), cout, cin, current_stream()),
              "plume_pack_conv3x3")
        self.launches += 1

    def pack_convT(self, w, wf, wd):
        _, cout, cin = w.shape
        check(self.lib.plume_pack_convT2x2(_f32(w, "w"), ptr(wf), ptr(wd), cout, cin, current_stream()),
              "plume_pack_convT2x2")
        self.launches += 1

    def pack_batch(self, jobs):
        """jobs: sequence of (kind, w, wf, wd) with kind "conv3x3" (w [Cout,3,3,Cin]) or "convT" (w [4,Cout,Cin]);
        wf / wd may be None.  One launch for all of them; the descriptor table (plume_pack_desc, device
        memory) is built once per distinct set of buffers."""
        # the shapes belong to the key: a freed buffer's address can come back for a layer of another shape
        key = tuple((k, tuple(w.shape), w.data_ptr(), ptr(wf).value, ptr(wd).value) for k, w, wf, wd in jobs)
        cached = getattr(self, "_pack_tables", None)
        if cached is None:
            cached = self._pack_tables = {}
        if key not in cached:
            import ctypes
            import struct

            blob, first = b"", 0
            for kind, w, wf, wd in jobs:
                k = 0 if kind == "conv3x3" else 1
                cout, cin = (w.shape[0], w.shape[3]) if k == 0 else (w.shape[1], w.shape[2])
                for t in (wf, wd):
                    if t is not None and t.numel() != self.planes * w.numel():
                        raise ValueError("pack_batch: operand copy has the wrong size for this precision")
                k |= 2 if self.planes == 2 else 0   # bf16x3: hi matrix followed by the lo matrix
                _f32(w, "w")
                blob += struct.pack("<QQQiiii", w.data_ptr(), ptr(wf).value or 0, ptr(wd).value or 0, k, cout, cin,
                                    first)
                first += self.lib.plume_pack_blocks(k, cout, cin)
            assert ctypes.sizeof(ctypes.c_void_p) == 8 and len(blob) == 40 * len(jobs)
            table = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(jobs[0][1].device)
            cached[key] = (table, len(jobs), first)
        table, n, total = cached[key]
        check(self.lib.plume_pack_batch(ptr(table), n, total, current_stream()), "plume_pack_batch")
        self.launches += 1

    # ------------------------------------------------------------------ bandwidth kernels
    def pad_channels(self, x, out):
        if not (x.is_cuda and out.is_cuda):
            raise TypeError("pad_channels: expected CUDA tensors (there is no CPU path)")
        assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype == torch.bfloat16
        assert x.dim() == 4 and out.dim() == 3 + self.planes     # the source is plain bf16 in either mode
        pixels = x.numel() // x.shape[-1]
        check(self._fn("plume_pad_channels")(ptr(x), x.shape[-1], ptr(out), out.shape[-1], pixels,
                                          current_stream()), "plume_pad_channels")
        self.launches += 1

    def bn_finalize(self, ssum, ssq, count, gamma, beta, eps, momentum, running_mean, running_var,
                    scale, shift, mean, invstd):
        c = ssum.numel()
        check(self.lib.plume_bn_finalize(_f64(ssum, "sum"), _f64(ssq, "sq"), int(count), _f32(gamma, "gamma"),
                                         _f32(beta, "beta"), float(eps), float(momentum),
                                         _f32(running_mean, "rm"), _f32(running_var, "rv"),
                                         _f32(scale, "scale"), _f32(shift, "shift"), _f32(mean, "mean"),
                                         _f32(invstd, "invstd"), c, current_stream()), "plume_bn_finalize")
        self.launches += 1

    def bn_fold_eval(self, gamma, beta, running_mean, running_var, conv_bias, eps, scale, shift):
        c = scale.numel()
        check(self.lib.plume_bn_fold_eval(_f32(gamma, "gamma"), _f32(beta, "beta"), _f32(running_mean, "rm"),
                                          _f32(running_var, "rv"), _f32(conv_bias, "bias"), float(eps),
                                          _f32(scale, "scale"), _f32(shift, "shift"), c, current_stream()),
              "plume_bn_fold_eval")
        self.launches += 1

    def scale_shift_act(self, y, scale, shift, relu, a):
        yp, ldy, n, h, w, c = self._a(y, "y")
        ap, lda, *_ = self._a(a, "a")
        check(self._fn("plume_scale_shift_act")(yp, ldy, _f32(scale, "scale", c), _f32(shift, "shift", c),
                                             int(bool(relu)), ap, lda, n * h * w, c, current_stream()),
              "plume_scale_shift_act")
        self.launches += 1

    def scale_shift_act_pool(self, y, scale, shift, relu, skip, pooled, argmax):
        yp, ldy, n, h, w, c = self._a(y, "y")
        sp, lds = (ptr(None), 0)
        if skip is not None:
            sp, lds, *_ = self._a(skip, "skip")
        pp, ldp, n2, h2, w2, _ = self._a(pooled, "pooled")
        assert (n2, h2, w2) == (n, h // 2, w // 2) and argmax.dtype == torch.uint8 and argmax.is_contiguous()
        check(self._fn("plume_scale_shift_act_pool")(yp, ldy, _f32(scale, "scale", c), _f32(shift, "shift", c),
                                                  int(bool(relu)), sp, lds, pp, ldp, ptr(argmax), n, h, w, c,
                                                  current_stream()), "plume_scale_shift_act_pool")
        self.launches += 1

    def maxpool_fwd(self, x, y, argmax):
        xp, ldx, n, h, w, c = self._a(x, "x")
        yp, ldy, *_ = self._a(y, "y")
        check(self._fn("plume_maxpool2x2_fwd")(xp, ldx, yp, ldy, ptr(argmax), n, h, w, c, current_stream()),
              "plume_maxpool2x2_fwd")
        self.launches += 1

    def _bn_args(self, bn, c):
        """bn = (y, scale, shift, mean, invstd, relu, sum_g, sum_gx): the BatchNorm layer whose `da` a producer
        kernel writes; returns the C-ABI argument tuple of the fused reduction."""
        y, scale, shift, mean, invstd, relu, sum_g, sum_gx = bn
        yp, ldy, *_ = self._a(y, "bn.y")
        return (yp, ldy, _f32(scale, "scale", c), _f32(shift, "shift", c), _f32(mean, "mean", c),
                _f32(invstd, "invstd", c), int(bool(relu)), _f32(sum_g, "sum_g", c), _f32(sum_gx, "sum_gx", c))

    def maxpool_bwd(self, dy, argmax, dskip, dx, bn=None):
        """bn: optional (y, scale, shift, mean, invstd, relu, sum_g, sum_gx) -- also accumulate the BatchNorm-backward
        sums of dx (what bn_bwd_reduce(dx, y, ...) would add)."""
        dyp, lddy, *_ = self._a(dy, "dy")
        dxp, lddx, n, h, w, c = self._a(dx, "dx")
        sp, lds = (ptr(None), 0)
        if dskip is not None:
            sp, lds, *_ = self._a(dskip, "dskip")
        if bn is None:
            check(self._fn("plume_maxpool2x2_bwd")(dyp, lddy, ptr(argmax), sp, lds, dxp, lddx, n, h, w, c,
                                                   current_stream()), "plume_maxpool2x2_bwd")
        else:
            check(self._fn("plume_maxpool2x2_bwd_bn")(dyp, lddy, ptr(argmax), sp, lds, dxp, lddx, *self._bn_args(bn, c),
                                                      n, h, w, c, current_stream()), "plume_maxpool2x2_bwd_bn")
        self.launches += 1

    def bn_bwd_reduce(self, da, y, scale, shift, mean, invstd, relu, sum_g, sum_gx):
        dap, ldda, n, h, w, c = self._a(da, "da")
        yp, ldy, *_ = self._a(y, "y")
        check(self._fn("plume_bn_bwd_reduce")(dap, ldda, yp, ldy, _f32(scale, "scale"), _f32(shift, "shift"),
                                           _f32(mean, "mean"), _f32(invstd, "invstd"), int(bool(relu)),
                                           _f32(sum_g, "sum_g", c), _f32(sum_gx, "sum_gx", c), n * h * w, c,
                                           current_stream()), "plume_bn_bwd_reduce")
        self.launches += 1

    def bn_bwd_apply(self, da, y, scale, shift, mean, invstd, relu, sum_g, sum_gx, dy, sum_dy,
                     dgamma=None, dbeta=None, accumulate=False):
        """sum_g / sum_gx: this batch's sums from bn_bwd_reduce (a scratch zeroed per backward pass); dgamma /
        dbeta: where the parameter gradients go (added to when accumulate)."""
        dap, ldda, n, h, w, c = self._a(da, "da")
        yp, ldy, *_ = self._a(y, "y")
        dyp, lddy, *_ = self._a(dy, "dy")
        check(self._fn("plume_bn_bwd_apply")(dap, ldda, yp, ldy, _f32(scale, "scale"), _f32(shift, "shift"),
                                          _f32(mean, "mean"), _f32(invstd, "invstd"), int(bool(relu)),
                                          _f32(sum_g, "sum_g", c), _f32(sum_gx, "sum_gx", c), dyp, lddy,
                                          _f32(sum_dy, "sum_dy"), _f32(dgamma, "dgamma", c),
                                          _f32(dbeta, "dbeta", c), int(bool(accumulate)), n * h * w, c,
                                          current_stream()), "plume_bn_bwd_apply")
        self.launches += 1

    def relu_bwd(self, da, a, dy, sum_dy):
        dap, ldda, n, h, w, c = self._a(da, "da")
        ap, lda, *_ = self._a(a, "a")
        dyp, lddy, *_ = self._a(dy, "dy")
        check(self._fn("plume_relu_bwd")(dap, ldda, ap, lda, dyp, lddy, _f32(sum_dy, "sum_dy"), n * h * w, c,
                                      current_stream()), "plume_relu_bwd")
        self.launches += 1

    def channel_sum(self, x, out):
        xp, ldx, n, h, w, c = self._a(x, "x")
        check(self._fn("plume_channel_sum")(xp, ldx, _f32(out, "out", c), n * h * w, c, current_stream()),
              "plume_channel_sum")
        self.launches += 1

    def head_fwd(self, feat, w, b, target, logits, sums):
        fp, ldf, n, h, wd, c = self._a(feat, "feat")
        check(self._fn("plume_head_fwd")(fp, ldf, _f32(w, "w", c), _f32(b, "b", 1), ptr(target),
                                      _f32(logits, "logits", n * h * wd), _f32(sums, "sums"), n * h * wd, c,
                                      current_stream()), "plume_head_fwd")
        self.launches += 1

    def head_loss(self, sums, pixels, bce_w, dice_w, eps, loss_out):
        check(self.lib.plume_head_loss(_f32(sums, "sums", 4), int(pixels), float(bce_w), float(dice_w),
                                       float(eps), _f32(loss_out, "loss", 3), current_stream()),
              "plume_head_loss")
        self.launches += 1

    def head_bwd(self, feat, w, logits, target, sums, bce_w, dice_w, eps, grad_scale, dfeat, dw, db, bn=None):
        """bn: see maxpool_bwd -- the BatchNorm layer that produced `feat`; its sums are accumulated from dfeat."""
        fp, ldf, n, h, wd, c = self._a(feat, "feat")
        dfp, lddf, *_ = self._a(dfeat, "dfeat")
        head = (fp, ldf, _f32(w, "w", c), _f32(logits, "logits"), ptr(target), _f32(sums, "sums", 4), float(bce_w),
                float(dice_w), float(eps), float(grad_scale), dfp, lddf, _f32(dw, "dw", c), _f32(db, "db", 1))
        if bn is None:
            check(self._fn("plume_head_bwd")(*head, n * h * wd, c, current_stream()), "plume_head_bwd")
        else:
            check(self._fn("plume_head_bwd_bn")(*head, *self._bn_args(bn, c), n * h * wd, c, current_stream()),
                  "plume_head_bwd_bn")
        self.launches += 1

    def cast_f32_bf16(self, src, dst):
        assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
        check(self.lib.plume_cast_f32_bf16(_f32(src, "src"), ptr(dst), src.numel(), current_stream()), "plume_cast_f32_bf16")
        self.launches += 1

    def cast_bf16_f32(self, src, dst):
        assert src.dtype == torch.bfloat16 and dst.dtype == torch.float32 and src.numel() == dst.numel()
        check(self.lib.plume_cast_bf16_f32(ptr(src), _f32(dst, "dst"), src.numel(), current_stream()), "plume_cast_bf16_f32")
        self.launches += 1

    def adam(self, param, grad, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
        n = param.numel()
        check(self.lib.plume_adam(_f32(param, "param"), _f32(grad, "grad", n), _f32(m, "m", n),
                                  _f32(v, "v", n), n, float(lr), float(beta1), float(beta2), float(eps),
                                  int(step), float(grad_scale), current_stream()), "plume_adam")
        self.launches += 1

    def adam_dev(self, param, grad, m, v, coef):
        n = param.numel()
        check(self.lib.plume_adam_dev(_f32(param, "param"), _f32(grad, "grad", n), _f32(m, "m", n),
                                      _f32(v, "v", n), n, _f32(coef, "coef", 8), current_stream()),
              "plume_adam_dev")
        self.launches += 1

    # ------------------------------------------------------------------ tiled inference
    def extract_tiles(self, scene, ys, xs, tile, tiles):
        hs, ws, cs = scene.shape
        count = ys.numel()
        assert tiles.shape[0] >= count and tiles.is_contiguous() and scene.is_contiguous()
        assert scene.dtype == torch.bfloat16 and tiles.dim() == 3 + self.planes   # the scene is plain bf16
        check(self._fn("plume_extract_tiles")(ptr(scene), hs, ws, cs, ptr(ys), ptr(xs), count, tile, ptr(tiles),
                                           tiles.shape[-1], current_stream()), "plume_extract_tiles")
        self.launches += 1

    def stitch_threshold(self, logits, ys, xs, tile, margin, logit_threshold, mask, prob=None):
        hs, ws = mask.shape
        check(self.lib.plume_stitch_threshold(_f32(logits, "logits"), ptr(ys), ptr(xs), ys.numel(), tile,
                                              margin, float(logit_threshold), ptr(mask), _f32(prob, "prob"),
                                              hs, ws, current_stream()), "plume_stitch_threshold")
        self.launches += 1

    # ------------------------------------------------------------------ label geometry
    def rasterize_hulls(self, verts, offsets, bbox, masks, ys=None, xs=None):
        """verts int32 [nv, 2] (x, y) counter-clockwise per polygon, offsets int32 [n+1], bbox int32 [n, 4]
        (xmin, ymin, xmax, ymax); masks uint8 [count, Hm, Wm] (or [Hm, Wm] without ys / xs), written 0 / 1."""
        for t, name in ((verts, "verts"), (offsets, "offsets"), (bbox, "bbox")):
            if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
                raise TypeError(f"{name}: expected a contiguous int32 CUDA tensor")
        if masks.dtype != torch.uint8 or not masks.is_cuda or not masks.is_contiguous():
            raise TypeError("masks: expected a contiguous uint8 CUDA tensor")
        n = offsets.numel() - 1
        if bbox.numel() != 4 * n or verts.dim() != 2 or verts.shape[1] != 2:
            raise ValueError("rasterize_hulls: polygon arrays do not agree")
        if ys is None:
            if masks.dim() != 2:
                raise ValueError("masks must be [Hm, Wm] when no window origins are given")
            count, (hm, wm) = 1, masks.shape
        else:
            count, hm, wm = masks.shape
            if ys.numel() != count or xs.numel() != count or ys.dtype != torch.int32 or xs.dtype != torch.int32:
                raise ValueError("ys / xs: one int32 origin per window")
        check(self.lib.plume_rasterize_hulls(ptr(verts), ptr(offsets), ptr(bbox), n, ptr(ys), ptr(xs), count, hm, wm,
                                             ptr(masks), current_stream()), "plume_rasterize_hulls")
        self.launches += 1

    def locate_fires(self, lats, lons, fire_lat, fire_lon, half_box_deg, out_rc):
        """lats / lons float64 [H, W], fire_lat / fire_lon float64 [n] -> out_rc int32 [n, 2] (row, col; -1, -1 when
        no pixel lies inside the fire's +-half_box_deg box)."""
        for t, name in ((lats, "lats"), (lons, "lons"), (fire_lat, "fire_lat"), (fire_lon, "fire_lon")):
            if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
                raise TypeError(f"{name}: expected a contiguous float64 CUDA tensor")
        if lats.dim() != 2 or lats.shape != lons.shape:
            raise ValueError("lats / lons must be [H, W] grids of the same shape")
        n = fire_lat.numel()
        if fire_lon.numel() != n or out_rc.dtype != torch.int32 or tuple(out_rc.shape) != (n, 2) or not out_rc.is_cuda \
                or not out_rc.is_contiguous():
            raise ValueError("fire arrays / out_rc do not agree")
        nbytes = int(self.lib.plume_locate_fires_workspace_bytes(n))
        ws = self._workspace(nbytes, lats.device)  # grown on demand, reused across calls (512-byte aligned)
        h, w = lats.shape
        check(self.lib.plume_locate_fires(ptr(lats), ptr(lons), h, w, ptr(fire_lat), ptr(fire_lon), n,
                                          float(half_box_deg), ptr(ws), nbytes, ptr(out_rc), current_stream()),
              "plume_locate_fires")
        self.launches += 4 if n else 0

    # ------------------------------------------------------------------ threshold sweep
    @staticmethod
    def _dev(t, dtype, name):
        if t.dtype != dtype or not t.is_cuda or not t.is_contiguous():
            raise TypeError(f"{name}: expected a contiguous {dtype} CUDA tensor")
        return ptr(t)

    def _image(self, aod, name):
        """(pointer, '_f64' or '') of a float32 / float64 image: the float64 entry points compare in float64."""
        if aod.dtype == torch.float64:
            return self._dev(aod, torch.float64, "aod"), name + "_f64"
        return self._dev(aod, torch.float32, "aod"), name

    def threshold_masks(self, aod, thresholds, masks):
        """aod float32 / float64 [H, W], thresholds float64 [T] -> masks uint8 [T, H, W] = dilate(erode(aod > t))."""
        h, w = aod.shape
        t = thresholds.numel()
        if tuple(masks.shape) != (t, h, w):
            raise ValueError("masks must be [T, H, W]")
        a, fn = self._image(aod, "plume_threshold_masks")
        check(getattr(self.lib, fn)(a, h, w, self._dev(thresholds, torch.float64, "thresholds"), t,
                                    self._dev(masks, torch.uint8, "masks"), current_stream()), fn)
        self.launches += 1

    def label_components(self, masks, labels, sizes):
        """masks uint8 [T, H, W] -> labels int32 [T, H, W] (-1 background, else the component's smallest pixel
        index), sizes int32 [T, H, W] (component size at its canonical index)."""
        t, h, w = masks.shape
        if labels.shape != masks.shape or sizes.shape != masks.shape:
            raise ValueError("labels / sizes must have the masks' shape")
        check(self.lib.plume_label_components(self._dev(masks, torch.uint8, "masks"), t, h, w,
                                              self._dev(labels, torch.int32, "labels"),
                                              self._dev(sizes, torch.int32, "sizes"), current_stream()),
              "plume_label_components")
        self.launches += 3

    def fire_extents(self, labels, sizes, fire_rc, win, extents):
        """fire_rc int32 [n, 2] -> extents int32 [T, n]: size of the component nearest to each fire in its window."""
        t, h, w = labels.shape
        n = fire_rc.shape[0]
        if tuple(extents.shape) != (t, n) or fire_rc.dim() != 2 or fire_rc.shape[1] != 2:
            raise ValueError("extents must be [T, n_fires], fire_rc [n_fires, 2]")
        check(self.lib.plume_fire_extents(self._dev(labels, torch.int32, "labels"), self._dev(sizes, torch.int32, "sizes"),
                                          t, h, w, self._dev(fire_rc, torch.int32, "fire_rc"), n, int(win),
                                          self._dev(extents, torch.int32, "extents"), current_stream()),
              "plume_fire_extents")
        self.launches += 1

    # bit-plane form: masks packed 32 pixels per word (int32 tensors [T, H, ceil(W / 32)], bit i of a word = pixel 32 s + i)
    def sweep_workspace_bytes(self, h, w, t) -> int:
        return int(self.lib.plume_sweep_workspace_bytes(int(h), int(w), int(t)))

    def threshold_mask_bits(self, aod, thresholds, bits):
        """aod float32 / float64 [H, W], thresholds float64 [T] -> bits int32 [T, H, ceil(W / 32)]."""
        h, w = aod.shape
        t = thresholds.numel()
        if tuple(bits.shape) != (t, h, (w + 31) // 32):
            raise ValueError("bits must be [T, H, ceil(W / 32)]")
        a, fn = self._image(aod, "plume_threshold_mask_bits")
        check(getattr(self.lib, fn)(a, h, w, self._dev(thresholds, torch.float64, "thresholds"), t,
                                    self._dev(bits, torch.int32, "bits"), current_stream()), fn)
        self.launches += 1

    def pack_mask_bits(self, masks, bits):
        """masks uint8 [T, H, W] -> bits int32 [T, H, ceil(W / 32)]."""
        t, h, w = masks.shape
        if tuple(bits.shape) != (t, h, (w + 31) // 32):
            raise ValueError("bits must be [T, H, ceil(W / 32)]")
        check(self.lib.plume_pack_mask_bits(self._dev(masks, torch.uint8, "masks"), t, h, w,
                                            self._dev(bits, torch.int32, "bits"), current_stream()),
              "plume_pack_mask_bits")
        self.launches += 1

    def bits_extents(self, bits, w, fire_rc, win, workspace, extents):
        """bits int32 [T, H, ceil(W / 32)] -> extents int32 [T, n]: components over runs of set bits + nearest
        component per fire; workspace: uint8 tensor of at least sweep_workspace_bytes(H, W, T)."""
        t, h, segs = bits.shape
        n = fire_rc.shape[0]
        if segs != (w + 31) // 32 or tuple(extents.shape) != (t, n) or fire_rc.dim() != 2 or fire_rc.shape[1] != 2:
            raise ValueError("bits must be [T, H, ceil(W / 32)], extents [T, n_fires], fire_rc [n_fires, 2]")
        check(self.lib.plume_bits_extents(self._dev(bits, torch.int32, "bits"), t, h, int(w),
                                          self._dev(fire_rc, torch.int32, "fire_rc"), n, int(win),
                                          self._dev(workspace, torch.uint8, "workspace"), workspace.numel(),
                                          self._dev(extents, torch.int32, "extents"), current_stream()),
              "plume_bits_extents")
        self.launches += 4 if n and t else 0

    def fire_components(self, bits, w, fire_rc, plane_of_fire, win, workspace, comp, stats):
        """After bits_extents / sweep_extents on the same bits and workspace: comp int32 [n, H, ceil(W / 32)] = the component
        nearest to fire f in plane plane_of_fire[f] (int32 [n], negative = none), stats int32 [n, 8]."""
        t, h, segs = bits.shape
        n = fire_rc.shape[0]
        if (segs != (w + 31) // 32 or tuple(comp.shape) != (n, h, segs) or tuple(stats.shape) != (n, 8)
                or tuple(plane_of_fire.shape) != (n,) or fire_rc.dim() != 2 or fire_rc.shape[1] != 2):
            raise ValueError("comp must be [n_fires, H, ceil(W / 32)], stats [n_fires, 8], plane_of_fire [n_fires]")
        check(self.lib.plume_fire_components(self._dev(bits, torch.int32, "bits"), t, h, int(w),
                                             self._dev(fire_rc, torch.int32, "fire_rc"),
                                             self._dev(plane_of_fire, torch.int32, "plane_of_fire"), n, int(win),
                                             self._dev(workspace, torch.uint8, "workspace"), workspace.numel(),
                                             self._dev(comp, torch.int32, "comp"), self._dev(stats, torch.int32, "stats"),
                                             current_stream()), "plume_fire_components")
        self.launches += 1 if n and t else 0

    def sweep_extents(self, aod, thresholds, fire_rc, win, workspace, extents):
        """aod float32 / float64 [H, W], thresholds float64 [T], fire_rc int32 [n, 2] -> extents int32 [T, n]."""
        h, w = aod.shape
        t = thresholds.numel()
        n = fire_rc.shape[0]
        if tuple(extents.shape) != (t, n) or fire_rc.dim() != 2 or fire_rc.shape[1] != 2:
            raise ValueError("extents must be [T, n_fires], fire_rc [n_fires, 2]")
        a, fn = self._image(aod, "plume_sweep_extents")
        check(getattr(self.lib, fn)(a, h, w, self._dev(thresholds, torch.float64, "thresholds"), t,
                                    self._dev(fire_rc, torch.int32, "fire_rc"), n, int(win),
                                    self._dev(workspace, torch.uint8, "workspace"), workspace.numel(),
                                    self._dev(extents, torch.int32, "extents"), current_stream()), fn)
        self.launches += 4 if n and t else 0

    # ------------------------------------------------------------------ nearest-valid fill (interpolate_aod_nearest)
    def fill_nearest_workspace_bytes(self, h, w) -> int:
        return int(self.lib.plume_fill_nearest_workspace_bytes(int(h), int(w)))

    def fill_nearest(self, aod, null_value, workspace, out):
        """aod float32 / float64 [H, W] -> out (same dtype): null pixels take the value of the nearest valid pixel."""
        h, w = aod.shape
        if out.shape != aod.shape or out.dtype != aod.dtype or out.data_ptr() == aod.data_ptr():
            raise ValueError("out must be a distinct tensor with the image's shape and dtype")
        a, fn = self._image(aod, "plume_fill_nearest")
        check(getattr(self.lib, fn)(a, h, w, float(null_value), self._dev(workspace, torch.uint8, "workspace"),
                                    workspace.numel(), self._dev(out, aod.dtype, "out"), current_stream()), fn)
        self.launches += 3

    # ------------------------------------------------------------------ UTM projection / nearest-neighbour resampling
    def utm_zone_histogram(self, lons, hist):
        """lons float64 (any shape) -> hist int32 [64]: hist[z] = pixels whose UTM zone is z (tools.py:27-28)."""
        check(self.lib.plume_utm_zone_histogram(self._dev(lons, torch.float64, "lons"), lons.numel(),
                                                self._dev(hist, torch.int32, "hist"), current_stream()),
              "plume_utm_zone_histogram")
        self.launches += 1

    def sinusoidal_grid_latlon(self, x_start, x_stop, y_start, y_stop, radius, lat, lon):
        """lat / lon float64 [ny, nx] (degrees) of the sinusoidal grid linspace(x_start, x_stop, nx) x linspace(y_start, y_stop, ny)."""
        ny, nx = lat.shape
        if lon.shape != lat.shape:
            raise ValueError("lat and lon must have the same [ny, nx] shape")
        check(self.lib.plume_sinusoidal_grid_latlon(float(x_start), float(x_stop), float(y_start), float(y_stop), ny, nx,
                                                    float(radius), self._dev(lat, torch.float64, "lat"),
                                                    self._dev(lon, torch.float64, "lon"), current_stream()),
              "plume_sinusoidal_grid_latlon")
        self.launches += 1

    def utm_forward(self, lats, lons, zone, x, y):
        check(self.lib.plume_utm_forward(self._dev(lats, torch.float64, "lats"), self._dev(lons, torch.float64, "lons"),
                                         lats.numel(), int(zone), self._dev(x, torch.float64, "x"),
                                         self._dev(y, torch.float64, "y"), current_stream()), "plume_utm_forward")
        self.launches += 1

    def utm_inverse(self, x, y, zone, lats, lons):
        check(self.lib.plume_utm_inverse(self._dev(x, torch.float64, "x"), self._dev(y, torch.float64, "y"), x.numel(),
                                         int(zone), self._dev(lats, torch.float64, "lats"),
                                         self._dev(lons, torch.float64, "lons"), current_stream()), "plume_utm_inverse")
        self.launches += 1

    def resample_nearest_index(self, src_lats, src_lons, zone, extent, x_size, y_size, radius, out_idx):
        """extent = (min_x, min_y, max_x, max_y) outer edges of the x_size x y_size target area in UTM metres of `zone`;
        out_idx int32 [y_size, x_size]: flat index of the nearest swath pixel within `radius` metres, else -1."""
        n = src_lats.numel()
        if src_lons.numel() != n or tuple(out_idx.shape) != (y_size, x_size):
            raise ValueError("resample_nearest_index: swath arrays / out_idx do not agree")
        ex = [float(v) for v in extent]
        nbytes = int(self.lib.plume_resample_workspace_bytes(n, ex[0], ex[1], ex[2], ex[3], float(radius)))
        ws = self._workspace(nbytes, src_lats.device)
        check(self.lib.plume_resample_nearest_index(self._dev(src_lats, torch.float64, "src_lats"),
                                                    self._dev(src_lons, torch.float64, "src_lons"), n, int(zone),
                                                    ex[0], ex[1], ex[2], ex[3], int(x_size), int(y_size), float(radius),
                                                    ptr(ws), nbytes, self._dev(out_idx, torch.int32, "out_idx"),
                                                    current_stream()), "plume_resample_nearest_index")
        self.launches += 4

    def gather_fill(self, src, idx, fill_value, out):
        if src.dtype not in (torch.float32, torch.float64) or out.dtype != src.dtype:
            raise TypeError("gather_fill: float32 or float64 images")
        check(self.lib.plume_gather_fill(self._dev(src, src.dtype, "src"), src.element_size(),
                                         self._dev(idx, torch.int32, "idx"), idx.numel(), float(fill_value),
                                         self._dev(out, out.dtype, "out"), current_stream()), "plume_gather_fill")
        self.launches += 1
