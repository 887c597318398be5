"""Host side of the B200 UNet: buffers, forward / backward / optimizer schedule over the operator layer.

There is no autograd here: the backward pass is an explicit schedule of the dgrad / wgrad / BatchNorm /
pool / head kernels, mirroring the forward one.  All activations are NHWC bf16 and preallocated per
(N, H, W); the decoder's concat is zero-copy (the encoder writes the skip into channels [0, C) of the
concat buffer, the transposed conv writes [C, 2C)).  Parameters, gradients and Adam moments each live in
ONE flat fp32 buffer laid out in reverse execution order (spec.build_layout), so the optimizer is a
single launch and data-parallel gradient buckets are contiguous slices that become ready front to back
while the backward pass is still running.

The definition being implemented is UNetSpec (spec.py); the reference repository has no model code
(SURVEY.md section 0).  ``state_dict()`` / ``load_state_dict()`` use the oracle's key names and NCHW fp32
tensors, so weights round-trip with ``oracle/unet_ref.py`` and ``torch.save`` files are interchangeable.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from .spec import UNetSpec, build_layout, conv_layers, up_layers

BF16 = torch.bfloat16


class _Buffers:
    """Activation / gradient workspace for one input geometry."""

    def __init__(self, n: int, h: int, w: int):
        self.n, self.h, self.w = n, h, w


class UNetB200:
    def __init__(self, spec: UNetSpec = UNetSpec(), ops=None, device="cuda", seed: Optional[int] = 0,
                 process_group=None, bucket_mb: float = 25.0):
        if ops is None:
            from .ops import CudaOps  # fails loudly without the library / a GPU

            ops = CudaOps(precision=spec.precision)
        self.spec, self.ops = spec, ops
        # bf16 on the GPU; the CPU operator oracle may ask for fp32 buffers to check the schedule exactly
        self.act_dtype = getattr(ops, "act_dtype", BF16)
        # bf16x3 precision: every activation is [N, H, W, 2, C] (hi / lo planes) and every bf16 weight copy a hi
        # matrix followed by a lo matrix; the schedule below is the same (channel slices index the last dim)
        self.planes = getattr(ops, "planes", 1)
        if spec.precision == "bf16x3" and self.planes != 2 and getattr(ops, "name", "") == "b200":
            raise ValueError("spec.precision is 'bf16x3' but the operator layer was built for bf16")
        self.device = torch.device(device)
        self.layout = build_layout(spec)
        self.convs, self.ups = conv_layers(spec), up_layers(spec)
        self.use_bn = spec.norm == "batch"
        dev = self.device
        T = self.layout.total
        self.params = torch.zeros(T, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(T, dtype=torch.float32, device=dev)
        self.adam_m = torch.zeros(T, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(T, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.num_batches_tracked = 0
        self.training = True

        # bf16 operand copies of the GEMM weights (forward layout and dgrad layout)
        off = 0
        self._wslot: Dict[str, tuple] = {}
        for name, L in self.convs.items():
            n = self.planes * L.cout * 9 * L.cin_k
            self._wslot[name] = (off, n)
            off += (n + 127) // 128 * 128
        for name, U in self.ups.items():
            n = self.planes * 4 * U.cout * U.cin
            self._wslot[name] = (off, n)
            off += (n + 127) // 128 * 128
        self.w_fwd = torch.zeros(off, dtype=self.act_dtype, device=dev)
        self.w_dgrad = torch.zeros(off, dtype=self.act_dtype, device=dev)

        # per-BatchNorm state: running statistics + per-step statistics / fused coefficients
        self._bn: Dict[str, Dict[str, torch.Tensor]] = {}
        zoff = 0          # fp64 region zeroed every step: [sum | sq] per BN layer, then the head's loss sums (as fp32)
        soff = 0          # persistent per-layer vectors
        zmap, smap = {}, {}
        for name, L in self.convs.items():
            zmap[name] = zoff
            zoff += 2 * L.cout
            smap[name] = soff
            soff += 8 * L.cout
        self._head_sums_off = zoff
        zoff += 4         # 4 doubles = room for 8 floats
        # the BatchNorm statistics accumulate in fp64 (E[y^2] - E[y]^2 cancels in fp32 when |mean| >> std); the head's
        # four loss sums are fp32 views into the tail of the same buffer so that one memset zeroes everything
        self._zero_region = torch.zeros(zoff, dtype=torch.float64, device=dev)
        self._stat_region = torch.zeros(soff, dtype=torch.float32, device=dev)
        # BatchNorm-backward sums of THIS backward pass ([sum_g | sum_gx] per layer, same offsets as the forward
        # statistics): zeroed at the start of every backward, never the gradient slots themselves -- with
        # micro-batch accumulation those hold the previous slices' sums, which must not enter bn_bwd_apply
        self._bwd_region = torch.zeros(max(self._head_sums_off, 1), dtype=torch.float32, device=dev)
        for name, L in self.convs.items():
            c, z, s = L.cout, zmap[name], smap[name]
            st = self._stat_region
            self._bn[name] = {
                "sum": self._zero_region[z:z + c], "sq": self._zero_region[z + c:z + 2 * c],
                "scale": st[s:s + c], "shift": st[s + c:s + 2 * c], "mean": st[s + 2 * c:s + 3 * c],
                "invstd": st[s + 3 * c:s + 4 * c], "running_mean": st[s + 4 * c:s + 5 * c],
                "running_var": st[s + 5 * c:s + 6 * c], "fold_scale": st[s + 6 * c:s + 7 * c],
                "fold_shift": st[s + 7 * c:s + 8 * c],
                "sum_g": self._bwd_region[z:z + c], "sum_gx": self._bwd_region[z + c:z + 2 * c],
            }
            self._bn[name]["running_var"].fill_(1.0)
        self.head_sums = self._zero_region[self._head_sums_off:self._head_sums_off + 4].view(torch.float32)[:4]
        self.loss_out = torch.zeros(3, dtype=torch.float32, device=dev)

        self._buf: Optional[_Buffers] = None
        self._packed_version = -1
        self._param_version = 0
        self._stats_version = 0     # advances with every training forward (the running statistics change)
        self._folded_version = (-1, -1)

        # data parallel
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(process_group)
        self._buckets = self._make_buckets(bucket_mb)
        self._pending: List = []
        self._pack_jobs = None
        # Weight-gradient kernels go to a second stream: they are tensor/L2 bound and leave registers and
        # shared memory for one block of the (HBM bound) BatchNorm-backward kernels of the next layer, so
        # the two overlap on the same SMs.  PLUME_NO_WGRAD_OVERLAP=1 keeps everything on one stream.
        self.overlap_wgrad = self.device.type == "cuda" and not os.environ.get("PLUME_NO_WGRAD_OVERLAP")
        self._side: Optional["torch.cuda.Stream"] = None   # weight gradients (lowest priority)
        self._chain: Optional["torch.cuda.Stream"] = None  # the backward pass's critical chain (high priority)
        self._gy_busy: Dict[int, "torch.cuda.Event"] = {}
        self._tail_event: Optional["torch.cuda.Event"] = None  # side stream: all wgrads before enc0's are done
        self._tail_open = False                                # backward returned without joining the side stream
        self._defer_ok = not os.environ.get("PLUME_NO_DEFER_TAIL")
        self.fuse_bn_reduce = not os.environ.get("PLUME_NO_FUSED_BN_REDUCE")   # A/B switch
        self.fuse_head_bn = bool(os.environ.get("PLUME_FUSE_HEAD_BN"))
        # PLUME_GRAD_COMM=bf16: gradient buckets are rounded to bf16 for the all-reduce (half the bytes on the wire,
        # summed in bf16 by NCCL) and widened back into the fp32 gradient buffer; default fp32
        self.grad_comm_bf16 = os.environ.get("PLUME_GRAD_COMM", "fp32").lower() == "bf16" and self.world > 1
        # PLUME_ADAM_PER_BUCKET=1: Adam per gradient bucket as its all-reduce completes instead of one launch after the
        # last one (opt-in: measured at 2 GPUs, configs[1] 11.59 -> 11.52 ms, wide model 26.7 -> 26.3 ms, both inside
        # the run-to-run spread; gpurun_out/r2ad)
        self.adam_per_bucket = os.environ.get("PLUME_ADAM_PER_BUCKET", "0") == "1"
        self._comm_buf: Optional[torch.Tensor] = None
        self._tail_offset = min(s.offset for k, s in self.layout.slots.items() if k.startswith("enc0."))

        if seed is not None:
            self.init_parameters(seed)

    # ------------------------------------------------------------------ parameter access
    def p(self, key: str) -> torch.Tensor:
        s = self.layout.slots[key]
        return self.params[s.offset:s.offset + s.numel].view(s.shape)

    def g(self, key: str) -> torch.Tensor:
        s = self.layout.slots[key]
        return self.grads[s.offset:s.offset + s.numel].view(s.shape)

    def wf(self, name: str) -> torch.Tensor:
        o, n = self._wslot[name]
        return self.w_fwd[o:o + n]

    def wd(self, name: str) -> torch.Tensor:
        o, n = self._wslot[name]
        return self.w_dgrad[o:o + n]

    def num_parameters(self) -> int:
        """Logical parameter count (the oracle's), excluding channel padding."""
        n = 0
        for k, t in self.state_dict().items():
            if "running_" not in k and "num_batches" not in k:
                n += t.numel()
        return n

    # ------------------------------------------------------------------ init / state dict
    def init_parameters(self, seed: int = 0) -> None:
        """PyTorch default initialisation (kaiming-uniform a=sqrt(5), bias U(+-1/sqrt(fan_in)); BN 1/0)
        drawn in the oracle's construction order under torch.manual_seed(seed), so that this model and
        ``UNetRef`` built under the same seed hold identical weights."""
        import torch.nn as nn

        spec, d = self.spec, self.spec.depth
        torch.manual_seed(seed)
        sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

        def double(prefix, cin, cout):
            for i, (ci, co) in enumerate(((cin, cout), (cout, cout)), start=1):
                conv = nn.Conv2d(ci, co, 3, padding=1)
                sd[f"{prefix}.conv{i}.weight"], sd[f"{prefix}.conv{i}.bias"] = conv.weight.data, conv.bias.data
            # note: nn.Module construction order inside DoubleConv is conv1, conv2, then the BNs (no RNG)

        for l in range(d):
            double(f"enc{l}", spec.in_channels if l == 0 else spec.channels(l - 1), spec.channels(l))
        double("bottleneck", spec.channels(d - 1), spec.channels(d))
        for l in reversed(range(d)):
            c = spec.channels(l)
            up = nn.ConvTranspose2d(2 * c, c, 2, stride=2)
            sd[f"up{l}.weight"], sd[f"up{l}.bias"] = up.weight.data, up.bias.data
            double(f"dec{l}", 2 * c, c)
        head = nn.Conv2d(spec.base_filters, 1, 1)
        sd["head.weight"], sd["head.bias"] = head.weight.data, head.bias.data
        self.load_state_dict(sd, strict=False)

    def load_state_dict(self, sd, strict: bool = True) -> None:
        """Accepts the oracle's state_dict (NCHW fp32).  With strict=False missing BatchNorm entries keep
        their defaults (weight 1, bias 0, running mean 0 / var 1)."""
        dev = self.device
        with torch.no_grad():
            self.params.zero_()
            for name, L in self.convs.items():
                w = sd[f"{name}.weight"].to(dev, torch.float32)  # [Cout][Cin][3][3]
                if tuple(w.shape) != (L.cout, L.cin, 3, 3):
                    raise ValueError(f"{name}.weight: expected {(L.cout, L.cin, 3, 3)}, got {tuple(w.shape)}")
                self.p(f"{name}.weight")[..., :L.cin].copy_(w.permute(0, 2, 3, 1))
                self.p(f"{name}.bias").copy_(sd[f"{name}.bias"].to(dev, torch.float32))
                if self.use_bn:
                    bn = L.bn
                    st = self._bn[name]
                    if f"{bn}.weight" in sd:
                        self.p(f"{bn}.weight").copy_(sd[f"{bn}.weight"].to(dev, torch.float32))
                        self.p(f"{bn}.bias").copy_(sd[f"{bn}.bias"].to(dev, torch.float32))
                        st["running_mean"].copy_(sd[f"{bn}.running_mean"].to(dev, torch.float32))
                        st["running_var"].copy_(sd[f"{bn}.running_var"].to(dev, torch.float32))
                        if f"{bn}.num_batches_tracked" in sd:
                            self.num_batches_tracked = int(sd[f"{bn}.num_batches_tracked"])
                    elif strict:
                        raise KeyError(f"{bn}.weight")
                    else:
                        self.p(f"{bn}.weight").fill_(1.0)
                        self.p(f"{bn}.bias").zero_()
                        st["running_mean"].zero_()
                        st["running_var"].fill_(1.0)
            for name, U in self.ups.items():
                w = sd[f"{name}.weight"].to(dev, torch.float32)  # ConvTranspose2d: [Cin][Cout][2][2]
                if tuple(w.shape) != (U.cin, U.cout, 2, 2):
                    raise ValueError(f"{name}.weight: expected {(U.cin, U.cout, 2, 2)}, got {tuple(w.shape)}")
                self.p(f"{name}.weight").copy_(w.permute(2, 3, 1, 0).reshape(4, U.cout, U.cin))
                self.p(f"{name}.bias").copy_(sd[f"{name}.bias"].to(dev, torch.float32))
            self.p("head.weight").copy_(sd["head.weight"].to(dev, torch.float32).reshape(-1))
            self.p("head.bias").copy_(sd["head.bias"].to(dev, torch.float32).reshape(-1))
        self._param_version += 1

    def _export(self, flat: torch.Tensor, with_buffers: bool) -> "OrderedDict[str, torch.Tensor]":
        """A flat buffer (parameters or gradients) in the oracle's key names / NCHW shapes."""
        spec, d = self.spec, self.spec.depth
        out: "OrderedDict[str, torch.Tensor]" = OrderedDict()

        def get(key):
            sl = self.layout.slots[key]
            return flat[sl.offset:sl.offset + sl.numel].view(sl.shape)

        def double(prefix):
            for i in (1, 2):
                L = self.convs[f"{prefix}.conv{i}"]
                w = get(f"{L.name}.weight")[..., :L.cin].permute(0, 3, 1, 2).contiguous()
                out[f"{L.name}.weight"] = w.detach().cpu().clone()
                out[f"{L.name}.bias"] = get(f"{L.name}.bias").detach().cpu().clone()
            if self.use_bn:
                for i in (1, 2):
                    L = self.convs[f"{prefix}.conv{i}"]
                    st = self._bn[L.name]
                    out[f"{L.bn}.weight"] = get(f"{L.bn}.weight").detach().cpu().clone()
                    out[f"{L.bn}.bias"] = get(f"{L.bn}.bias").detach().cpu().clone()
                    if with_buffers:
                        out[f"{L.bn}.running_mean"] = st["running_mean"].detach().cpu().clone()
                        out[f"{L.bn}.running_var"] = st["running_var"].detach().cpu().clone()
                        out[f"{L.bn}.num_batches_tracked"] = torch.tensor(self.num_batches_tracked,
                                                                          dtype=torch.long)

        for l in range(d):
            double(f"enc{l}")
        double("bottleneck")
        for l in reversed(range(d)):
            U = self.ups[f"up{l}"]
            w = get(f"up{l}.weight").view(2, 2, U.cout, U.cin).permute(3, 2, 0, 1).contiguous()
            out[f"up{l}.weight"] = w.detach().cpu().clone()
            out[f"up{l}.bias"] = get(f"up{l}.bias").detach().cpu().clone()
            double(f"dec{l}")
        out["head.weight"] = get("head.weight").detach().cpu().clone().view(1, spec.base_filters, 1, 1)
        out["head.bias"] = get("head.bias").detach().cpu().clone()
        return out

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """Oracle-compatible state dict (same keys, shapes, dtypes and ordering as ``UNetRef``)."""
        return self._export(self.params, with_buffers=True)

    def grad_dict(self) -> "OrderedDict[str, torch.Tensor]":
        """Current gradients under the oracle's parameter names (for parity checks)."""
        self._close_tail()
        return self._export(self.grads, with_buffers=False)

    def optimizer_state(self) -> dict:
        return {"step": self.step_count, "m": self.adam_m.detach().cpu().clone(),
                "v": self.adam_v.detach().cpu().clone()}

    def load_optimizer_state(self, st: dict) -> None:
        self.step_count = int(st["step"])
        self.adam_m.copy_(st["m"].to(self.device))
        self.adam_v.copy_(st["v"].to(self.device))

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ------------------------------------------------------------------ buffers
    def _ensure_buffers(self, n: int, h: int, w: int) -> _Buffers:
        b = self._buf
        if b is not None and (b.n, b.h, b.w) == (n, h, w):
            return b
        spec, d, dev = self.spec, self.spec.depth, self.device
        div = spec.divisor()
        if h % div or w % div:
            raise ValueError(f"tile height/width must be multiples of {div} (depth {d}); got {h}x{w}")

        def act(hh, ww, c):
            if self.planes == 2:
                return torch.empty(n, hh, ww, 2, c, dtype=self.act_dtype, device=dev)
            return torch.empty(n, hh, ww, c, dtype=self.act_dtype, device=dev)

        b = _Buffers(n, h, w)
        b.x0 = act(h, w, spec.cin_padded) if (spec.cin_padded != spec.in_channels or self.planes == 2) else None
        b.y1, b.a1, b.y2, b.cat, b.pool, b.am = [], [], [], [], [], []
        b.dy1, b.da1, b.dy2, b.da2 = [], [], [], []          # decoder activations (raw / activated)
        b.g_a, b.g_y, b.g_cat, b.g_pool = [], [], [], []       # gradient scratch per level
        for l in range(d):
            c, hh, ww = spec.channels(l), h >> l, w >> l
            b.y1.append(act(hh, ww, c) if self.use_bn else None)
            b.a1.append(act(hh, ww, c))
            b.y2.append(act(hh, ww, c) if self.use_bn else None)
            b.cat.append(act(hh, ww, 2 * c))
            b.pool.append(act(hh // 2, ww // 2, c))
            b.am.append(torch.empty(n, hh // 2, ww // 2, c, dtype=torch.uint8, device=dev))
            b.dy1.append(act(hh, ww, c) if self.use_bn else None)
            b.da1.append(act(hh, ww, c))
            b.dy2.append(act(hh, ww, c) if self.use_bn else None)
            b.da2.append(act(hh, ww, c))
            b.g_a.append(act(hh, ww, c))
            b.g_y.append((act(hh, ww, c), act(hh, ww, c)))   # ping-pong: wgrad reads one while the next is written
            b.g_cat.append(act(hh, ww, 2 * c))
            b.g_pool.append(act(hh // 2, ww // 2, c))
        cb, hb, wb = spec.channels(d), h >> d, w >> d
        b.by1 = act(hb, wb, cb) if self.use_bn else None
        b.ba1 = act(hb, wb, cb)
        b.by2 = act(hb, wb, cb) if self.use_bn else None
        b.ba2 = act(hb, wb, cb)
        b.bg_a = act(hb, wb, cb)
        b.bg_y = (act(hb, wb, cb), act(hb, wb, cb))
        b.logits = torch.empty(n, h, w, dtype=torch.float32, device=dev)
        self._buf = b
        return b

    def activation_bytes(self) -> int:
        b = self._buf
        if b is None:
            return 0
        tot = 0
        for v in vars(b).values():
            for t in (v if isinstance(v, list) else [v]):
                for u in (t if isinstance(t, tuple) else (t,)):
                    if isinstance(u, torch.Tensor):
                        tot += u.numel() * u.element_size()
        return tot

    # ------------------------------------------------------------------ weights -> bf16 operands
    def pack_weights(self) -> None:
        if self._packed_version == self._param_version:
            return
        if self._pack_jobs is None:
            jobs = []
            for name in self.convs:
                need_dgrad = name != "enc0.conv1"
                jobs.append(("conv3x3", self.p(f"{name}.weight"), self.wf(name), self.wd(name) if need_dgrad else None))
            for name in self.ups:
                jobs.append(("convT", self.p(f"{name}.weight"), self.wf(name), self.wd(name)))
            self._pack_jobs = jobs
        self.ops.pack_batch(self._pack_jobs)  # every layer in one launch
        self._packed_version = self._param_version

    def _fold_bn(self) -> None:
        version = (self._param_version, self._stats_version)
        if not self.use_bn or self._folded_version == version:
            return
        for name, L in self.convs.items():
            st = self._bn[name]
            self.ops.bn_fold_eval(self.p(f"{L.bn}.weight"), self.p(f"{L.bn}.bias"), st["running_mean"],
                                  st["running_var"], self.p(f"{name}.bias"), self.spec.bn_eps,
                                  st["fold_scale"], st["fold_shift"])
        self._folded_version = version

    # ------------------------------------------------------------------ forward
    def _prep_input(self, x: torch.Tensor, b: _Buffers) -> torch.Tensor:
        spec = self.spec
        if x.dim() == 5 and self.planes == 2 and x.shape[-1] == spec.cin_padded and x.is_contiguous():
            return x  # bf16x3 tiles already in the split format (ScenePredictor cuts them that way)
        if x.dtype != self.act_dtype or x.dim() != 4:
            raise ValueError(f"input must be a 4-D NHWC {self.act_dtype} tensor, got {x.dtype} {tuple(x.shape)}")
        if x.shape[-1] == spec.cin_padded and x.is_contiguous() and self.planes == 1:
            return x  # already zero padded to the first layer's K
        if x.shape[-1] != spec.in_channels:
            raise ValueError(f"input must have {spec.in_channels} channels, got {x.shape[-1]}")
        if spec.in_channels % 8:
            raise ValueError("in_channels must be a multiple of 8 (16-byte NHWC vectors); pad on the host")
        # The C ABI could read the 8-band tensor directly (plume_conv3x3_fwd with ldx < Cin: TMA zero-fills
        # the missing channels), but boxes with 16 valid bytes per pixel load slowly: measured 205 us vs
        # 166 us for the first convolution and 226 us vs 155 us for its weight gradient, more than the
        # 58 us this padded copy costs.
        self.ops.pad_channels(x.contiguous(), b.x0)
        return b.x0

    def _conv_block_train(self, name: str, x, y, a, pool=None):
        """conv3x3 (+bias) -> [batch statistics -> BatchNorm] -> ReLU, optionally fused with the 2x2 pool.
        pool = (skip_view, pooled, argmax): `a` is then the skip view itself."""
        ops, spec = self.ops, self.spec
        L = self.convs[name]
        bias = self.p(f"{name}.bias")
        if not self.use_bn:
            ops.conv3x3_fwd(x, self.wf(name), None, bias, 1, a)
            if pool is not None:
                ops.maxpool_fwd(a, pool[1], pool[2])
            return
        st = self._bn[name]
        ops.conv3x3_fwd(x, self.wf(name), None, bias, 0, y, st["sum"], st["sq"])
        count = y.shape[0] * y.shape[1] * y.shape[2]
        ops.bn_finalize(st["sum"], st["sq"], count, self.p(f"{L.bn}.weight"), self.p(f"{L.bn}.bias"),
                        spec.bn_eps, spec.bn_momentum, st["running_mean"], st["running_var"], st["scale"],
                        st["shift"], st["mean"], st["invstd"])
        if pool is None:
            ops.scale_shift_act(y, st["scale"], st["shift"], 1, a)
        else:
            ops.scale_shift_act_pool(y, st["scale"], st["shift"], 1, pool[0], pool[1], pool[2])

    def _conv_block_eval(self, name: str, x, a):
        if self.use_bn:
            st = self._bn[name]
            self.ops.conv3x3_fwd(x, self.wf(name), st["fold_scale"], st["fold_shift"], 1, a)
        else:
            self.ops.conv3x3_fwd(x, self.wf(name), None, self.p(f"{name}.bias"), 1, a)

    def forward(self, x: torch.Tensor, target: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: [N,H,W,C_in] bf16 (device).  Training mode: batch statistics, activations kept for backward;
        if `target` (uint8 [N,H,W]) is given the loss sums are accumulated.  Returns fp32 logits [N,H,W]."""
        spec, d, ops = self.spec, self.spec.depth, self.ops
        self._close_tail()
        n, h, w = x.shape[:3]
        b = self._ensure_buffers(n, h, w)
        self.pack_weights()
        train = self.training
        if train:
            self._zero_region.zero_()
            if self.use_bn:
                self.num_batches_tracked += 1
                self._stats_version += 1
        else:
            self._fold_bn()
        cur = self._prep_input(x, b)
        b.x_in = cur
        for l in range(d):
            c = spec.channels(l)
            skip = b.cat[l][..., :c]
            if train:
                self._conv_block_train(f"enc{l}.conv1", cur, b.y1[l], b.a1[l])
                self._conv_block_train(f"enc{l}.conv2", b.a1[l], b.y2[l], skip, pool=(skip, b.pool[l], b.am[l]))
            else:
                self._conv_block_eval(f"enc{l}.conv1", cur, b.a1[l])
                self._conv_block_eval(f"enc{l}.conv2", b.a1[l], skip)
                ops.maxpool_fwd(skip, b.pool[l], b.am[l])
            cur = b.pool[l]
        if train:
            self._conv_block_train("bottleneck.conv1", cur, b.by1, b.ba1)
            self._conv_block_train("bottleneck.conv2", b.ba1, b.by2, b.ba2)
        else:
            self._conv_block_eval("bottleneck.conv1", cur, b.ba1)
            self._conv_block_eval("bottleneck.conv2", b.ba1, b.ba2)
        cur = b.ba2
        for l in reversed(range(d)):
            c = spec.channels(l)
            ops.convT_fwd(cur, self.wf(f"up{l}"), self.p(f"up{l}.bias"), b.cat[l][..., c:])
            if train:
                self._conv_block_train(f"dec{l}.conv1", b.cat[l], b.dy1[l], b.da1[l])
                self._conv_block_train(f"dec{l}.conv2", b.da1[l], b.dy2[l], b.da2[l])
            else:
                self._conv_block_eval(f"dec{l}.conv1", b.cat[l], b.da1[l])
                self._conv_block_eval(f"dec{l}.conv2", b.da1[l], b.da2[l])
            cur = b.da2[l]
        b.target = target
        ops.head_fwd(cur, self.p("head.weight"), self.p("head.bias"), target, b.logits,
                     self.head_sums if target is not None else None)
        if target is not None:
            ops.head_loss(self.head_sums, n * h * w, spec.bce_weight, spec.dice_weight, spec.dice_eps,
                          self.loss_out)
        return b.logits

    # ------------------------------------------------------------------ backward
    def _on_side(self, after: Optional["torch.cuda.Event"], fn, tail: bool = False) -> Optional["torch.cuda.Event"]:
        """Run `fn` (weight-gradient launches) on the side stream once `after` has happened; returns an event
        that fires when they are done.  Without overlap: runs inline, returns None."""
        if not self.overlap_wgrad:
            fn()
            return None
        self._side.wait_event(after)
        with torch.cuda.stream(self._side):
            if tail and self._tail_event is None:
                self._tail_event = torch.cuda.Event()
                self._tail_event.record(self._side)
            fn()
            done = torch.cuda.Event()
            done.record(self._side)
        return done

    def _mark(self) -> Optional["torch.cuda.Event"]:
        if not self.overlap_wgrad:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def _bn_reduce_args(self, name: str, y):
        """The (y, scale, shift, mean, invstd, relu, sum_g, sum_gx) tuple a producer kernel needs to accumulate the
        BatchNorm-backward sums of layer `name` while it writes that layer's incoming gradient (None: not fused)."""
        if not self.use_bn or not self.fuse_bn_reduce:
            return None
        st = self._bn[name]
        return (y, st["scale"], st["shift"], st["mean"], st["invstd"], 1, st["sum_g"], st["sum_gx"])

    def _conv_block_bwd(self, name: str, x_in, y, a, g_a, g_y, g_in, acc=False, reduced=False):
        """Backward of conv -> [BN] -> ReLU.  g_a: gradient w.r.t. the activated output; g_y: scratch for
        the gradient w.r.t. the raw conv output; g_in: where the input gradient goes (None = not needed).
        reduced=True: the kernel that produced g_a has already accumulated this layer's sum_g / sum_gx."""
        ops = self.ops
        L = self.convs[name]
        busy = self._gy_busy.pop(g_y.data_ptr(), None)
        if busy is not None:  # an earlier layer's weight-gradient kernel may still be reading this scratch
            torch.cuda.current_stream(self.device).wait_event(busy)
        if self.use_bn:
            st = self._bn[name]
            if not reduced:
                ops.bn_bwd_reduce(g_a, y, st["scale"], st["shift"], st["mean"], st["invstd"], 1, st["sum_g"],
                                  st["sum_gx"])
            ops.bn_bwd_apply(g_a, y, st["scale"], st["shift"], st["mean"], st["invstd"], 1, st["sum_g"],
                             st["sum_gx"], g_y, self.g(f"{name}.bias"), self.g(f"{L.bn}.weight"),
                             self.g(f"{L.bn}.bias"), acc)
        else:
            ops.relu_bwd(g_a, a, g_y, self.g(f"{name}.bias"))
        ready = self._mark()
        # data gradient first (host launch order): it is the critical chain, the weight gradient fills in
        if g_in is not None:
            ops.conv3x3_dgrad(g_y, self.wd(name), g_in)
        # the flat gradient buffer was zeroed (or holds the previous micro-batches), so always accumulate
        done = self._on_side(ready, lambda: ops.conv3x3_wgrad(x_in, g_y, self.g(f"{name}.weight"), True),
                             tail=name.startswith("enc0."))
        if done is not None:
            self._gy_busy[g_y.data_ptr()] = done

    def backward(self, accumulate: bool = False, sync: bool = True, loss_scale: float = 1.0,
                 defer_tail: bool = False) -> None:
        """Gradient of the loss computed by the last training forward (with target) into ``self.grads``.
        Under data parallelism the loss is pre-scaled by 1/world and finished gradient buckets are
        all-reduced (sum) asynchronously while the rest of the backward pass runs.
        accumulate=True adds to the existing gradients (micro-batching); sync=False skips the all-reduce
        (all but the last micro-batch); loss_scale multiplies the loss (1/num_micro_batches).
        defer_tail=True (internal, train_step): return with the first encoder block's weight gradients
        possibly still running on the side stream; only ``optimizer_step*`` may follow, which updates
        every other parameter meanwhile (see _adam_launch)."""
        spec, d, ops = self.spec, self.spec.depth, self.ops
        b = self._buf
        if b is None or b.target is None:
            raise RuntimeError("backward() needs a preceding training forward(x, target)")
        self._close_tail()
        if not self.overlap_wgrad:
            self._backward_body(b, bool(accumulate), bool(sync), loss_scale)
            return
        # Fork: the critical chain (BN backward -> data gradient -> ...) runs on a high-priority stream and
        # the weight gradients on a lowest-priority one, so that a bandwidth kernel of the chain is placed
        # beside the resident weight-gradient CTAs ahead of that kernel's still-pending CTAs.
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=0)
            self._chain = torch.cuda.Stream(device=self.device, priority=-1)
        caller = torch.cuda.current_stream(self.device)
        self._chain.wait_stream(caller)
        with torch.cuda.stream(self._chain):
            self._backward_body(b, bool(accumulate), bool(sync), loss_scale)
        caller.wait_stream(self._chain)   # join: the optimizer / next forward see every gradient
        self._gy_busy = {}
        if defer_tail and self.world == 1 and self._tail_event is not None and self._defer_ok:
            caller.wait_event(self._tail_event)   # every weight gradient except the tail block's
            self._tail_open = True
        else:
            caller.wait_stream(self._side)

    def _backward_body(self, b: _Buffers, acc: bool, sync: bool, loss_scale: float) -> None:
        spec, d, ops = self.spec, self.spec.depth, self.ops
        if not acc:
            self.grads.zero_()
        if self.use_bn:
            self._bwd_region.zero_()
        self._pending = []
        self._next_bucket = 0
        self._sync = sync
        self._gy_busy = {}
        self._tail_event = None
        feat = b.da2[0]
        # the head's and the pools' backward kernels write the incoming gradient of a BatchNorm layer: they also
        # accumulate that layer's backward sums (one read of y instead of a separate pass over the gradient and y)
        # (measured on configs[1], one box: pools 0.42 -> 0.36 ms per step fused; the head kernel is compute-heavier
        # and LOSES 0.03 ms when it also carries the reduction, so it stays separate unless PLUME_FUSE_HEAD_BN=1)
        bn0 = self._bn_reduce_args("dec0.conv2", b.dy2[0]) if self.fuse_head_bn else None
        ops.head_bwd(feat, self.p("head.weight"), b.logits, b.target, self.head_sums, spec.bce_weight,
                     spec.dice_weight, spec.dice_eps, loss_scale / self.world, b.g_a[0], self.g("head.weight"),
                     self.g("head.bias"), **({"bn": bn0} if bn0 else {}))
        self._grads_ready("head.bias")
        for l in range(d):
            c = spec.channels(l)
            self._conv_block_bwd(f"dec{l}.conv2", b.da1[l], b.dy2[l], b.da2[l], b.g_a[l], b.g_y[l][0], b.g_a[l], acc,
                                 reduced=(l == 0 and bn0 is not None))
            self._conv_block_bwd(f"dec{l}.conv1", b.cat[l], b.dy1[l], b.da1[l], b.g_a[l], b.g_y[l][1], b.g_cat[l], acc)
            du = b.g_cat[l][..., c:]
            x_up = b.ba2 if l == d - 1 else b.da2[l + 1]
            g_up = b.bg_a if l == d - 1 else b.g_a[l + 1]
            ops.channel_sum(du, self.g(f"up{l}.bias"))
            ready = self._mark()
            ops.convT_dgrad(du, self.wd(f"up{l}"), g_up)
            self._on_side(ready, lambda: ops.convT_wgrad(x_up, du, self.g(f"up{l}.weight"), True))
            self._grads_ready(f"up{l}.bias")
        self._conv_block_bwd("bottleneck.conv2", b.ba1, b.by2, b.ba2, b.bg_a, b.bg_y[0], b.bg_a, acc)
        self._conv_block_bwd("bottleneck.conv1", b.pool[d - 1], b.by1, b.ba1, b.bg_a, b.bg_y[1], b.g_pool[d - 1], acc)
        self._grads_ready("bottleneck.conv1.bias")
        for l in reversed(range(d)):
            c = spec.channels(l)
            skip = b.cat[l][..., :c]
            bnl = self._bn_reduce_args(f"enc{l}.conv2", b.y2[l])
            ops.maxpool_bwd(b.g_pool[l], b.am[l], b.g_cat[l][..., :c], b.g_a[l], **({"bn": bnl} if bnl else {}))
            self._conv_block_bwd(f"enc{l}.conv2", b.a1[l], b.y2[l], skip, b.g_a[l], b.g_y[l][0], b.g_a[l], acc,
                                 reduced=bnl is not None)
            x_in = b.x_in if l == 0 else b.pool[l - 1]
            g_in = None if l == 0 else b.g_pool[l - 1]
            self._conv_block_bwd(f"enc{l}.conv1", x_in, b.y1[l], b.a1[l], b.g_a[l], b.g_y[l][1], g_in, acc)
            self._grads_ready(None if l == 0 else f"enc{l}.conv1.bias")

    # ------------------------------------------------------------------ data parallel buckets
    def _make_buckets(self, bucket_mb: float) -> List[tuple]:
        """Contiguous [begin, end) element ranges of the flat gradient buffer, cut at module boundaries
        once a bucket holds at least `bucket_mb` MB; ordered as the backward pass produces them."""
        limit = int(bucket_mb * (1 << 20) / 4)
        lay = self.layout
        # module boundaries in layout order
        ends, cur_mod = [], None
        for key in lay.order:
            mod = key.split(".")[0]
            if cur_mod is not None and mod != cur_mod:
                ends.append(lay.slots[key].offset)
            cur_mod = mod
        ends.append(lay.total)
        buckets, begin = [], 0
        for e in ends:
            if e - begin >= limit or e == lay.total:
                buckets.append((begin, e))
                begin = e
        # The last bucket is reduced after the backward pass has ended, i.e. fully exposed: keep it small (the
        # modules finished last -- the first encoder blocks -- hold little) by cutting it at the earliest module
        # boundary that leaves at most `tail_mb` for the end (default spec: 18.9 MB -> 14.2 MB early + 4.7 MB... the
        # final piece [enc1 | enc0] is 1.2 MB, latency bound).
        tail = int(2.0 * (1 << 20) / 4)
        b0, total = buckets[-1]
        cut = [e for e in ends if b0 < e < total and total - e <= tail]
        if cut:
            buckets[-1:] = [(b0, cut[0]), (cut[0], total)]
        return buckets

    def _grads_ready(self, last_key: Optional[str]) -> None:
        """Called by backward() after the gradients up to and including the module holding `last_key`
        (None = everything) are complete; launches the all-reduce of every bucket that is now finished."""
        if self.world == 1 or not self._sync:
            return
        import torch.distributed as dist

        if last_key is None:
            upto = self.layout.total
        else:
            mod = last_key.split(".")[0]
            upto = 0
            for key in self.layout.order:
                if key.split(".")[0] == mod:
                    s = self.layout.slots[key]
                    upto = max(upto, s.offset + s.numel)
        while self._next_bucket < len(self._buckets):
            a, e = self._buckets[self._next_bucket]
            if e > upto and not (last_key is None):
                break
            def reduce_bucket():
                if not self.grad_comm_bf16:
                    return dist.all_reduce(self.grads[a:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True), (a, e), False
                if self._comm_buf is None:
                    self._comm_buf = torch.empty(self.layout.total, dtype=torch.bfloat16, device=self.device)
                self.ops.cast_f32_bf16(self.grads[a:e], self._comm_buf[a:e])
                return (dist.all_reduce(self._comm_buf[a:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True),
                        (a, e), True)

            if self.overlap_wgrad:
                # the bucket's weight gradients are on the side stream, the rest on this one
                self._side.wait_event(self._mark())
                with torch.cuda.stream(self._side):
                    wk = reduce_bucket()
            else:
                wk = reduce_bucket()
            self._pending.append(wk)
            self._next_bucket += 1

    def _wait_bucket(self, wk, rng, narrow) -> None:
        wk.wait()                   # the current stream waits for this bucket's all-reduce (no host block under NCCL)
        if narrow:                  # bf16 wire format: back into the fp32 gradient buffer
            self.ops.cast_bf16_f32(self._comm_buf[rng[0]:rng[1]], self.grads[rng[0]:rng[1]])

    def wait_grads(self) -> None:
        for wk, rng, narrow in self._pending:
            self._wait_bucket(wk, rng, narrow)
        self._pending = []

    # ------------------------------------------------------------------ optimizer / step
    def _close_tail(self) -> None:
        """Join the side stream if backward(defer_tail=True) left it running (anything but the optimizer
        follows): the tail weight gradients read activations and write ``grads``."""
        if self._tail_open:
            self._tail_open = False
            torch.cuda.current_stream(self.device).wait_stream(self._side)

    def _adam_launch(self, launch) -> None:
        """launch(lo, hi) enqueues Adam on the flat range [lo, hi).  After backward(defer_tail=True) the first
        encoder block's weight gradients (the last ones the backward pass produces, at the end of the flat
        buffer) may still be running: everything before them is updated now, overlapping those kernels."""
        total = self.layout.total
        if self._pending and self.adam_per_bucket:
            # data parallel: the buckets' all-reduces finish front to back; each bucket is updated as soon as ITS
            # all-reduce is done, under the all-reduces still in flight (the wide model's 2 GB of gradients take
            # longer on the wire than Adam needs for them).  The last bucket's all-reduce was queued behind the
            # tail weight gradients on the side stream, so waiting for it also joins that stream.
            done = 0
            for wk, rng, narrow in self._pending:
                self._wait_bucket(wk, rng, narrow)
                assert rng[0] == done
                launch(rng[0], rng[1])
                done = rng[1]
            self._pending = []
            if self._tail_open:
                self._tail_open = False
                torch.cuda.current_stream(self.device).wait_stream(self._side)
            if done < total:
                launch(done, total)
            return
        self.wait_grads()
        if self._tail_open:
            self._tail_open = False
            launch(0, self._tail_offset)
            torch.cuda.current_stream(self.device).wait_stream(self._side)
            launch(self._tail_offset, total)
        else:
            launch(0, total)

    def optimizer_step(self) -> None:
        spec = self.spec
        self.step_count += 1
        self._adam_launch(lambda lo, hi: self.ops.adam(
            self.params[lo:hi], self.grads[lo:hi], self.adam_m[lo:hi], self.adam_v[lo:hi], spec.lr,
            spec.betas[0], spec.betas[1], spec.adam_eps, self.step_count))
        self._param_version += 1

    def adam_coefficients(self, step: int) -> torch.Tensor:
        """The 8 floats ``plume_adam_dev`` reads (host tensor), formed in double like torch.optim.Adam."""
        spec = self.spec
        b1, b2 = float(spec.betas[0]), float(spec.betas[1])
        bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
        return torch.tensor([spec.lr / bc1, b1, b2, 1.0 - b1, 1.0 - b2, spec.adam_eps, 1.0 / math.sqrt(bc2), 1.0],
                            dtype=torch.float32)

    def optimizer_step_dev(self, coef: torch.Tensor) -> None:
        """Adam with coefficients already in device memory (CUDA-graph capturable); the caller advances
        ``step_count`` and refreshes `coef` before every replay."""
        self._adam_launch(lambda lo, hi: self.ops.adam_dev(
            self.params[lo:hi], self.grads[lo:hi], self.adam_m[lo:hi], self.adam_v[lo:hi], coef))
        self._param_version += 1

    def train_step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One optimisation step; returns the device tensor [loss, bce, dice] (no host sync)."""
        self.train(True)
        self.forward(x, target)
        self.backward(defer_tail=True)
        self.optimizer_step()
        return self.loss_out

    @torch.no_grad()
    def predict_logits(self, x: torch.Tensor) -> torch.Tensor:
        was = self.training
        self.eval()
        out = self.forward(x)
        self.train(was)
        return out

    def predict_mask(self, x: torch.Tensor) -> torch.Tensor:
        """uint8 [N,H,W]: sigmoid(logit) >= mask_threshold  (== logit >= logit(threshold))."""
        thr = self.spec.mask_threshold
        return (self.predict_logits(x) >= math.log(thr / (1 - thr))).to(torch.uint8)
