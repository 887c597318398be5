#!/usr/bin/env python
"""Benchmark of the UNet training hot path (BASELINE.json: "UNet train tiles/sec (256^2 px)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--configs 2,3,4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one optimisation step (forward, backward, gradient
all-reduce, Adam) of the default UNetSpec on this rank's 32 synthetic 256x256x8 bf16 tiles
(BASELINE.json configs[1]); with N GPUs every rank keeps 32 tiles (weak scaling) and the value is the
whole-job tiles/s.  `value` is timed with inputs resident in HBM; `e2e` runs the same step through the
public Trainer API from pinned host buffers with the host->device copies and the loss read-back inside
the timed region.  `roofline` is measured live with CUDA events around every launch of the dominant
kernel family (the tcgen05 implicit-GEMM kernels: conv fwd + dgrad + convT) during eagerly launched steps
of the same workload; `per_layer` breaks the same measurement down by layer and pass, `roofline_hbm` by
bandwidth kernel.  `other_configs` carries the remaining BASELINE.json configs measured at the same N:
configs[2] (global batch 256 of 512^2 tiles, micro-batched), configs[3] (tiled inference over 4096^2
scenes sharded round-robin) and configs[4] (wide UNet, 1024^2 tiles).
`--impl reference` times the CPU oracle (oracle/unet_ref.py -- the reference repository itself has no
model code) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "unet_train_tiles_per_sec_256px"
UNIT = "tiles/s"
TILE, BATCH = 256, 32


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def headline_config(world: int, graphed: bool = True) -> dict:
    """The `config` object of the JSON line; the reference arm prints the same one (it times a bounded sample
    of this workload on the host cores)."""
    return {"workload": "BASELINE.json configs[1]: default UNetSpec (in 8, base 64, depth 4, BatchNorm) "
                        "training step, 32 tiles of 256x256 per GPU, bf16 activations / fp32 accumulate, Adam",
            "tile": TILE, "per_gpu_batch": BATCH, "global_batch": BATCH * world, "parallelism": f"dp{world}",
            "params": 31046401,
            "launch": "CUDA graph replay (Trainer.step_graphed)" if graphed else "eager launches",
            "l2_policy": "inputs larger than L2: one step streams several GB of activations (8 GB workspace) "
                         "through a 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:  # noqa: BLE001  (no nvidia-smi: report nulls)
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (ts, ln) in self.lines
                  if self.t0 is None or (self.t0 + 0.05 <= ts <= (self.t1 or ts) + 0.05)]
        if not inside:  # region shorter than the sampling period: fall back to every sample taken
            inside = [ln for (_, ln) in self.lines]
        for ln in inside:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ----------------------------------------------------------------------------------------------------
def time_cpu_oracle(steps: int, warmup: int, tiles_per_step: int, budget_s: float):
    """fp32 PyTorch-CPU UNet (the parity oracle) doing full training steps on `tiles_per_step` synthetic
    256^2 tiles; stops early if the time budget is exhausted.  Returns (tiles/s, s/step, steps done)."""
    import torch

    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec
    from oracle.unet_ref import UNetRef, make_optimizer, plume_loss

    torch.set_num_threads(os.cpu_count() or 1)
    spec = UNetSpec()
    torch.manual_seed(0)
    model = UNetRef(spec).train()
    opt = make_optimizer(model, spec)
    x, t = synthetic_batch(tiles_per_step, TILE, TILE, spec.in_channels, seed=1234, dtype=torch.float32)
    x = x.permute(0, 3, 1, 2).contiguous()

    def one():
        opt.zero_grad(set_to_none=True)
        loss = plume_loss(model(x)[:, 0], t, spec)
        loss.backward()
        opt.step()
        return float(loss.detach())

    t_start = time.perf_counter()
    warm_done = 0
    for _ in range(warmup):
        one()
        warm_done += 1
        if time.perf_counter() - t_start > budget_s / 2:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    return tiles_per_step / sec, sec, len(times), torch.get_num_threads(), warm_done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: 2 tiles of the 32-tile batch per step keeps K+W steps within a few minutes
    tiles = 2
    v, sec, done, threads, warm = time_cpu_oracle(max(args.steps, 1), args.warmup, tiles, budget_s=240.0)
    sample = (f"{tiles} of the {BATCH} tiles of a step per timed step, {done} timed full training steps after "
              f"{warm} warm-up steps (fwd+bwd+Adam, fp32, PyTorch CPU oracle on {threads} host threads; the "
              f"reference repo has no model code)")
    cb = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(max(args.gpus, 1)),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# Per-launch timing of the operator layer
# ----------------------------------------------------------------------------------------------------
def conv_variant(op: str, cin: int, cout: int, h: int, w: int) -> str:
    """Name of the kernel instantiation csrc/igemm.cu picks for a layer (mirrors try_launch_conv3 / wgrad_config)."""
    if op == "wgrad":
        if h >= 8 and w >= 8 and (cin == 64 or cin % 128 == 0):
            bn = 64 if cin == 64 else (128 if cout % 128 == 0 else 64)
            return f"igemm_wgrad3_kernel<{bn}>" + (" 9 taps/CTA" if cin == 64 else "")
        bn = 256 if cout % 256 == 0 else (128 if cout % 128 == 0 else 64)
        return f"igemm_wgrad_kernel<{bn}>"
    n = cout  # GEMM N of the launch (dgrad: the layer's input channels, passed as cout here)
    bn = 256 if n % 256 == 0 else (128 if n % 128 == 0 else 64)
    if h >= 16 and w >= 8:
        # CTA pairs (cta_group::2) whenever the M tiles pair up -- always for the even batches measured here; a CTA
        # of a pair holds half of every weight tile, which decides whether the slice is resident
        kb = cin // 64
        resident = n == bn and bn <= 128 and \
            (4 * 18432 + 9 * kb * (bn // 2) * 128 + 2 * 16384 + 32 * bn + 8 * 40 + 16 + 1024) <= 232448
        mode = 0 if resident else (1 if bn <= 128 else 2)
        return f"igemm_conv3_kernel<{bn},{mode}> CTA pair"
    return f"igemm_fwd_kernel<{bn}>"


class OpTimer:
    """Wraps the launches of an operator layer (CudaOps, or the CPU RefOps in tests) and attributes algorithmic
    work to each: FLOPs (logical channels) for the tensor-core GEMMs, bytes for the bandwidth kernels.  With
    `events` every call is bracketed by a CUDA event pair recorded on the launching stream."""

    FWD = ("conv3x3_fwd", "conv3x3_dgrad", "convT_fwd", "convT_dgrad")
    WGRAD = ("conv3x3_wgrad", "convT_wgrad")
    STREAM = ("scale_shift_act", "scale_shift_act_pool", "maxpool_bwd", "bn_bwd_reduce", "bn_bwd_apply",
              "channel_sum", "head_fwd", "head_bwd", "adam", "adam_dev", "pad_channels", "pack_batch")
    BN_FAMILY = ("scale_shift_act", "bn_bwd_reduce", "bn_bwd_apply")

    def __init__(self, model, events: bool = True):
        self.model, self.ops, self.events = model, model.ops, events
        self.records = []  # (op name, layer, work, e0, e1, meta)
        self.enabled = False
        if events:
            import torch
            self.torch = torch
        # weight / gradient buffer pointers -> layer names
        self.by_ptr = {}
        for name in list(model.convs) + list(model.ups):
            self.by_ptr[model.wf(name).data_ptr()] = name
            self.by_ptr[model.wd(name).data_ptr()] = name
            self.by_ptr[model.g(f"{name}.weight").data_ptr()] = name
        for name in self.FWD + self.WGRAD + self.STREAM:
            if hasattr(self.ops, name):
                setattr(self.ops, name, self._wrap(name, getattr(self.ops, name)))

    # -- algorithmic work ---------------------------------------------------------------------------
    def _gemm(self, name, a):
        """(layer, pass, FLOPs, Cin, Cout, H, W) with LOGICAL input channels: the zero padding of the first
        layer (keyed on the layer, not on its padded width) earns nothing."""
        m = self.model
        if name == "conv3x3_fwd":
            x, y, layer = a[0], a[5], self.by_ptr.get(a[1].data_ptr(), "?")
            cin = m.convs[layer].cin if layer in m.convs else x.shape[-1]
            return layer, "fwd", 2.0 * 9 * cin * y.shape[-1] * y.shape[0] * y.shape[1] * y.shape[2], cin, y.shape[-1], \
                y.shape[1], y.shape[2]
        if name == "conv3x3_dgrad":
            dy, dx, layer = a[0], a[2], self.by_ptr.get(a[1].data_ptr(), "?")
            return layer, "dgrad", 2.0 * 9 * dy.shape[-1] * dx.shape[-1] * dx.shape[0] * dx.shape[1] * dx.shape[2], \
                dy.shape[-1], dx.shape[-1], dx.shape[1], dx.shape[2]
        if name == "conv3x3_wgrad":
            x, dy, layer = a[0], a[1], self.by_ptr.get(a[2].data_ptr(), "?")
            cin = m.convs[layer].cin if layer in m.convs else x.shape[-1]
            return layer, "wgrad", 2.0 * 9 * cin * dy.shape[-1] * x.shape[0] * x.shape[1] * x.shape[2], cin, \
                dy.shape[-1], x.shape[1], x.shape[2]
        if name == "convT_fwd":
            x, u, layer = a[0], a[3], self.by_ptr.get(a[1].data_ptr(), "?")
            return layer, "fwd", 2.0 * 4 * x.shape[-1] * u.shape[-1] * x.shape[0] * x.shape[1] * x.shape[2], \
                x.shape[-1], u.shape[-1], x.shape[1], x.shape[2]
        if name == "convT_wgrad":
            x, u, layer = a[0], a[1], self.by_ptr.get(a[2].data_ptr(), "?")
            return layer, "wgrad", 2.0 * 4 * x.shape[-1] * u.shape[-1] * x.shape[0] * x.shape[1] * x.shape[2], \
                x.shape[-1], u.shape[-1], x.shape[1], x.shape[2]
        if name == "convT_dgrad":
            du, dx, layer = a[0], a[2], self.by_ptr.get(a[1].data_ptr(), "?")
            return layer, "dgrad", 2.0 * 4 * du.shape[-1] * dx.shape[-1] * dx.shape[0] * dx.shape[1] * dx.shape[2], \
                du.shape[-1], dx.shape[-1], dx.shape[1], dx.shape[2]
        raise KeyError(name)

    @staticmethod
    def _bytes(name, a):
        """Algorithmic HBM bytes of a bandwidth kernel (bf16 activations = 2 B)."""
        n0 = a[0].numel() if hasattr(a[0], "numel") else 0
        if name == "scale_shift_act":
            return 4.0 * n0                                   # read y, write a
        if name == "scale_shift_act_pool":
            return (2 + 2 + 0.5 + 0.25) * n0                  # read y; write skip, pooled, argmax
        if name == "maxpool_bwd":
            return (0.5 + 0.25 + 2 + 2) * a[3].numel()        # per full-resolution element: dy, argmax, dskip, dx
        if name == "bn_bwd_reduce":
            return 4.0 * n0                                   # read da, y
        if name == "bn_bwd_apply":
            return 6.0 * n0                                   # read da, y; write dy
        if name == "channel_sum":
            return 2.0 * n0
        if name == "head_fwd":
            pixels, c = n0 // a[0].shape[-1], a[0].shape[-1]
            return pixels * (2.0 * c + 1 + 4)                 # feat, target, logits
        if name == "head_bwd":
            pixels, c = n0 // a[0].shape[-1], a[0].shape[-1]
            return pixels * (2.0 * c + 4 + 1 + 2.0 * c)       # feat, logits, target, dfeat
        if name in ("adam", "adam_dev"):
            return 28.0 * n0                                  # p rw, g r, m rw, v rw (fp32)
        if name == "pad_channels":
            pixels = n0 // a[0].shape[-1]
            return pixels * 2.0 * (a[0].shape[-1] + a[1].shape[-1])
        if name == "pack_batch":
            return float(sum(w.numel() * (4 + (2 if wf is not None else 0) + (2 if wd is not None else 0))
                             for _, w, wf, wd in a[0]))
        raise KeyError(name)

    def _wrap(self, name, fn):
        def timed(*a, **kw):
            if not self.enabled:
                return fn(*a, **kw)
            if name in self.STREAM:
                layer, meta, work = None, None, self._bytes(name, a)
            else:
                layer, pas, work, cin, cout, h, w = self._gemm(name, a)
                meta = (pas, cin, cout, h, w)
            e0 = e1 = None
            if self.events:
                e0 = self.torch.cuda.Event(enable_timing=True)
                e1 = self.torch.cuda.Event(enable_timing=True)
                e0.record()
            r = fn(*a, **kw)
            if self.events:
                e1.record()
            self.records.append((name, layer, work, e0, e1, meta))
            return r

        return timed

    # -- summaries -----------------------------------------------------------------------------------
    def _ms(self, r):
        return r[3].elapsed_time(r[4]) if self.events else 0.0

    def total_flops(self) -> float:
        return sum(r[2] for r in self.records if r[0] in self.FWD + self.WGRAD)

    def family(self, names):
        rec = [r for r in self.records if r[0] in names]
        ms = sum(self._ms(r) for r in rec)
        work = sum(r[2] for r in rec)
        return {"launches": len(rec), "ms": ms, "work": work, "rate": (work / (ms * 1e-3)) if ms > 0 else 0.0}

    def per_layer(self, steps: int, peak_sust: float, peak_burst: float):
        """One row per (layer, pass): kernel variant, average us per launch, TFLOP/s, fractions of the sustained
        and burst bf16 peaks; plus the FLOP-weighted means."""
        rows = {}
        for r in self.records:
            if r[0] not in self.FWD + self.WGRAD:
                continue
            key = (r[1], r[5][0])
            d = rows.setdefault(key, {"flops": 0.0, "ms": 0.0, "n": 0, "meta": r[5], "op": r[0]})
            d["flops"] += r[2]
            d["ms"] += self._ms(r)
            d["n"] += 1
        out, fl_sum, w_sust, w_burst = [], 0.0, 0.0, 0.0
        for (layer, pas), d in rows.items():
            _, cin, cout, h, w = d["meta"]
            tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
            # fwd / dgrad rows carry the GEMM's own (K channels -> N channels); the first layer runs zero padded to 64
            variant = conv_variant("wgrad" if pas == "wgrad" else "fwd", max(cin, 64), cout, h, w) \
                if d["op"].startswith("conv3x3") else \
                ("igemm_wgrad_kernel (convT)" if pas == "wgrad" else "igemm_fwd_kernel (convT)")
            out.append({"layer": layer, "pass": pas, "kernel": variant, "gemm": f"{cin}->{cout} @{h}x{w}",
                        "us": d["ms"] * 1e3 / max(d["n"], 1), "gflop": d["flops"] / max(d["n"], 1) / 1e9,
                        "tflops": tf, "frac_sustained": tf / peak_sust if peak_sust else None,
                        "frac_burst": tf / peak_burst if peak_burst else None})
            fl_sum += d["flops"]
            w_sust += d["flops"] * (tf / peak_sust if peak_sust else 0.0)
            w_burst += d["flops"] * (tf / peak_burst if peak_burst else 0.0)
        summary = {"flop_weighted_frac_sustained": w_sust / fl_sum if fl_sum else None,
                   "flop_weighted_frac_burst": w_burst / fl_sum if fl_sum else None,
                   "gemm_flops_per_step": fl_sum / max(steps, 1)}
        return out, summary

    def hbm_by_kernel(self, steps: int, peak_gbs: float):
        out = {}
        for name in self.STREAM:
            f = self.family((name,))
            if not f["launches"]:
                continue
            gbs = f["rate"] / 1e9
            out[name] = {"achieved": gbs, "unit": "GB/s", "frac": gbs / peak_gbs if peak_gbs else None,
                         "launches_per_step": f["launches"] / max(steps, 1), "ms_per_step": f["ms"] / max(steps, 1),
                         "algorithmic_mb_per_step": f["work"] / max(steps, 1) / 1e6}
        return out


class StdoutToStderr:
    """Everything other than the final JSON line goes to stderr, including what native libraries print on
    file descriptor 1 (NCCL prints its version banner there)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_gpu(args):
    with StdoutToStderr():
        rc, line = _run_gpu(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return rc


class Ranks:
    """Barrier / max-over-ranks helpers (no-ops on one GPU)."""

    def __init__(self, pg, dev):
        self.pg, self.dev = pg, dev

    def barrier(self):
        import torch
        if self.pg is not None:
            import torch.distributed as dist
            dist.barrier(group=self.pg)
        torch.cuda.synchronize()

    def max(self, v: float) -> float:
        if self.pg is None:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.pg)
        return float(t.item())


def timed_steps(ranks: Ranks, fn, warmup: int, steps: int) -> float:
    """ms per step of fn(i): CUDA events on the current stream, barrier + synchronize on both sides, max over ranks."""
    import torch
    for i in range(warmup):
        fn(i)
    ranks.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    ranks.barrier()
    return ranks.max(e0.elapsed_time(e1)) / steps


# ----------------------------------------------------------------------------------------------------
# The other BASELINE.json configs (extra keys of the same JSON line)
# ----------------------------------------------------------------------------------------------------
def run_config2(ranks, rank, world, pg, dev, peak_tf):
    """configs[2]: data-parallel training, GLOBAL batch 256 of 512x512 tiles (strong scaling: 256/N per GPU),
    processed in micro-batches of 32 with gradient accumulation (BatchNorm sees one micro-batch at a time)."""
    import torch
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec, train_flops_per_tile
    from kcl_ltss_bioatm_b200.trainer import Trainer

    spec = UNetSpec()
    per_gpu = 256 // world
    micro = max(1, per_gpu // 32)
    tr = Trainer(spec, device=dev, process_group=pg, seed=0, micro_batches=micro)
    # 8 distinct tiles per rank tiled up to the shard: generating 256 tiles of 512^2 on the CPU would take minutes
    xs, ts = synthetic_batch(8, 512, 512, spec.in_channels, seed=300 + rank)
    x, t = xs.repeat(per_gpu // 8, 1, 1, 1).to(dev), ts.repeat(per_gpu // 8, 1, 1).to(dev)
    ms = timed_steps(ranks, lambda i: tr.step(x, t), 1, 3)
    loss = float(tr.model.loss_out[0].item())
    tf = train_flops_per_tile(spec, 512, 512) * per_gpu / (ms * 1e-3) / 1e12
    out = {"workload": "BASELINE.json configs[2]: global batch 256 of 512x512x8 tiles, data parallel, micro-batches "
                       "of 32 with gradient accumulation, bucketed all-reduce on the last micro-batch",
           "n_gpus": world, "scaling": "strong", "global_batch": 256, "per_gpu_batch": per_gpu, "micro_batches": micro,
           "tile": 512, "ms_per_step": ms, "tiles_per_s": 256 / (ms * 1e-3), "tflops_per_gpu": tf,
           "frac_of_sustained_peak": tf / peak_tf if peak_tf else None, "steps": 3, "warmup": 1,
           "launch": "eager launches (Trainer.step)", "final_loss": loss,
           "workspace_gb": tr.model.activation_bytes() / 1e9}
    del tr, x, t
    torch.cuda.empty_cache()
    return out


def run_config3(ranks, rank, world, dev, peak_tf):
    """configs[3]: tiled inference over 4096x4096 scenes (256^2 tiles, stride 224, overlap-stitched, thresholded),
    scenes dealt round-robin to the GPUs, no collective."""
    import torch
    from kcl_ltss_bioatm_b200.data import synthetic_scene
    from kcl_ltss_bioatm_b200.predict import ScenePredictor, shard_round_robin
    from kcl_ltss_bioatm_b200.spec import UNetSpec, fwd_flops_per_tile
    from kcl_ltss_bioatm_b200.unet import UNetB200

    spec = UNetSpec()
    model = UNetB200(spec, device=dev, seed=0)
    pred = ScenePredictor(model, tile=256, margin=16, batch_tiles=64)
    scenes_total = 2 * world
    mine = shard_round_robin(scenes_total, rank, world)
    host = synthetic_scene(4096, 4096, spec.in_channels, seed=40 + rank).pin_memory()
    scene = host.to(dev)
    tiles = pred.num_tiles(4096, 4096)
    masks = []

    def resident(i):
        masks.append(pred.predict_scene(scene))

    ms = timed_steps(ranks, resident, 1, len(mine))          # ms per scene per GPU
    plume = float(masks[-1].float().mean().item())
    masks.clear()
    # end to end: scene from pinned host memory, mask back to the host, every scene
    mask_host = torch.empty(4096, 4096, dtype=torch.uint8).pin_memory()

    def e2e(i):
        dscene = host.to(dev, non_blocking=True)
        mask_host.copy_(pred.predict_scene(dscene), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e(0)
    ranks.barrier()
    t0 = time.perf_counter()
    for i in range(len(mine)):
        e2e(i)
    ranks.barrier()
    e2e_ms = ranks.max((time.perf_counter() - t0) * 1e3) / len(mine)
    tf = fwd_flops_per_tile(spec, 256, 256)["total"] * tiles / (ms * 1e-3) / 1e12
    out = {"workload": "BASELINE.json configs[3]: tiled inference over 4096x4096x8 scenes, 256^2 tiles at stride 224 "
                       "(361 tiles/scene), overlap-stitched + thresholded, scenes sharded round-robin, no collective",
           "n_gpus": world, "scaling": "weak", "scenes": scenes_total, "scenes_per_gpu": len(mine),
           "tiles_per_scene": tiles, "ms_per_scene": ms, "tiles_per_s": world * tiles / (ms * 1e-3),
           "scenes_per_s": world / (ms * 1e-3), "tflops_per_gpu": tf,
           "frac_of_sustained_peak": tf / peak_tf if peak_tf else None,
           "e2e": {"ms_per_scene": e2e_ms, "tiles_per_s": world * tiles / (e2e_ms * 1e-3),
                   "h2d_bytes_per_scene": host.numel() * host.element_size(), "d2h_bytes_per_scene": mask_host.numel()},
           "plume_fraction": plume}
    del model, pred, scene, host
    torch.cuda.empty_cache()
    return out


def run_config4(ranks, rank, world, pg, dev, peak_tf):
    """configs[4]: wide UNet (2x base filters, depth 5) training on 1024x1024 tiles, one tile per GPU per step."""
    import torch
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec, num_parameters, train_flops_per_tile
    from kcl_ltss_bioatm_b200.trainer import Trainer

    spec = UNetSpec.wide()
    tr = Trainer(spec, device=dev, process_group=pg, seed=0)
    x, t = synthetic_batch(1, 1024, 1024, spec.in_channels, seed=500 + rank)
    x, t = x.to(dev), t.to(dev)
    graphed = pg is None or tr.graph_dp
    step = tr.step_graphed if graphed else tr.step
    ms = timed_steps(ranks, lambda i: step(x, t), 2, 5)
    loss = float(tr.model.loss_out[0].item())
    tf = train_flops_per_tile(spec, 1024, 1024) / (ms * 1e-3) / 1e12
    out = {"workload": "BASELINE.json configs[4]: wide UNet (base 128, depth 5, 497.5 M parameters) training step, "
                       "1 tile of 1024x1024x8 per GPU, bf16 activations / fp32 accumulate, Adam, bucketed fp32 "
                       "gradient all-reduce (1.99 GB)",
           "n_gpus": world, "scaling": "weak", "per_gpu_batch": 1, "global_batch": world, "tile": 1024,
           "ms_per_step": ms, "tiles_per_s": world / (ms * 1e-3), "tflops_per_gpu": tf,
           "frac_of_sustained_peak": tf / peak_tf if peak_tf else None, "steps": 5, "warmup": 2,
           "launch": "CUDA graph replay (Trainer.step_graphed)" if graphed else "eager launches",
           "params": num_parameters(spec), "final_loss": loss,
           "workspace_gb": tr.model.activation_bytes() / 1e9}
    tr.release_graphs()
    del tr, x, t
    torch.cuda.empty_cache()
    return out


def run_modes(ranks, rank, world, pg, dev):
    """The headline step (configs[1]) in the two optional modes, at this N: deterministic reductions
    (PLUME_DETERMINISTIC) and the bf16x3 high-precision arithmetic (north_star's "tf32 mode" bar)."""
    import torch

    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec
    from kcl_ltss_bioatm_b200.trainer import Trainer

    out = {}
    for name, spec, det, n in (("deterministic", UNetSpec(), True, BATCH), ("bf16x3", UNetSpec(precision="bf16x3"), False, 8)):
        tr = Trainer(spec, device=dev, process_group=pg, seed=0)
        tr.model.ops.set_deterministic(det)
        try:
            x, t = synthetic_batch(n, TILE, TILE, spec.in_channels, seed=900 + rank)
            x, t = x.to(dev), t.to(dev)
            graphed = pg is None or tr.graph_dp
            step = tr.step_graphed if graphed else tr.step
            ms = timed_steps(ranks, lambda i: step(x, t), 3, 6)
            out[name] = {"workload": f"configs[1] network, {n} tiles of {TILE}x{TILE} per GPU, one training step",
                         "n_gpus": world, "per_gpu_batch": n, "ms_per_step": ms, "tiles_per_s": world * n / (ms * 1e-3),
                         "steps": 6, "warmup": 3, "final_loss": float(tr.model.loss_out[0].item())}
            tr.release_graphs()
        finally:
            tr.model.ops.set_deterministic(False)
        del tr, x, t
        torch.cuda.empty_cache()
    out["deterministic"]["note"] = ("ordered partial-sum reductions instead of fp atomics, one MMA issuer in the 3x3 "
                                    "weight gradient: bit-identical runs (tests/test_gpu_determinism.py)")
    out["bf16x3"]["note"] = ("activations / operands as bf16 hi+lo pairs, three tcgen05 passes per GEMM through the generic "
                             "kernels: logits within 2e-5 of the fp32 oracle (tests/test_gpu_precise.py)")
    return out


def run_label_generation(dev, peak_bw):
    """the label-generation kernels next to the UNet path, on one GPU (no collective: timestamps are
    independent).  One MAIAC-sized timestamp (1200 x 1200 float64 AOD, 64 fire clusters): nearest-valid fill
    (interpolate_aod_nearest), then the reference's three threshold sweeps (75 thresholds: masks, components, per-fire
    plume extents) in one call on bit planes; CUDA events, inputs resident, 10 repetitions after 2."""
    import numpy as np
    import torch

    from kcl_ltss_bioatm_b200 import sweep
    from tests.sweep_data import synthetic_aod, synthetic_null_aod

    h = w = 1200
    sw = sweep.ThresholdSweep(dev)
    aod = torch.from_numpy(synthetic_aod(h, w, 5)[0].astype(np.float64)).to(dev)
    nul = torch.from_numpy(synthetic_null_aod(h, w, 9)).to(dev)
    rng = np.random.default_rng(3)
    rc = torch.tensor(np.stack([rng.integers(16, h - 16, 64), rng.integers(16, w - 16, 64)], 1), dtype=torch.int32).to(dev)
    thr = torch.tensor(np.concatenate([np.abs(np.arange(0, tmax, step) - tmax)
                                       for step, tmax in [(0.02, 0.5), (0.03, 0.75), (0.04, 1)]])).to(dev)
    ws = torch.empty(sw.ops.sweep_workspace_bytes(h, w, thr.numel()), dtype=torch.uint8, device=dev)
    ext = torch.empty(thr.numel(), 64, dtype=torch.int32, device=dev)
    fws = torch.empty(sw.ops.fill_nearest_workspace_bytes(h, w), dtype=torch.uint8, device=dev)
    filled = torch.empty_like(nul)

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_sweep = timed(lambda: sw.ops.sweep_extents(aod, thr, rc, sweep.P_ID_WIN_SIZE, ws, ext))
    ms_fill = timed(lambda: sw.ops.fill_nearest(nul, sweep.NULL_VALUE, fws, filled))
    segs = (w + 31) // 32
    sweep_bytes = h * w * 8 + thr.numel() * h * segs * 4 * 4          # image once; bit plane written, read by merge / flatten / extents
    fill_bytes = h * w * (8 + 8 + 4 + 4)
    return {"workload": f"{h}x{w} float64 AOD timestamp, 64 fire clusters: nearest-valid fill + 75-threshold sweep (one call)",
            "n_gpus": 1, "sweep_ms": ms_sweep, "fill_ms": ms_fill, "timestamps_per_s": 1e3 / (ms_sweep + ms_fill),
            "sweep_bit_plane_gbps": sweep_bytes / (ms_sweep * 1e-3) / 1e9, "fill_gbps": fill_bytes / (ms_fill * 1e-3) / 1e9,
            "hbm_peak_gbps": peak_bw or None, "extents_checksum": int(ext.sum().item()),
            "null_fraction": float((nul == sweep.NULL_VALUE).float().mean().item()),
            "note": "latency / atomic bound union-find, not streaming: the bit planes are 32 x smaller than label planes "
                    "(profiles/r2_bench_sweep.json holds the comparison with the dense-plane path and the C oracle)"}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def _run_gpu(args):
    import torch

    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec, fwd_flops_per_tile, train_flops_per_tile
    from kcl_ltss_bioatm_b200.trainer import Trainer, init_distributed

    if not torch.cuda.is_available():
        return 2, {"metric": METRIC, "error": "no CUDA device: the B200 path has no CPU fallback"}
    rank, world, local, pg = init_distributed("cuda")
    dev = torch.device("cuda", local)
    ranks = Ranks(pg, dev)
    spec = UNetSpec()
    trainer = Trainer(spec, device=dev, process_group=pg, seed=0)
    model, ops = trainer.model, trainer.model.ops
    timer = OpTimer(model)
    peaks, peak_kind = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    peak_burst = float(peaks.get("bf16_tflops", peak_tf))
    peak_bw = float(peaks.get("hbm_gbs", 0) or 0)

    # distinct synthetic shard per rank; host copies live in pinned memory for the end-to-end leg
    nbuf = 2
    host = [synthetic_batch(BATCH, TILE, TILE, spec.in_channels, seed=1234 + 97 * rank + i) for i in range(nbuf)]
    host = [(x.pin_memory(), t.pin_memory()) for x, t in host]
    dev_batches = [(x.to(dev), t.to(dev)) for x, t in host]

    # The step is replayed from a captured CUDA graph (Trainer.step_graphed, one graph launch per step); under
    # data parallelism the bucketed NCCL all-reduces are captured with it (PLUME_GRAPH_DP=0: eager steps).
    graphed = pg is None or trainer.graph_dp
    step_fn = trainer.step_graphed if graphed else trainer.step

    # ---------------- device-resident leg
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step_fn(*dev_batches[i % nbuf])
    ranks.barrier()
    sampler.mark_begin()
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_fn(*dev_batches[i % nbuf])
    e1.record()
    ranks.barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ranks.max(e0.elapsed_time(e1))
    launches = ops.launches - launches0
    loss_now = float(model.loss_out[0].item())

    # ---------------- end-to-end leg: pinned host buffers -> H2D -> step -> loss read-back, every step.
    # The user-facing input path is DevicePrefetcher: each batch is copied once, inside the timed region,
    # on a copy stream while the previous step computes.  Runs directly after the device-resident leg: the
    # same loop measured later in a long run reads 2-3 % slower for BOTH legs (the part heats up and the power
    # cap lowers the clock; scripts/diag_e2e.py: resident 11.49 -> 11.79 ms/step from 20 to 100 steps, the
    # prefetch + loss-read loop 11.70), which is drift, not a cost of the input path.
    from kcl_ltss_bioatm_b200.data import DevicePrefetcher

    h2d = host[0][0].numel() * host[0][0].element_size() + host[0][1].numel()
    d2h = 3 * 4
    from kcl_ltss_bioatm_b200.trainer import LossLog

    def e2e_run(n):
        # every step: H2D of its batch (copy stream, two batches ahead) and a D2H read of its [loss, bce, dice];
        # the loss of step i is read on the host while step i+1 runs (LossLog), the last one before the clock stops
        ring, seen = LossLog(), 0
        for x, t in DevicePrefetcher((host[i % nbuf] for i in range(n)), dev, depth=2):
            out = step_fn(x, t)  # graphed: copies into its static buffers, then one graph launch
            seen += ring.push(out) is not None
        seen += ring.flush() is not None
        assert seen == n, "every step's loss must have been read on the host"

    e2e_run(3)
    ranks.barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    ranks.barrier()
    e2e_s = ranks.max(time.perf_counter() - t0)

    # ---------------- roofline leg: the same steps run eagerly with a CUDA-event pair around every launch of
    # ours (events cannot be recorded inside a graph replay).  The product overlaps the weight-gradient kernels
    # with the BatchNorm-backward kernels on a second stream; a kernel's own duration is only defined when it
    # has the GPU to itself, so this leg launches everything on one stream.
    roof_steps = min(args.steps, 20)
    overlap_was, model.overlap_wgrad = model.overlap_wgrad, False
    for i in range(2):
        trainer.step(*dev_batches[i % nbuf])
    ranks.barrier()
    timer.enabled = True
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(roof_steps):
        trainer.step(*dev_batches[i % nbuf])
    r1.record()
    ranks.barrier()
    timer.enabled = False
    model.overlap_wgrad = overlap_was
    eager_ms = r0.elapsed_time(r1) / roof_steps

    trainer.release_graphs()
    workspace_gb = model.activation_bytes() / 1e9
    del dev_batches
    model._buf = None   # the headline legs are done: hand their activation workspace back before the other configs
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE.json configs at the same N (each failure is reported, not fatal)
    others = {}
    wanted = {c.strip() for c in args.configs.split(",") if c.strip()}
    for key, fn in (("configs[2]", lambda: run_config2(ranks, rank, world, pg, dev, peak_tf)),
                    ("configs[3]", lambda: run_config3(ranks, rank, world, dev, peak_tf)),
                    ("configs[4]", lambda: run_config4(ranks, rank, world, pg, dev, peak_tf)),
                    ("modes", lambda: run_modes(ranks, rank, world, pg, dev)),
                    ("label_generation", lambda: run_label_generation(dev, peak_bw))):
        if (key[8] if key.startswith("configs") else key[0]) not in wanted:
            continue
        if key == "label_generation" and rank != 0:
            continue
        try:
            t0 = time.perf_counter()
            others[key] = fn()
            others[key]["bench_wall_s"] = time.perf_counter() - t0
        except Exception as e:  # noqa: BLE001
            others[key] = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()

    if pg is not None:
        import torch.distributed as dist
        dist.barrier(group=pg)
        dist.destroy_process_group()
    if rank != 0:
        return 0, None

    tiles = world * BATCH * args.steps
    value = tiles / (ms_total * 1e-3)
    e2e_value = tiles / e2e_s
    fk = timer.family(OpTimer.FWD)
    fk_tf = fk["rate"] / 1e12
    # DRAM traffic of the same kernel family from the committed ncu launch list (per launch, like `achieved`)
    traffic, traffic_src = None, None
    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f).get("fwd_kernel", {})
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
            break
    roof = {"bound": "tensor", "kernel": "igemm_conv3_kernel + igemm_fwd_kernel (conv3x3 fwd + dgrad, convT fwd + dgrad)",
            "achieved": fk_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": fk_tf / peak_tf if peak_tf else None, "traffic": traffic,
            "traffic_unit": "bytes of DRAM read+write per launch (ncu)", "traffic_source": traffic_src,
            "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)",
            "frac_of_burst_peak": fk_tf / peak_burst if peak_burst else None,
            "launches_timed": fk["launches"], "ms_per_step": fk["ms"] / max(roof_steps, 1),
            "algorithmic_tflop_per_step": fk["work"] / max(roof_steps, 1) / 1e12,
            "measured_in": f"{roof_steps} eagerly launched single-stream steps of the same workload "
                           f"({eager_ms:.2f} ms/step; the timed product step overlaps wgrad with BN backward)"}
    wk = timer.family(OpTimer.WGRAD)
    wk_tf = wk["rate"] / 1e12
    roof_w = {"bound": "tensor", "kernel": "igemm_wgrad3_kernel / igemm_wgrad_kernel (split-K, fp32 atomics into dW)",
              "achieved": wk_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": wk_tf / peak_tf if peak_tf else None,
              "launches_timed": wk["launches"], "ms_per_step": wk["ms"] / max(roof_steps, 1),
              "algorithmic_tflop_per_step": wk["work"] / max(roof_steps, 1) / 1e12}
    sk = timer.family(OpTimer.BN_FAMILY)
    gbs = sk["rate"] / 1e9
    roof_s = {"bound": "hbm", "kernel": "scale_shift_act + bn_bwd_reduce + bn_bwd_apply (BatchNorm apply / backward)",
              "achieved": gbs, "peak": peak_bw or None, "unit": "GB/s", "frac": gbs / peak_bw if peak_bw else None,
              "launches_timed": sk["launches"], "ms_per_step": sk["ms"] / max(roof_steps, 1),
              "by_kernel": timer.hbm_by_kernel(roof_steps, peak_bw)}
    per_layer, pl_summary = timer.per_layer(roof_steps, peak_tf, peak_burst)
    # the timer's FLOPs must be the algorithmic FLOPs of a step: everything but the 1x1 head (a bandwidth kernel)
    f = fwd_flops_per_tile(spec, TILE, TILE)
    expect = (train_flops_per_tile(spec, TILE, TILE) - 3.0 * f["head"]) * BATCH
    pl_summary["expected_gemm_flops_per_step"] = expect
    pl_summary["flops_accounted"] = pl_summary["gemm_flops_per_step"] / expect if expect else None
    step_tf = train_flops_per_tile(spec, TILE, TILE) * BATCH / (ms_total / args.steps * 1e-3) / 1e12
    if args.per_layer_out:
        with open(args.per_layer_out, "w") as fo:
            json.dump({"summary": pl_summary, "rows": per_layer, "roofline_hbm": roof_s,
                       "peaks": {"bf16_tflops_sustained": peak_tf, "bf16_tflops": peak_burst, "hbm_gbs": peak_bw}},
                      fo, indent=1)

    # ---------------- CPU baseline: config 1 on the host cores (bounded to ~20 s)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, done, threads, _ = time_cpu_oracle(steps=8, warmup=1, tiles_per_step=1, budget_s=25.0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"BASELINE.json configs[0]: 1 tile of 256x256x8 per step, fp32, {done} timed full "
                         f"training steps of the PyTorch-CPU oracle ({sec:.3f} s/step)"}

    cfg = headline_config(world, graphed)
    cfg["workspace_gb"] = workspace_gb
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "path": "pinned host batches -> DevicePrefetcher (H2D on a copy stream, two batches ahead) -> "
                        "Trainer.step_graphed -> LossLog (every step's [loss, bce, dice] copied D2H and read on "
                        "the host one step late, the last one before the clock stops)"},
        "gpu_launches": launches,
        "roofline": roof, "roofline_wgrad": roof_w, "roofline_hbm": roof_s,
        "per_layer_summary": pl_summary, "per_layer": per_layer,
        "step_tflops": step_tf, "step_frac_of_peak": step_tf / peak_tf if peak_tf else None,
        "other_configs": others,
        "clocks": clocks, "cpu_baseline": cpu, "final_loss": loss_now,
    }
    return 0, line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="2,3,4,m,l",
                    help="which of the other BASELINE.json configs (indices 2, 3, 4; m = the deterministic and bf16x3 modes; "
                         "l = the label-generation kernels of SURVEY 8(f)) to measure as extra keys; empty = none")
    ap.add_argument("--per-layer-out", default=None, help="also write the per-layer table to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup),
               "--configs", args.configs]
        return subprocess.call(cmd)
    if args.warmup < 3:
        args.warmup = 3
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
