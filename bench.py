#!/usr/bin/env python
"""Benchmark of the UNet training hot path (BASELINE.json: "UNet train tiles/sec (256^2 px)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  A "step" is one optimisation step (forward, backward, gradient
all-reduce, Adam) of the default UNetSpec on this rank's 32 synthetic 256x256x8 bf16 tiles
(BASELINE.json configs[1]); with N GPUs every rank keeps 32 tiles (weak scaling) and the value is the
whole-job tiles/s.  `value` is timed with inputs resident in HBM; `e2e` runs the same step through the
public Trainer API from pinned host buffers with the host->device copies and the loss read-back inside
the timed region.  `roofline` is measured live with CUDA events around every launch of the dominant
kernel (the tcgen05 implicit-GEMM forward kernel: conv fwd + dgrad + convT) during the timed steps.
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
    for _ in range(warmup):
        one()
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
    return tiles_per_step / sec, sec, len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: 2 tiles of the 32-tile batch per step keeps K+W steps within a few minutes
    tiles = 2
    v, sec, done, threads = time_cpu_oracle(args.steps, min(args.warmup, 2), tiles, budget_s=240.0)
    sample = (f"{tiles} of the {BATCH} tiles per step, {done} timed full training steps "
              f"(fwd+bwd+Adam, fp32, PyTorch CPU oracle; the reference repo has no model code)")
    cb = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "default UNetSpec training step, 256x256x8 tiles (BASELINE.json configs[1] "
                               "sampled at 2 tiles/step on host cores)", "tile": TILE, "tiles_per_step": tiles,
                   "host_threads": threads},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class GemmTimer:
    """Wraps the tensor-core launches of CudaOps with CUDA event pairs (recorded on the launching stream)
    and attributes algorithmic FLOPs to them."""

    FWD_KERNEL = ("conv3x3_fwd", "conv3x3_dgrad", "convT_fwd", "convT_dgrad")
    WGRAD = ("conv3x3_wgrad", "convT_wgrad")
    # HBM-bound BatchNorm kernels: algorithmic bytes per bf16 element (reads + writes), DESIGN.md section 3.4
    STREAM = {"scale_shift_act": 4.0, "bn_bwd_reduce": 4.0, "bn_bwd_apply": 6.0}

    def __init__(self, ops, logical_cin0: int, padded_cin0: int):
        import torch

        self.torch, self.ops = torch, ops
        self.records = []  # (family, flops, ev0, ev1)
        self.enabled = False
        self.cin0, self.cin0_pad = logical_cin0, padded_cin0
        for name in self.FWD_KERNEL + self.WGRAD + tuple(self.STREAM):
            setattr(ops, name, self._wrap(name, getattr(ops, name)))

    def _flops(self, name, a):
        # algorithmic FLOPs with LOGICAL channels (the zero padding of the first layer earns nothing)
        def lc(c):
            return self.cin0 if c == self.cin0_pad and self.cin0 != self.cin0_pad else c
        if name == "conv3x3_fwd":
            x, y = a[0], a[5]
            return 2.0 * 9 * lc(x.shape[-1]) * y.shape[-1] * y.shape[0] * y.shape[1] * y.shape[2]
        if name == "conv3x3_dgrad":
            dy, dx = a[0], a[2]
            return 2.0 * 9 * dy.shape[-1] * dx.shape[-1] * dx.shape[0] * dx.shape[1] * dx.shape[2]
        if name == "conv3x3_wgrad":
            x, dy = a[0], a[1]
            return 2.0 * 9 * lc(x.shape[-1]) * dy.shape[-1] * x.shape[0] * x.shape[1] * x.shape[2]
        if name in ("convT_fwd", "convT_wgrad"):
            x, u = a[0], (a[3] if name == "convT_fwd" else a[1])
            return 2.0 * 4 * x.shape[-1] * u.shape[-1] * x.shape[0] * x.shape[1] * x.shape[2]
        if name == "convT_dgrad":
            du, dx = a[0], a[2]
            return 2.0 * 4 * du.shape[-1] * dx.shape[-1] * dx.shape[0] * dx.shape[1] * dx.shape[2]
        raise KeyError(name)

    def _wrap(self, name, fn):
        fam = "fwd_kernel" if name in self.FWD_KERNEL else ("wgrad_kernel" if name in self.WGRAD else "stream_kernel")

        def timed(*a, **kw):
            if not self.enabled:
                return fn(*a, **kw)
            e0 = self.torch.cuda.Event(enable_timing=True)
            e1 = self.torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **kw)
            e1.record()
            work = self.STREAM[name] * a[0].numel() if name in self.STREAM else self._flops(name, a)
            self.records.append((fam, work, e0, e1))
            return r

        return timed

    def summary(self):
        out = {}
        for fam in ("fwd_kernel", "wgrad_kernel", "stream_kernel"):
            rec = [r for r in self.records if r[0] == fam]
            ms = sum(r[2].elapsed_time(r[3]) for r in rec)
            fl = sum(r[1] for r in rec)
            out[fam] = {"launches": len(rec), "ms": ms, "flops": fl,
                        "tflops": (fl / (ms * 1e-3) / 1e12) if ms > 0 else 0.0}
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


def _run_gpu(args):
    import torch

    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec, train_flops_per_tile
    from kcl_ltss_bioatm_b200.trainer import Trainer, init_distributed

    if not torch.cuda.is_available():
        return 2, {"metric": METRIC, "error": "no CUDA device: the B200 path has no CPU fallback"}
    rank, world, local, pg = init_distributed("cuda")
    dev = torch.device("cuda", local)
    spec = UNetSpec()
    trainer = Trainer(spec, device=dev, process_group=pg, seed=0)
    model, ops = trainer.model, trainer.model.ops
    timer = GemmTimer(ops, spec.in_channels, spec.cin_padded)

    # distinct synthetic shard per rank; host copies live in pinned memory for the end-to-end leg
    nbuf = 2
    host = [synthetic_batch(BATCH, TILE, TILE, spec.in_channels, seed=1234 + 97 * rank + i) for i in range(nbuf)]
    host = [(x.pin_memory(), t.pin_memory()) for x, t in host]
    dev_batches = [(x.to(dev), t.to(dev)) for x, t in host]

    def barrier():
        if pg is not None:
            import torch.distributed as dist
            dist.barrier(group=pg)
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if pg is None:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
        return float(t.item())

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
    barrier()
    sampler.mark_begin()
    launches0 = ops.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_fn(*dev_batches[i % nbuf])
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launches - launches0
    loss_now = float(model.loss_out[0].item())

    # ---------------- roofline leg: the same steps run eagerly with a CUDA-event pair around every
    # tensor-core launch (events cannot be recorded inside a graph replay).  The product overlaps the
    # weight-gradient kernels with the BatchNorm-backward kernels on a second stream; a kernel's own duration
    # is only defined when it has the GPU to itself, so this leg launches everything on one stream.
    roof_steps = min(args.steps, 20)
    overlap_was, model.overlap_wgrad = model.overlap_wgrad, False
    for i in range(2):
        trainer.step(*dev_batches[i % nbuf])
    barrier()
    timer.enabled = True
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(roof_steps):
        trainer.step(*dev_batches[i % nbuf])
    r1.record()
    barrier()
    timer.enabled = False
    model.overlap_wgrad = overlap_was
    eager_ms = r0.elapsed_time(r1) / roof_steps
    gemm = timer.summary()

    # ---------------- end-to-end leg: pinned host buffers -> H2D -> step -> loss read-back, every step.
    # The user-facing input path is DevicePrefetcher: each batch is copied once, inside the timed region,
    # on a copy stream while the previous step computes.
    from kcl_ltss_bioatm_b200.data import DevicePrefetcher

    h2d = host[0][0].numel() * host[0][0].element_size() + host[0][1].numel()
    d2h = 3 * 4
    loss_host = torch.empty(3, dtype=torch.float32).pin_memory()

    def e2e_run(n):
        for x, t in DevicePrefetcher((host[i % nbuf] for i in range(n)), dev, depth=2):
            out = step_fn(x, t)  # graphed: copies into its static buffers, then one graph launch
            loss_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the user reads the loss every step

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    if pg is not None:
        import torch.distributed as dist
        dist.barrier(group=pg)
        trainer.release_graphs()   # captured NCCL kernels must be gone before the communicator is
        dist.destroy_process_group()
    if rank != 0:
        return 0, None

    peaks, peak_kind = load_peaks()
    tiles = world * BATCH * args.steps
    value = tiles / (ms_total * 1e-3)
    e2e_value = tiles / e2e_s
    fk = gemm["fwd_kernel"]
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    # DRAM traffic of the same kernel family from the committed ncu launch list (per launch, like `achieved`)
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get("fwd_kernel", {})
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    roof = {"bound": "tensor", "kernel": "igemm_conv3_kernel + igemm_fwd_kernel (conv3x3 fwd + dgrad, convT fwd + dgrad)",
            "achieved": fk["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
            "frac": fk["tflops"] / peak_tf if peak_tf else None, "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu)",
            "traffic_source": traffic_src,
            "peak_source": f"{peak_kind} bf16_tflops_sustained (kernel timed inside a long step)",
            "launches_timed": fk["launches"], "ms_per_step": fk["ms"] / max(roof_steps, 1),
            "measured_in": f"{roof_steps} eagerly launched single-stream steps of the same workload "
                           f"({eager_ms:.2f} ms/step; the timed product step overlaps wgrad with BN backward)"}
    wk = gemm["wgrad_kernel"]
    roof_w = {"bound": "tensor", "kernel": "igemm_wgrad3_kernel / igemm_wgrad_kernel (split-K, fp32 atomics into dW)", "achieved": wk["tflops"],
              "peak": peak_tf, "unit": "TFLOP/s", "frac": wk["tflops"] / peak_tf if peak_tf else None,
              "launches_timed": wk["launches"], "ms_per_step": wk["ms"] / max(roof_steps, 1)}
    sk = gemm["stream_kernel"]   # for this family "flops" holds algorithmic bytes
    peak_bw = float(peaks.get("hbm_gbs", 0) or 0)
    gbs = sk["flops"] / (sk["ms"] * 1e-3) / 1e9 if sk["ms"] > 0 else 0.0
    roof_s = {"bound": "hbm", "kernel": "scale_shift_act + bn_bwd_reduce + bn_bwd_apply (BatchNorm apply / backward)",
              "achieved": gbs, "peak": peak_bw or None, "unit": "GB/s", "frac": gbs / peak_bw if peak_bw else None,
              "launches_timed": sk["launches"], "ms_per_step": sk["ms"] / max(roof_steps, 1)}
    step_tf = train_flops_per_tile(spec, TILE, TILE) * BATCH / (ms_total / args.steps * 1e-3) / 1e12

    # ---------------- CPU baseline: config 1 on the host cores (bounded to ~20 s)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sec, done, threads = time_cpu_oracle(steps=8, warmup=1, tiles_per_step=1, budget_s=25.0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"BASELINE.json configs[0]: 1 tile of 256x256x8 per step, fp32, {done} timed full "
                         f"training steps of the PyTorch-CPU oracle ({sec:.3f} s/step)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[1]: default UNetSpec (in 8, base 64, depth 4, BatchNorm) "
                               "training step, 32 tiles of 256x256 per GPU, bf16 activations / fp32 accumulate, "
                               "Adam", "tile": TILE, "per_gpu_batch": BATCH, "global_batch": BATCH * world,
                   "parallelism": f"dp{world}", "params": 31046401,
                   "launch": "CUDA graph replay (Trainer.step_graphed)" if graphed else "eager launches",
                   "l2_policy": "inputs larger than L2: one step streams several GB of activations "
                                f"({model.activation_bytes() / 1e9:.1f} GB workspace) through a 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": launches,
        "roofline": roof, "roofline_wgrad": roof_w, "roofline_hbm": roof_s,
        "step_tflops": step_tf, "step_frac_of_peak": step_tf / peak_tf if peak_tf else None,
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    if args.warmup < 3:
        args.warmup = 3
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
