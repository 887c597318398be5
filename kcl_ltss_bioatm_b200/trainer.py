"""Training driver: data-parallel setup, micro-batching, checkpoint / resume, logging.

One process per GPU (``torch.distributed``, NCCL over NVLink; gloo on CPU for tests).  Every rank holds a
full replica and a distinct shard of the tile batch; BatchNorm statistics stay per GPU (the
DistributedDataParallel default), the loss is scaled by 1/world and gradient buckets are summed with
asynchronous all-reduces launched from inside the backward pass (unet.UNetB200.backward).

Checkpoint layout (defined here; the reference only reserves the directory,
``src/config/filepaths.py:33`` ``path_to_model_folder``):
    <dir>/<name>.pt       torch.save(state_dict) with the oracle's keys, NCHW fp32 -- loadable by
                          ``UNetRef.load_state_dict`` and by ``UNetB200.load_state_dict``
    <dir>/<name>.opt.pt   {"spec": UNetSpec dict, "step": int, "adam_m": flat fp32, "adam_v": flat fp32}
Log format follows the reference's scripts (``src/features/plume_identifier_rg.py:23-25``).
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Optional, Tuple

import torch

from .spec import UNetSpec
from .unet import UNetB200

LOG_FMT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"


def init_distributed(device_type: str = "cuda"):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun).  Returns (rank, world, local_rank, process_group)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        if device_type == "cuda":
            torch.cuda.set_device(local)
        return rank, world, local, None
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if device_type == "cuda":
        torch.cuda.set_device(local)
        # PLUME_NCCL_HIGH_PRIORITY=1: NCCL's stream outranks the compute streams, so an all-reduce kernel gets its
        # CTAs resident at the next kernel boundary instead of waiting behind the backward pass's pending CTAs
        opts = None
        if os.environ.get("PLUME_NCCL_HIGH_PRIORITY", "0") == "1":
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local),
                                pg_options=opts)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local, dist.group.WORLD


class LossLog:
    """Reads every step's [loss, bce, dice] on the host without idling the GPU.  ``push(out)`` enqueues the
    device->host copy of this step's result into a pinned ring slot (behind the step, on the current stream) and
    returns the PREVIOUS step's values, waiting only for that older copy's event; the host can therefore launch
    step i+1 before step i has finished.  ``flush()`` returns the last step's values.  Every step is read exactly
    once, one step late."""

    def __init__(self, depth: int = 2):
        self._slots = [(torch.empty(3, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(max(2, depth))]
        self._n = 0
        self.bytes_per_step = 12

    def _read(self, i: int):
        buf, ev = self._slots[i % len(self._slots)]
        ev.synchronize()
        return buf.tolist()

    def push(self, out: torch.Tensor):
        prev = self._read(self._n - 1) if self._n else None
        buf, ev = self._slots[self._n % len(self._slots)]
        buf.copy_(out, non_blocking=True)
        ev.record()
        self._n += 1
        return prev

    def flush(self):
        return self._read(self._n - 1) if self._n else None


class Trainer:
    def __init__(self, spec: UNetSpec = UNetSpec(), device="cuda", process_group=None, ops=None,
                 seed: int = 0, micro_batches: int = 1, bucket_mb: float = 25.0):
        self.spec = spec
        self.model = UNetB200(spec, ops=ops, device=device, seed=seed, process_group=process_group,
                              bucket_mb=bucket_mb)
        self.micro_batches = micro_batches
        self.pg = process_group
        # Data-parallel steps are captured too (the bucketed NCCL all-reduces live inside the step's CUDA graph):
        # verified on 2 and 8 x B200 (8 GPUs: 12.04 vs 12.53 ms/step eager).  PLUME_GRAPH_DP=0 keeps them eager.
        # Two requirements found the hard way: capture in "thread_local" error mode (the NCCL watchdog thread
        # touches the CUDA API) and release_graphs() before destroy_process_group() -- tearing the communicator
        # down while a graph still holds its kernels hangs.
        self.graph_dp = os.environ.get("PLUME_GRAPH_DP", "1") != "0"
        self.log = logging.getLogger("train_model")
        if process_group is not None:
            import torch.distributed as dist

            # replicas must start identical even if a caller seeded ranks differently
            dist.broadcast(self.model.params, src=0, group=process_group)
            self.model._param_version += 1

    # ------------------------------------------------------------------ one optimisation step
    def step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """x [n,h,w,c] / target [n,h,w] are this rank's shard, already on the device.  With
        micro_batches = k the shard is processed in k slices with gradient accumulation (BatchNorm and the
        Dice term then see one slice at a time).  Returns the device tensor [loss, bce, dice] of the last
        slice; no host synchronisation happens here."""
        m, k = self.model, self.micro_batches
        if k == 1:
            return m.train_step(x, target)
        n = x.shape[0]
        if n % k:
            raise ValueError("batch must divide into micro_batches")
        mb = n // k
        m.train(True)
        for i in range(k):
            m.forward(x[i * mb:(i + 1) * mb], target[i * mb:(i + 1) * mb])
            m.backward(accumulate=i > 0, sync=(i == k - 1), loss_scale=1.0 / k)
        m.optimizer_step()
        return m.loss_out

    # ------------------------------------------------------------------ CUDA-graph step
    def _snapshot(self):
        """Everything an optimisation step mutates, so that the un-captured warm-up step can be undone."""
        m = self.model
        return (m.params.clone(), m.adam_m.clone(), m.adam_v.clone(), m._stat_region.clone(), m.step_count,
                m.num_batches_tracked)

    def _restore(self, snap) -> None:
        m = self.model
        m.params.copy_(snap[0])
        m.adam_m.copy_(snap[1])
        m.adam_v.copy_(snap[2])
        m._stat_region.copy_(snap[3])
        m.step_count, m.num_batches_tracked = snap[4], snap[5]
        m._param_version += 1

    def step_graphed(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Same optimisation step replayed from a captured CUDA graph (forward, backward, the bucketed gradient
        all-reduces under data parallelism, Adam: ~170 kernel launches become one graph launch).  The batch is
        copied into static device buffers, the Adam bias-correction coefficients into an 8-float device buffer,
        then the graph is replayed.  Captured on first use per (N, H, W); every call, the capturing one
        included, performs exactly ONE optimisation step: capture is preceded by one un-captured warm-up step
        (one-time work such as kernel attributes and buffer allocation must not be captured) whose effect on
        parameters, Adam moments, running statistics and counters is rolled back from a snapshot.  Each captured
        shape keeps its own activation buffers alive (the graph holds raw pointers into them)."""
        m = self.model
        if self.micro_batches != 1 or m.device.type != "cuda" or (self.pg is not None and not self.graph_dp):
            return self.step(x, target)
        key = tuple(x.shape)
        g = getattr(self, "_graphs", None)
        if g is None:
            g = self._graphs = {}
        if key not in g:
            xs, ts = torch.empty_like(x), torch.empty_like(target)
            coef = torch.zeros(8, dtype=torch.float32, device=x.device)
            # ring of pinned staging slots for the coefficients: a slot is rewritten only after the copy that
            # read it has executed (its event), so the host may run many steps ahead of the GPU
            coef_host = [(torch.zeros(8, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            xs.copy_(x)
            ts.copy_(target)
            m._buf = None  # a fresh activation set for this shape: other shapes' graphs keep theirs
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                m.train(True)
                m.forward(xs, ts)
                m.backward(defer_tail=True)
                coef.copy_(m.adam_coefficients(m.step_count + 1).to(x.device))
                m.optimizer_step_dev(coef)
                self._restore(snap)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            m.train(True)
            launches0 = m.ops.launches
            # thread_local: the NCCL watchdog thread may touch the CUDA API while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.pg is not None else "global"):
                m.forward(xs, ts)
                m.backward(defer_tail=True)
                m.optimizer_step_dev(coef)
            m.num_batches_tracked = snap[5]  # capture runs the Python bookkeeping once without executing kernels
            self.graph_launches = m.ops.launches - launches0  # kernels of ours inside one replay
            m.ops.launches = launches0
            g[key] = (graph, xs, ts, coef, coef_host, m._buf)
        graph, xs, ts, coef, coef_host, bufs = g[key]
        m._buf = bufs  # the activations this graph writes (grad_dict / a following eager backward read them)
        xs.copy_(x, non_blocking=True)
        ts.copy_(target, non_blocking=True)
        m.step_count += 1
        if m.use_bn:
            m.num_batches_tracked += 1
        slot, ev = coef_host[m.step_count % len(coef_host)]
        ev.synchronize()
        slot.copy_(m.adam_coefficients(m.step_count))
        coef.copy_(slot, non_blocking=True)
        ev.record()
        graph.replay()
        m.ops.launches += self.graph_launches
        # the replay updated the parameters and the running statistics on the device: the packed bf16 weights
        # and the folded eval coefficients cached on the host side are stale
        m._param_version += 1
        m._stats_version += 1
        return m.loss_out

    def release_graphs(self) -> None:
        """Drop the captured CUDA graphs.  Must be called before ``destroy_process_group()`` when data-parallel
        steps were captured: destroying the NCCL communicator while a graph still holds its kernels hangs."""
        g = getattr(self, "_graphs", None)
        if g:
            torch.cuda.synchronize()
            g.clear()

    def fit(self, steps: int, batch_fn: Callable[[int], Tuple[torch.Tensor, torch.Tensor]],
            log_every: int = 10, on_step: Optional[Callable[[int, float], None]] = None, graphed: bool = False):
        """batch_fn(i) returns this rank's shard of batch i: device tensors, or (pinned) host tensors, which are
        then copied `depth` batches ahead on a copy stream (DevicePrefetcher).  graphed=True replays the step from
        a CUDA graph (step_graphed).  On a GPU every step's loss is read through a LossLog, one step late, so the
        host never waits for the step it has just launched; `on_step(i, loss)` sees every step."""
        from .data import DevicePrefetcher

        losses = []
        first = batch_fn(0)

        def batches():
            yield first
            for it in range(1, steps):
                yield batch_fn(it)

        stream = batches()
        on_gpu = self.model.device.type == "cuda"
        if first[0].device.type == "cpu" and on_gpu:
            stream = DevicePrefetcher(stream, self.model.device)
        step_fn = self.step_graphed if graphed else self.step
        ring = LossLog() if on_gpu else None

        def report(it, v):
            if log_every and (it % log_every == 0 or it == steps - 1):
                losses.append((it, v[0]))
                self.log.info("step %d loss %.5f (bce %.5f dice %.5f)", it, v[0], v[1], v[2])
            if on_step:
                on_step(it, v[0])

        for it, (x, t) in enumerate(stream):
            out = step_fn(x, t)
            if ring is None:
                report(it, out.detach().float().cpu().tolist())
            else:
                prev = ring.push(out)
                if prev is not None:
                    report(it - 1, prev)
        if ring is not None and steps > 0:
            report(steps - 1, ring.flush())
        return losses

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, directory: str, name: str = "unet_plume") -> str:
        os.makedirs(directory, exist_ok=True)
        path = os.path.join(directory, f"{name}.pt")
        torch.save(self.model.state_dict(), path)
        opt = self.model.optimizer_state()
        torch.save({"spec": self.spec.to_dict(), "step": opt["step"], "adam_m": opt["m"], "adam_v": opt["v"]},
                   os.path.join(directory, f"{name}.opt.pt"))
        return path

    def load_checkpoint(self, directory: str, name: str = "unet_plume", with_optimizer: bool = True) -> None:
        sd = torch.load(os.path.join(directory, f"{name}.pt"), map_location="cpu")
        self.model.load_state_dict(sd)
        opt_path = os.path.join(directory, f"{name}.opt.pt")
        if with_optimizer and os.path.exists(opt_path):
            st = torch.load(opt_path, map_location="cpu")
            if UNetSpec.from_dict(st["spec"]) != self.spec:
                raise ValueError("checkpoint was written for a different UNetSpec")
            self.model.load_optimizer_state({"step": st["step"], "m": st["adam_m"], "v": st["adam_v"]})
