"""GPU, world_size 2 over NCCL (skipped on a one-GPU box): the data-parallel path with the CUDA kernels --
weight gradients on the side stream, bucketed all-reduce issued behind them, 1/world loss scaling.  Two
ranks with half the batch each must reproduce one GPU that sees both halves (BatchNorm-free, Dice-free
spec: the only legitimate difference is the order of the fp32 split-K reductions)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _spec():
    from kcl_ltss_bioatm_b200.spec import UNetSpec

    return UNetSpec(base_filters=64, depth=2, norm="none", dice_weight=0.0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.trainer import Trainer

    spec = _spec()
    tr = Trainer(spec, device=f"cuda:{rank}", process_group=dist.group.WORLD, seed=rank, bucket_mb=0.5)
    assert len(tr.model._buckets) >= 3
    x, t = synthetic_batch(4, 64, 64, spec.in_channels, seed=11)
    half = slice(rank * 2, rank * 2 + 2)
    m = tr.model
    m.train(True)
    m.forward(x[half].cuda(), t[half].cuda())
    m.backward()
    m.wait_grads()
    torch.cuda.synchronize()
    torch.save({"grads": m.grads.cpu(), "params0": m.params.cpu()}, os.path.join(out_dir, f"r{rank}.pt"))
    m.optimizer_step()
    torch.cuda.synchronize()
    torch.save(m.params.cpu(), os.path.join(out_dir, f"p{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_nccl_matches_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29900 + os.getpid() % 90
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["params0"], r1["params0"])       # broadcast from rank 0
    assert torch.equal(r0["grads"], r1["grads"])           # all-reduced gradients identical on both ranks
    assert torch.equal(torch.load(tmp_path / "p0.pt"), torch.load(tmp_path / "p1.pt"))

    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.unet import UNetB200

    spec = _spec()
    single = UNetB200(spec, device="cuda:0", seed=None)
    single.params.copy_(r0["params0"].cuda())
    single._param_version += 1
    x, t = synthetic_batch(4, 64, 64, spec.in_channels, seed=11)
    single.train(True)
    single.forward(x.cuda(), t.cuda())
    single.backward()
    torch.cuda.synchronize()
    g = single.grads.cpu()
    err = (r0["grads"] - g).abs().max() / g.abs().max()
    assert err < 1e-4, float(err)


def _worker_graphed(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec
    from kcl_ltss_bioatm_b200.trainer import Trainer

    spec = UNetSpec(base_filters=64, depth=2)          # with BatchNorm (per-GPU statistics) and Dice
    batches = [synthetic_batch(4, 64, 64, spec.in_channels, seed=40 + 7 * rank + i, device=f"cuda:{rank}")
               for i in range(3)]
    eager = Trainer(spec, device=f"cuda:{rank}", process_group=dist.group.WORLD, seed=1, bucket_mb=0.5)
    graphed = Trainer(spec, device=f"cuda:{rank}", process_group=dist.group.WORLD, seed=1, bucket_mb=0.5)
    assert graphed.graph_dp and len(graphed.model._buckets) >= 3
    le, lg = [], []
    for i in range(6):
        x, t = batches[i % 3]
        le.append(float(eager.step(x, t)[0].item()))
        lg.append(float(graphed.step_graphed(x, t)[0].item()))
    torch.cuda.synchronize()
    torch.save({"le": le, "lg": lg, "pe": eager.model.params.cpu(), "pg": graphed.model.params.cpu(),
                "steps": (eager.model.step_count, graphed.model.step_count)}, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    graphed.release_graphs()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_graphed_steps_match_eager_steps(tmp_path):
    """Data-parallel steps replayed from a CUDA graph (bucketed NCCL all-reduces captured inside) against the same
    steps launched eagerly: same losses step by step, same parameters on both ranks, parameters of the two modes
    equal up to what the unordered fp32 reductions allow."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29800 + os.getpid() % 90
    mp.spawn(_worker_graphed, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert r0["steps"] == r1["steps"] == (6, 6)
    assert torch.equal(r0["pg"], r1["pg"]) and torch.equal(r0["pe"], r1["pe"])     # replicas stay identical
    for r in (r0, r1):
        assert abs(r["le"][0] - r["lg"][0]) <= 1e-4 * abs(r["le"][0])              # step 1: identical weights
        for a, b in zip(r["le"], r["lg"]):
            assert abs(a - b) <= 1e-2 * abs(a), (r["le"], r["lg"])
    d = (r0["pe"] - r0["pg"]).abs()
    print(f"graphed vs eager after 6 steps: median |dp| {d.median().item():.3e}, rel L2 {(d.norm() / r0['pe'].norm()).item():.3e}")
    # two runs of unordered fp32 reductions compared with each other after six Adam steps (lr 1e-3): measured over five
    # runs on 2 GPUs median 7.3e-5 ... 1.3e-4, rel L2 ~1e-2 (gpurun_out/r2ad, r2ad2) -- the bars sit above that spread
    assert d.median().item() <= 3e-4 and (d.norm() / r0["pe"].norm()).item() <= 3e-2


def _worker_bf16_wire(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      PLUME_GRAD_COMM="bf16")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.trainer import Trainer

    spec = _spec()
    tr = Trainer(spec, device=f"cuda:{rank}", process_group=dist.group.WORLD, seed=rank, bucket_mb=0.5)
    assert tr.model.grad_comm_bf16
    x, t = synthetic_batch(4, 64, 64, spec.in_channels, seed=11)
    half = slice(rank * 2, rank * 2 + 2)
    m = tr.model
    m.train(True)
    m.forward(x[half].cuda(), t[half].cuda())
    m.backward()
    m.wait_grads()
    torch.cuda.synchronize()
    g_eager = m.grads.cpu().clone()
    # and the same through the captured-graph step (casts + all-reduces inside the graph)
    losses = [float(tr.step_graphed(x[half].cuda(), t[half].cuda())[0].item()) for _ in range(3)]
    torch.cuda.synchronize()
    torch.save({"grads": g_eager, "params": m.params.cpu(), "losses": losses}, os.path.join(out_dir, f"b{rank}.pt"))
    dist.barrier()
    tr.release_graphs()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_bf16_gradient_wire_format(tmp_path):
    """PLUME_GRAD_COMM=bf16: buckets are rounded to bf16 for the all-reduce and widened back.  Both ranks end with the
    same gradients; they equal the single-GPU gradient of the full batch to bf16 rounding (2^-8 relative per
    element); graphed steps run and keep the replicas identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + os.getpid() % 90
    mp.spawn(_worker_bf16_wire, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert torch.equal(r0["grads"], r1["grads"]) and torch.equal(r0["params"], r1["params"])
    assert r0["losses"][-1] < r0["losses"][0] and r1["losses"][-1] < r1["losses"][0]   # three steps on one batch
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.unet import UNetB200

    spec = _spec()
    single = UNetB200(spec, device="cuda:0", seed=0)     # rank 0's seed: the replicas were broadcast from rank 0
    x, t = synthetic_batch(4, 64, 64, spec.in_channels, seed=11)
    single.train(True)
    single.forward(x.cuda(), t.cuda())
    single.backward()
    torch.cuda.synchronize()
    g = single.grads.cpu()
    rel = ((r0["grads"] - g).norm() / g.norm()).item()
    print(f"bf16 wire format: all-reduced gradient vs single-GPU gradient rel L2 {rel:.2e}")
    assert rel < 8e-3
