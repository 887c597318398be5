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
