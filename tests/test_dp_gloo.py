"""CPU, world_size 2 over gloo: the data-parallel path of UNetB200 / Trainer (bucketed asynchronous
all-reduce launched from inside backward, 1/world loss scaling, parameter broadcast) driven by the CPU
operator oracle.  Two ranks with half the batch each must reproduce a single process that sees both
halves -- exactly for a BatchNorm-free, Dice-free spec (BCE is a mean over pixels), which isolates the
collective logic from the per-GPU statistics that DP legitimately changes."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, per_bucket_adam=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec
    from kcl_ltss_bioatm_b200.trainer import Trainer
    from oracle.ops_ref import RefOps

    if per_bucket_adam:
        os.environ["PLUME_ADAM_PER_BUCKET"] = "1"
    spec = UNetSpec(base_filters=64, depth=2, norm="none", dice_weight=0.0)
    # deliberately different seeds: the constructor's broadcast must make the replicas identical
    tr = Trainer(spec, device="cpu", process_group=dist.group.WORLD, ops=RefOps(torch.float32), seed=rank,
                 bucket_mb=0.5)
    assert len(tr.model._buckets) >= 3          # several buckets: exercises the ordering logic
    x, t = synthetic_batch(4, 16, 16, spec.in_channels, seed=11, dtype=torch.float32)
    half = slice(rank * 2, rank * 2 + 2)
    m = tr.model
    m.train(True)
    p0 = m.params.clone()
    m.forward(x[half], t[half])
    m.backward()
    if per_bucket_adam:
        # optimizer_step() itself takes the pending all-reduce handles one by one and updates each bucket as its
        # all-reduce completes (the product path; wait_grads() first = one Adam launch over everything)
        assert len(m._pending) == len(m._buckets) and m.adam_per_bucket
        m.optimizer_step()
        assert not m._pending
        torch.save({"grads": m.grads.clone(), "params0": p0}, os.path.join(out_dir, f"r{rank}.pt"))
    else:
        m.wait_grads()
        torch.save({"grads": m.grads.clone(), "params0": p0}, os.path.join(out_dir, f"r{rank}.pt"))
        m.optimizer_step()
    torch.save(m.params.clone(), os.path.join(out_dir, f"p{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("per_bucket_adam", [False, True])
def test_two_rank_gloo_matches_single_process(tmp_path, per_bucket_adam):
    port = 29600 + os.getpid() % 300 + (300 if per_bucket_adam else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), per_bucket_adam), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(r0["params0"], r1["params0"])       # broadcast from rank 0
    assert torch.equal(r0["grads"], r1["grads"])           # all-reduced gradients identical on both ranks
    assert torch.equal(torch.load(tmp_path / "p0.pt"), torch.load(tmp_path / "p1.pt"))

    sys.path.insert(0, ROOT)
    from kcl_ltss_bioatm_b200.data import synthetic_batch
    from kcl_ltss_bioatm_b200.spec import UNetSpec
    from kcl_ltss_bioatm_b200.unet import UNetB200
    from oracle.ops_ref import RefOps

    spec = UNetSpec(base_filters=64, depth=2, norm="none", dice_weight=0.0)
    single = UNetB200(spec, ops=RefOps(torch.float32), device="cpu", seed=None)
    single.params.copy_(r0["params0"])
    single._param_version += 1
    x, t = synthetic_batch(4, 16, 16, spec.in_channels, seed=11, dtype=torch.float32)
    single.forward(x, t)
    single.backward()
    err = (r0["grads"] - single.grads).abs().max() / single.grads.abs().max()
    assert err < 1e-5, float(err)
    single.grads.copy_(r0["grads"])             # same gradients in: Adam per bucket == Adam over the whole buffer
    single.optimizer_step()
    assert torch.equal(single.params, torch.load(tmp_path / "p0.pt"))


def test_bucket_ranges_tile_the_gradient_buffer():
    sys.path.insert(0, ROOT)
    from kcl_ltss_bioatm_b200.spec import UNetSpec, build_layout
    from kcl_ltss_bioatm_b200.unet import UNetB200

    lay = build_layout(UNetSpec())

    class Dummy:
        layout = lay

    buckets = UNetB200._make_buckets(Dummy(), 25.0)
    assert buckets[0][0] == 0 and buckets[-1][1] == lay.total
    for (a, b), (c, d) in zip(buckets, buckets[1:]):
        assert b == c and a < b
    assert 2 <= len(buckets) <= 8   # 124 MB of gradients in ~25 MB+ buckets cut at module boundaries
