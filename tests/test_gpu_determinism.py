"""Deterministic mode (PLUME_DETERMINISTIC=1 / plume_set_deterministic): two runs of the same training steps give
BIT-IDENTICAL parameters, Adam moments and running statistics; the mode changes results only at rounding level."""
import pytest
import torch

from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(spec, steps, n, hw, det, graphed=False):
    from kcl_ltss_bioatm_b200.trainer import Trainer

    tr = Trainer(spec, device=DEV, seed=5)
    tr.model.ops.set_deterministic(det)
    try:
        losses = []
        for i in range(steps):
            x, t = synthetic_batch(n, hw, hw, spec.in_channels, seed=300 + i)
            out = (tr.step_graphed if graphed else tr.step)(x.to(DEV), t.to(DEV))
            losses.append(out.clone())
        torch.cuda.synchronize()
        m = tr.model
        state = (m.params.clone(), m.adam_m.clone(), m.adam_v.clone(), m._stat_region.clone(), torch.stack(losses))
        tr.release_graphs()
        return state
    finally:
        tr.model.ops.set_deterministic(False)


@pytest.mark.parametrize("spec,n,hw", [(UNetSpec(base_filters=64, depth=2), 4, 32),      # small images: generic kernels
                                       (UNetSpec(), 4, 128),                             # default spec: halo kernels
                                       (UNetSpec(norm="none", depth=3), 2, 64)])
def test_ten_training_steps_are_bit_identical(spec, n, hw):
    a = run(spec, 10, n, hw, det=True)
    b = run(spec, 10, n, hw, det=True)
    for name, u, v in zip(("params", "adam_m", "adam_v", "bn state", "losses"), a, b):
        assert torch.equal(u, v), f"{name} differ between two deterministic runs ({(u != v).sum().item()} elements)"
    c = run(spec, 10, n, hw, det=False)
    rel = ((a[4][:, 0] - c[4][:, 0]).abs() / c[4][:, 0].abs()).max().item()
    print(f"deterministic vs default mode: max loss deviation over 10 steps {rel:.2e}")
    assert rel < 2e-2   # same arithmetic up to summation order


def test_graphed_steps_are_bit_identical_and_match_eager_deterministic():
    spec = UNetSpec(base_filters=64, depth=2)
    a = run(spec, 6, 4, 64, det=True, graphed=True)
    b = run(spec, 6, 4, 64, det=True, graphed=True)
    e = run(spec, 6, 4, 64, det=True, graphed=False)
    for u, v, w in zip(a, b, e):
        assert torch.equal(u, v)
        assert torch.equal(u, w)      # the captured graph runs the very same kernels in the same order


def test_bf16x3_mode_is_deterministic_too():
    spec = UNetSpec(base_filters=64, depth=2, precision="bf16x3")
    a = run(spec, 4, 2, 32, det=True)
    b = run(spec, 4, 2, 32, det=True)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
