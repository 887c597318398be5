"""Committed golden vectors (tests/golden/, made by scripts/make_golden.py from the in-repo oracle).
CPU: the oracle and the synthetic-data generator still reproduce them (guards against accidental edits of
the frozen spec).  GPU: the CUDA path reproduces them without touching /root/reference or recomputing the
oracle.  The reference repository itself holds no vectors for this path."""
import os

import numpy as np
import pytest
import torch

from kcl_ltss_bioatm_b200.data import synthetic_batch
from kcl_ltss_bioatm_b200.spec import UNetSpec

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unet_d2_f64_seed0.npz")
SPEC = UNetSpec(base_filters=64, depth=2)


def load():
    g = np.load(GOLD)
    x = torch.from_numpy(g["x_bf16_bits"]).view(torch.bfloat16)
    return g, x, torch.from_numpy(g["mask"])


def test_synthetic_data_is_reproducible():
    g, x, t = load()
    x2, t2 = synthetic_batch(2, 32, 32, SPEC.in_channels, seed=42)
    assert torch.equal(x2, x) and torch.equal(t2, t)
    assert 0.02 < t.float().mean().item() < 0.5


def test_oracle_reproduces_golden():
    from oracle.unet_ref import UNetRef, plume_loss

    g, x, t = load()
    torch.manual_seed(0)
    ref = UNetRef(SPEC).train()
    logits = ref(x.float().permute(0, 3, 1, 2))[:, 0]
    loss = plume_loss(logits, t, SPEC)
    loss.backward()
    assert np.allclose(logits.detach().numpy(), g["logits_train"], rtol=1e-4, atol=1e-5)
    assert abs(float(loss.detach()) - float(g["loss"][0])) < 1e-5
    assert np.allclose(ref.head.weight.grad.numpy().reshape(-1), g["grad_head_weight"], rtol=1e-3, atol=1e-6)
    assert np.allclose(ref.enc0.conv1.weight.detach().numpy()[:4, :, 1, 1], g["enc0_conv1_weight_sample"])
    ref.eval()
    with torch.no_grad():
        le = ref(x.float().permute(0, 3, 1, 2))[:, 0]
    assert np.allclose(le.numpy(), g["logits_eval"], rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    from kcl_ltss_bioatm_b200.unet import UNetB200

    g, x, t = load()
    net = UNetB200(SPEC, device="cuda:0", seed=0)
    z = net.forward(x.cuda(), t.cuda())
    net.backward()
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["logits_train"])
    rel = ((z.cpu() - ref).norm() / ref.norm()).item()
    assert rel <= 1e-2, rel                       # bf16 bound; depth-2 network
    assert abs(net.loss_out[0].item() - float(g["loss"][0])) <= 1e-2 * float(g["loss"][0])
    gh = net.grad_dict()["head.weight"].reshape(-1)
    gr = torch.from_numpy(g["grad_head_weight"])
    assert ((gh - gr).norm() / gr.norm()).item() < 2e-2
    gb = net.grad_dict()["enc0.bn1.weight"]
    gbr = torch.from_numpy(g["grad_enc0_bn1_weight"])
    assert (torch.dot(gb, gbr) / (gb.norm() * gbr.norm())).item() > 0.95
