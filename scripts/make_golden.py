"""Generate the committed golden fixtures under tests/golden/ from the in-repo oracle.

The reference repository has no model code and no test vectors for this path (SURVEY.md sections 0, 4), so
these fixtures do not pin the oracle to the reference; they freeze the oracle (and the synthetic data
generator) against accidental edits, and give the GPU tests inputs/outputs that do not depend on
/root/reference.  Run on CPU:  python scripts/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kcl_ltss_bioatm_b200.data import synthetic_batch  # noqa: E402
from kcl_ltss_bioatm_b200.spec import UNetSpec  # noqa: E402
from oracle.unet_ref import UNetRef, plume_loss  # noqa: E402

torch.set_num_threads(1)  # bit-stable summation order
spec = UNetSpec(base_filters=64, depth=2)
torch.manual_seed(0)
ref = UNetRef(spec).train()
x, t = synthetic_batch(2, 32, 32, spec.in_channels, seed=42)
xr = x.float().permute(0, 3, 1, 2).contiguous()
logits = ref(xr)[:, 0]
loss = plume_loss(logits, t, spec)
loss.backward()
ref.eval()
with torch.no_grad():
    logits_eval = ref(xr)[:, 0]
out = {
    "x_bf16_bits": x.view(torch.int16).numpy(),          # exact input bits
    "mask": t.numpy(),
    "logits_train": logits.detach().numpy(),
    "logits_eval": logits_eval.numpy(),
    "loss": np.array([float(loss)], dtype=np.float64),
    "grad_head_weight": ref.head.weight.grad.numpy().reshape(-1),
    "grad_enc0_bn1_weight": ref.enc0.bn1.weight.grad.numpy(),
    "enc0_conv1_weight_sample": ref.enc0.conv1.weight.detach().numpy()[:4, :, 1, 1],
}
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
path = os.path.join(ROOT, "tests", "golden", "unet_d2_f64_seed0.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()}, "loss", float(loss))
