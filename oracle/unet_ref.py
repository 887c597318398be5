"""TEST INFRASTRUCTURE -- the whole-network oracle: a plain PyTorch fp32 CPU UNet.

PARITY UNPINNED.  The reference repository has no UNet to restate: ``/root/reference/src/models/`` holds
two empty files and ``train_model.py`` / ``predict_model.py`` exist only in the README's directory
diagram (README.md:44-47).  No golden vector, fixture or reference output exists for
this path, so this module *is* the frozen definition (UNetSpec in ``kcl_ltss_bioatm_b200/spec.py``):
Ronneberger-style encoder-decoder, DoubleConv = [Conv3x3(pad 1, bias) -> BatchNorm2d -> ReLU] x 2,
MaxPool2d(2), ConvTranspose2d(k=2, s=2), cat([skip, up]), Conv1x1 head, BCEWithLogits + Dice loss, Adam.
The committed fixtures under tests/golden/ are generated from this module by
``scripts/make_golden.py`` and pin the oracle against accidental edits, nothing more.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module; the product (``src/models``, ``kcl_ltss_bioatm_b200``) never does.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from kcl_ltss_bioatm_b200.spec import UNetSpec


class DoubleConv(nn.Module):
    def __init__(self, cin: int, cout: int, spec: UNetSpec):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, bias=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=True)
        self.use_bn = spec.norm == "batch"
        if self.use_bn:
            self.bn1 = nn.BatchNorm2d(cout, eps=spec.bn_eps, momentum=spec.bn_momentum)
            self.bn2 = nn.BatchNorm2d(cout, eps=spec.bn_eps, momentum=spec.bn_momentum)

    def forward(self, x):
        x = self.conv1(x)
        if self.use_bn:
            x = self.bn1(x)
        x = F.relu(x)
        x = self.conv2(x)
        if self.use_bn:
            x = self.bn2(x)
        return F.relu(x)


class UNetRef(nn.Module):
    """Construction order (and therefore RNG consumption under a fixed seed) is: enc0..enc{D-1},
    bottleneck, then for l = D-1..0: up{l}, dec{l}; finally head.  ``UNetB200.init_parameters`` follows
    the same order so both start from identical weights under the same seed."""

    def __init__(self, spec: UNetSpec = UNetSpec()):
        super().__init__()
        self.spec = spec
        d = spec.depth
        for l in range(d):
            cin = spec.in_channels if l == 0 else spec.channels(l - 1)
            setattr(self, f"enc{l}", DoubleConv(cin, spec.channels(l), spec))
        self.bottleneck = DoubleConv(spec.channels(d - 1), spec.channels(d), spec)
        for l in reversed(range(d)):
            c = spec.channels(l)
            setattr(self, f"up{l}", nn.ConvTranspose2d(2 * c, c, 2, stride=2))
            setattr(self, f"dec{l}", DoubleConv(2 * c, c, spec))
        self.head = nn.Conv2d(spec.base_filters, spec.n_classes, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [N, C_in, H, W] fp32 -> logits [N, 1, H, W]."""
        d = self.spec.depth
        skips = []
        for l in range(d):
            x = getattr(self, f"enc{l}")(x)
            skips.append(x)
            x = F.max_pool2d(x, 2)
        x = self.bottleneck(x)
        for l in reversed(range(d)):
            x = getattr(self, f"up{l}")(x)
            x = torch.cat([skips[l], x], dim=1)
            x = getattr(self, f"dec{l}")(x)
        return self.head(x)


def plume_loss(logits: torch.Tensor, target: torch.Tensor, spec: UNetSpec) -> torch.Tensor:
    """bce_weight * mean BCE-with-logits + dice_weight * (1 - (2*sum(p t)+eps)/(sum p + sum t + eps)),
    sums taken over the whole batch.  target: {0,1} of the logits' shape."""
    t = target.to(logits.dtype)
    bce = F.binary_cross_entropy_with_logits(logits, t)
    p = torch.sigmoid(logits)
    dice = 1 - (2 * (p * t).sum() + spec.dice_eps) / (p.sum() + t.sum() + spec.dice_eps)
    return spec.bce_weight * bce + spec.dice_weight * dice


def make_optimizer(model: nn.Module, spec: UNetSpec) -> torch.optim.Optimizer:
    return torch.optim.Adam(model.parameters(), lr=spec.lr, betas=spec.betas, eps=spec.adam_eps)


def predict_mask(model: nn.Module, x: torch.Tensor, spec: UNetSpec) -> torch.Tensor:
    """uint8 [N, H, W] mask, sigmoid(logit) >= threshold, eval mode (running BN statistics)."""
    was = model.training
    model.eval()
    with torch.no_grad():
        logits = model(x)[:, 0]
    model.train(was)
    return (torch.sigmoid(logits) >= spec.mask_threshold).to(torch.uint8)


def with_bf16_storage(model: nn.Module) -> nn.Module:
    """Copy of `model` that rounds to bf16 exactly where a bf16-storage implementation must: the GEMM
    weights, every conv / transposed-conv output and every BatchNorm output (ReLU and max-pool commute
    with rounding).  Arithmetic stays fp32.  The gap between this and the fp32 model is the error floor of
    ANY bf16-activation implementation of the network; the gap between this and the CUDA path is what the
    kernels themselves add."""
    import copy

    m2 = copy.deepcopy(model)

    def _round(_m, _i, out):
        return out.to(torch.bfloat16).float()

    for name, mod in m2.named_modules():
        if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
            if name != "head":
                mod.weight.data = mod.weight.data.to(torch.bfloat16).float()
                mod.register_forward_hook(_round)
        elif isinstance(mod, nn.BatchNorm2d):
            mod.register_forward_hook(_round)
    return m2
