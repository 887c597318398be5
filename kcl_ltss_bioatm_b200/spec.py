"""Frozen model definition of the smoke-plume UNet: hyper-parameters, layer table, parameter layout.

DEFINED HERE -- the reference pins nothing.  gridl/kcl-ltss-bioatm names a UNet (README.md:1-4) and
reserves ``src/models/train_model.py`` / ``predict_model.py`` for it (README.md:44-47), but ships no
model code (``src/models/__init__.py`` is empty).  The values below are the
canonical Ronneberger-style UNet recommended in SURVEY.md section 8(c); ``oracle/unet_ref.py`` is the
same definition as a plain PyTorch module and is the parity oracle.

Layer naming (= state_dict keys of the oracle):
    enc{l}.conv1/.bn1/.conv2/.bn2   l = 0..depth-1     C_l = base_filters * 2**l
    bottleneck.conv1/.bn1/.conv2/.bn2                  C = base_filters * 2**depth
    up{l}  (ConvTranspose2d 2C_l -> C_l, k=2, s=2)     dec{l}.conv1 (2C_l -> C_l) /.bn1/.conv2/.bn2
    head   (Conv2d C_0 -> 1, 1x1)
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class UNetSpec:
    in_channels: int = 8          # MODIS/VIIRS-like multi-band tile
    n_classes: int = 1            # plume / no plume
    base_filters: int = 64
    depth: int = 4
    norm: str = "batch"           # "batch": Conv -> BatchNorm -> ReLU;  "none": Conv(+bias) -> ReLU
    bn_eps: float = 1e-5
    bn_momentum: float = 0.1
    bce_weight: float = 1.0       # loss = bce_weight * BCEWithLogits(mean) + dice_weight * Dice
    dice_weight: float = 1.0
    dice_eps: float = 1.0         # Dice = 1 - (2*sum(p*t) + eps) / (sum(p) + sum(t) + eps), over the batch
    mask_threshold: float = 0.5   # mask = sigmoid(logit) >= threshold, uint8
    lr: float = 1e-3              # Adam
    betas: Tuple[float, float] = (0.9, 0.999)
    adam_eps: float = 1e-8
    # Arithmetic of the GPU path (not part of the model definition; the oracle ignores it):
    #   "bf16"  : bf16 activations and GEMM operands, fp32 accumulate -- north_star's 1e-2 logits bar;
    #   "bf16x3": activations / operands as bf16 hi + lo pairs, three MMA passes (hi*hi + hi*lo + lo*hi), fp32
    #             accumulate -- north_star's "tf32 mode" bar of 1e-3 ("tf32" is accepted as an alias; a single
    #             kind::tf32 pass keeps 11 significant bits and does not reach 1e-3 on the 23-layer network).
    precision: str = "bf16"

    def __post_init__(self):
        if self.precision == "tf32":
            object.__setattr__(self, "precision", "bf16x3")
        if self.precision not in ("bf16", "bf16x3"):
            raise ValueError("precision must be 'bf16' or 'bf16x3' (alias 'tf32')")
        if self.n_classes != 1:
            raise ValueError("the plume segmenter is single-class (n_classes == 1)")
        if self.norm not in ("batch", "none"):
            raise ValueError("norm must be 'batch' or 'none'")
        if self.base_filters % 64 != 0:
            raise ValueError("base_filters must be a multiple of 64 (tensor-core tile width)")
        if self.depth < 1:
            raise ValueError("depth must be >= 1")

    # ------------------------------------------------------------------ derived quantities
    @property
    def cin_padded(self) -> int:
        """First-layer input channels as seen by the implicit-GEMM kernel (zero padded to 64)."""
        return (self.in_channels + 63) // 64 * 64

    def channels(self, level: int) -> int:
        return self.base_filters * (2 ** level)

    def to_dict(self) -> dict:
        d = asdict(self)
        d["betas"] = list(self.betas)
        return d

    @staticmethod
    def from_dict(d: dict) -> "UNetSpec":
        d = dict(d)
        if "betas" in d:
            d["betas"] = tuple(d["betas"])
        return UNetSpec(**d)

    @staticmethod
    def wide() -> "UNetSpec":
        """BASELINE.json config 5: 2x base filters, depth 5."""
        return UNetSpec(base_filters=128, depth=5)

    def divisor(self) -> int:
        """Tile height/width must be a multiple of this (depth poolings)."""
        return 2 ** self.depth


@dataclass
class ConvLayer:
    name: str        # e.g. "enc0.conv1"
    bn: str          # e.g. "enc0.bn1"
    cin: int         # logical input channels (oracle)
    cin_k: int       # input channels seen by the kernel (>= cin, multiple of 64)
    cout: int
    level: int       # spatial level (0 = full resolution)


@dataclass
class UpLayer:
    name: str        # "up{l}"
    cin: int
    cout: int
    level: int       # output level


@dataclass
class ParamSlot:
    key: str                 # oracle state_dict key
    shape: Tuple[int, ...]   # shape in the flat buffer (kernel layout)
    offset: int              # element offset in the flat fp32 buffer
    numel: int
    kind: str                # conv_w | conv_b | bn_w | bn_b | up_w | up_b | head_w | head_b


@dataclass
class Layout:
    """Flat fp32 parameter buffer, ordered in REVERSE execution order (head first, enc0 last) so that
    gradient buckets become ready front to back during the backward pass."""
    slots: Dict[str, ParamSlot] = field(default_factory=dict)
    order: List[str] = field(default_factory=list)
    total: int = 0


def conv_layers(spec: UNetSpec) -> Dict[str, ConvLayer]:
    out: Dict[str, ConvLayer] = {}
    d = spec.depth
    for l in range(d):
        c = spec.channels(l)
        cin = spec.in_channels if l == 0 else spec.channels(l - 1)
        cin_k = spec.cin_padded if l == 0 else cin
        out[f"enc{l}.conv1"] = ConvLayer(f"enc{l}.conv1", f"enc{l}.bn1", cin, cin_k, c, l)
        out[f"enc{l}.conv2"] = ConvLayer(f"enc{l}.conv2", f"enc{l}.bn2", c, c, c, l)
    cb = spec.channels(d)
    out["bottleneck.conv1"] = ConvLayer("bottleneck.conv1", "bottleneck.bn1", spec.channels(d - 1), spec.channels(d - 1), cb, d)
    out["bottleneck.conv2"] = ConvLayer("bottleneck.conv2", "bottleneck.bn2", cb, cb, cb, d)
    for l in range(d):
        c = spec.channels(l)
        out[f"dec{l}.conv1"] = ConvLayer(f"dec{l}.conv1", f"dec{l}.bn1", 2 * c, 2 * c, c, l)
        out[f"dec{l}.conv2"] = ConvLayer(f"dec{l}.conv2", f"dec{l}.bn2", c, c, c, l)
    return out


def up_layers(spec: UNetSpec) -> Dict[str, UpLayer]:
    return {f"up{l}": UpLayer(f"up{l}", 2 * spec.channels(l), spec.channels(l), l) for l in range(spec.depth)}


def execution_order(spec: UNetSpec) -> List[str]:
    """Module names in forward execution order."""
    d = spec.depth
    names = [f"enc{l}" for l in range(d)] + ["bottleneck"]
    for l in reversed(range(d)):
        names += [f"up{l}", f"dec{l}"]
    return names + ["head"]


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


def build_layout(spec: UNetSpec) -> Layout:
    lay = Layout()
    convs, ups = conv_layers(spec), up_layers(spec)

    def add(key, shape, kind):
        numel = 1
        for s in shape:
            numel *= s
        lay.slots[key] = ParamSlot(key, tuple(shape), lay.total, numel, kind)
        lay.order.append(key)
        lay.total = _align(lay.total + numel)  # 256-byte aligned slots (TMA / vector loads)

    def add_block(block):
        for cv in ("conv2", "conv1"):
            L = convs[f"{block}.{cv}"]
            add(f"{L.name}.weight", (L.cout, 3, 3, L.cin_k), "conv_w")
            add(f"{L.name}.bias", (L.cout,), "conv_b")
            if spec.norm == "batch":
                add(f"{L.bn}.weight", (L.cout,), "bn_w")
                add(f"{L.bn}.bias", (L.cout,), "bn_b")

    add("head.weight", (spec.base_filters,), "head_w")
    add("head.bias", (1,), "head_b")
    for l in range(spec.depth):
        add_block(f"dec{l}")
        U = ups[f"up{l}"]
        add(f"up{l}.weight", (4, U.cout, U.cin), "up_w")
        add(f"up{l}.bias", (U.cout,), "up_b")
    add_block("bottleneck")
    for l in reversed(range(spec.depth)):
        add_block(f"enc{l}")
    return lay


def fwd_flops_per_tile(spec: UNetSpec, h: int, w: int) -> Dict[str, float]:
    """Algorithmic forward FLOPs for one tile (logical channels)."""
    conv = 0.0
    for L in conv_layers(spec).values():
        hh, ww = h >> L.level, w >> L.level
        conv += 2.0 * 9 * L.cin * L.cout * hh * ww
    up = 0.0
    for U in up_layers(spec).values():
        hh, ww = h >> (U.level + 1), w >> (U.level + 1)
        up += 2.0 * 4 * U.cin * U.cout * hh * ww
    head = 2.0 * spec.base_filters * h * w
    return {"conv3x3": conv, "convT": up, "head": head, "total": conv + up + head}


def train_flops_per_tile(spec: UNetSpec, h: int, w: int) -> float:
    """fwd + dgrad + wgrad; the first layer has no dgrad."""
    f = fwd_flops_per_tile(spec, h, w)
    first = 2.0 * 9 * spec.in_channels * spec.base_filters * h * w
    return 3.0 * f["total"] - first


def num_parameters(spec: UNetSpec) -> int:
    """Logical (oracle) parameter count: conv / transposed-conv / head weights and biases, BatchNorm affine."""
    n = 0
    for L in conv_layers(spec).values():
        n += L.cout * 9 * L.cin + L.cout + (2 * L.cout if spec.norm == "batch" else 0)
    for U in up_layers(spec).values():
        n += 4 * U.cin * U.cout + U.cout
    return n + spec.base_filters * spec.n_classes + spec.n_classes
