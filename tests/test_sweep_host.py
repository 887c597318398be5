"""CPU: the host part of the threshold-sweep path (find_threshold_index) against the reference's outputs."""
import os

import numpy as np

from kcl_ltss_bioatm_b200 import sweep

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sweep_cases.npz"))


def test_find_threshold_index_matches_reference():
    idx = [-1 if v is None else v for v in sweep.find_threshold_index(G["tables"])]
    assert idx == G["tables_index"].tolist()
    for i in range(int(G["n_cases"])):
        got = [-1 if v is None else v for v in sweep.find_threshold_index(G[f"c{i}_extents"])]
        assert got == G[f"c{i}_index"].tolist()


def test_reference_named_module_reexports():
    import src.features.plume_identifier_gaussian_profile as ref_named

    assert ref_named.find_threshold_index is sweep.find_threshold_index
    assert ref_named.generate_mask_dict is sweep.generate_mask_dict
    assert ref_named.find_plume_extents is sweep.find_plume_extents
    assert ref_named.cluster_fires is sweep.cluster_fires


def test_mask_dict_is_a_lazy_dict_over_bit_planes():
    """MaskDict without a GPU: bit planes in a CPU tensor; dict protocol, lazy unpacking, duplicate thresholds,
    host edits."""
    import torch
    rng = np.random.default_rng(0)
    h, w = 9, 70
    masks = rng.random((4, h, w)) < 0.5
    segs = (w + 31) // 32
    padded = np.zeros((4, h, segs * 32), dtype=bool)
    padded[:, :, :w] = masks
    bits = torch.from_numpy(np.ascontiguousarray(np.packbits(padded, axis=2, bitorder="little").view("<u4").view(np.int32)))
    thr = [0.5, 0.25, 0.5, 0.1]                                   # 0.5 twice: one key, the later plane (same mask in practice)
    d = sweep.MaskDict(thr, bits, w)
    assert isinstance(d, dict) and list(d) == [0.5, 0.25, 0.1] and len(d) == 3 and 0.25 in d
    assert all(v is None for v in dict.values(d))
    assert np.array_equal(d[0.25], masks[1]) and np.array_equal(d[0.5], masks[2]) and d.get(0.7) is None
    assert np.array_equal(np.stack([d[k] for k in d]), masks[[2, 1, 3]])
    assert np.array_equal(np.stack(d.values()), masks[[2, 1, 3]]) and [k for k, _ in d.items()] == [0.5, 0.25, 0.1]
    fresh = sweep.MaskDict(thr, bits, w)
    assert all(isinstance(v, np.ndarray) for v in dict(fresh).values()) and np.array_equal(fresh.copy()[0.25], masks[1])
    merged = {}
    merged.update(sweep.MaskDict(thr, bits, w))
    assert np.array_equal(merged[0.1], masks[3])
    popped = sweep.MaskDict(thr, bits, w)
    assert np.array_equal(popped.pop(0.25), masks[1]) and list(popped) == [0.5, 0.1] and popped.pop(9.0, None) is None
    assert popped.device_planes().shape == (2, h, segs)
    planes = d.device_planes()
    assert planes.shape == (3, h, segs) and torch.equal(planes, bits[[2, 1, 3]])
    assert np.array_equal(sweep.ThresholdSweep.unpack_bits(planes, w), masks[[2, 1, 3]])
    d[0.1] = np.zeros((h, w), dtype=bool)
    assert d.device_planes() is None and not d[0.1].any()
