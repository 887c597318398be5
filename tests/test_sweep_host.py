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
