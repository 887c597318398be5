"""Drop-in for the fire-geolocation and threshold-sweep functions of the reference's
``src/features/plume_identifier_gaussian_profile.py`` (same names, same arguments): the per-fire nearest-pixel
search runs on the GPU through ``plume_locate_fires``, mask generation / labelling / plume extents through
``plume_sweep_extents`` (bit planes; ``plume_threshold_masks``, ``plume_label_components`` and ``plume_fire_extents``
for dense planes), fire clustering through ``plume_label_components``, the nearest-valid fill of the AOD grid through ``plume_fill_nearest``.  The rest of the reference file
(plume acceptance tests, hull extraction, :243-649) is not part of this path."""
from kcl_ltss_bioatm_b200.fires import (P_ID_WIN_SIZE, grid_indexes, haversine, locate_fire_in_image,  # noqa: F401
                                        subset_fires_to_image)
from kcl_ltss_bioatm_b200.sweep import (NULL_VALUE, cluster_fires, find_plume_extents,  # noqa: F401,E402
                                        find_threshold_index, fire_cluster_centroids, generate_mask_dict,
                                        interpolate_aod_nearest)
