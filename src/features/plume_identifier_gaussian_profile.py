"""Drop-in for the fire-geolocation functions of the reference's
``src/features/plume_identifier_gaussian_profile.py`` (same names, same arguments): the per-fire nearest-pixel
search runs on the GPU through ``plume_locate_fires``.  The threshold-sweep plume identification of the reference
file (:126-649) is not part of this path."""
from kcl_ltss_bioatm_b200.fires import (P_ID_WIN_SIZE, grid_indexes, haversine, locate_fire_in_image,  # noqa: F401
                                        subset_fires_to_image)
