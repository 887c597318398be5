"""Drop-in for the geometry functions of the reference's ``src/features/plume_selector.py`` (same names, same
arguments): the hull -> pixel decisions run on the GPU through ``plume_rasterize_hulls``.  The interactive curation
loop of the reference file (matplotlib key presses, plume_selector.py:118-237) is not part of this path."""
from kcl_ltss_bioatm_b200.labels import (find_plume_aod, in_hull, remove_duplicated_plumes,  # noqa: F401
                                         subset_plume)
