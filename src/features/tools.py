"""Reference layout: ``src.features.tools`` (``/root/reference/src/features/tools.py``).  ``utm_resampler`` is the
GPU implementation; ``read_modis_aod`` needs pyhdf (HDF4), which this image does not have: its file parsing is not
rebuilt, its geolocation half (tools.py:97-128) is ``modis_grid_latlon``."""
from kcl_ltss_bioatm_b200.resample import modis_grid_latlon, utm_resampler  # noqa: F401
