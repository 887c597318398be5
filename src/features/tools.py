"""Reference layout: ``src.features.tools`` (``/root/reference/src/features/tools.py``).  ``utm_resampler`` is the
GPU implementation; ``read_modis_aod`` needs pyhdf (HDF4), which this image does not have, and is not rebuilt."""
from kcl_ltss_bioatm_b200.resample import utm_resampler  # noqa: F401
