"""Data locations used by the model entry points.

Same constant names as the reference's ``src/config/filepaths.py:7-33`` so that code written against
the reference keeps working; the root, hard-coded there to the author's external drive
(``/Volumes/INTENSO/kcl-ltss-bioatm``, filepaths.py:7), comes from the environment here:
    KCL_LTSS_BIOATM_ROOT   (default: ./data under the current working directory)
"""
import os

root_path = os.environ.get("KCL_LTSS_BIOATM_ROOT", os.path.join(os.getcwd(), "data"))

_LAYOUT = {
    # raw data
    "path_to_viirs_sdr": "raw/viirs/sdr",
    "path_to_viirs_sdr_reprojected_tcc": "raw/reprojected_viirs/tcc",
    "path_to_viirs_sdr_reprojected_blue": "raw/reprojected_viirs/blue",
    "path_to_viirs_sdr_reprojected_h5": "raw/reprojected_viirs/h5",
    "path_to_viirs_aod": "raw/viirs/aod",
    "path_to_viirs_geo": "raw/viirs/geo",
    "path_to_viirs_masks": "raw/viirs/masks",
    # machine-learning inputs (imagery, plume masks rasterised from the hull CSVs)
    "path_to_viirs_ml_sdr": "raw/ml_data_viirs/sdr",
    "path_to_viirs_ml_reprojected_tcc": "raw/ml_data_viirs/tcc",
    "path_to_viirs_ml_reprojected_h5": "raw/ml_data_viirs/h5",
    "path_to_viirs_ml_plume_masks": "raw/ml_data_viirs/mask_full_plume",
    "path_to_fire": "raw/fires",
    # tiles ready for the model, and trained models
    "path_to_model_data_folder": "interim/model_input",
    "path_to_model_folder": "interim/models",
}

globals().update({name: os.path.join(root_path, rel) for name, rel in _LAYOUT.items()})
__all__ = ["root_path"] + list(_LAYOUT)
