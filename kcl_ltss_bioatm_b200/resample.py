"""UTM nearest-neighbour resampler on the GPU: the drop-in for the reference's ``utm_resampler``
(``/root/reference/src/features/tools.py:9-64``) -- same constructor arguments, attributes (``zone``, ``extent``,
``x_size``, ``y_size``) and methods (``resample_image``, ``resample_points_to_utm``, ``resample_point_to_geo``).

The reference builds a pyproj UTM projection and a pyresample ``AreaDefinition`` and calls
``pr.kd_tree.resample_nearest(swath, image, area, radius_of_influence=10000, fill_value=...)``.  Neither library is
needed here: the projection (Krueger series, fp64) and the neighbour search (counting sort into buckets + scan) are
CUDA kernels behind the C ABI (``csrc/resample.cu``); what the libraries compute is restated and anchored in
``oracle/resample_ref.py``.  numpy arrays in, numpy arrays out, like the reference; the neighbour INDEX map of the last
swath geometry is cached on the device, so resampling further images on the same swath costs one gather.

``read_modis_aod`` (``tools.py:67-130``) reads HDF4 through pyhdf, which is absent: the file parsing is not rebuilt
(DESIGN.md section 7); its geolocation half (``:97-128``, the lat / lon arrays of the sinusoidal grid) is
``modis_grid_latlon`` below.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

RADIUS_OF_INFLUENCE = 10000.0   # tools.py:57
MODIS_SPHERE_RADIUS = 6371007.181   # tools.py:124 (+proj=sinu +R=6371007.181)


class utm_resampler(object):
    def __init__(self, lats, lons, pixel_size, device="cuda", ops=None):
        if ops is None:
            from .ops import CudaOps  # fails loudly without the library / a GPU

            ops = CudaOps()
        self.ops, self.device = ops, torch.device(device)
        self.lats = np.asarray(lats, dtype=np.float64)
        self.lons = np.asarray(lons, dtype=np.float64)
        self.pixel_size = pixel_size
        self._lat_d, self._lon_d = self._to_dev(self.lats), self._to_dev(self.lons)
        self.zone = self.__utm_zone()
        self.extent = self.__utm_extent()
        self.x_size, self.y_size = self.__utm_grid_size()
        self._cache_key: Optional[tuple] = None
        self._cache_idx: Optional[torch.Tensor] = None

    def _to_dev(self, a) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device).reshape(-1)

    def __utm_zone(self):
        """tools.py:20-28: the zone in which most of the data falls (smallest zone among ties, as scipy's mode)."""
        hist = torch.empty(64, dtype=torch.int32, device=self.device)
        self.ops.utm_zone_histogram(self._lon_d, hist)
        return int(torch.argmax(hist).item())   # argmax returns the first maximum

    def __utm_extent(self):
        """tools.py:33-37: bounding box of the projected swath coordinates."""
        x, y = torch.empty_like(self._lat_d), torch.empty_like(self._lat_d)
        self.ops.utm_forward(self._lat_d, self._lon_d, self.zone, x, y)
        return (float(x.min()), float(y.min()), float(x.max()), float(y.max()))

    def __utm_grid_size(self):
        """tools.py:39-42."""
        x_size = int(np.round((self.extent[2] - self.extent[0]) / self.pixel_size))
        y_size = int(np.round((self.extent[3] - self.extent[1]) / self.pixel_size))
        return x_size, y_size

    def neighbour_index(self, image_lats, image_lons) -> torch.Tensor:
        """int32 [y_size, x_size] on the device: flat swath index per target cell (-1 = no pixel within 10 km)."""
        la, lo = np.asarray(image_lats, dtype=np.float64), np.asarray(image_lons, dtype=np.float64)
        key = (la.shape, la.tobytes(), lo.tobytes())
        if self._cache_key != key:
            idx = torch.empty(self.y_size, self.x_size, dtype=torch.int32, device=self.device)
            self.ops.resample_nearest_index(self._to_dev(la), self._to_dev(lo), self.zone, self.extent, self.x_size,
                                            self.y_size, RADIUS_OF_INFLUENCE, idx)
            self._cache_key, self._cache_idx = key, idx
        return self._cache_idx

    def resample_image(self, image, image_lats, image_lons, fill_value=-999):
        """tools.py:52-58.  image: float32 / float64 array of the swath's shape -> [y_size, x_size] array."""
        img = np.asarray(image)
        if img.dtype not in (np.float32, np.float64):
            img = img.astype(np.float64)
        if img.shape != np.shape(image_lats):
            raise ValueError("image and image_lats / image_lons must have the same shape")
        idx = self.neighbour_index(image_lats, image_lons)
        src = torch.from_numpy(np.ascontiguousarray(img)).to(self.device).reshape(-1)
        out = torch.empty(self.y_size, self.x_size, dtype=src.dtype, device=self.device)
        self.ops.gather_fill(src, idx, fill_value, out)
        return out.cpu().numpy()

    def resample_points_to_utm(self, point_lats, point_lons):
        """tools.py:60-61: [(x, y), ...] in metres."""
        la, lo = self._to_dev(point_lats), self._to_dev(point_lons)
        x, y = torch.empty_like(la), torch.empty_like(la)
        self.ops.utm_forward(la, lo, self.zone, x, y)
        return list(zip(x.cpu().tolist(), y.cpu().tolist()))

    def resample_point_to_geo(self, point_y, point_x):
        """tools.py:63-64: (lon, lat) of a UTM point."""
        x, y = self._to_dev([point_x]), self._to_dev([point_y])
        la, lo = torch.empty_like(x), torch.empty_like(x)
        self.ops.utm_inverse(x, y, self.zone, la, lo)
        return float(lo.item()), float(la.item())


def modis_grid_latlon(x0, y0, x1, y1, ny, nx, device="cuda", ops=None):
    """(lat, lon) float64 [ny, nx] device tensors of a MODIS sinusoidal grid from its corner coordinates in metres
    (``UpperLeftPointMtrs`` = (x0, y0), ``LowerRightMtrs`` = (x1, y1) of StructMetadata.0) -- the second half of
    ``read_modis_aod`` (tools.py:103-128): ``xinc = (x1 - x0) / nx``, ``x = linspace(x0, x0 + xinc * nx, nx)`` (likewise
    y), meshgrid, ``pyproj.transform(sinu, wgs84, xv, yv)`` with ``+proj=sinu +R=6371007.181 +nadgrids=@null``."""
    if ops is None:
        from .ops import CudaOps

        ops = CudaOps()
    x0, y0, x1, y1 = float(x0), float(y0), float(x1), float(y1)
    xinc, yinc = (x1 - x0) / nx, (y1 - y0) / ny                      # tools.py:115-116
    lat = torch.empty(int(ny), int(nx), dtype=torch.float64, device=device)
    lon = torch.empty_like(lat)
    ops.sinusoidal_grid_latlon(x0, x0 + xinc * nx, y0, y0 + yinc * ny, MODIS_SPHERE_RADIUS, lat, lon)
    return lat, lon
