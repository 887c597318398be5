"""TEST INFRASTRUCTURE -- CPU oracle of the reference's UTM nearest-neighbour resampler
(``/root/reference/src/features/tools.py:9-64``, class ``utm_resampler``).

PARITY UNPINNED AGAINST THE LIBRARIES: the reference delegates the two computations to third-party packages that are
absent from this container and not vendored in /root/reference -- ``pyproj`` (the UTM projection, ``tools.py:30-31``;
requirements.txt pins no version) and ``pyresample`` (``pr.kd_tree.resample_nearest``, ``tools.py:52-58``).  What is
restated here is their published behaviour:

* UTM = transverse Mercator on WGS84 (a = 6378137, 1/f = 298.257223563), k0 = 0.9996, false easting 500 km, no false
  northing (``pyproj.Proj(proj='utm', zone=z)`` without ``south``), evaluated with the Krueger series in the third
  flattening n (Karney 2011, "Transverse Mercator with an accuracy of a few nanometers", eqs. 7-11 and 35-36) -- what
  PROJ's ``utm`` evaluates to well below a millimetre everywhere a swath can reach.  Pinned by the published anchors
  in ``tests/test_resample_oracle.py`` (easting 833 978.557 m at the equator 3 degrees off the central meridian, northing
  9 328 093.831 m at 84 N on the central meridian) and by an independent restatement with Snyder's series (USGS PP 1395 eqs. 8-9/8-10).
* ``resample_nearest``: every cell CENTRE of the target area (extent = outer edges, row 0 at the top), mapped back to
  lon/lat, takes the value of the swath pixel nearest in 3-D Cartesian distance on the sphere R = 6 370 997 m
  (pyresample ``geometry`` / ``kd_tree``), provided that distance is below ``radius_of_influence`` (10 km, strict as in
  pykdtree's ``distance_upper_bound``); all other cells get ``fill_value``.  Ties (not defined by a kd-tree) go to the
  smallest flat swath index.

Each function cites the reference lines it follows.  Only tests/ and bench legs may import this module.
"""
from __future__ import annotations

import numpy as np

A_WGS84 = 6378137.0
F_WGS84 = 1.0 / 298.257223563
K0, E0 = 0.9996, 500000.0
R_SPHERE = 6370997.0          # pyresample's sphere for the Cartesian neighbour search


def _series():
    n = F_WGS84 / (2.0 - F_WGS84)
    n2, n3, n4, n5, n6 = n ** 2, n ** 3, n ** 4, n ** 5, n ** 6
    A = A_WGS84 / (1 + n) * (1 + n2 / 4 + n4 / 64 + n6 / 256)
    alpha = [n / 2 - 2 * n2 / 3 + 5 * n3 / 16 + 41 * n4 / 180 - 127 * n5 / 288 + 7891 * n6 / 37800,
             13 * n2 / 48 - 3 * n3 / 5 + 557 * n4 / 1440 + 281 * n5 / 630 - 1983433 * n6 / 1935360,
             61 * n3 / 240 - 103 * n4 / 140 + 15061 * n5 / 26880 + 167603 * n6 / 181440,
             49561 * n4 / 161280 - 179 * n5 / 168 + 6601661 * n6 / 7257600,
             34729 * n5 / 80640 - 3418889 * n6 / 1995840,
             212378941 * n6 / 319334400]
    beta = [n / 2 - 2 * n2 / 3 + 37 * n3 / 96 - n4 / 360 - 81 * n5 / 512 + 96199 * n6 / 604800,
            n2 / 48 + n3 / 15 - 437 * n4 / 1440 + 46 * n5 / 105 - 1118711 * n6 / 3870720,
            17 * n3 / 480 - 37 * n4 / 840 - 209 * n5 / 4480 + 5569 * n6 / 90720,
            4397 * n4 / 161280 - 11 * n5 / 504 - 830251 * n6 / 7257600,
            4583 * n5 / 161280 - 108847 * n6 / 3991680,
            20648693 * n6 / 638668800]
    return A, alpha, beta


def utm_zone_ref(lons) -> int:
    """tools.py:20-28: the modal zone of the longitudes wrapped to [-180, 180); scipy.stats.mode returns the
    smallest of equally common values."""
    lons = np.asarray(lons, dtype=np.float64)
    lons = (lons + 180) - np.floor((lons + 180) / 360) * 360 - 180
    zones = (np.floor((lons + 180) / 6) + 1).astype(np.int64).ravel()
    counts = np.bincount(zones, minlength=62)
    return int(np.argmax(counts))             # first maximum = smallest zone among ties


def utm_forward_ref(lats, lons, zone: int):
    """tools.py:30-31, 34: proj(lons, lats) -> (x, y) in metres.  Complex-variable form of the Krueger series."""
    A, alpha, _ = _series()
    e = np.sqrt(F_WGS84 * (2 - F_WGS84))
    phi = np.radians(np.asarray(lats, dtype=np.float64))
    lam = np.radians(np.asarray(lons, dtype=np.float64) - (6.0 * zone - 183.0))
    lam = (lam + np.pi) % (2 * np.pi) - np.pi
    tau = np.tan(phi)
    sigma = np.sinh(e * np.arctanh(e * tau / np.sqrt(1 + tau * tau)))
    taup = tau * np.sqrt(1 + sigma * sigma) - sigma * np.sqrt(1 + tau * tau)
    zeta = np.arctan2(taup, np.cos(lam)) + 1j * np.arcsinh(np.sin(lam) / np.sqrt(taup * taup + np.cos(lam) ** 2))
    out = zeta.copy()
    for j, a in enumerate(alpha, start=1):
        out = out + a * np.sin(2 * j * zeta)
    return E0 + K0 * A * out.imag, K0 * A * out.real


def utm_inverse_ref(x, y, zone: int):
    """tools.py:63-64: proj(x, y, inverse=True) -> (lon, lat) in degrees."""
    A, _, beta = _series()
    e2 = F_WGS84 * (2 - F_WGS84)
    e = np.sqrt(e2)
    zeta = (np.asarray(y, dtype=np.float64) + 1j * (np.asarray(x, dtype=np.float64) - E0)) / (K0 * A)
    zp = zeta.copy()
    for j, b in enumerate(beta, start=1):
        zp = zp - b * np.sin(2 * j * zeta)
    xip, etap = zp.real, zp.imag
    taup = np.sin(xip) / np.sqrt(np.sinh(etap) ** 2 + np.cos(xip) ** 2)
    lam = np.arctan2(np.sinh(etap), np.cos(xip))
    tau = taup.copy()
    for _ in range(5):                         # Newton on tau'(tau) = taup (Karney eqs. 19-21)
        sigma = np.sinh(e * np.arctanh(e * tau / np.sqrt(1 + tau * tau)))
        tpi = tau * np.sqrt(1 + sigma * sigma) - sigma * np.sqrt(1 + tau * tau)
        tau = tau + (taup - tpi) / np.sqrt(1 + tpi * tpi) * (1 + (1 - e2) * tau * tau) / ((1 - e2) * np.sqrt(1 + tau * tau))
    return np.degrees(lam) + (6.0 * zone - 183.0), np.degrees(np.arctan(tau))


def utm_forward_snyder(lats, lons, zone: int):
    """Independent restatement (Snyder, USGS PP 1395, eqs. 3-21, 8-9, 8-10, 8-12..8-15): millimetre-accurate within
    a few degrees of the central meridian.  Only used to cross-check the Krueger coefficients."""
    a, e2 = A_WGS84, F_WGS84 * (2 - F_WGS84)
    ep2 = e2 / (1 - e2)
    phi = np.radians(np.asarray(lats, dtype=np.float64))
    lam = np.radians(np.asarray(lons, dtype=np.float64) - (6.0 * zone - 183.0))
    N = a / np.sqrt(1 - e2 * np.sin(phi) ** 2)
    T, C, Aa = np.tan(phi) ** 2, ep2 * np.cos(phi) ** 2, lam * np.cos(phi)
    M = a * ((1 - e2 / 4 - 3 * e2 ** 2 / 64 - 5 * e2 ** 3 / 256) * phi
             - (3 * e2 / 8 + 3 * e2 ** 2 / 32 + 45 * e2 ** 3 / 1024) * np.sin(2 * phi)
             + (15 * e2 ** 2 / 256 + 45 * e2 ** 3 / 1024) * np.sin(4 * phi) - (35 * e2 ** 3 / 3072) * np.sin(6 * phi))
    x = K0 * N * (Aa + (1 - T + C) * Aa ** 3 / 6 + (5 - 18 * T + T * T + 72 * C - 58 * ep2) * Aa ** 5 / 120)
    y = K0 * (M + N * np.tan(phi) * (Aa ** 2 / 2 + (5 - T + 9 * C + 4 * C * C) * Aa ** 4 / 24
                                     + (61 - 58 * T + T * T + 600 * C - 330 * ep2) * Aa ** 6 / 720))
    return E0 + x, y


def area_from_swath_ref(lats, lons, pixel_size: float):
    """tools.py:33-50: zone, extent (min_x, min_y, max_x, max_y) of the projected swath, grid size."""
    zone = utm_zone_ref(lons)
    x, y = utm_forward_ref(lats, lons, zone)
    extent = (float(np.min(x)), float(np.min(y)), float(np.max(x)), float(np.max(y)))
    x_size = int(np.round((extent[2] - extent[0]) / pixel_size))
    y_size = int(np.round((extent[3] - extent[1]) / pixel_size))
    return zone, extent, x_size, y_size


def cartesian_ref(lats, lons):
    lat, lon = np.radians(np.asarray(lats, dtype=np.float64)), np.radians(np.asarray(lons, dtype=np.float64))
    return np.stack([R_SPHERE * np.cos(lat) * np.cos(lon), R_SPHERE * np.cos(lat) * np.sin(lon),
                     R_SPHERE * np.sin(lat)], axis=-1)


def target_lonlats_ref(zone: int, extent, x_size: int, y_size: int):
    """Cell centres of the area definition (pyresample AreaDefinition.get_lonlats): row 0 is the top row."""
    psx, psy = (extent[2] - extent[0]) / x_size, (extent[3] - extent[1]) / y_size
    xc = extent[0] + (np.arange(x_size) + 0.5) * psx
    yc = extent[3] - (np.arange(y_size) + 0.5) * psy
    xx, yy = np.meshgrid(xc, yc)
    return utm_inverse_ref(xx, yy, zone)


def nearest_index_ref(image_lats, image_lons, zone, extent, x_size, y_size, radius=10000.0, chunk=256):
    """tools.py:52-58 (pr.kd_tree.resample_nearest): flat swath index of the nearest pixel per target cell, -1 where
    no pixel lies within `radius`.  Brute force over ALL swath pixels, squared distances as sums of squared
    differences (the same fp64 operations as the kernel, no cancellation), first minimum = smallest flat index."""
    src = cartesian_ref(np.ravel(image_lats), np.ravel(image_lons))
    valid = (np.abs(np.ravel(image_lats)) <= 90) & (np.abs(np.ravel(image_lons)) <= 180)
    tlon, tlat = target_lonlats_ref(zone, extent, x_size, y_size)
    tgt = cartesian_ref(tlat.ravel(), tlon.ravel())
    out = np.full(tgt.shape[0], -1, dtype=np.int64)
    for a in range(0, tgt.shape[0], chunk):
        t = tgt[a:a + chunk]
        dx = src[None, :, 0] - t[:, None, 0]
        dy = src[None, :, 1] - t[:, None, 1]
        dz = src[None, :, 2] - t[:, None, 2]
        d2 = dx * dx + dy * dy + dz * dz
        d2[:, ~valid] = np.inf
        j = np.argmin(d2, axis=1)
        out[a:a + chunk] = np.where(d2[np.arange(len(j)), j] < radius * radius, j, -1)
    return out.reshape(y_size, x_size)


def resample_image_ref(image, image_lats, image_lons, zone, extent, x_size, y_size, fill_value=-999, radius=10000.0):
    idx = nearest_index_ref(image_lats, image_lons, zone, extent, x_size, y_size, radius)
    flat = np.ravel(image)
    out = np.full(idx.shape, fill_value, dtype=flat.dtype)
    hit = idx >= 0
    out[hit] = flat[idx[hit]]
    return out


MODIS_SPHERE_RADIUS = 6371007.181   # tools.py:124


def modis_grid_latlon_ref(x0, y0, x1, y1, ny, nx):
    """tools.py:103-128 with pyproj's spherical sinusoidal inverse written out (PARITY UNPINNED against pyproj, which is
    absent: PROJ's gn_sinu spherical inverse is phi = y / R, lam = (x / R) / cos(phi); ``+nadgrids=@null`` suppresses
    the datum step, so the sphere's latitude / longitude are the EPSG:4326 output; longitudes wrap into [-180, 180]).
    Anchored on the published MODIS tile grid: tiles are 10 degrees = R * pi / 18 metres
    (tests/test_resample_oracle.py)."""
    xinc, yinc = (x1 - x0) / nx, (y1 - y0) / ny
    x = np.linspace(x0, x0 + xinc * nx, nx)
    y = np.linspace(y0, y0 + yinc * ny, ny)
    xv, yv = np.meshgrid(x, y)
    inv_r = 1.0 / MODIS_SPHERE_RADIUS
    phi = yv * inv_r
    lam = (xv * inv_r) / np.cos(phi)
    wrap = np.abs(lam) > 3.14159265359
    lam = np.where(wrap, (lam + np.pi) - 2 * np.pi * np.floor((lam + np.pi) / (2 * np.pi)) - np.pi, lam)
    return phi * 57.295779513082321, lam * 57.295779513082321
