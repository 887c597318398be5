"""GPU parity of the UTM projection and the nearest-neighbour resampler (csrc/resample.cu, through the C ABI and the
reference-named ``utm_resampler`` class) against oracle/resample_ref.py.  The neighbour INDEX map must be identical
(integer work: bit-exact); projected coordinates agree to 1e-6 m / 1e-11 degrees (fp64 series, libm vs CUDA math)."""
import numpy as np
import pytest
import torch

from oracle import resample_ref as rr
from tests.resample_data import swath

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from kcl_ltss_bioatm_b200.ops import CudaOps

    return CudaOps()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(DEV).reshape(-1)


def test_utm_forward_inverse_and_zone_histogram(ops):
    rng = np.random.default_rng(3)
    lat, lon = rng.uniform(-80, 84, 50000), -69 + rng.uniform(-12, 12, 50000)
    x, y = torch.empty(50000, dtype=torch.float64, device=DEV), torch.empty(50000, dtype=torch.float64, device=DEV)
    ops.utm_forward(dev(lat), dev(lon), 19, x, y)
    xr, yr = rr.utm_forward_ref(lat, lon, 19)
    assert np.abs(x.cpu().numpy() - xr).max() < 1e-6 and np.abs(y.cpu().numpy() - yr).max() < 1e-6
    la, lo = torch.empty_like(x), torch.empty_like(x)
    ops.utm_inverse(x, y, 19, la, lo)
    assert np.abs(la.cpu().numpy() - lat).max() < 1e-11 and np.abs(lo.cpu().numpy() - lon).max() < 1e-11
    lons = np.concatenate([rng.uniform(-200, 200, 10000), [179.0, 181.0, 182.0, -180.0, 180.0]])
    hist = torch.empty(64, dtype=torch.int32, device=DEV)
    ops.utm_zone_histogram(dev(lons), hist)
    w = (lons + 180) - np.floor((lons + 180) / 360) * 360 - 180
    assert np.array_equal(hist.cpu().numpy(), np.bincount((np.floor((w + 180) / 6) + 1).astype(int), minlength=64))


@pytest.mark.parametrize("h,w,lat0,lon0,step,px,rot", [
    (60, 80, 45.0, 10.0, 1.0, 750.0, 12.0),        # target finer than the swath
    (90, 70, -12.5, -63.0, 1.0, 2500.0, -35.0),    # coarser target, southern hemisphere (negative northings)
    (33, 47, 68.0, 27.9, 0.75, 1000.0, 60.0),      # high latitude, ragged sizes
    (40, 40, 5.0, 11.9, 3.0, 1000.0, 0.0),         # sparse swath (3 km) straddling a zone boundary: holes get filled
])
def test_resampler_matches_the_oracle(ops, h, w, lat0, lon0, step, px, rot):
    from kcl_ltss_bioatm_b200.resample import utm_resampler

    lat, lon = swath(h, w, lat0, lon0, step_km=step, rot_deg=rot, seed=h, jitter=0.2)
    zone, extent, xs, ys = rr.area_from_swath_ref(lat, lon, px)
    rs = utm_resampler(lat, lon, px, device=DEV, ops=ops)
    assert rs.zone == zone and (rs.x_size, rs.y_size) == (xs, ys)
    assert np.allclose(rs.extent, extent, rtol=0, atol=1e-6)
    img = np.random.default_rng(1).normal(size=(h, w))
    idx_ref = rr.nearest_index_ref(lat, lon, zone, rs.extent, xs, ys)      # the same extent on both sides
    idx = rs.neighbour_index(lat, lon).cpu().numpy()
    assert np.array_equal(idx, idx_ref), f"{(idx != idx_ref).sum()} of {idx.size} cells differ"
    for dt in (np.float64, np.float32):
        got = rs.resample_image(img.astype(dt), lat, lon, fill_value=-999)
        ref = rr.resample_image_ref(img.astype(dt), lat, lon, zone, rs.extent, xs, ys, fill_value=-999)
        assert got.dtype == dt and np.array_equal(got, ref)
    pts = rs.resample_points_to_utm(lat[0, :5], lon[0, :5])
    xr, yr = rr.utm_forward_ref(lat[0, :5], lon[0, :5], zone)
    assert np.allclose(np.array(pts), np.stack([xr, yr], 1), rtol=0, atol=1e-6)
    lo, la = rs.resample_point_to_geo(pts[2][1], pts[2][0])
    assert abs(lo - lon[0, 2]) < 1e-10 and abs(la - lat[0, 2]) < 1e-10


def test_image_on_another_geometry_invalid_pixels_and_radius(ops):
    """resample_image(image, image_lats, image_lons) takes the image's OWN geolocation (tools.py:52-53): an image that
    covers only part of the area leaves cells further than 10 km at fill_value; swath pixels with invalid coordinates
    are ignored."""
    from kcl_ltss_bioatm_b200.resample import utm_resampler

    lat, lon = swath(80, 80, 40.0, -3.0, step_km=1.0, rot_deg=5.0)
    rs = utm_resampler(lat, lon, 1000.0, device=DEV, ops=ops)
    sub_lat, sub_lon = lat[10:40, 20:50].copy(), lon[10:40, 20:50].copy()
    sub_lat[3, 4], sub_lon[7, 7] = 1e30, -999.0                          # invalid geolocation
    img = np.random.default_rng(2).uniform(0, 2, size=sub_lat.shape)
    got = rs.resample_image(img, sub_lat, sub_lon, fill_value=-999)
    ref = rr.resample_image_ref(img, sub_lat, sub_lon, rs.zone, rs.extent, rs.x_size, rs.y_size, fill_value=-999)
    assert np.array_equal(got, ref)
    assert (got == -999).mean() > 0.3 and (got != -999).mean() > 0.2


@pytest.mark.parametrize("h,v,ny,nx", [(18, 4, 1200, 1200), (0, 8, 64, 48), (35, 9, 1, 7), (17, 0, 33, 1), (20, 11, 240, 240)])
def test_modis_grid_latlon_matches_the_oracle(ops, h, v, ny, nx):
    """Geolocation half of read_modis_aod (tools.py:97-128) for MODIS tiles incl. the outermost ones (wrapped
    longitudes) and degenerate shapes; fp64, operation order of the oracle; cos differs from libm's by an ulp."""
    from kcl_ltss_bioatm_b200.resample import modis_grid_latlon
    tile = rr.MODIS_SPHERE_RADIUS * np.pi / 18
    x0, y0 = (h - 18) * tile, (9 - v) * tile
    lat, lon = modis_grid_latlon(x0, y0, x0 + tile, y0 - tile, ny, nx, device="cuda", ops=ops)
    rlat, rlon = rr.modis_grid_latlon_ref(x0, y0, x0 + tile, y0 - tile, ny, nx)
    assert lat.dtype == torch.float64 and tuple(lat.shape) == (ny, nx)
    assert np.array_equal(lat.cpu().numpy(), rlat)                                  # no transcendental: bit-equal
    # longitude: x / (R cos(lat)) before wrapping; an ulp of cos is amplified by that magnitude, and towards the pole
    # (top / bottom tile rows, fill area outside the projection) it grows without bound -- compare where it is sane
    xv = np.linspace(x0, x0 + tile, nx)[None, :] / rr.MODIS_SPHERE_RADIUS
    raw = np.degrees(np.abs(xv / np.cos(np.radians(rlat))))
    sane = (raw < 3600.0) & (np.abs(np.abs(rlon) - 180.0) > 1e-6)
    assert sane.mean() > 0.8
    dlon = np.abs(lon.cpu().numpy() - rlon)
    assert (dlon[sane] <= 4e-13 * np.maximum(1.0, raw[sane])).all()
    from src.features import tools
    assert tools.modis_grid_latlon is modis_grid_latlon
