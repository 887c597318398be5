"""CPU checks that pin oracle/resample_ref.py (the restatement of pyproj's UTM and pyresample's nearest-neighbour
resampling used by /root/reference/src/features/tools.py:9-64; both libraries are absent here)."""
import numpy as np

from oracle import resample_ref as rr
from tests.resample_data import swath


def test_utm_published_anchor_values():
    # the textbook UTM zone-edge constants (WGS84): at the equator, 3 degrees from the central meridian
    x, y = rr.utm_forward_ref([0.0, 0.0], [6.0, 0.0], 31)            # zone 31: central meridian 3 E
    assert abs(x[0] - 833978.5569) < 1e-3 and abs(x[1] - 166021.4431) < 1e-3 and np.all(np.abs(y) < 1e-9)
    x, y = rr.utm_forward_ref([84.0], [3.0], 31)                      # northern limit of UTM on the central meridian
    assert abs(x[0] - 500000.0) < 1e-9 and abs(y[0] - 9328093.831) < 1e-2
    x, y = rr.utm_forward_ref([-30.0], [3.0], 31)                     # no false northing without `south`
    assert y[0] < 0


def test_krueger_series_agrees_with_snyder_series_inside_the_zone():
    rng = np.random.default_rng(0)
    lat, lon = rng.uniform(-80, 84, 20000), 3 + rng.uniform(-3, 3, 20000)
    xk, yk = rr.utm_forward_ref(lat, lon, 31)
    xs, ys = rr.utm_forward_snyder(lat, lon, 31)
    assert np.abs(xk - xs).max() < 2e-3 and np.abs(yk - ys).max() < 2e-3   # Snyder's own truncation is ~1 mm


def test_utm_round_trip_far_outside_the_zone():
    rng = np.random.default_rng(1)
    lat, lon = rng.uniform(-80, 84, 20000), -69 + rng.uniform(-12, 12, 20000)   # zone 19 +- 12 degrees
    x, y = rr.utm_forward_ref(lat, lon, 19)
    lo, la = rr.utm_inverse_ref(x, y, 19)
    assert np.abs(lo - lon).max() < 1e-11 and np.abs(la - lat).max() < 1e-11


def test_zone_is_the_modal_zone_with_wraparound():
    assert rr.utm_zone_ref([2.9, 3.1, 4.0, 6.1]) == 31
    assert rr.utm_zone_ref([179.0, 181.0, 182.0]) == 1                # 181 E wraps to 179 W = zone 1
    assert rr.utm_zone_ref([-3.0, 3.0]) == 30                         # tie -> smallest zone (scipy.stats.mode)


def test_area_definition_follows_the_reference():
    lat, lon = swath(40, 50, 45.0, 10.0, step_km=1.0)
    zone, extent, xs, ys = rr.area_from_swath_ref(lat, lon, 750.0)
    x, y = rr.utm_forward_ref(lat, lon, zone)
    assert zone == 32 and extent == (x.min(), y.min(), x.max(), y.max())
    assert xs == int(np.round((x.max() - x.min()) / 750.0)) and ys == int(np.round((y.max() - y.min()) / 750.0))


def test_nearest_resampling_of_a_grid_onto_itself_is_the_identity():
    """Swath = the cell centres of the target area itself -> every cell picks its own pixel; cells further than the
    radius from any pixel are filled."""
    zone, extent, xs, ys = 33, (400000.0, 5000000.0, 430000.0, 5020000.0), 30, 20
    tlon, tlat = rr.target_lonlats_ref(zone, extent, xs, ys)
    idx = rr.nearest_index_ref(tlat, tlon, zone, extent, xs, ys)
    assert np.array_equal(idx, np.arange(xs * ys).reshape(ys, xs))
    img = np.arange(xs * ys, dtype=np.float64).reshape(ys, xs)
    assert np.array_equal(rr.resample_image_ref(img, tlat, tlon, zone, extent, xs, ys), img)
    # a swath covering only the western 5 km: cells more than 10 km from it get the fill value
    keep = slice(0, 5)
    out = rr.resample_image_ref(img[:, keep], tlat[:, keep], tlon[:, keep], zone, extent, xs, ys, fill_value=-999)
    assert np.array_equal(out[:, :5], img[:, :5])
    assert np.all(out[:, 15:] == -999) and np.all(out[:, 5:14] == img[:, 4:5])   # nearest column is the 5th


def test_modis_sinusoidal_grid_against_the_published_tile_grid():
    """MODIS sinusoidal tiles are 10 x 10 degrees at the equator: tile size = R * pi / 18 = 1111950.5197 m (published
    grid constant), tile h18v04's upper-left corner is (0 m, 5559752.598 m) = (lon 0, lat 50 N), h17v03's lower-right
    the same point.  The oracle must put the grid's first sample on the corner and follow x / (R cos(lat))."""
    tile = rr.MODIS_SPHERE_RADIUS * np.pi / 18
    assert abs(tile - 1111950.5197) < 1e-3
    x0, y0 = 0.0, 5 * tile                                   # h18v04
    lat, lon = rr.modis_grid_latlon_ref(x0, y0, x0 + tile, y0 - tile, 1200, 1200)
    assert lat.shape == lon.shape == (1200, 1200)
    assert abs(lat[0, 0] - 50.0) < 1e-9 and abs(lon[0, 0]) < 1e-12
    assert abs(lat[-1, 0] - 40.0) < 1e-9                                          # linspace includes the far edge (tools.py:118-119)
    assert abs(lon[-1, -1] - 10.0 / np.cos(np.radians(40.0))) < 1e-9             # 10 degrees of x at latitude 40
    assert abs(lon[0, -1] - 10.0 / np.cos(np.radians(50.0))) < 1e-9
    assert np.all(np.diff(lat[:, 0]) < 0) and np.all(np.diff(lon[0]) > 0)
    # beyond the edge of the projection (fill area of the outermost tiles) longitudes wrap like PROJ's adjlon
    lat2, lon2 = rr.modis_grid_latlon_ref(17 * tile, 8 * tile, 18 * tile, 7 * tile, 5, 5)
    assert np.all(np.abs(lon2) <= 180.0 + 1e-9)
