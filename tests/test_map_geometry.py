"""Map geometry (obstacle.py:126-141 through check_condition.py:48-119) pinned
against EXACT rational arithmetic on an adversarial point set (tests/golden/map_adversarial.npz, made by
tests/golden/make_map_adversarial.py; the exact predicates are tests/exact_geometry.py).

Shapely / GEOS is not installed in the build container, so real GEOS never ran.  What is pinned instead: GEOS's
documented semantics -- Polygon.contains(Point) is the strict interior, decided with robust (exact) orientation
predicates; exterior.distance is the Euclidean distance to the closed ring -- evaluated exactly.  The float
even-odd rule / point-segment distance used by the Shapely stand-in, the C oracle and the CUDA kernels must give
the exact answer on every DECISIVE point (farther than 1e-9 m from every ring; all 2000 random points, every
scan-line-through-a-vertex and horizontal-edge case are of this kind); the points within rounding distance of a
ring (on a vertex, on an edge, one ulp off either) are the measure-zero set where a float rule may differ from GEOS,
and the tests state how often it does."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_harness as H

import exact_geometry as X
from helpers import golden

DECISIVE_M = 1e-9          # a point farther than this from every ring is decisive
DIST_ABS_TOL = 4e-12       # float point-segment distance vs exact, map coordinates up to 2e4 m


def _fixture():
    g = golden("map_adversarial")
    decisive = g["exact_distance"] > DECISIVE_M
    return g, decisive


def test_exact_predicates_on_hand_cases():
    sq = [(0.0, 0.0), (4.0, 0.0), (4.0, 4.0), (0.0, 4.0)]
    assert X.contains(sq, 2.0, 2.0) and not X.contains(sq, 5.0, 2.0)
    for x, y in ((0.0, 0.0), (2.0, 0.0), (4.0, 1.0), (0.0, 3.0), (4.0, 4.0)):       # vertices and edge points
        assert X.on_boundary(sq, x, y) and not X.contains(sq, x, y)
    assert not X.contains(sq, np.nextafter(0.0, -1.0), 2.0) and X.contains(sq, np.nextafter(0.0, 1.0), 2.0)
    assert X.ring_distance2(sq, 2.0, 2.0) == 4 and X.ring_distance2(sq, 7.0, 8.0) == 25
    assert X.sqrt_fraction(X.ring_distance2(sq, 7.0, 8.0)) == 5.0
    # a scan line through a vertex of a concave polygon
    poly = [(0.0, 0.0), (2.0, 2.0), (4.0, 0.0), (4.0, 4.0), (0.0, 4.0)]
    assert X.contains(poly, 1.0, 2.0) and X.contains(poly, 3.0, 2.0) and not X.contains(poly, 2.0, 1.0)
    assert X.on_boundary(poly, 2.0, 2.0)


def test_standin_and_oracle_against_exact_arithmetic():
    g, decisive = _fixture()
    polys = [H._Polygon(p) for p in H.MAP_DATA]
    m = O.make_map(H.MAP_DATA)
    lib = O.lib()
    n = len(g["east"])
    st_in = np.zeros(n, np.int32); st_d = np.zeros(n); or_in = np.zeros(n, np.int32); or_d = np.zeros(n)
    for i in range(n):
        pt = H._Point(g["east"][i], g["north"][i])
        st_in[i] = any(p.contains(pt) for p in polys)
        st_d[i] = min(p.exterior.distance(pt) for p in polys)
        or_in[i] = lib.orc_map_contains(C.byref(m), g["north"][i], g["east"][i])
        or_d[i] = lib.orc_map_distance(C.byref(m), g["north"][i], g["east"][i])
    # the fixture was made with the committed stand-in, and the oracle is the same float algorithm: bit-identical
    assert np.array_equal(st_in, g["standin_contains"]) and np.array_equal(st_d, g["standin_distance"])
    assert np.array_equal(or_in, st_in)
    assert np.array_equal(or_d, st_d)
    # decisive points: the float rule gives the exact (= GEOS's documented) answer
    assert decisive.sum() > 2500
    assert np.array_equal(or_in[decisive], g["exact_contains"][decisive])
    assert np.abs(or_d - g["exact_distance"]).max() < DIST_ABS_TOL
    # the measure-zero rest, stated: boundary points a float rule calls "contained" (GEOS: never), and points within
    # 1e-9 m of a ring that it puts on the wrong side
    on = g["on_boundary"] == 1
    near = ~decisive & ~on
    wrong_on = int(or_in[on].sum())
    wrong_near = int((or_in[near] != g["exact_contains"][near]).sum())
    print(f"map geometry: {int(decisive.sum())} decisive points exact; of {int(on.sum())} points exactly on a ring the float "
          f"rule calls {wrong_on} contained (GEOS: 0); of {int(near.sum())} points within {DECISIVE_M} m of a ring "
          f"{wrong_near} fall on the wrong side")
    assert wrong_on == 136 and wrong_near == 187          # the numbers DESIGN.md quotes


@pytest.mark.gpu
@pytest.mark.parametrize("math_mode", ["strict", "fast"])
def test_cuda_geometry_against_exact_arithmetic(math_mode):
    """The env kernel's own routines (culling grid, bounding boxes, even-odd rule, four-corner test, clipped ring
    distance) through shipenv_map_query."""
    import torch
    from ast_sac_b200 import _lib as L
    from ast_sac_b200 import scenarios as S
    g, decisive = _fixture()
    args = S.get_env_args(time_step=4)
    env, _ = S.prepare_multiship_rl_env(args, num_envs=1, math_mode=math_mode)
    n = len(g["east"])
    north = torch.from_numpy(g["north"]).cuda()
    east = torch.from_numpy(g["east"]).cuda()
    contains = torch.zeros(n, dtype=torch.int32, device="cuda")
    square = torch.zeros(n, dtype=torch.int32, device="cuda")
    dist = torch.zeros(n, dtype=torch.float64, device="cuda")
    ship_length = 80.0
    L.check(L.load().shipenv_map_query(env._handle, n, north.data_ptr(), east.data_ptr(), ship_length,
                                       contains.data_ptr(), square.data_ptr(), dist.data_ptr(), env._stream_ptr()))
    torch.cuda.synchronize()
    contains, square, dist = contains.cpu().numpy(), square.cpu().numpy(), dist.cpu().numpy()
    # point test: the oracle's answer everywhere (same float rule; the bounding-box / grid culling is exact), hence
    # the exact answer on every decisive point
    assert np.array_equal(contains, g["standin_contains"])
    assert np.array_equal(contains[decisive], g["exact_contains"][decisive])
    # ring distance inside the reward's clip: the exact distance to 4e-12 m; beyond the clip only "> 1000"
    near = g["exact_distance"] <= 1000.0
    assert np.abs(dist[near] - g["exact_distance"][near]).max() < DIST_ABS_TOL
    assert (dist[~near] > 1000.0).all()
    # four-corner test of the ship square against the oracle's four point tests
    m = O.make_map(H.MAP_DATA)
    lib = O.lib()
    half = ship_length / 2
    for i in range(0, n, 3):
        want = any(lib.orc_map_contains(C.byref(m), g["north"][i] + dn, g["east"][i] + de)
                   for dn in (-half, half) for de in (-half, half))
        assert bool(square[i]) == bool(want), (i, g["north"][i], g["east"][i])
    env.close()


@pytest.mark.gpu
def test_safe_radius_bounds_the_four_corner_test():
    """The quiet steps of the env kernel skip the grounding test of a ship that has moved less than the safe radius
    it read at some earlier position.  The invariant behind that, checked against the kernel's own
    four-corner test: around every point with a positive radius, no position within the radius is grounded -- 64
    positions per point, half of them right at the rim -- and the radius is not vacuous (open water reads hundreds of
    metres, points inside an island read 0)."""
    import torch
    from ast_sac_b200 import _lib as L
    from ast_sac_b200 import scenarios as S
    args = S.get_env_args(time_step=4)
    env, assets = S.prepare_colav_env(args, iw=True, num_envs=1)
    lib = L.load()
    ship_length = max(a.ship_model.l_ship for a in assets)
    gen = torch.Generator().manual_seed(3)
    n = 200_000
    north = (torch.rand(n, generator=gen, dtype=torch.float64) * 10400.0 - 200.0).cuda()     # a margin outside the map too
    east = (torch.rand(n, generator=gen, dtype=torch.float64) * 20400.0 - 200.0).cuda()
    radius = torch.zeros(n, dtype=torch.float32, device="cuda")
    L.check(lib.shipenv_map_safe_radius(env._handle, n, north.data_ptr(), east.data_ptr(), radius.data_ptr(),
                                        env._stream_ptr()))
    torch.cuda.synchronize()
    outside = (north < 0) | (north >= 10000.0) | (east < 0) | (east >= 20000.0)
    assert bool((radius[outside] == 0).all())
    pos = radius > 0
    assert 0.25 < float(pos.double().mean()) < 0.75           # the islands cover ~40 % of the map
    assert float(radius.max()) > 1500.0
    pn, pe, r = north[pos], east[pos], radius[pos].double()
    m = int(pos.sum())
    dummy_i = torch.zeros(m, dtype=torch.int32, device="cuda")
    dummy_d = torch.zeros(m, dtype=torch.float64, device="cuda")
    square = torch.zeros(m, dtype=torch.int32, device="cuda")
    grounded = 0
    for j in range(64):
        ang = torch.rand(m, generator=gen, dtype=torch.float64).cuda() * (2 * np.pi)
        frac = torch.ones(m, dtype=torch.float64, device="cuda") if j % 2 else torch.rand(m, generator=gen, dtype=torch.float64).cuda()
        qn = pn + r * frac * torch.cos(ang)
        qe = pe + r * frac * torch.sin(ang)
        L.check(lib.shipenv_map_query(env._handle, m, qn.data_ptr(), qe.data_ptr(), float(ship_length),
                                      dummy_i.data_ptr(), square.data_ptr(), dummy_d.data_ptr(), env._stream_ptr()))
        torch.cuda.synchronize()
        grounded += int(square.sum())
    assert grounded == 0
    # a point inside an island reads 0 (MAP_DATA polygon 0 contains (east 2000, north 8000))
    one_n = torch.tensor([8000.0], dtype=torch.float64, device="cuda")
    one_e = torch.tensor([2000.0], dtype=torch.float64, device="cuda")
    one_r = torch.ones(1, dtype=torch.float32, device="cuda")
    L.check(lib.shipenv_map_safe_radius(env._handle, 1, one_n.data_ptr(), one_e.data_ptr(), one_r.data_ptr(),
                                        env._stream_ptr()))
    torch.cuda.synchronize()
    assert float(one_r[0]) == 0.0
    env.close()
