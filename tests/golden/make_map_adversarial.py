"""Adversarial fixture for the map geometry: tests/golden/map_adversarial.npz.

    python tests/golden/make_map_adversarial.py

Shapely / GEOS is not installed here, so the env-level goldens were made with the stand-in of
oracle/ref_harness.py.  This fixture pins the stand-in, the oracle and the CUDA kernels against EXACT rational
arithmetic (tests/exact_geometry.py) on the points where a float implementation can disagree with GEOS's documented
semantics: polygon vertices, points on edges, points one ulp off an edge, points on the scan line through a vertex,
points on the line of a horizontal edge -- for all six map polygons (run/env_setup.py:145-152) -- plus seeded random
points.  Stored per point: the coordinates, whether it lies exactly on the boundary, the exact contains() answer, the
exact ring distance, and the stand-in's answers.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import exact_geometry as X  # noqa: E402
from oracle import ref_harness as H  # noqa: E402


def points():
    rng = np.random.default_rng(16)
    pts, kind = [], []

    def add(x, y, k):
        pts.append((float(x), float(y)))
        kind.append(k)

    def ulp_ring(x, y, k):
        for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1)):
            xx = np.nextafter(x, np.inf * dx) if dx else x
            yy = np.nextafter(y, np.inf * dy) if dy else y
            add(xx, yy, k)

    for poly in H.MAP_DATA:
        n = len(poly)
        for i in range(n):
            ax, ay = map(float, poly[i])
            bx, by = map(float, poly[(i + 1) % n])
            add(ax, ay, 0)                                      # 0: a vertex
            ulp_ring(ax, ay, 1)                                 # 1: one ulp around a vertex
            for t in (0.5, 0.25, 1.0 / 3.0, 0.9):               # 2: on / next to an edge (t = 1/3 is rounded)
                x, y = ax + t * (bx - ax), ay + t * (by - ay)
                add(x, y, 2)
                ulp_ring(x, y, 3)                               # 3: one ulp around a point of an edge
            for d in (1e-9, 1e-3, 1.0, 100.0, 5000.0, 30000.0):  # 4: on the scan line through a vertex
                add(ax - d, ay, 4)
                add(ax + d, ay, 4)
            if ay == by:                                        # 5: the line of a horizontal edge
                lo, hi = min(ax, bx), max(ax, bx)
                for x in (lo - 10.0, lo, 0.5 * (lo + hi), hi, hi + 10.0):
                    for y in (ay, np.nextafter(ay, np.inf), np.nextafter(ay, -np.inf)):
                        add(x, y, 5)
    for _ in range(2000):                                       # 6: seeded random points over the map and its rim
        add(rng.uniform(-500, 20500), rng.uniform(-500, 10500), 6)
    return np.array(pts), np.array(kind, dtype=np.int32)


def main():
    pts, kind = points()
    polys = [[(float(e), float(n)) for e, n in p] for p in H.MAP_DATA]
    standin = [H._Polygon(p) for p in H.MAP_DATA]
    n = len(pts)
    boundary = np.zeros(n, np.int32); exact_in = np.zeros(n, np.int32); exact_d = np.zeros(n)
    st_in = np.zeros(n, np.int32); st_d = np.zeros(n)
    for i, (x, y) in enumerate(pts):
        boundary[i] = X.map_on_boundary(polys, x, y)
        exact_in[i] = X.map_contains(polys, x, y)
        exact_d[i] = X.sqrt_fraction(X.map_distance2(polys, x, y))
        pt = H._Point(x, y)
        st_in[i] = any(p.contains(pt) for p in standin)
        st_d[i] = min(p.exterior.distance(pt) for p in standin)
    np.savez_compressed(os.path.join(HERE, "map_adversarial.npz"), east=pts[:, 0], north=pts[:, 1], kind=kind,
                        on_boundary=boundary, exact_contains=exact_in, exact_distance=exact_d,
                        standin_contains=st_in, standin_distance=st_d)
    off = boundary == 0
    print(f"{n} points, {boundary.sum()} exactly on a boundary")
    print("stand-in vs exact contains(), points NOT on a boundary: mismatches", int((st_in[off] != exact_in[off]).sum()),
          "by kind", {int(k): int(((st_in != exact_in) & off & (kind == k)).sum()) for k in np.unique(kind)})
    print("stand-in contains() == True on boundary points (GEOS: False):", int(st_in[~off].sum()), "of", int((~off).sum()))
    rel = np.abs(st_d - exact_d) / np.maximum(exact_d, 1e-300)
    big = exact_d > 1e-6
    print("ring distance: max relative error of the stand-in where d > 1e-6 m:", rel[big].max(),
          "; max absolute error where d <= 1e-6 m:", np.abs(st_d - exact_d)[~big].max() if (~big).any() else 0.0)


if __name__ == "__main__":
    main()
