"""Exact-arithmetic (fractions.Fraction) restatement of the two Shapely / GEOS operations the simulator uses on
the map polygons (obstacle.py:126-141):

  Polygon.contains(Point)        -- True iff the point lies in the polygon's INTERIOR (a point on the boundary is
                                    not contained; DE-9IM `T*****FF*`, the documented semantics GEOS implements
                                    with robust orientation predicates);
  polygon.exterior.distance(pt)  -- the Euclidean distance from the point to the closed ring.

Every float is converted to a Fraction exactly, so the answers below are the mathematically exact ones for the
given double-precision coordinates: the yardstick for the float implementations (the Shapely stand-in of
oracle/ref_harness.py, the C oracle, the CUDA kernels).  Test infrastructure only.
"""
from __future__ import annotations

from fractions import Fraction as F


def _fr(v):
    return F(float(v))


def on_boundary(poly, x, y) -> bool:
    """Is (x, y) exactly on one of the ring's segments."""
    x, y = _fr(x), _fr(y)
    n = len(poly)
    for i in range(n):
        ax, ay = _fr(poly[i][0]), _fr(poly[i][1])
        bx, by = _fr(poly[(i + 1) % n][0]), _fr(poly[(i + 1) % n][1])
        cross = (bx - ax) * (y - ay) - (by - ay) * (x - ax)
        if cross == 0 and min(ax, bx) <= x <= max(ax, bx) and min(ay, by) <= y <= max(ay, by):
            return True
    return False


def contains(poly, x, y) -> bool:
    """Exact Polygon.contains(Point(x, y)): strict interior."""
    if on_boundary(poly, x, y):
        return False
    x, y = _fr(x), _fr(y)
    n = len(poly)
    inside = False
    for i in range(n):
        xi, yi = _fr(poly[i][0]), _fr(poly[i][1])
        xj, yj = _fr(poly[i - 1][0]), _fr(poly[i - 1][1])
        if (yi > y) != (yj > y):
            if x < (xj - xi) * (y - yi) / (yj - yi) + xi:
                inside = not inside
    return inside


def ring_distance2(poly, x, y) -> F:
    """Exact squared distance from (x, y) to the closed ring."""
    x, y = _fr(x), _fr(y)
    n = len(poly)
    best = None
    for i in range(n):
        ax, ay = _fr(poly[i][0]), _fr(poly[i][1])
        bx, by = _fr(poly[(i + 1) % n][0]), _fr(poly[(i + 1) % n][1])
        dx, dy = bx - ax, by - ay
        l2 = dx * dx + dy * dy
        t = F(0) if l2 == 0 else ((x - ax) * dx + (y - ay) * dy) / l2
        t = min(F(1), max(F(0), t))
        cx, cy = ax + t * dx, ay + t * dy
        d2 = (x - cx) ** 2 + (y - cy) ** 2
        if best is None or d2 < best:
            best = d2
    return best


def map_contains(polys, x, y) -> bool:
    return any(contains(p, x, y) for p in polys)


def map_on_boundary(polys, x, y) -> bool:
    return any(on_boundary(p, x, y) for p in polys)


def map_distance2(polys, x, y) -> F:
    return min(ring_distance2(p, x, y) for p in polys)


def sqrt_fraction(q: F) -> float:
    """Correctly rounded-ish float sqrt of a non-negative Fraction (integer sqrt on a scaled numerator: the error
    is far below one ulp of the result)."""
    from math import isqrt
    if q == 0:
        return 0.0
    scale = 1 << 200
    v = isqrt((q.numerator * scale * scale) // q.denominator)
    return float(F(v, scale))
