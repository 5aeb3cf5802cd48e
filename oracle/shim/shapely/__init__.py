"""TEST-ONLY exact-geometry stand-in for the 8 `shapely` names the reference LIDAR env uses.

`shapely`/GEOS is an un-vendored, un-pinned third-party dependency of the reference
(`/root/reference/pyproject.toml:20`) and is absent from this image.  This module lets the
reference's own `ap_gym/envs/lidar_localization2d.py` run UNMODIFIED so that its python logic
(pose integration, wall slide, contact-cell rule, odometry alias, TimeLimit, autoreset) pins the
oracle.  The geometry itself -- `LineString([p, t]).intersection(union_all(boxes))` -- is a
restatement of what GEOS OverlayNG computes in exact arithmetic for a segment against a union of
closed unit squares; it is NOT GEOS.  LIDAR parity against real GEOS therefore stays UNPINNED
(SURVEY.md section 8c).

Geometry model (shared with oracle/lidar_oracle.c and the CUDA kernel):
  * wall cells are CLOSED unit squares [cx, cx+1] x [cy, cy+1];
  * the segment p->t is clipped against every wall cell near its bounding box with Liang-Barsky in
    float64, every line parameter being computed as (grid_line - p) * (1 / d)  (the reciprocal of each
    direction component is taken once per segment; the same expression the DDA uses, so both produce
    bit-identical parameters);
  * touching/overlapping parameter intervals are merged; a merged interval with t0 == t1 is a Point,
    otherwise a LineString; pieces are ordered and oriented along the ray;
  * the entry vertex of a piece is p + t0 * d with the coordinate(s) on the crossed grid line(s)
    set to the exact integer (GEOS computes crossings in double-double, so they come out integral).
Result typing: none -> empty LineString; one piece -> Point | LineString; several pieces ->
MultiPoint | MultiLineString | GeometryCollection by type mix.

Not product code: nothing under `active-perception-gym_b200/` imports it.
"""
from __future__ import annotations

import math

import numpy as np


class _G:
    is_empty = False


class _XY:
    def __init__(self, xs, ys):
        self._d = (list(xs), list(ys))

    def __getitem__(self, i):
        return self._d[i]


class LineString(_G):
    def __init__(self, coords=()):
        self.c = [(float(a), float(b)) for a, b in coords]
        self.is_empty = len(self.c) == 0

    @property
    def xy(self):
        return _XY([p[0] for p in self.c], [p[1] for p in self.c])

    def intersection(self, other):
        return other._clip(self)


class Point(_G):
    def __init__(self, x, y):
        self.c = [(float(x), float(y))]

    @property
    def xy(self):
        return _XY([self.c[0][0]], [self.c[0][1]])


class _Multi(_G):
    def __init__(self, geoms):
        self.geoms = list(geoms)


class MultiPoint(_Multi):
    pass


class MultiLineString(_Multi):
    pass


class GeometryCollection(_Multi):
    pass


class _Cells(_G):
    """Union of closed unit squares keyed by integer (x, y)."""

    def __init__(self, cells):
        self.cells = set(cells)

    def _clip(self, line):
        (px, py), (tx, ty) = line.c
        dx, dy = tx - px, ty - py
        x0, x1 = sorted((px, tx))
        y0, y1 = sorted((py, ty))
        pieces = []  # (t0, t1, entry_x, entry_y)
        for cx in range(int(math.floor(x0)) - 1, int(math.floor(x1)) + 2):
            for cy in range(int(math.floor(y0)) - 1, int(math.floor(y1)) + 2):
                if (cx, cy) not in self.cells:
                    continue
                t0, t1, ok = 0.0, 1.0, True
                par = []  # per axis: (a, ga, b, gb)
                for p, d, lo, hi in ((px, dx, cx, cx + 1), (py, dy, cy, cy + 1)):
                    if d == 0.0:
                        if p < lo or p > hi:
                            ok = False
                            break
                        par.append(None)
                    else:
                        inv = 1.0 / d
                        a, b = (lo - p) * inv, (hi - p) * inv
                        ga, gb = float(lo), float(hi)
                        if a > b:
                            a, b, ga, gb = b, a, gb, ga
                        par.append((a, ga, b, gb))
                        t0, t1 = max(t0, a), min(t1, b)
                        if t0 > t1:
                            ok = False
                            break
                if not ok:
                    continue
                e = [px + t0 * dx, py + t0 * dy]
                for k, q in enumerate(par):
                    if q is None:
                        continue
                    a, ga, b, gb = q
                    if a == t0:
                        e[k] = ga  # entered through this grid line
                    elif t0 == t1 and b == t0:
                        e[k] = gb  # point touch that also lies on the exit line
                pieces.append((t0, t1, e[0], e[1]))
        if not pieces:
            return LineString()
        pieces.sort(key=lambda q: (q[0], q[1]))
        merged = [list(pieces[0])]
        for t0, t1, ex, ey in pieces[1:]:
            if t0 <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], t1)
            else:
                merged.append([t0, t1, ex, ey])
        geoms = []
        for t0, t1, ex, ey in merged:
            if t0 == t1:
                geoms.append(Point(ex, ey))
            else:
                geoms.append(LineString([(ex, ey), (px + t1 * dx, py + t1 * dy)]))
        if len(geoms) == 1:
            return geoms[0]
        if all(isinstance(g, Point) for g in geoms):
            return MultiPoint(geoms)
        if all(isinstance(g, LineString) for g in geoms):
            return MultiLineString(geoms)
        return GeometryCollection(geoms)


def box(x0, y0, x1, y1):
    return (int(x0), int(y0))


def union_all(cells):
    return _Cells(cells)
