"""TEST INFRASTRUCTURE ONLY — the subset of shapely.geometry the reference's Python uses, in plain Python (not GEOS)."""
import math


def _xy(p):
    if isinstance(p, Point):
        return p.x, p.y
    return float(p[0]), float(p[1])


def _seg_point(ax, ay, bx, by, px, py):
    """(distance, parameter t in [0,1]) of the point of segment AB nearest to P"""
    dx, dy = bx - ax, by - ay
    d2 = dx * dx + dy * dy
    t = 0.0 if d2 == 0.0 else max(0.0, min(1.0, ((px - ax) * dx + (py - ay) * dy) / d2))
    qx, qy = ax + t * dx, ay + t * dy
    return math.hypot(px - qx, py - qy), t


def _orient(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def _segments_intersect(a, b, c, d):
    o1, o2 = _orient(*a, *b, *c), _orient(*a, *b, *d)
    o3, o4 = _orient(*c, *d, *a), _orient(*c, *d, *b)
    if ((o1 > 0) != (o2 > 0)) and ((o3 > 0) != (o4 > 0)) and o1 != 0 and o2 != 0 and o3 != 0 and o4 != 0:
        return True
    def on(p, q, r):   # r on segment pq (collinear assumed)
        return min(p[0], q[0]) <= r[0] <= max(p[0], q[0]) and min(p[1], q[1]) <= r[1] <= max(p[1], q[1])
    return (o1 == 0 and on(a, b, c)) or (o2 == 0 and on(a, b, d)) or (o3 == 0 and on(c, d, a)) or (o4 == 0 and on(c, d, b))


def _seg_seg_distance(a, b, c, d):
    if _segments_intersect(a, b, c, d):
        return 0.0
    return min(_seg_point(*a, *b, *c)[0], _seg_point(*a, *b, *d)[0], _seg_point(*c, *d, *a)[0], _seg_point(*c, *d, *b)[0])


class Point:
    def __init__(self, *args):
        if len(args) == 1:
            self.x, self.y = _xy(args[0])
        else:
            self.x, self.y = float(args[0]), float(args[1])

    @property
    def coords(self):
        return [(self.x, self.y)]

    def distance(self, other):
        if isinstance(other, Point):
            return math.hypot(self.x - other.x, self.y - other.y)
        return other.distance(self)

    def within(self, polygon):
        return polygon.contains_point(self.x, self.y)


class LineString:
    def __init__(self, coords):
        self.coords = [(float(p[0]), float(p[1])) for p in coords]

    @property
    def length(self):
        return sum(math.hypot(b[0] - a[0], b[1] - a[1]) for a, b in zip(self.coords, self.coords[1:]))

    def segments(self):
        return list(zip(self.coords, self.coords[1:]))

    def project(self, point):
        px, py = _xy(point)
        best, best_s, run = None, 0.0, 0.0
        for a, b in self.segments():
            d, t = _seg_point(*a, *b, px, py)
            seg = math.hypot(b[0] - a[0], b[1] - a[1])
            if best is None or d < best:
                best, best_s = d, run + t * seg
            run += seg
        return best_s

    def interpolate(self, s):
        run = 0.0
        for a, b in self.segments():
            seg = math.hypot(b[0] - a[0], b[1] - a[1])
            if s <= run + seg or (a, b) == self.segments()[-1]:
                t = 0.0 if seg == 0 else (s - run) / seg
                return Point(a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]))
            run += seg

    def distance(self, other):
        if isinstance(other, Point):
            return min(_seg_point(*a, *b, other.x, other.y)[0] for a, b in self.segments())
        if isinstance(other, LineString):
            return min(_seg_seg_distance(a, b, c, d) for a, b in self.segments() for c, d in other.segments())
        return other.distance(self)

    def simplify(self, tolerance, preserve_topology=True):
        def dp(pts):
            if len(pts) < 3:
                return pts
            a, b = pts[0], pts[-1]
            imax, dmax = 0, -1.0
            for i in range(1, len(pts) - 1):
                d = _seg_point(*a, *b, *pts[i])[0]
                if d > dmax:
                    imax, dmax = i, d
            if dmax > tolerance:
                return dp(pts[:imax + 1])[:-1] + dp(pts[imax:])
            return [a, b]
        return LineString(dp(self.coords))


class Polygon:
    def __init__(self, shell):
        pts = [(float(p[0]), float(p[1])) for p in shell]
        if pts and pts[0] == pts[-1]:
            pts = pts[:-1]
        self.shell = pts

    def edges(self):
        return [(self.shell[i], self.shell[(i + 1) % len(self.shell)]) for i in range(len(self.shell))]

    def contains_point(self, x, y):
        inside = False
        for (x1, y1), (x2, y2) in self.edges():
            if _seg_point(x1, y1, x2, y2, x, y)[0] == 0.0:
                return False          # on the boundary: not `within`
            if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1:
                inside = not inside
        return inside

    def distance(self, other):
        if isinstance(other, Point):
            return 0.0 if self.contains_point(other.x, other.y) else min(_seg_point(*a, *b, other.x, other.y)[0] for a, b in self.edges())
        if isinstance(other, LineString):
            if any(self.contains_point(*p) for p in other.coords):
                return 0.0
            return min(_seg_seg_distance(a, b, c, d) for a, b in self.edges() for c, d in other.segments())
        raise TypeError("unsupported")
