"""Terminal routing reward of the pin environments (oracle; test infra).

Follows dummy_env_rectangular_pin.py:643-759 (intersections, wirelength),
1282-1460 (centroid, outlier, beam search, routes).  The float64 operation
order is the reference's: numpy ``mean`` / ``linalg.norm`` for centroid and
distances, plain Python arithmetic for the Cramer intersection point.  The beam
search relies on CPython ``set`` iteration order and ``heapq`` tuple ordering
exactly as the reference does, so it uses the real
``set`` / ``heapq`` objects rather than an emulation.
"""
import heapq

import numpy as np


def distance(p, q):
    """euclidean_distance, :1341-1354."""
    return np.linalg.norm(np.array(p) - np.array(q))


def centroid(points):
    """get_centroid, :1282-1294."""
    cx, cy = np.mean(np.array(points), axis=0)
    return (cx, cy)


def outlier(points):
    """pin_outlier, :1326-1339: first farthest point from the centroid."""
    c = centroid(points)
    d = [np.linalg.norm(np.array(p) - np.array(c)) for p in points]
    return points[int(np.argmax(d))]


def segments_cross(s, t):
    """is_intersect, :687-739."""
    (x1, y1), (x2, y2) = s
    (x3, y3), (x4, y4) = t
    if s[0] == t[0] or s[0] == t[1] or s[1] == t[0] or s[1] == t[1]:
        return True
    det = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
    if det == 0:
        return False
    x = ((x1 * y2 - y1 * x2) * (x3 - x4) - (x1 - x2) * (x3 * y4 - y3 * x4)) / det
    y = ((x1 * y2 - y1 * x2) * (y3 - y4) - (y1 - y2) * (x3 * y4 - y3 * x4)) / det
    return bool(
        min(x1, x2) <= x <= max(x1, x2)
        and min(x3, x4) <= x <= max(x3, x4)
        and min(y1, y2) <= y <= max(y1, y2)
        and min(y3, y4) <= y <= max(y3, y4)
    )


def count_crossings(route):
    """find_num_intersection, :663-685: segment pairs of different nets."""
    total = 0
    for a in range(len(route)):
        for b in range(a + 1, len(route)):
            for s in route[a]:
                for t in route[b]:
                    if segments_cross(s, t):
                        total += 1
    return total


def wirelength(route):
    """find_wirelength, :741-759: sequential float64 sum."""
    acc = 0.0
    for net in route:
        for s in net:
            acc += distance(s[0], s[1])
    return acc


def route_centroid(net_positions):
    """route_pins_centroid, :1296-1324."""
    route = []
    for pts in net_positions:
        if len(pts) == 2:
            route.append([(pts[0], pts[1])])
        else:
            c = centroid(pts)
            route.append([(p, c) for p in pts])
    return route


def beam_path(start, points, width):
    """beam_search, :1356-1423 (level-synchronous beam over a heap)."""
    targets = set(points)
    queue = [(0, [start], set())]
    path = [start]
    while queue:
        nxt = []
        for _ in range(min(width, len(queue))):
            prio, path, seen = heapq.heappop(queue)
            here = path[-1]
            if seen == set(targets):
                return path
            near = sorted(targets - seen, key=lambda p: distance(here, p))
            for p in near[:width]:
                heapq.heappush(nxt, (prio + distance(p, here), path + [p], seen | {p}))
        queue = nxt
    return path


def route_beam(net_positions, width):
    """route_pins_beam_search, :1425-1460."""
    route = []
    for pts in net_positions:
        pts = list(pts)
        start = outlier(pts)
        pts.remove(start)
        path = beam_path(start, pts, width)
        route.append([(path[i], path[i + 1]) for i in range(len(path) - 1)])
    return route
