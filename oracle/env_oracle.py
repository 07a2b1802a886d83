"""TEST INFRASTRUCTURE ONLY (imported by tests/ and __graft_entry__.smoke(); nothing under the package may import it).

CPU restatement, in plain Python floats (IEEE double, no fused multiply-add — the arithmetic the CUDA kernels of
csrc/env_kernels.cuh are compiled to with -fmad=false), of the OUTER environment's geometry path of the reference:

    environment.py:40-44    Map.corridor_match
    environment.py:85-101   Environment.match_agents_on_corridor
    environment.py:135-143  agents_in_corridor_by_arc_length
    environment.py:145-156  sorted_agents_on_corridor_within_range_of_vehicle
    environment.py:166-195  front_agent / following_agent
    environment.py:197-251  neighbor_around_agent
    type.py:180-206         Corridor.distance_to_corridor_end / distance_to_border (environment.py:113-115)
    type.py:255-317         LineSimplified (project, get_point_at_length, get_yaw_at_length, is_left_of,
                            signed_distance, to_frenet_frame)
    utility.py:83-98        offset_point, step_along_path
    utility.py:111-126      extend_line, extend_line_both_sides

The geometry the reference takes from shapely (LineString.project / distance / simplify, Point.within) is a third-party
dependency absent from /root/reference and from this image (shapely, any version — the reference pins none, README.md);
its published algorithms are restated here: nearest point of each segment, first segment wins ties; even-odd ray
crossing with boundary points outside; Douglas-Peucker.

Parity pin: tests/golden/env_*.json.gz are written by the reference's OWN environment.py / type.py / mcs_wrapper.py
(tests/golden/make_env_golden.py; shapely replaced by tests/refshim): lanes, per-lane orders and neighbour ids are
compared exactly, arc lengths / distances / Frenet coordinates within 1e-9 (the reference evaluates hypot() and
arctan2/sin where this file evaluates sqrt(dx*dx+dy*dy) and a cross product; see side_of()).
"""
import math

NEIGHBORS = 10
LEFT_FF, LEFT_F, LEFT_B, LEFT_BB, FRONT, BACK, RIGHT_FF, RIGHT_F, RIGHT_B, RIGHT_BB = range(10)


# ---- shapely's primitives -------------------------------------------------------------------------------------------

def seg_point(ax, ay, bx, by, px, py):
    """(distance, t, squared segment length) of the point of segment AB nearest to P"""
    dx, dy = bx - ax, by - ay
    d2 = dx * dx + dy * dy
    if d2 == 0.0:
        t = 0.0
    else:
        t = ((px - ax) * dx + (py - ay) * dy) / d2
        t = t if t < 1.0 else 1.0
        t = t if t > 0.0 else 0.0
    qx, qy = ax + t * dx, ay + t * dy
    ex, ey = px - qx, py - qy
    return math.sqrt(ex * ex + ey * ey), t, d2


def project_and_distance(line, px, py):
    """LineString.project + LineString.distance: arc length of the nearest point (first segment wins ties), distance"""
    best, best_s, run = None, 0.0, 0.0
    for (ax, ay), (bx, by) in zip(line, line[1:]):
        d, t, d2 = seg_point(ax, ay, bx, by, px, py)
        seg = math.sqrt(d2)
        if best is None or d < best:
            best, best_s = d, run + t * seg
        run = run + seg
    return best_s, best


def line_length(line):
    run = 0.0
    for (ax, ay), (bx, by) in zip(line, line[1:]):
        dx, dy = bx - ax, by - ay
        run = run + math.sqrt(dx * dx + dy * dy)
    return run


def within(shell, px, py):
    """Point.within(Polygon): even-odd crossing, a point on the boundary is outside"""
    inside = False
    n = len(shell)
    for i in range(n):
        (x1, y1), (x2, y2) = shell[i], shell[(i + 1) % n]
        if seg_point(x1, y1, x2, y2, px, py)[0] == 0.0:
            return False
        if (y1 > py) != (y2 > py) and px < (x2 - x1) * (py - y1) / (y2 - y1) + x1:
            inside = not inside
    return inside


def simplify(pts, tol):
    """LineString.simplify (Douglas-Peucker)"""
    if len(pts) < 3:
        return list(pts)
    a, b = pts[0], pts[-1]
    imax, dmax = 0, -1.0
    for i in range(1, len(pts) - 1):
        d = seg_point(a[0], a[1], b[0], b[1], pts[i][0], pts[i][1])[0]
        if d > dmax:
            imax, dmax = i, d
    if dmax > tol:
        return simplify(pts[:imax + 1], tol)[:-1] + simplify(pts[imax:], tol)
    return [a, b]


# ---- utility.py / type.py -------------------------------------------------------------------------------------------

def _dist(p, q):                                            # utility.py:8-9
    dx, dy = p[0] - q[0], p[1] - q[1]
    return math.sqrt(dx * dx + dy * dy)


def extend_line(line, d):                                   # utility.py:111-115
    k = d / _dist(line[-1], line[-2])
    return [tuple(p) for p in line] + [(line[-1][0] + k * (line[-1][0] - line[-2][0]), line[-1][1] + k * (line[-1][1] - line[-2][1]))]


def extend_line_both_sides(line, d):                        # utility.py:118-126
    kf = d / _dist(line[0], line[1])
    kb = d / _dist(line[-1], line[-2])
    f = (line[0][0] + kf * (line[0][0] - line[1][0]), line[0][1] + kf * (line[0][1] - line[1][1]))
    b = (line[-1][0] + kb * (line[-1][0] - line[-2][0]), line[-1][1] + kb * (line[-1][1] - line[-2][1]))
    return [f] + [tuple(p) for p in line] + [b]


class LineSimplified:
    """type.py:255-317.  yaw_list is kept as segment directions: the reference compares angles
    (sin(arctan2(p - foot) - yaw) > 0), this file and the kernels take the sign of the cross product of the segment
    direction with (p - foot) — the same sign unless p lies on the line to rounding, where the reference's own answer
    is decided by the last bit of arctan2."""

    def __init__(self, path):
        self.path = [(float(p[0]), float(p[1])) for p in path]
        self.line = simplify(self.path, 0.05)
        self.path_length = line_length(self.line)
        self.dis_list = [0.0]
        run = 0.0
        for p, q in zip(self.path, self.path[1:]):
            run = run + _dist(p, q)
            self.dis_list.append(run)
        seg0 = (self.path[1][0] - self.path[0][0], self.path[1][1] - self.path[0][1])
        self.dir_list = [seg0] + [(q[0] - p[0], q[1] - p[1]) for p, q in zip(self.path, self.path[1:])]
        self.yaw_list = [math.atan2(d[1], d[0]) for d in self.dir_list]

    def _search(self, length):                               # np.searchsorted(dis_list, length), side="left"
        i = 0
        while i < len(self.dis_list) and self.dis_list[i] < length:
            i += 1
        return i

    def get_yaw_index(self, length):                         # type.py:276-280
        i = self._search(length)
        return i if i < len(self.yaw_list) else len(self.yaw_list) - 1

    def get_point_at_length(self, length):                   # type.py:282-295
        P = self.path
        if length == 0:
            return P[0]
        if length < self.path_length:
            i = self._search(length)
            seg = self.dis_list[i] - self.dis_list[i - 1]
            r = (self.dis_list[i] - length) / seg
            return (P[i][0] - r * (P[i][0] - P[i - 1][0]), P[i][1] - r * (P[i][1] - P[i - 1][1]))
        end = self.dis_list[-1] - self.dis_list[-2]
        r = (length - self.path_length) / end
        return (P[-1][0] + r * (P[-1][0] - P[-2][0]), P[-1][1] + r * (P[-1][1] - P[-2][1]))

    def side_of(self, s, px, py):
        """+1 left / -1 right of the line at arc length s (type.py:297-315 without the angles)"""
        fx, fy = self.get_point_at_length(s)
        dx, dy = self.dir_list[self.get_yaw_index(s)]
        return 1 if dx * (py - fy) - dy * (px - fx) > 0.0 else -1

    def to_frenet_frame(self, px, py):                       # type.py:307-317
        s, dis = project_and_distance(self.line, px, py)
        if dis <= 0.0:
            return s, 0.0
        return s, self.side_of(s, px, py) * dis


class PreparedMap:
    """What Map / Corridor constructors derive (environment.py:21-32, type.py:140-169) for corridors given as dicts
    {id, left, right, center, left_index, right_index} in dict-insertion order."""

    def __init__(self, corridors):
        self.n = len(corridors)
        self.left = [c["left_index"] for c in corridors]
        self.right = [c["right_index"] for c in corridors]
        self.center = [[(float(p[0]), float(p[1])) for p in c["center"]] for c in corridors]
        self.center_ext = [extend_line(c, 100) for c in self.center]                  # type.py:154
        self.shell = []
        for c in corridors:                                                           # type.py:160-164
            pts = [(float(p[0]), float(p[1])) for p in c["left"]] + [(float(p[0]), float(p[1])) for p in reversed(c["right"])]
            if pts[0] == pts[-1]:
                pts = pts[:-1]
            self.shell.append(pts)
        self.ref = [LineSimplified(extend_line_both_sides(c, 1000)) for c in self.center]   # mcs_wrapper.py:42,109,211
        self.center_simplified = [LineSimplified(c) for c in self.center]             # type.py:153
        self.border = [[LineSimplified(c["left"]), LineSimplified(c["right"])] for c in corridors]   # type.py:151-152
        self.raw = [{k: [(float(p[0]), float(p[1])) for p in c[k]] for k in ("left", "right", "center")} for c in corridors]


def match(pm, xs, ys):
    """environment.py:85-101 + 135-143 for one scene.  -> (lane[n], arc[n], lane_order[n], lane_start[L+1])"""
    n = len(xs)
    lane, arc = [-1] * n, [0.0] * n
    for i in range(n):
        if xs[i] != xs[i]:
            continue
        for c in range(pm.n):                                # corridor_match: first corridor in dict order
            if within(pm.shell[c], xs[i], ys[i]):
                lane[i] = c
                arc[i] = project_and_distance(pm.center[c], xs[i], ys[i])[0]
                break
    order, start = [], [0]
    for c in range(pm.n):
        members = [i for i in range(n) if lane[i] == c]
        members.sort(key=lambda i: arc[i])                   # list.sort is stable: ties stay in dict order
        order += members
        start.append(len(order))
    order += [-1] * (n - len(order))
    start += [start[-1]] * (9 - len(start))
    return lane, arc, order, start


def arc_on_lanes(pm, xs, ys):
    return [[project_and_distance(pm.center[c], x, y)[0] if x == x else 0.0 for x, y in zip(xs, ys)] for c in range(pm.n)]


def neighbors(pm, lane, arc, order, start, i, qx, qy):
    """front_agent, following_agent, neighbor_around_agent (environment.py:166-251) of slot i standing at (qx, qy),
    against the per-lane vectors of the last match.  -> ([10 slots or -1], [10 dis])"""
    idx, dis = [-1] * NEIGHBORS, [0.0] * NEIGHBORS
    c = lane[i]
    if c < 0:
        return idx, dis

    def lane_list(cc):
        return [(arc[j], j) for j in order[start[cc]:start[cc + 1]]]

    e = project_and_distance(pm.center[c], qx, qy)[0]
    own = lane_list(c)
    for key, j in own:                                       # :166-180
        if key > e:
            idx[FRONT], dis[FRONT] = j, key - e
            break
    for key, j in reversed(own):                             # :182-195
        if key < e:
            idx[BACK], dis[BACK] = j, key - e
            break
    for side, ff, f, b, bb in ((pm.left[c], LEFT_FF, LEFT_F, LEFT_B, LEFT_BB), (pm.right[c], RIGHT_FF, RIGHT_F, RIGHT_B, RIGHT_BB)):
        if side < 0:
            continue
        lst = lane_list(side)
        es = project_and_distance(pm.center_ext[side], qx, qy)[0]
        for key, j in lst:                                   # :213-219
            if idx[f] < 0 and key >= es:
                idx[f], dis[f] = j, key - es
                continue
            if idx[f] >= 0 and key > dis[f] + es:
                idx[ff], dis[ff] = j, key - es
                break
        for key, j in reversed(lst):                         # :220-226
            if idx[b] < 0 and key < es:
                idx[b], dis[b] = j, key - es
                continue
            if idx[b] >= 0 and key < dis[b] + es:
                idx[bb], dis[bb] = j, key - es
                break
    return idx, dis


def within_range(pm, lane, xs, ys, corridor, ego, max_dis):
    """environment.py:145-156: slots of `corridor` within max_dis (arc length) of slot `ego`, sorted by arc length"""
    e = project_and_distance(pm.center[corridor], xs[ego], ys[ego])[0]
    out = []
    for j in range(len(xs)):
        if lane[j] == corridor:
            a = project_and_distance(pm.center[corridor], xs[j], ys[j])[0]
            if abs(a - e) <= max_dis:
                out.append((a, j))
    out.sort(key=lambda t: t[0])
    return [j for _, j in out]


def step_along_path(pm, ref_lane, x, y, v, ax, vy, dt):
    """utility.py:88-98 on corridors[ref_lane].centerSimplified"""
    ls = pm.center_simplified[ref_lane]
    s, dis = project_and_distance(ls.line, x, y)
    sign = 1 if (dis > 0.0 and ls.side_of(s, x, y) > 0) else -1       # is_left_of: on the line -> False
    d = sign * dis
    s_new = s + v * dt
    d_new = d + vy * dt
    v_new = v + ax * dt
    v_new = v_new if v_new > 0 else 0.0
    k = ls.get_yaw_index(s_new)
    yaw = ls.yaw_list[k]
    px, py = ls.get_point_at_length(s_new)
    return px - d_new * math.sin(yaw), py + d_new * math.cos(yaw), v_new, yaw


def center_distance(pm, lane, x, y):
    """(centerSimplified.signed_distance type.py:307-315, Corridor.distance_to_corridor_end type.py:180-181) on corridors[lane]"""
    ls = pm.center_simplified[lane]
    s, dis = project_and_distance(ls.line, x, y)
    d = 0.0 if dis <= 0.0 else ls.side_of(s, x, y) * dis
    return d, ls.path_length - s


def _orient(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def _on_box(p, q, r):
    return min(p[0], q[0]) <= r[0] <= max(p[0], q[0]) and min(p[1], q[1]) <= r[1] <= max(p[1], q[1])


def segments_intersect(a, b, c, d):
    """do the closed segments AB and CD share a point (orientation test)"""
    o1, o2 = _orient(*a, *b, *c), _orient(*a, *b, *d)
    o3, o4 = _orient(*c, *d, *a), _orient(*c, *d, *b)
    if ((o1 > 0) != (o2 > 0)) and ((o3 > 0) != (o4 > 0)) and o1 != 0 and o2 != 0 and o3 != 0 and o4 != 0:
        return True
    return (o1 == 0 and _on_box(a, b, c)) or (o2 == 0 and _on_box(a, b, d)) or (o3 == 0 and _on_box(c, d, a)) or \
        (o4 == 0 and _on_box(c, d, b))


def seg_seg_distance(a, b, c, d):
    if segments_intersect(a, b, c, d):
        return 0.0
    return min(seg_point(*a, *b, *c)[0], seg_point(*a, *b, *d)[0], seg_point(*c, *d, *a)[0], seg_point(*c, *d, *b)[0])


def border_distance(pm, lane, side, hull):
    """Corridor.distance_to_border(hull, "left" if side == 0 else "right") type.py:183-206; hull = 4 corners"""
    ls = pm.border[lane][side]
    passed, worst = False, 0.0
    for (px, py) in hull:
        s, dis = project_and_distance(ls.line, px, py)
        left_of = dis > 0.0 and ls.side_of(s, px, py) > 0                      # is_left_of type.py:297-305
        if left_of if side == 0 else not left_of:
            passed = True
            worst = dis if dis > worst else worst
    if passed:
        return -worst
    shell = [(float(p[0]), float(p[1])) for p in hull]
    if any(within(shell, qx, qy) for qx, qy in ls.line):                        # Polygon.distance(LineString)
        return 0.0
    best = None
    for e in range(4):
        a, b = shell[e], shell[(e + 1) % 4]
        for c, d in zip(ls.line, ls.line[1:]):
            v = seg_seg_distance(a, b, c, d)
            if best is None or v < best:
                best = v
    return best


# ---- mcs_wrapper.py:8-283: the three builders of the inner simulator's scene ------------------------------------------
BEH_IDM, BEH_IDM_LC, BEH_MERGING, BEH_EXIT = 0, 1, 2, 3          # include/mcsim.h MCS_BEH_*
LANE_MAIN, LANE_MERGING, LANE_EXIT = 0, 1, 2                      # MCS_LANE_*


def build_scene(pm, info, tables, xs, ys, attrs, ids, kind, ego, ego_rec, others_merging, noise):
    """merging_ (kind 0, mcs_wrapper.py:24-75), exit_ (1, :78-190), lane_change_mcs_sim_from_env (2, :193-283) for slot
    `ego` at the vehicles' current positions xs / ys against the match `tables` = (lane, arc, order, start).
    info: per corridor {"id", "type" (LANE_*), "v_limit"}; attrs: per slot {"v", "vy", "a", "l", "w", "vd"}; ego_rec: the
    ego's {"behavior", "vd", "idm", "rss", "yp", "commit", "keep", "target"} (memory-order tuples); noise: the
    np.random.normal(0., 1) draws in append order (None: plain vd, the caller adds them).  -> (corridors, agents, possible_actions) as plain records:
    corridor = (id, type, [l.p1.x, l.p1.y, l.p2.x, l.p2.y], [r...], [c...], v_limit), agent = dict of mcs_agent fields."""
    lane, arc, order, start = tables
    ego_lane = lane[ego]
    items, corr, actions = [], [], [-1]

    def nb(i):
        return neighbors(pm, lane, arc, order, start, i, xs[i], ys[i])[0]

    def gap(j):
        if j >= 0 and len(actions) < 4:
            actions.append(ids[j])

    def in_range(l, beh):
        for j in within_range(pm, lane, xs, ys, l, ego, 80):
            items.append((j, beh))

    idx = nb(ego)
    if kind == 0:
        left = pm.left[ego_lane]
        ref = left                                                 # :42 reference line: the main lane next to the ramp
        beh = BEH_MERGING if others_merging else BEH_IDM_LC
        for k in (LEFT_FF, LEFT_F, LEFT_B, LEFT_BB, FRONT):         # neighbor.lefts_and_front, :47-54
            if idx[k] >= 0:
                items.append((idx[k], beh))
        for k in (LEFT_BB, LEFT_B, LEFT_F, LEFT_FF):                # reversed(neighbor.lefts), :57-59
            gap(idx[k])
        corr += [ego_lane, left]
        ll = pm.left[left]
        if ll >= 0:                                                # :64-72
            in_range(ll, BEH_IDM_LC)
            corr.append(ll)
    else:
        ref = ego_lane
        if kind == 1:                                              # :96-103 the exit lane somewhere to the right
            r, ex = pm.right[ego_lane], -1
            while r >= 0:
                if info[r]["type"] == LANE_EXIT:
                    ex = r
                    break
                r = pm.right[r]
            if ex < 0:
                raise ValueError("no exit lane to the right of the ego's lane")
            corr.append(ex)
        corr.append(ego_lane)
        if idx[FRONT] >= 0:                                        # :120-137 / :219-236 front, front of the front, follower
            items.append((idx[FRONT], BEH_IDM_LC))
            ff = nb(idx[FRONT])[FRONT]
            if ff >= 0:
                items.append((ff, BEH_IDM_LC))
        if idx[BACK] >= 0:
            items.append((idx[BACK], BEH_IDM_LC))
        left = pm.left[ego_lane]
        if left >= 0:                                              # :139-160 / :238-258
            for k in (LEFT_FF, LEFT_F, LEFT_B, LEFT_BB):
                if idx[k] >= 0:
                    items.append((idx[k], BEH_IDM_LC))
            corr.append(left)
            ll = pm.left[left]
            if ll >= 0:
                in_range(ll, BEH_IDM_LC)
                corr.append(ll)
        if kind == 1:                                              # :155-157 candidate gaps on the right
            for k in (RIGHT_BB, RIGHT_B, RIGHT_F, RIGHT_FF):
                gap(idx[k])
        right = pm.right[ego_lane]
        if right >= 0:                                             # :162-189 / :260-282
            beh = BEH_MERGING if info[right]["type"] == LANE_MERGING else BEH_IDM_LC
            for k in (RIGHT_FF, RIGHT_F, RIGHT_B, RIGHT_BB):
                if idx[k] >= 0:
                    items.append((idx[k], beh))
            corr.append(right)
            rr = pm.right[right]
            if rr >= 0:
                in_range(rr, BEH_MERGING if info[rr]["type"] == LANE_MERGING else BEH_IDM_LC)
                corr.append(rr)
    line = pm.ref[ref]
    ego_s, ego_d = line.to_frenet_frame(xs[ego], ys[ego])

    def corridor_ms(l, x_limit=None):                              # Corridor.to_frenet_corridor[_limit_length], type.py:207-237
        raw = pm.raw[l]
        pts = [raw["left"][0], raw["left"][-1], raw["right"][0], raw["right"][-1], raw["center"][0], raw["center"][-1]]
        sd = [line.to_frenet_frame(p[0], p[1]) for p in pts]
        far = [sd[1][0], sd[3][0], sd[5][0]]
        if x_limit is not None:
            far = [min(v, x_limit) for v in far]
        return (info[l]["id"], info[l]["type"], [sd[0][0], sd[0][1], far[0], sd[0][1]], [sd[2][0], sd[2][1], far[1], sd[2][1]],
                [sd[4][0], sd[4][1], far[2], sd[4][1]], info[l]["v_limit"])

    if kind == 1:
        exit_ms = corridor_ms(corr[0])
        el = corr[1]
        max_length = min(info[el]["v_limit"] ** 2 / 2.5 ** 2,
                         exit_ms[4][2] - (abs(exit_ms[0] - info[el]["id"]) - 1) * 50 - ego_s)              # :113-116
        corridors = [corridor_ms(el, ego_s + max_length)] + [corridor_ms(l) for l in corr[2:]]
    else:
        corridors = [corridor_ms(l) for l in corr]
    a = attrs[ego]
    agents = [dict(id=ids[ego], behavior=ego_rec["behavior"], x=ego_s, y=ego_d, v=a["v"], vy=a["vy"], vd=ego_rec["vd"], a=a["a"],
                   l=a["l"], w=a["w"], idm=tuple(ego_rec["idm"]), rss=tuple(ego_rec["rss"]), yp=tuple(ego_rec["yp"]),
                   commit=ego_rec["commit"], keep=ego_rec["keep"], target=ego_rec["target"])]
    for q, (j, beh) in enumerate(items):
        a = attrs[j]
        s_, d_ = line.to_frenet_frame(xs[j], ys[j])
        idm = (0.8, 2.0, -1.5, -2.0, -8.0, 1.2) if a["l"] > 7 else (1.6, 2.0, -3.0, -2.0, -8.0, 1.2)     # idm_param(l), :8-13
        rss = (0.7, 0.4, -8.0, -10.0, 1.8, 1.2) if beh == BEH_MERGING else (0.7, 0.4, -6.0, -6.0, 1.8, 1.2)   # rss_param, :16-21
        agents.append(dict(id=ids[j], behavior=beh, x=s_, y=d_, v=a["v"], vy=a["vy"], vd=a["vd"] + noise[q] if noise is not None else a["vd"], a=a["a"], l=a["l"],
                           w=a["w"], idm=idm, rss=rss, yp=(-0.5, 1.81, -4.8, -1.1, 0.9, 0.5, 0.51), commit=0, keep=0, target=-1))
    return corridors, agents, actions
