"""Host mirror of the reference's outer Map / Environment (environment.py:21-251) over the device kernels of
include/mcsim_env.h: one launch per `match_agents_on_corridor()` answers the lane match, the per-lane ordering and the
front / following / eight side neighbours of EVERY vehicle; the reference's per-vehicle query methods read those tables.

Same names, arguments and return shapes as the reference (`Map`, `Environment`, `IdmMatch`, `Neighbor`,
`MatchedAgent`), so the builders of mcs_wrapper.py and the behaviours of behaviors.py run on it unchanged.  Vehicles are
duck-typed: any object with the reference Agent's attributes (id, x, y, v, vy, a, l, w, idm_param.vd, behavior_model,
corridor_history, set_behavior_model) works.

The reference plans and moves its vehicles one after the other between two matches (environment.py:316-323): a query
for a vehicle that has moved since the match is answered by `mce_env_neighbors` against the resident (stale) per-lane
keys, exactly the reference's semantics.  No CPU fallback: every geometric answer comes from the kernels.
"""
import ctypes as C

import numpy as np

from . import _abi as abi
from . import backend
from .type import YieldingIntention

MCE_MAX_LANES, MCE_MAX_POINTS, MCE_MAX_AGENTS, MCE_NEIGHBORS = 8, 16, 1024, 10
LEFT_FF, LEFT_F, LEFT_B, LEFT_BB, FRONT, BACK, RIGHT_FF, RIGHT_F, RIGHT_B, RIGHT_BB = range(10)


class CorridorPOD(C.Structure):
    """mce_corridor"""
    _fields_ = [("id", C.c_int32), ("left_index", C.c_int32), ("right_index", C.c_int32),
                ("n_left", C.c_int32), ("n_right", C.c_int32), ("n_center", C.c_int32),
                ("left", C.c_double * 2 * MCE_MAX_POINTS), ("right", C.c_double * 2 * MCE_MAX_POINTS),
                ("center", C.c_double * 2 * MCE_MAX_POINTS)]


class BuildRequest(C.Structure):
    """mce_build_request"""
    _fields_ = [("kind", C.c_int32), ("scene", C.c_int32), ("ego_slot", C.c_int32), ("ego_behavior", C.c_int32),
                ("others_merging", C.c_int32), ("ego_commit", C.c_int32), ("ego_keep_step", C.c_int32),
                ("ego_target_lane", C.c_int32), ("ego_vd", C.c_double), ("ego_idm", C.c_double * 6),
                ("ego_rss", C.c_double * 6), ("ego_yp", C.c_double * 7)]


class WalkVehicle(C.Structure):
    """mce_walk_vehicle"""
    _fields_ = [("id", C.c_int32), ("behavior", C.c_int32), ("flags", C.c_int32), ("seed", C.c_int32),
                ("commit", C.c_int32), ("keep_step", C.c_int32), ("target_lane", C.c_int32), ("yi_n", C.c_int32),
                ("x", C.c_double), ("y", C.c_double), ("v", C.c_double), ("vy", C.c_double), ("a", C.c_double), ("yaw", C.c_double),
                ("l", C.c_double), ("w", C.c_double), ("hull", C.c_double * 8), ("idm", C.c_double * 6), ("rss", C.c_double * 6),
                ("mobil", C.c_double * 3), ("yp", C.c_double * 13), ("threshold", C.c_double), ("ax_last", C.c_double),
                ("yi_id", C.c_int32 * 12), ("yi_yielding", C.c_int32 * 12), ("yi_p0", C.c_double * 12)]


class WalkResult(C.Structure):
    """mce_walk_result"""
    _fields_ = [("stopped_at", C.c_int32), ("status", C.c_int32), ("noise_used", C.c_int32), ("mobil_calls", C.c_int32),
                ("random_used", C.c_int32), ("random_seed", C.c_int32), ("moved", C.c_int32), ("reserved", C.c_int32)]


WALK_BEHAVIORS = {"idm": 1, "idm_mobil_lane_change": 2, "idm_mobil_lane_change_safe": 3, "merging_closest_gap_policy": 4, "exit": 5}
_random_first = None


def random_first_table():
    """random.seed(s); random.random() for s = 0..99 (every consumer of Python's `random` in the outer step re-seeds with
    a vehicle's yielding seed and draws once: type.py:75-77, utility.py:136-138); the global stream is left untouched"""
    global _random_first
    if _random_first is None:
        import random
        state = random.getstate()
        out = []
        for s_ in range(100):
            random.seed(s_)
            out.append(random.random())
        random.setstate(state)
        _random_first = (C.c_double * 100)(*out)
    return _random_first


BUILD_MERGING, BUILD_EXIT, BUILD_LANE_CHANGE = 0, 1, 2
MCE_BUILD_MAX_AGENTS = 48


class EnvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mcsim_env error %d: %s" % (code, msg))
        self.code = code


_bound = False
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _lib():
    global _bound
    L = backend.lib()
    if not _bound:
        V, I, D = C.c_void_p, C.c_int, C.c_double
        L.mce_last_error.restype = C.c_char_p
        L.mce_map_create.argtypes = [C.POINTER(CorridorPOD), I, I, C.POINTER(V)]
        L.mce_map_destroy.argtypes = [V]
        L.mce_env_match.argtypes = [V, _pd, _pd, I, I, V, V, V, V, V, V, V]
        L.mce_env_neighbors.argtypes = [V, _pi, _pi, _pd, _pd, I, _pi, _pd]
        L.mce_frenet.argtypes = [V, I, _pd, _pd, I, _pd, _pd]
        L.mce_project_center.argtypes = [V, I, _pd, _pd, I, _pd, _pd]
        L.mce_step_along_path.argtypes = [V, V, V, V, V, V, V, D, I, V, V, V, V]
        L.mce_last_kernel_ms.argtypes = [V, C.POINTER(C.c_float)]
        L.mce_center_distance.argtypes = [V, _pi, _pd, _pd, I, _pd, _pd]
        L.mce_border_distance.argtypes = [V, _pi, _pi, _pd, I, _pd]
        L.mce_map_set_lane_info.argtypes = [V, _pi, _pi, _pd, I]
        L.mce_walk_step.argtypes = [V, C.POINTER(WalkVehicle), I, V, V, I, V, I, I, D, C.c_uint64, C.POINTER(WalkResult)]
        L.mce_walk_step.restype = I
        L.mce_build_rollout_scene.argtypes = [V, C.POINTER(BuildRequest), V, V, V, I, V, C.POINTER(C.c_int32), V,
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), I, C.POINTER(V)]
        for f in ("mce_map_set_lane_info", "mce_build_rollout_scene", "mce_center_distance", "mce_border_distance", "mce_map_create", "mce_map_destroy", "mce_env_match", "mce_env_neighbors", "mce_frenet", "mce_project_center",
                  "mce_step_along_path", "mce_last_kernel_ms"):
            getattr(L, f).restype = I
        _bound = True
    return L


def _check(rc):
    if rc != 0:
        raise EnvError(rc, _lib().mce_last_error().decode("utf-8", "replace"))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def corridor_pods(corridors):
    """corridors: sequence of objects or dicts with id, left, right, center, left_id, right_id (dict-insertion order)."""
    get = (lambda c, k: c[k]) if isinstance(corridors[0], dict) else getattr
    pos = {get(c, "id"): k for k, c in enumerate(corridors)}
    out = (CorridorPOD * len(corridors))()
    for k, c in enumerate(corridors):
        p = out[k]
        p.id = int(get(c, "id"))
        li, ri = get(c, "left_id"), get(c, "right_id")
        p.left_index = pos[li] if li is not None else -1
        p.right_index = pos[ri] if ri is not None else -1
        for name in ("left", "right", "center"):
            pts = get(c, name)
            if len(pts) > MCE_MAX_POINTS:
                raise EnvError(-3, "polyline with more than MCE_MAX_POINTS points")
            setattr(p, "n_" + name, len(pts))
            arr = getattr(p, name)
            for i, q in enumerate(pts):
                arr[i][0], arr[i][1] = float(q[0]), float(q[1])
    return out


class MatchTables:
    """Outputs of one mce_env_match launch (numpy, host)."""
    __slots__ = ("lane", "arc", "order", "start", "arc_on_lane", "nb_index", "nb_dis", "n_agents", "n_scenes")


class DeviceMap:
    """mce_map: the corridors of one map resident on a GPU."""

    def __init__(self, corridors, device=0):
        self._pods = corridor_pods(list(corridors))
        self.n_lanes = len(self._pods)
        self.center_cache = {}                               # (lane, x, y) -> (signed centre distance, distance to the end)
        self.center_fill = None                              # pending batch of the last match (Environment), run on first use
        self._build_out = None
        self._step_io = None
        self._h = C.c_void_p()
        _check(_lib().mce_map_create(self._pods, self.n_lanes, device, C.byref(self._h)))

    def close(self):
        if self._h:
            _lib().mce_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def match(self, x, y, n_scenes=1, arc_on_lanes=False, fetch=("lane", "arc", "order", "start", "neighbors")):
        """mce_env_match.  `fetch`: which tables come back to the host (the others stay None; all of them stay resident on
        the device for mce_env_neighbors / mce_build_rollout_scene either way — a caller that only needs the lanes of a
        large batch does not pay for 120 bytes of neighbour tables per vehicle)."""
        x, y = _f64(x).reshape(n_scenes, -1), _f64(y).reshape(n_scenes, -1)
        n = x.shape[1]
        t = MatchTables()
        t.n_agents, t.n_scenes = n, n_scenes
        t.lane = np.empty((n_scenes, n), np.int32) if "lane" in fetch else None
        t.arc = np.empty((n_scenes, n), np.float64) if "arc" in fetch else None
        t.order = np.empty((n_scenes, n), np.int32) if "order" in fetch else None
        t.start = np.empty((n_scenes, MCE_MAX_LANES + 1), np.int32) if "start" in fetch else None
        t.arc_on_lane = np.empty((n_scenes, self.n_lanes, n), np.float64) if arc_on_lanes else None
        t.nb_index = np.empty((n_scenes, MCE_NEIGHBORS, n), np.int32) if "neighbors" in fetch else None
        t.nb_dis = np.empty((n_scenes, MCE_NEIGHBORS, n), np.float64) if "neighbors" in fetch else None
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        _check(_lib().mce_env_match(self._h, x, y, n, n_scenes, p(t.lane), p(t.arc), p(t.order), p(t.start), p(t.arc_on_lane),
                                    p(t.nb_index), p(t.nb_dis)))
        return t

    def neighbors(self, scene, slot, qx, qy):
        scene, slot = np.ascontiguousarray(scene, np.int32), np.ascontiguousarray(slot, np.int32)
        qx, qy = _f64(qx), _f64(qy)
        q = len(slot)
        idx, dis = np.empty((MCE_NEIGHBORS, q), np.int32), np.empty((MCE_NEIGHBORS, q), np.float64)
        _check(_lib().mce_env_neighbors(self._h, scene, slot, qx, qy, q, idx, dis))
        return idx, dis

    def frenet(self, ref_lane, x, y):
        x, y = _f64(x), _f64(y)
        s, d = np.empty(len(x)), np.empty(len(x))
        _check(_lib().mce_frenet(self._h, int(ref_lane), x, y, len(x), s, d))
        return s, d

    def project_center(self, lane, x, y):
        x, y = _f64(x), _f64(y)
        s, d = np.empty(len(x)), np.empty(len(x))
        _check(_lib().mce_project_center(self._h, int(lane), x, y, len(x), s, d))
        return s, d

    def step_along_path(self, ref_lane, x, y, v, ax, vy, dt):
        ref_lane = np.ascontiguousarray(ref_lane, np.int32)
        n = len(ref_lane)
        out = [np.empty(n) for _ in range(4)]
        ins = [ref_lane, _f64(x), _f64(y), _f64(v), _f64(ax), _f64(vy)]
        _check(_lib().mce_step_along_path(self._h, *[a.ctypes.data for a in ins], float(dt), n, *[a.ctypes.data for a in out]))
        return out

    def step_one(self, ref_lane, x, y, v, ax, vy, dt):
        """step_along_path for one vehicle (the sequential outer loop moves them one by one) without numpy in between"""
        if self._step_io is None:
            self._step_io = ((C.c_int32 * 1)(), [(C.c_double * 1)() for _ in range(9)])
        lane, d = self._step_io
        lane[0] = ref_lane
        d[0][0], d[1][0], d[2][0], d[3][0], d[4][0] = x, y, v, ax, vy
        _check(_lib().mce_step_along_path(self._h, C.addressof(lane), *[C.addressof(a) for a in d[:5]], dt, 1,
                                          *[C.addressof(a) for a in d[5:]]))
        return d[5][0], d[6][0], d[7][0], d[8][0]

    def center_distance(self, lane, x, y):
        """(centerSimplified.signed_distance, distance_to_corridor_end) of every point on its own corridor index"""
        lane, x, y = np.ascontiguousarray(lane, np.int32), _f64(x), _f64(y)
        d, e = np.empty(len(x)), np.empty(len(x))
        _check(_lib().mce_center_distance(self._h, lane, x, y, len(x), d, e))
        return d, e

    def border_distance(self, lane, side, hull):
        """Corridor.distance_to_border of hulls [n][4][2]; side 0 = left, 1 = right"""
        lane, side = np.ascontiguousarray(lane, np.int32), np.ascontiguousarray(side, np.int32)
        hull = _f64(hull).reshape(len(lane), 8)
        out = np.empty(len(lane))
        _check(_lib().mce_border_distance(self._h, lane, side, hull, len(lane), out))
        return out

    def set_lane_info(self, ids, types, v_limits):
        """Corridor.id / type (MCS_LANE_*) / v_limit per corridor, for the builders (mce_map_set_lane_info)"""
        _check(_lib().mce_map_set_lane_info(self._h, np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(types, np.int32),
                                            _f64(v_limits), len(ids)))

    def build_scene(self, request, ids, attrs, noise, device=0, records=True, scene=True):
        """mce_build_rollout_scene: one of mcs_wrapper's builders on the device.  Returns (corridors, agents, actions, handle):
        ctypes arrays of abi.Corridor / abi.Agent cut to their lengths (None unless `records`), the candidate gap ids, and
        the handle of the marshalled inner scene (None unless `scene`; the caller owns it: backend.DeviceScene.adopt)."""
        # `ids` / `attrs`: numpy arrays (int32 [n], float64 [8][n]) or raw addresses of such arrays (a caller in a loop
        # caches them).  The record arrays are the map's own and are overwritten by the next call.
        if self._build_out is None:
            self._build_out = ((abi.Corridor * MCE_MAX_LANES)(), (abi.Agent * MCE_BUILD_MAX_AGENTS)(), C.c_int32(), C.c_int32(),
                               C.c_int32(), (C.c_int32 * 4)())
        cs, ag, nc, na, nact, acts = self._build_out
        if not isinstance(ids, int):
            ids = np.ascontiguousarray(ids, np.int32).ctypes.data
        if not isinstance(attrs, int):
            attrs = _f64(attrs).ctypes.data
        pn, nn = None, 0
        if noise is not None:
            noise = _f64(noise)
            pn, nn = noise.ctypes.data, len(noise)
        h = C.c_void_p()
        _check(_lib().mce_build_rollout_scene(self._h, request, ids, attrs, pn, nn, cs, nc, ag, na, acts, nact, device,
                                              C.byref(h) if scene else None))
        return cs, nc.value, ag, na.value, list(acts[:nact.value]), (h if scene else None)

    def walk_step(self, vehicles, upload, ids_ptr, noise, first, last, dt, seed_base):
        """mce_walk_step: the sequential plan-and-move walk over slots [first, last) on the device -> WalkResult"""
        res = WalkResult()
        pn, nn = (noise.ctypes.data, len(noise)) if noise is not None and len(noise) else (None, 0)
        _check(_lib().mce_walk_step(self._h, vehicles, 1 if upload else 0, ids_ptr, pn, nn, C.addressof(random_first_table()),
                                    first, last, float(dt), seed_base, C.byref(res)))
        return res

    @staticmethod
    def create_scene(corridors, n_corridors, agents, n_agents, device=0):
        """mcs_scene_create over the first n_corridors / n_agents records of the builder's arrays -> mcs_scene*"""
        handle = C.c_void_p()
        backend.check(backend.lib().mcs_scene_create(corridors, n_corridors, agents, n_agents, device, C.byref(handle)))
        return handle

    def last_kernel_ms(self):
        ms = (C.c_float * 2)()
        _check(_lib().mce_last_kernel_ms(self._h, ms))
        return float(ms[0]), float(ms[1])


# ---- the reference's value classes (agent.py:203-247, type.py:140-176) ----------------------------------------------

class _CenterLine:
    """corridor.centerSimplified (type.py:153) as far as the behaviours read it: signed_distance, on the device"""

    def __init__(self, corridor):
        self._c = corridor

    def signed_distance(self, point):                        # type.py:307-315
        return self._c._center_distance(point)[0]


class Corridor:
    """type.py:140-176 reduced to data (the polygons / shapely lines live on the device; `Map` attaches the handle)."""

    def __init__(self, idx, t, left, right, center, v_limit):
        self.id, self.type, self.left, self.right, self.center, self.v_limit = idx, t, left, right, center, v_limit
        self.left_id = self.right_id = None
        self.device_map, self.lane_index = None, -1
        self.centerSimplified = _CenterLine(self)

    def set_left_and_right_corridor_id(self, left_id, right_id):
        self.left_id, self.right_id = left_id, right_id

    # ---- what the reference's Corridor derives in its constructor (type.py:157-169), as cheap host values.  Not on the
    # hot path (plot limits, sanity checks); the device holds its own copies of the geometry.
    @property
    def border(self):
        """left + reversed right: the shell of the corridor polygon"""
        return [list(p) for p in self.left] + [list(p) for p in reversed(self.right)]

    @property
    def length(self):
        """centerSimplified.path_length = the distance from the first centre point to the corridor end"""
        return self.distance_to_corridor_end(self.center[0])

    @property
    def width(self):
        """leftLine.distance(rightLine): smallest distance between the two border polylines"""
        return min(_segment_distance(self.left[i], self.left[i + 1], self.right[j], self.right[j + 1])
                   for i in range(len(self.left) - 1) for j in range(len(self.right) - 1))

    def get_x_y_limit(self):                                 # type.py:208-215
        pts = list(self.left) + list(self.right)
        return min(p[0] for p in pts), max(p[0] for p in pts), min(p[1] for p in pts), max(p[1] for p in pts)

    x_min = property(lambda self: self.get_x_y_limit()[0])
    x_max = property(lambda self: self.get_x_y_limit()[1])
    y_min = property(lambda self: self.get_x_y_limit()[2])
    y_max = property(lambda self: self.get_x_y_limit()[3])

    def contains(self, point):
        """Point(point).within(self.polygon): even-odd rule, points on the boundary are outside (the rule of the device's
        lane match, env_kernels.cuh `within`)"""
        px, py = float(point[0]), float(point[1])
        shell = self.border
        inside = False
        for k in range(len(shell)):
            (x1, y1), (x2, y2) = shell[k], shell[(k + 1) % len(shell)]
            if _point_segment_distance(px, py, x1, y1, x2, y2) == 0.0:
                return False
            if (y1 > py) != (y2 > py) and px < (x2 - x1) * (py - y1) / (y2 - y1) + x1:
                inside = not inside
        return inside

    def _center_distance(self, point):
        """(signed distance to the centre line, distance to the corridor end); answers are kept per position — the
        environment fills them for every vehicle of a step in one launch (Environment.match_agents_on_corridor)"""
        key = (self.lane_index, float(point[0]), float(point[1]))
        cache = self.device_map.center_cache
        if key not in cache:
            fill = getattr(self.device_map, "center_fill", None)
            if fill is not None:                             # the step's batch for every matched vehicle, on first use
                self.device_map.center_fill = None
                fill()
        if key not in cache:
            if len(cache) > 65536:                           # a map used without an Environment never gets its cache reset
                cache.clear()
            d, e = self.device_map.center_distance([key[0]], [key[1]], [key[2]])
            cache[key] = (float(d[0]), float(e[0]))
        return cache[key]

    def distance_to_corridor_end(self, position):            # type.py:180-181
        return self._center_distance(position)[1]

    def distance_to_border(self, hull, direction):           # type.py:183-206
        return float(self.device_map.border_distance([self.lane_index], [0 if direction == "left" else 1], [hull])[0])

    def to_dict(self):                                       # environment.py:59-67 (one entry of Map.to_dict)
        return {"id": self.id, "type": self.type, "left": self.left, "right": self.right, "center": self.center,
                "v_limit": self.v_limit, "left_id": "none" if self.left_id is None else self.left_id,
                "right_id": "none" if self.right_id is None else self.right_id}


def _point_segment_distance(px, py, ax, ay, bx, by):
    dx, dy = bx - ax, by - ay
    d2 = dx * dx + dy * dy
    t = 0.0 if d2 == 0.0 else min(1.0, max(0.0, ((px - ax) * dx + (py - ay) * dy) / d2))
    ex, ey = px - (ax + t * dx), py - (ay + t * dy)
    return (ex * ex + ey * ey) ** 0.5


def _segment_distance(a, b, c, d):
    """distance between the closed segments ab and cd (0 when they touch)"""
    def orient(p, q, r):
        return (q[0] - p[0]) * (r[1] - p[1]) - (q[1] - p[1]) * (r[0] - p[0])
    o1, o2, o3, o4 = orient(a, b, c), orient(a, b, d), orient(c, d, a), orient(c, d, b)
    if (o1 > 0) != (o2 > 0) and (o3 > 0) != (o4 > 0) and 0 not in (o1, o2, o3, o4):
        return 0.0
    return min(_point_segment_distance(c[0], c[1], a[0], a[1], b[0], b[1]), _point_segment_distance(d[0], d[1], a[0], a[1], b[0], b[1]),
               _point_segment_distance(a[0], a[1], c[0], c[1], d[0], d[1]), _point_segment_distance(b[0], b[1], c[0], c[1], d[0], d[1]))


def create_corridor_from_dict(d):                            # environment.py:12-18
    c = Corridor(d["id"], d["type"], d["left"], d["right"], d["center"], d["v_limit"])
    c.set_left_and_right_corridor_id(d["left_id"] if d["left_id"] != "none" else None,
                                     d["right_id"] if d["right_id"] != "none" else None)
    return c


class MatchedAgent:                                          # agent.py:229-232
    def __init__(self, agent, corridor_id):
        self.agent, self.corridor_id = agent, corridor_id


class IdmMatch:                                              # agent.py:235-247
    def __init__(self, dis, agent):
        self.id, self.agent, self.dis = agent.id, agent, dis
        self.x, self.y, self.v, self.vy, self.a, self.l, self.w = agent.x, agent.y, agent.v, agent.vy, agent.a, agent.l, agent.w
        self.vd = agent.idm_param.vd


class SimpleIdmMatch:                                        # agent.py:250-253
    def __init__(self, dis, v):
        self.dis, self.v = dis, v


class Neighbor:                                              # agent.py:203-226
    def __init__(self, left_ff, left_f, left_b, left_bb, f, b, right_ff, right_f, right_b, right_bb):
        self.left_ff, self.left_f, self.left_b, self.left_bb, self.f, self.b = left_ff, left_f, left_b, left_bb, f, b
        self.right_ff, self.right_f, self.right_b, self.right_bb = right_ff, right_f, right_b, right_bb
        self.lefts = [left_ff, left_f, left_b, left_bb]
        self.rights = [right_ff, right_f, right_b, right_bb]
        self.lefts_and_front = [left_ff, left_f, left_b, left_bb, f]


class Map:
    """environment.py:21-56"""

    def __init__(self, corridors, device=0):
        self.corridors = {}
        self.corridors_by_type = {}
        for c in corridors:
            self.corridors[c.id] = c
            self.corridors_by_type.setdefault(c.type, []).append(c)
        self.index_of = {cid: k for k, cid in enumerate(self.corridors)}
        self.id_at = list(self.corridors)
        self.device = device
        self.device_map = DeviceMap(list(self.corridors.values()), device)
        for k, c in enumerate(self.corridors.values()):
            c.device_map, c.lane_index = self.device_map, k
        self.device_map.set_lane_info([int(c.id) for c in self.corridors.values()],
                                      [abi.LANE_TYPES.get(c.type, abi.LANE_OTHER) for c in self.corridors.values()],
                                      [float(c.v_limit) for c in self.corridors.values()])

    def to_dict(self):                                       # environment.py:58-68
        return {i: c.to_dict() for i, c in self.corridors.items()}

    def corridor_match(self, vehicle):                       # environment.py:35-39 (one vehicle, host; the scene-wide match is the kernel's)
        for corridor in self.corridors.values():
            if corridor.contains([vehicle.x, vehicle.y]):
                return corridor
        return False

    def on_merging_lane(self, point):                        # environment.py:41-45
        return any(c.contains(point) for c in self.corridors_by_type["merging"])

    def truck_lane_id(self):                                 # environment.py:46-56
        if "merging" in self.corridors_by_type:
            return self.corridors_by_type["merging"][0].left_id
        if "exit" in self.corridors_by_type:
            return self.corridors_by_type["exit"][0].left_id
        for i, c in self.corridors.items():
            if c.right_id is None:
                return i


class Environment:
    """environment.py:70-251: same attributes (`agents`, `removed_agents`, `matched_agents`,
    `matched_agents_sorted_by_arc_length`) and query methods."""

    def __init__(self, m, agents):
        self.map = m
        self.agents = {a.id: a for a in agents}
        self.removed_agents = {}
        self.matched_agents = {}
        self.matched_agents_sorted_by_arc_length = {}
        self.sim_time = 0
        self._walk = None
        self._walk_objs = None
        self.match_agents_on_corridor()

    def match_agents_on_corridor(self):                      # environment.py:85-101 — one launch
        objs = list(self.agents.values())
        self.matched_agents = {}
        self._border_cache = {}
        self.matched_agents_sorted_by_arc_length = {c: [] for c in self.map.corridors}
        self._slot, self._objs = {a.id: i for i, a in enumerate(objs)}, objs
        if not objs:
            self._tables = None
            return
        self._x0 = np.array([a.x for a in objs], np.float64)
        self._y0 = np.array([a.y for a in objs], np.float64)
        # structure-of-arrays mirror of what the builders read of every vehicle (mce_build_rollout_scene): x y v vy a l w vd
        self._ids = np.array([a.id for a in objs], np.int32)
        self._attr = np.array([self._x0, self._y0, [a.v for a in objs], [a.vy for a in objs], [a.a for a in objs],
                               [a.l for a in objs], [a.w for a in objs], [a.idm_param.vd for a in objs]], np.float64)
        self._ids_p, self._attr_p = self._ids.ctypes.data, self._attr.ctypes.data
        self._mirror_live = False
        old = self._walk_objs
        if old is None or len(old) != len(objs) or any(p is not q for p, q in zip(old, objs)):
            self._walk = None                                 # the slots changed (first match, or a vehicle left the map)
        t = self._tables = self.map.device_map.match(self._x0, self._y0, 1, arc_on_lanes=True)
        for i, a in enumerate(objs):
            lane = int(t.lane[0, i])
            if lane < 0:
                self.removed_agents[a.id] = self.agents.pop(a.id)
                continue
            c = self.map.corridors[self.map.id_at[lane]]
            a.corridor_history.append(c.id)
            if c.type == "main" and "merging" in a.behavior_model:
                a.set_behavior_model("idm_mobil_lane_change_safe")
            if c.type == "exit" and "exit" in a.behavior_model:
                a.set_behavior_model("idm")
            self.matched_agents[a.id] = MatchedAgent(a, c.id)
        for lane, cid in enumerate(self.map.id_at):
            members = t.order[0, t.start[0, lane]:t.start[0, lane + 1]]
            self.matched_agents_sorted_by_arc_length[cid] = [[float(t.arc[0, j]), objs[j]] for j in members]
        # centre-line distances of every matched vehicle at its matched position, one launch (read by idm_behavior and the
        # yielding model during the sequential loop on the host; a vehicle that has moved since is asked for again).  The
        # launch waits for its first reader: a step whose vehicles are all walked on the device never asks.
        dm = self.map.device_map
        dm.center_cache.clear()
        x0, y0 = self._x0, self._y0

        def fill():
            on = [i for i in range(len(objs)) if t.lane[0, i] >= 0]
            if on:
                lanes = [int(t.lane[0, i]) for i in on]
                d, e = dm.center_distance(lanes, x0[on], y0[on])
                for k, i in enumerate(on):
                    dm.center_cache[(lanes[k], float(x0[i]), float(y0[i]))] = (float(d[k]), float(e[k]))
        dm.center_fill = fill

    def corridor_for_agent(self, idx):                       # environment.py:103-112
        if idx not in self.agents or idx not in self.matched_agents:
            return None
        return self.map.corridors[self.matched_agents[idx].corridor_id]

    def _neighbor_row(self, idx):
        """(slots[10], dis[10]) of vehicle idx at its current position against the vectors of the last match"""
        i = self._slot[idx]
        a = self._objs[i]
        if a.x == self._x0[i] and a.y == self._y0[i]:
            return self._tables.nb_index[0, :, i], self._tables.nb_dis[0, :, i]
        nb, dis = self.map.device_map.neighbors([0], [i], [a.x], [a.y])
        return nb[:, 0], dis[:, 0]

    def _match_or_none(self, row, k):
        j = int(row[0][k])
        return IdmMatch(float(row[1][k]), self._objs[j]) if j >= 0 else None

    def front_agent(self, idx):                              # environment.py:166-180
        if idx not in self.agents or idx not in self.matched_agents:
            return None
        return self._match_or_none(self._neighbor_row(idx), FRONT)

    def following_agent(self, idx):                          # environment.py:182-195
        if idx not in self.agents or idx not in self.matched_agents:
            return None
        return self._match_or_none(self._neighbor_row(idx), BACK)

    def neighbor_around_agent(self, idx):                    # environment.py:197-251
        if idx not in self.agents or idx not in self.matched_agents:
            return Neighbor(*([None] * 10))
        row = self._neighbor_row(idx)
        return Neighbor(*[self._match_or_none(row, k) for k in range(MCE_NEIGHBORS)])

    def sorted_agents_on_corridor_within_range_of_vehicle(self, corridor_id, ego_id, max_l_dis):   # environment.py:145-156
        lane = self.map.index_of[corridor_id]
        members = [a.agent for a in self.matched_agents.values() if a.corridor_id == corridor_id]
        ego = self.agents[ego_id]
        pts = [ego] + members
        moved = any(p.x != self._x0[self._slot[p.id]] or p.y != self._y0[self._slot[p.id]] for p in pts)
        if moved:
            arcs, _ = self.map.device_map.project_center(lane, [p.x for p in pts], [p.y for p in pts])
        else:
            arcs = [self._tables.arc_on_lane[0, lane, self._slot[p.id]] for p in pts]
        e = arcs[0]
        near = [(arcs[k + 1], k) for k in range(len(members)) if abs(arcs[k + 1] - e) <= max_l_dis]
        near.sort(key=lambda t: t[0])
        return [members[k] for _, k in near]

    def agent_to_idm_match(self, ego_id, obj_id):            # environment.py:158-164
        ego, obj = self.agents[ego_id], self.agents[obj_id]
        lane = self.map.index_of[self.matched_agents[obj_id].corridor_id]
        arcs, _ = self.map.device_map.project_center(lane, [ego.x, obj.x], [ego.y, obj.y])
        return IdmMatch(float(arcs[1] - arcs[0]), obj)

    def dis_to_lane_border(self, idx, direction):            # environment.py:113-115
        a = self.matched_agents[idx].agent
        key = (idx, direction, tuple(map(tuple, a.hull)))
        if key not in self._border_cache:
            self._border_cache[key] = self.corridor_for_agent(idx).distance_to_border(a.hull, direction)
        return self._border_cache[key]

    def limit_of_vehicles(self):                             # environment.py:117-124
        xs, ys = [a.x for a in self.agents.values()], [a.y for a in self.agents.values()]
        return min(xs + [1e10]) - 30, max(xs + [-1e10]) + 30, min(ys + [1e10]) - 10, max(ys + [-1e10]) + 10

    def limit_of_maps(self):                                 # environment.py:126-133 over Corridor.get_x_y_limit type.py:208-215
        pts = [p for c in self.map.corridors.values() for p in list(c.left) + list(c.right)]
        return min(p[0] for p in pts), max(p[0] for p in pts), min(p[1] for p in pts), max(p[1] for p in pts)

    # ---- bookkeeping of the evaluation scripts (environment.py:253-313) ----
    def all_finish_merging(self):
        return all(a.merging_flag.finish_merging for a in self.agents.values() if "merging_flag" in a.__dict__)

    def all_finish_exit(self):
        return all(self.map.corridors[a.corridor_history[-1]].type == "exit" for a in self.agents.values() if "exit" in a.behavior_model)

    def agents_once_on_merging_lane(self):
        return [a for a in self.agents.values() if "merging_flag" in a.__dict__]

    def agents_never_on_merging_lane(self):
        return [a for a in self.agents.values() if "merging_flag" not in a.__dict__]

    def average_normalized_acc_agents(self, agents):
        total = 0
        for a in agents:
            average_acc = sum(a.acceleration_history) / len(a.acceleration_history)
            total += (average_acc - a.idm_param.acc_min) / (a.idm_param.acc_max - a.idm_param.acc_min)
        return total / max(len(agents), 1)

    def set_flag(self, agent):                               # environment.py:282-313
        for b in agent.behavior_model_history:
            if "merging" in b:
                if not agent.merging_flag.flag_freeze:
                    t = self.map.corridors[self.matched_agents[agent.id].corridor_id].type
                    if t == "main":
                        agent.merging_flag.finish_merging = True
                        agent.merging_flag.t_finish_merging = self.sim_time
                        agent.merging_flag.flag_freeze = True
                    if t == "merging" and agent.v <= 2 and agent.had_harsh_brake():
                        agent.merging_flag.once_fallback = True
                break
        for b in agent.behavior_model_history:
            if "exit" in b:
                if not agent.exit_flag.flag_freeze:
                    t = self.map.corridors[self.matched_agents[agent.id].corridor_id].type
                    if t == "exit":
                        agent.exit_flag.finish_exit = True
                        agent.exit_flag.t_finish_exit = self.sim_time
                        agent.exit_flag.flag_freeze = True
                    if t == "main" and agent.had_harsh_brake():
                        agent.merging_flag.once_fallback = True       # sic (environment.py:311)
                break

    def _mirror(self):
        """The vehicles' current x y v vy a in the builders' mirror: rows are kept up to date by step(); outside of it (a
        caller that moved vehicles by hand) they are read again."""
        if not self._mirror_live:
            at = self._attr
            for i, a in enumerate(self._objs):
                at[0, i] = a.x; at[1, i] = a.y; at[2, i] = a.v; at[3, i] = a.vy; at[4, i] = a.a
        return self._ids_p, self._attr_p

    def marshal(self, agent, kind, records=False, scene=True, poke_commit=False):
        """One of mcs_wrapper's builders for `agent` on the device (mce_build_rollout_scene): `kind` = BUILD_MERGING /
        BUILD_EXIT / BUILD_LANE_CHANGE.  Consumes numpy's global stream like the reference: one np.random.normal(0., 1) per
        collected vehicle.  Returns (corridors, n_corridors, agents, n_agents, possible_actions, scene handle): ctypes arrays
        of abi.Corridor / abi.Agent (valid up to their counts) and, with `scene`, the mcs_scene* of the marshalled inner scene
        (the caller owns it: backend.DeviceScene.adopt)."""
        if agent.id not in self._slot or agent.id not in self.matched_agents:
            raise EnvError(-1, "agent %r is not matched on any corridor" % (agent.id,))
        ip, mp, yp = agent.idm_param, agent.mobil_param, agent.yielding_param
        merging_model = "merging" in agent.behavior_model
        rq = BuildRequest()
        rq.kind, rq.scene, rq.ego_slot = kind, 0, self._slot[agent.id]
        if kind == BUILD_EXIT:
            rq.ego_behavior = abi.BEHAVIORS["exit"]
        else:
            rq.ego_behavior = abi.BEHAVIORS["merging" if merging_model else "idm"]
        rq.others_merging = 1 if merging_model else 0
        if poke_commit:       # lane_change_behavior_mobil hands the vehicle's commit state to the inner agent (behaviors.py:75-78)
            rq.ego_commit = abi.COMMITS[agent.commit_lane_change]
            rq.ego_keep_step, rq.ego_target_lane = int(agent.commit_keep_lane_step), int(agent.target_lane_id)
        else:
            rq.ego_commit, rq.ego_keep_step, rq.ego_target_lane = abi.COMMITS["not_decided"], 0, -1
        rq.ego_vd = ip.vd
        rq.ego_idm[:] = (ip.acc_max, ip.min_dis, ip.acc_min, ip.acc_com, -8., ip.thw)         # IdmParam memory order
        if kind == BUILD_LANE_CHANGE:
            rp = agent.rss_param
            rq.ego_rss[:] = (rp.r_t_other, rp.r_t_ego, rp.a_min, rp.a_max, rp.soft_acc, rp.soft_dcc)
        else:
            rq.ego_rss[:] = (0.7, 0.4, -8., -10., 1.8, 1.2)                                    # rss_param(True)
        rq.ego_yp[:] = (yp[0], yp[1], yp[2], yp[3], mp.politeness_factor, mp.delta_a, mp.bias_a)
        ids, attrs = self._mirror()
        cs, nc, ag, na, actions, _ = self.map.device_map.build_scene(rq, ids, attrs, None, device=self.map.device, records=True,
                                                                     scene=False)
        # the reference draws one np.random.normal(0., 1) per vehicle while it appends them: now that their number is known
        # the same draws come from the same stream and are added where the reference adds them (vd + noise, IEEE add)
        if na > 1:
            noise = np.random.normal(0., 1, na - 1)
            for q in range(1, na):
                ag[q].vd += noise[q - 1]
        handle = self.map.device_map.create_scene(cs, nc, ag, na, self.map.device) if scene else None
        return cs, nc, ag, na, actions, handle

    # ---- the sequential walk on the device (mce_walk_step, csrc/walk_kernel.cuh) ----
    walk_on_device = True          # class-wide switch (tests compare both paths)

    def _walk_record(self, rec, a):
        ip, rp, mp = a.idm_param, a.rss_param, a.mobil_param
        rec.id, rec.behavior = int(a.id), WALK_BEHAVIORS.get(a.behavior_model, 0)
        rec.flags = (1 if "merging" in a.behavior_model else 0) | (2 if "exit" in a.behavior_model else 0)
        rec.seed = int(a.random_yielding_seed)
        rec.commit, rec.keep_step, rec.target_lane = abi.COMMITS[a.commit_lane_change], int(a.commit_keep_lane_step), int(a.target_lane_id)
        rec.x, rec.y, rec.v, rec.vy, rec.a, rec.yaw, rec.l, rec.w = a.x, a.y, a.v, a.vy, a.a, a.yaw, a.l, a.w
        rec.hull[:] = [c for p in a.hull for c in p]
        rec.idm[:] = (ip.acc_max, ip.acc_min, ip.min_dis, ip.acc_com, ip.thw, ip.vd)
        rec.rss[:] = (rp.r_t_ego, rp.r_t_other, rp.a_min, rp.a_max, rp.soft_acc, rp.soft_dcc)
        rec.mobil[:] = (mp.politeness_factor, mp.delta_a, mp.bias_a)
        rec.yp[:] = tuple(a.yielding_model.param)
        rec.threshold = a.change_intention_threshold
        held = list(a.yielding_intention_to_others.values())
        if len(held) > 12 or rec.seed < 0 or rec.seed > 99:
            rec.behavior = 0                                  # beyond what the kernel's tables hold: this vehicle stays on the host
            held = []
        rec.yi_n = len(held)
        for k, it in enumerate(held):
            rec.yi_id[k], rec.yi_yielding[k], rec.yi_p0[k] = int(it.id), 1 if it.yielding else 0, it.initial_probability

    def _walk_array(self):
        """The vehicles as mce_walk_vehicle records, slot order.  Built once per set of slots; afterwards the device hands the
        moved records back every step, and a vehicle somebody else touched (host-planned, behaviour model changed, moved by
        hand) is written again."""
        objs = self._objs
        if self._walk is None:
            self._walk = (WalkVehicle * len(objs))()
            for rec, a in zip(self._walk, objs):
                self._walk_record(rec, a)
            self._walk_objs = objs
            self._walk_models = [a.behavior_model for a in objs]
            return self._walk, True
        stale = False
        for i, a in enumerate(objs):
            rec = self._walk[i]
            if a.x != rec.x or a.y != rec.y or a.v != rec.v or a.behavior_model != self._walk_models[i]:
                self._walk_record(rec, a)
                self._walk_models[i] = a.behavior_model
                stale = True
        return self._walk, stale

    def _walk_take(self, i):
        """vehicle i as the device moved it -> the Python object (what Agent.step and lane_change_behavior_mobil leave behind)"""
        rec, a = self._walk[i], self._objs[i]
        a.x, a.y, a.v, a.yaw, a.a, a.vy = rec.x, rec.y, rec.v, rec.yaw, rec.a, rec.vy
        a.acceleration_history.append(a.a)
        h = rec.hull
        a.hull = [[h[0], h[1]], [h[2], h[3]], [h[4], h[5]], [h[6], h[7]]]
        a.state_history.append([a.x, a.y, a.v, a.a])
        a.commit_lane_change, a.commit_keep_lane_step, a.target_lane_id = abi.COMMIT_NAMES[rec.commit], rec.keep_step, rec.target_lane
        if rec.yi_n or a.yielding_intention_to_others:
            held = {}
            for k in range(rec.yi_n):
                it = YieldingIntention.__new__(YieldingIntention)
                it.id, it.initial_probability, it.change_intention_threshold = rec.yi_id[k], rec.yi_p0[k], a.change_intention_threshold
                it.yielding = bool(rec.yi_yielding[k])
                held[it.id] = it
            a.yielding_intention_to_others = held
        at = self._attr
        at[0, i] = a.x; at[1, i] = a.y; at[2, i] = a.v; at[3, i] = a.vy; at[4, i] = a.a

    def _step_walk(self, dt):
        """environment.py:316-323 with the walk on the device: one launch per run of device-plannable vehicles; a vehicle that
        needs the host (learned behaviours: their rollouts are launches of their own) is planned and moved by the Python
        path in between, exactly in its turn.  The three host random streams end up where the reference leaves them."""
        import random
        from . import mc_sim_highway as ms
        objs, n = self._objs, len(self._objs)
        lanes = self._tables.lane[0]
        self._mirror()
        self._mirror_live = True
        try:
            first = 0
            while first < n:
                vehicles, upload = self._walk_array()
                state = np.random.get_state()
                noise = np.random.normal(0., 1, min(4096, 48 * (n - first)))
                res = self.map.device_map.walk_step(vehicles, upload, self._ids_p, noise, first, n, dt, ms._peek_seed())
                np.random.set_state(state)
                if res.status:
                    raise EnvError(-3 if res.status in (101, 300) else -1, "device walk of the outer step failed (status %d)" % res.status)
                if res.noise_used:
                    np.random.normal(0., 1, res.noise_used)
                if res.random_used:
                    random.seed(res.random_seed)
                    random.random()
                ms._advance_seeds(res.mobil_calls)
                stop = res.stopped_at if res.stopped_at >= 0 else n
                for i in range(first, stop):
                    if lanes[i] >= 0:
                        self._walk_take(i)
                if stop < n:                                  # this vehicle's plan needs the host
                    a = objs[stop]
                    a.step(a.plan(self), dt)
                    i, at = stop, self._attr
                    at[0, i] = a.x; at[1, i] = a.y; at[2, i] = a.v; at[3, i] = a.vy; at[4, i] = a.a
                first = stop + 1
            for a in list(self.agents.values()):
                self.set_flag(a)
        finally:
            self._mirror_live = False
        self.match_agents_on_corridor()
        self.sim_time += dt

    def step(self, dt):                                      # environment.py:316-323
        """Plan and move the vehicles one after the other against the lane vectors of the last match, then re-match."""
        if self.walk_on_device and self._tables is not None and hasattr(self.map.device_map, "walk_step"):
            return self._step_walk(dt)
        self._mirror()
        self._mirror_live = True
        try:
            for a in list(self.agents.values()):
                a.step(a.plan(self), dt)
                i, at = self._slot[a.id], self._attr
                at[0, i] = a.x; at[1, i] = a.y; at[2, i] = a.v; at[3, i] = a.vy; at[4, i] = a.a
                self.set_flag(a)
        finally:
            self._mirror_live = False
        self.match_agents_on_corridor()
        self.sim_time += dt
