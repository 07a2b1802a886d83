"""The reference's synthetic-scene generator for straight-lane maps (random_traffic_generator.py:5-136 together with the
geometry it touches: type.py LineSimplified.get_point_at_length / get_yaw_at_length / to_global_coordinate :255-310,
utility.py offset_point :83-85, environment.py Map.truck_lane_id :44-56, agent.py Agent.__init__ :73-76).

Same two global random streams in the same order — numpy's for the continuous quantities, Python's `random` for the
discrete choices and the per-agent yielding seed — so that `np.random.seed(s); random.seed(s)` followed by
`RandomTrafficGenerator().random_agents_on_map(...)` returns exactly the agents the reference generates for that seed
(tests/test_traffic_generator.py, against records written by the reference's own module).  Agents are plain dicts in the
reference's `Agent.to_dict()` layout (agent.py:116-132).  Lanes are two-point polylines (every map of the reference's
evaluation scripts is); the only shapely call on this path, `LineString.length`, is the Euclidean segment length.
"""
import math
import random

import numpy as np


class RandomDensityAndVelocityParameter:
    """random_traffic_generator.py:57-71"""

    def __init__(self, thw_mean, thw_var, mean_v_diff_to_limit, v_var, s_range=None, num_vehicles=None):
        self.thw_mean = thw_mean
        self.thw_var = thw_var
        self.mean_v_diff_to_limit = mean_v_diff_to_limit
        self.v_var = v_var
        self.s_range = s_range
        self.num_vehicles = num_vehicles


class StraightLane:
    """One corridor of a map: id, type ("main" / "merging" / "exit"), two-point left / right / centre lines, speed limit,
    neighbour ids (type.py Corridor :140-176 reduced to what the generator reads)."""

    def __init__(self, idx, typ, left, right, center, v_limit, left_id=None, right_id=None):
        self.id, self.type, self.left, self.right, self.center, self.v_limit = idx, typ, left, right, center, v_limit
        self.left_id, self.right_id = left_id, right_id
        # LineSimplified.__init__ (type.py:257-272)
        p0, p1 = center[0], center[1]
        self.dis_list = [0, math.sqrt((p0[0] - p1[0]) ** 2 + (p0[1] - p1[1]) ** 2)]
        yaw = np.arctan2(p1[1] - p0[1], p1[0] - p0[0])
        self.yaw_list = [yaw, yaw]
        self.path_length = math.sqrt((p1[0] - p0[0]) * (p1[0] - p0[0]) + (p1[1] - p0[1]) * (p1[1] - p0[1]))

    def get_yaw_at_length(self, length):
        index = np.searchsorted(self.dis_list, length)
        if index >= len(self.yaw_list):
            return self.yaw_list[-1]
        return self.yaw_list[index]

    def get_point_at_length(self, length):
        path = self.center
        if length == 0:
            return path[0]
        if length < self.path_length:
            index = np.searchsorted(self.dis_list, length)
            seg_dis = self.dis_list[index] - self.dis_list[index - 1]
            ratio = (self.dis_list[index] - length) / seg_dis
            return np.array(path[index]) - ratio * (np.array(path[index]) - np.array(path[index - 1]))
        dis_end = self.dis_list[-1] - self.dis_list[-2]
        ratio = (length - self.path_length) / dis_end
        return np.array(path[-1]) + ratio * (np.array(path[-1]) - np.array(path[-2]))

    def to_global_coordinate(self, l, d):
        yaw = self.get_yaw_at_length(l)
        p = self.get_point_at_length(l)
        return p[0] - d * math.sin(yaw), p[1] + d * math.cos(yaw), yaw


def truck_lane_id(lanes):
    """Map.truck_lane_id (environment.py:44-56); `lanes` = {id: StraightLane} in insertion order."""
    by_type = {}
    for c in lanes.values():
        by_type.setdefault(c.type, []).append(c)
    if "merging" in by_type:
        return by_type["merging"][0].left_id
    if "exit" in by_type:
        return by_type["exit"][0].left_id
    for i, c in lanes.items():
        if c.right_id is None:
            return i
    return None


# ---- parameter draws.  What matters for reproducing the reference's scenes is WHICH global stream each quantity comes
# from and in which order; the tables below carry the distributions, the functions only walk them in that order.

# IDM families (random_traffic_generator.py:8-35): acc_max ~ U, acc_min ~ U, then min_dis ~ max(2, N(4, 1)),
# acc_com ~ N(-2, 0.1), thw ~ max(1, N(thw_mean, 0.1)), vd = desired + N(0, vd_sigma) — six numpy draws in this order
IDM_FAMILIES = {
    "car": {"acc_max": (1.8, 2.2), "acc_min": (-2., -4.), "thw_mean": 1.5, "vd_sigma": 0.5},
    "truck": {"acc_max": (0.8, 1.1), "acc_min": (-1.0, -2.0), "thw_mean": 1.2, "vd_sigma": 1.},
    "merging": {"acc_max": (2.0, 2.5), "acc_min": (-3., -4.), "thw_mean": 1.5, "vd_sigma": 0.5},
}

# the 13 weights of the reference's yielding model (random_traffic_generator.py:39-41); a vehicle gets each one
# perturbed by 20 % (one numpy normal draw per weight, in order)
YIELDING_PARAMETER = [1.70968573, 0.00582201, 0.40742229, 1.02390871, -0.32091885, 1.48127062, -0.69499426, -0.34547462,
                      -0.56578542, -1.911662, -0.03998742, 0.02613949, -0.02392805]


def draw_idm(family, desired_v):
    f = IDM_FAMILIES[family]
    p = {"acc_max": np.random.uniform(*f["acc_max"]), "acc_min": np.random.uniform(*f["acc_min"])}
    p["min_dis"] = max(2, np.random.normal(4, 1))
    p["acc_com"] = np.random.normal(-2.0, 0.1)
    p["thw"] = max(1.0, np.random.normal(f["thw_mean"], 0.1))
    p["vd"] = desired_v + np.random.normal(0., f["vd_sigma"])
    return p


def draw_yielding():
    return [np.random.normal(w, abs(w) * 0.2) for w in YIELDING_PARAMETER]


def draw_mobil():
    """politeness ~ U(0, 1), delta_a ~ N(0.5, 0.2), bias_a = delta_a + N(0.1, 0.05) (random_traffic_generator.py:47-51)"""
    p = {"politeness_factor": np.random.uniform(0, 1), "delta_a": np.random.normal(0.5, 0.2)}
    p["bias_a"] = p["delta_a"] + np.random.normal(0.1, 0.05)
    return p


def _plain(d):
    return {k: float(v) for k, v in d.items()}


class RandomTrafficGenerator:
    """random_traffic_generator.py:74-136: vehicles are placed lane by lane from the lane's start, each a random time
    headway ahead of the previous one; ids count up across lanes and calls."""

    CAR_SHAPES = [[4.5, 1.9], [4.8, 2], [4.9, 2.1], [5.2, 2.2], [5.9, 2.4]]
    TRUCK_SHAPES = [[11.5, 2.5], [12, 2.5], [10, 2.5]]

    def __init__(self):
        self.vehicle_id = 0
        self.shapes = self.CAR_SHAPES
        self.truck_shape = self.TRUCK_SHAPES

    def random_agents_on_map(self, lanes, parameter, unsafe_lc_ratio=0., truck_ratio=0.):
        """lanes: {id: StraightLane} in the map's insertion order -> (agents on merging lanes, agents on main lanes).
        Only the truck lane (Map.truck_lane_id) gets trucks; exit lanes get nobody."""
        speed_limits = [lane.v_limit for lane in lanes.values() if lane.type == "main"]
        trucks_on = truck_lane_id(lanes)
        on_ramp, on_main = [], []
        for lane_id, lane in lanes.items():
            if lane.type == "main":
                on_main += self.random_agents_on_lane(lane, parameter[lane_id], speed_limits, "idm", unsafe_lc_ratio,
                                                      truck_ratio if lane_id == trucks_on else 0.)
            elif lane.type == "merging":
                on_ramp += self.random_agents_on_lane(lane, parameter[lane_id], speed_limits, "merging_closest_gap_policy")
        return on_ramp, on_main

    def random_agents_on_lane(self, corridor, params, v_limit_candidates, behavior_model, unsafe_lc_ratio=0., truck_ratio=0.):
        first, last = 0, corridor.path_length
        if params.s_range:
            first, last = max(first, params.s_range[0]), min(last, params.s_range[1])
        room = params.num_vehicles if params.num_vehicles else 1e10
        gap_lo = corridor.v_limit * (params.thw_mean - params.thw_var)
        gap_hi = corridor.v_limit * (params.thw_mean + params.thw_var)
        out = []
        s = np.random.uniform(first, 0.3 * corridor.v_limit * params.thw_mean)
        while s < last and len(out) < room:
            vehicle = self._spawn(corridor, params, s, v_limit_candidates, behavior_model, unsafe_lc_ratio, truck_ratio)
            out.append(vehicle)
            s += np.random.uniform(gap_lo, gap_hi) + vehicle["l"]
            self.vehicle_id += 1
        return out

    def _spawn(self, lane, params, s, speed_limits, base_model, unsafe_lc_ratio, truck_ratio):
        """One vehicle at arc length `s`.  Draw order (numpy unless noted): lateral offset, speed, truck?, shape (`random`),
        IDM (cars pick their desired-speed limit with `random`), intention threshold, behaviour model (main lanes only),
        13 yielding weights, 3 MOBIL values, yielding seed (`random`), and for ramps a second IDM set that replaces the first."""
        lateral = np.random.normal(0, 0.2)
        speed = np.random.normal(lane.v_limit + params.mean_v_diff_to_limit, params.v_var)
        x, y, yaw = lane.to_global_coordinate(s, lateral)
        if np.random.uniform(0, 1) < truck_ratio:
            length, width = random.choice(self.truck_shape)
            idm = draw_idm("truck", min(speed_limits))
        else:
            length, width = random.choice(self.shapes)
            idm = draw_idm("car", random.choice(speed_limits))
        threshold = np.random.uniform(0.4, 0.7)
        model = base_model
        if base_model == "idm":
            # a tenth stay plain IDM; the rest change lanes by MOBIL, `unsafe_lc_ratio` of them without the RSS check
            u = np.random.uniform(0, 1)
            if u > 1 - unsafe_lc_ratio:
                model = "idm_mobil_lane_change"
            elif u > 0.1:
                model = "idm_mobil_lane_change_safe"
        yielding = draw_yielding()
        mobil = draw_mobil()
        yielding_seed = int(random.random() * 100)            # drawn by Agent.__init__ (agent.py:73-76)
        if "merging" in base_model:
            idm = draw_idm("merging", random.choice(speed_limits))
        return {"id": int(self.vehicle_id), "x": float(x), "y": float(y), "v": float(speed), "vy": 0.0,
                "yaw": float(np.rad2deg(np.deg2rad(yaw))), "a": 0.0, "l": float(length), "w": float(width),
                "idm_p": _plain(idm), "mobil_param": _plain(mobil), "yielding_param": [float(w) for w in yielding],
                "behavior_model": model, "change_intention_threshold": float(threshold),
                "random_yielding_seed": yielding_seed}


def evaluation_map(name, v_limit=100 / 3.6):
    """The map, per-lane generator parameters, unsafe-lane-change ratio and truck ratio of one of the reference's evaluation
    scripts: "lane_change" (evaluation_lane_change.py:9-44), "merging" (evaluation_merging.py:8-37), "exit"
    (evaluation_exit.py:9-43).  Returns (list of corridor dicts, {lane id: parameter tuple}, unsafe ratio, truck ratio)."""
    def lane(idx, typ, left, right, center, v, left_id, right_id):
        return {"id": idx, "type": typ, "left": left, "right": right, "center": center, "v_limit": v, "left_id": left_id, "right_id": right_id}
    if name == "lane_change":
        m = v_limit ** 2 / 4 + 20
        spec = [lane(0, "merging", [[-10, 3.5], [m + 10, 3.5]], [[-10, 0], [m, 0]], [[-10, 1.75], [m + 5, 1.75]], v_limit, 1, None),
                lane(1, "main", [[-200, 7], [500, 7]], [[-200, 3.5], [500, 3.5]], [[-200, 5.25], [500, 5.25]], v_limit, 2, 0),
                lane(2, "main", [[-200, 10.5], [500, 10.5]], [[-200, 7], [500, 7]], [[-200, 8.75], [500, 8.75]], v_limit + 20 / 3.6, 3, 1),
                lane(3, "main", [[-200, 10.5 + 3.5], [500, 10.5 + 3.5]], [[-200, 10.5], [500, 10.5]], [[-200, 8.75 + 3.5], [500, 8.75 + 3.5]],
                     v_limit + 30 / 3.6, None, 2)]
        params = {0: (1.0, 0.1, -2, 2, None, 2), 1: (1.8, 0.4, 0, 2, [0, 500], None), 2: (2.0, 1.0, 0, 2, [0, 500], None),
                  3: (2.0, 1.0, 0, 2, [0, 500], None)}
    elif name == "merging":
        m = v_limit * 3.6 * 2 + 20
        spec = [lane(0, "merging", [[0, 3.5], [m + 10, 3.5]], [[0, 0], [m, 0]], [[0, 1.75], [m + 5, 1.75]], v_limit, 1, None),
                lane(1, "main", [[-100, 7], [500, 7]], [[-100, 3.5], [500, 3.5]], [[-100, 5.25], [500, 5.25]], v_limit, 2, 0),
                lane(2, "main", [[-100, 10.5], [500, 10.5]], [[-100, 7], [500, 7]], [[-100, 8.75], [500, 8.75]], v_limit + 20 / 3.6, None, 1)]
        params = {0: (1.0, 0.1, -2, 2, None, 2), 1: (1.6, 0.4, 0, 2, [0, 400], None), 2: (2.0, 1.0, 0, 2, [0, 400], None)}
    elif name == "exit":
        spec = [lane(0, "exit", [[245, 3.5], [500, 3.5]], [[255, 0], [500, 0]], [[250, 1.75], [500 + 5, 1.75]], v_limit, 1, None),
                lane(1, "main", [[-200, 7], [500, 7]], [[-200, 3.5], [500, 3.5]], [[-200, 5.25], [500, 5.25]], v_limit, 2, 0),
                lane(2, "main", [[-200, 10.5], [500, 10.5]], [[-200, 7], [500, 7]], [[-200, 8.75], [500, 8.75]], v_limit + 20 / 3.6, 3, 1),
                lane(3, "main", [[-200, 10.5 + 3.5], [500, 10.5 + 3.5]], [[-200, 10.5], [500, 10.5]], [[-200, 8.75 + 3.5], [500, 8.75 + 3.5]],
                     v_limit + 30 / 3.6, None, 2)]
        params = {0: (1.0, 0.1, -2, 2, None, 0), 1: (1.6, 0.4, 0, 2, [0, 500], None), 2: (2.0, 1.0, 0, 2, [0, 500], None),
                  3: (2.0, 1.0, 0, 2, [0, 500], None)}
    else:
        raise ValueError("unknown evaluation map %r" % name)
    return spec, params, 0., 0.4


def lanes_from_spec(spec):
    return {c["id"]: StraightLane(c["id"], c["type"], c["left"], c["right"], c["center"], c["v_limit"], c["left_id"], c["right_id"])
            for c in spec}
