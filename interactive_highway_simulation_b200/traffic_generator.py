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


def _at_least(mean, var, lo):
    return max(lo, np.random.normal(mean, var))


def random_idm_p(desired_v):
    acc_max = np.random.uniform(1.8, 2.2)
    acc_min = np.random.uniform(-2., -4.)
    min_dis = _at_least(4, 1, 2)
    acc_com = np.random.normal(-2.0, 0.1)
    thw = _at_least(1.5, 0.1, 1.0)
    vd = desired_v + np.random.normal(0., 0.5)
    return {"acc_max": acc_max, "acc_min": acc_min, "min_dis": min_dis, "acc_com": acc_com, "thw": thw, "vd": vd}


def random_idm_p_truck(desired_v):
    acc_max = np.random.uniform(0.8, 1.1)
    acc_min = np.random.uniform(-1.0, -2.0)
    min_dis = _at_least(4, 1, 2)
    acc_com = np.random.normal(-2.0, 0.1)
    thw = _at_least(1.2, 0.1, 1.0)
    vd = desired_v + np.random.normal(0., 1.)
    return {"acc_max": acc_max, "acc_min": acc_min, "min_dis": min_dis, "acc_com": acc_com, "thw": thw, "vd": vd}


def random_idm_p_merging(desired_v):
    acc_max = np.random.uniform(2.0, 2.5)
    acc_min = np.random.uniform(-3., -4.)
    min_dis = _at_least(4, 1, 2)
    acc_com = np.random.normal(-2.0, 0.1)
    thw = _at_least(1.5, 0.1, 1.0)
    vd = desired_v + np.random.normal(0., 0.5)
    return {"acc_max": acc_max, "acc_min": acc_min, "min_dis": min_dis, "acc_com": acc_com, "thw": thw, "vd": vd}


YIELDING_PARAMETER = [1.70968573, 0.00582201, 0.40742229, 1.02390871, -0.32091885, 1.48127062, -0.69499426, -0.34547462,
                      -0.56578542, -1.911662, -0.03998742, 0.02613949, -0.02392805]


def random_yielding_p():
    return [np.random.normal(p, abs(p) * 0.2) for p in YIELDING_PARAMETER]


def random_mobil_p():
    politeness_f = np.random.uniform(0, 1)
    delta_a = np.random.normal(0.5, 0.2)
    bias_a = delta_a + np.random.normal(0.1, 0.05)
    return {"politeness_factor": politeness_f, "delta_a": delta_a, "bias_a": bias_a}


class RandomTrafficGenerator:
    """random_traffic_generator.py:74-136"""

    def __init__(self):
        self.vehicle_id = 0
        self.shapes = [[4.5, 1.9], [4.8, 2], [4.9, 2.1], [5.2, 2.2], [5.9, 2.4]]
        self.truck_shape = [[11.5, 2.5], [12, 2.5], [10, 2.5]]

    def random_agents_on_map(self, lanes, parameter, unsafe_lc_ratio=0., truck_ratio=0.):
        """lanes: {id: StraightLane} in the map's insertion order -> (agents on merging lanes, agents on main lanes)"""
        agents_merging, agents_main = [], []
        truck_lane = truck_lane_id(lanes)
        v_limit_candidates = [c.v_limit for c in lanes.values() if c.type == "main"]
        for i, c in lanes.items():
            if c.type == "main":
                agents_main.extend(self.random_agents_on_lane(c, parameter[i], v_limit_candidates, "idm", unsafe_lc_ratio=unsafe_lc_ratio,
                                                              truck_ratio=truck_ratio if i == truck_lane else 0.))
            if c.type == "merging":
                agents_merging.extend(self.random_agents_on_lane(c, parameter[i], v_limit_candidates, "merging_closest_gap_policy"))
        return agents_merging, agents_main

    def random_agents_on_lane(self, corridor, params, v_limit_candidates, behavior_model, unsafe_lc_ratio=0., truck_ratio=0.):
        vehicles = []
        s_range = [0, corridor.path_length]
        num_vehicles = 1e10
        if params.s_range:
            s_range = [max(s_range[0], params.s_range[0]), min(s_range[1], params.s_range[1])]
        if params.num_vehicles:
            num_vehicles = params.num_vehicles
        x_frenet = np.random.uniform(s_range[0], 0.3 * corridor.v_limit * params.thw_mean)
        while x_frenet < s_range[1] and len(vehicles) < num_vehicles:
            y_frenet = np.random.normal(0, 0.2)
            v = np.random.normal(corridor.v_limit + params.mean_v_diff_to_limit, params.v_var)
            x, y, yaw = corridor.to_global_coordinate(x_frenet, y_frenet)
            random_num = np.random.uniform(0, 1)
            if random_num < truck_ratio:
                l, w = random.choice(self.truck_shape)
                idm_p = random_idm_p_truck(min(v_limit_candidates))
            else:
                l, w = random.choice(self.shapes)
                idm_p = random_idm_p(random.choice(v_limit_candidates))
            change_intention_threshold = np.random.uniform(0.4, 0.7)
            b_model = behavior_model
            if behavior_model == "idm":     # 10 % plain idm
                random_num = np.random.uniform(0, 1)
                if 1 - unsafe_lc_ratio >= random_num > 0.1:
                    b_model = "idm_mobil_lane_change_safe"
                elif random_num > 1 - unsafe_lc_ratio:
                    b_model = "idm_mobil_lane_change"
            yielding_p = random_yielding_p()
            mobil_p = random_mobil_p()
            yielding_seed = int(random.random() * 100)            # Agent.__init__ (agent.py:73-76)
            if "merging" in behavior_model:
                idm_p = random_idm_p_merging(random.choice(v_limit_candidates))
            vehicles.append({"id": int(self.vehicle_id), "x": float(x), "y": float(y), "v": float(v), "vy": 0.0,
                             "yaw": float(np.rad2deg(np.deg2rad(yaw))), "a": 0.0, "l": float(l), "w": float(w),
                             "idm_p": {k: float(val) for k, val in idm_p.items()},
                             "mobil_param": {k: float(val) for k, val in mobil_p.items()},
                             "yielding_param": [float(p) for p in yielding_p],
                             "behavior_model": b_model,
                             "change_intention_threshold": float(change_intention_threshold),
                             "random_yielding_seed": yielding_seed})
            x_frenet += np.random.uniform(corridor.v_limit * (params.thw_mean - params.thw_var),
                                          corridor.v_limit * (params.thw_mean + params.thw_var)) + l
            self.vehicle_id += 1
        return vehicles


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
