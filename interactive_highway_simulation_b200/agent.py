"""Host mirror of the reference's outer-loop vehicle (agent.py:8-208) and its evaluation statistics (agent.py:134-173,
265-496): same constructor order, attribute names, yaml dict format (`create_agent_from_dict` / `to_dict`), behaviour
dispatch (`plan`) and state update (`step`).  `step` moves the vehicle along the corridor centre line on the device
(`mce_step_along_path`); `plan` runs the behaviours of behaviors.py, whose decisions come from the rollout kernels.
No rendering state (the reference's shapely polygon per vehicle is only drawn).
"""
import os
import random

import numpy as np

from . import behaviors
from .learned_model import LearnedLaneChangeModel, LearnedMergingModel
from .outer_env import IdmMatch, MatchedAgent, Neighbor, SimpleIdmMatch  # noqa: F401  (agent.py:211-253 live with the environment)
from .type import DecoupledAction, IdmParameter, MOBILParameter, RSSParameter, YieldingModel
from .utility import step_along_path, vehicle_pose_to_rect


def create_agent_from_dict(data):                            # agent.py:8-28
    p, m = data["idm_p"], data["mobil_param"]
    agent = Agent(data["id"], data["x"], data["y"], data["v"], data["yaw"], data["a"], data["l"], data["w"],
                  IdmParameter(p["acc_max"], p["acc_min"], p["min_dis"], p["acc_com"], p["thw"], p["vd"]),
                  data["yielding_param"], MOBILParameter(m["politeness_factor"], m["delta_a"], m["bias_a"]),
                  data["behavior_model"], data["change_intention_threshold"], data["random_yielding_seed"])
    if "vy" in data:
        agent.set_lateral_velocity(data["vy"])
    return agent


def create_agents_from_dicts(agents_dicts):                  # agent.py:31-35
    return [create_agent_from_dict(a) for a in agents_dicts]


def create_dicts_from_agents(agents):                        # agent.py:38-42
    return [a.to_dict() for a in agents]


class MergingFlag:                                           # agent.py:256-266
    def __init__(self):
        self.finish_merging, self.t_finish_merging, self.once_fallback, self.flag_freeze = False, None, False, False


class ExitFlag:                                              # agent.py:269-279
    def __init__(self):
        self.finish_exit, self.t_finish_exit, self.once_fallback, self.flag_freeze = False, None, False, False


class Agent:
    def __init__(self, idx, x, y, v, yaw, a, l, w, idm_param, yielding_param, mobil_param, behavior_model,
                 change_intention_threshold=0.625, random_yielding_seed=None):                      # agent.py:45-96
        self.id, self.x, self.y, self.v, self.vy = idx, x, y, v, 0.
        self.yaw = np.deg2rad(yaw)
        self.a, self.l, self.w = a, l, w
        self.corridor_history = []
        self.acceleration_history = [a]
        self.state_history = [[x, y, v, a]]
        self.hull = vehicle_pose_to_rect(x, y, yaw, l, w)    # sic: degrees here, radians in step() (agent.py:64,205)
        self.rss_param = RSSParameter()
        self.idm_param, self.yielding_param, self.mobil_param = idm_param, yielding_param, mobil_param
        self.behavior_model = behavior_model
        self.behavior_model_history = [behavior_model]
        self.change_intention_threshold = change_intention_threshold
        if random_yielding_seed and random_yielding_seed >= 0:
            self.random_yielding_seed = random_yielding_seed
        else:
            self.random_yielding_seed = int(random.random() * 100)
        self.yielding_model = YieldingModel(yielding_param) if yielding_param is not None else YieldingModel()
        self.yielding_intention_to_others = {}
        self._models_for(behavior_model)
        self.commit_lane_change = "not_decided"
        self.commit_keep_lane_step = 0
        self.target_lane_id = -1

    def _models_for(self, model):                            # agent.py:82-91 / 106-114
        if "merging" in model or "exit" in model:
            self.learned_merging_model = LearnedMergingModel()
            self.merging_flag = MergingFlag()
            self.exit_flag = ExitFlag()
        if model == "lane_change_learned":
            self.learned_lane_change_model = LearnedLaneChangeModel()
            self.current_decision = "Initial"
        if "lane_change" in model:
            self.rss_param = RSSParameter(0.4, 0.7, -6, -6, 1.8, 1.2)

    def get_position(self):
        return [self.x, self.y]

    def set_lateral_velocity(self, vy):
        self.vy = vy

    def set_behavior_model(self, model):                     # agent.py:104-115
        self.behavior_model = model
        self._models_for(model)
        self.behavior_model_history.append(model)

    def reset_behavior_model(self, model):                   # agent.py:117-119
        self.set_behavior_model(model)
        self.behavior_model_history = [model]

    def to_dict(self):                                       # agent.py:121-138
        ip, mp = self.idm_param, self.mobil_param
        return {"id": int(self.id), "x": float(self.x), "y": float(self.y), "v": float(self.v), "vy": float(self.vy),
                "yaw": float(np.rad2deg(self.yaw)), "a": float(self.a), "l": float(self.l), "w": float(self.w),
                "idm_p": {"acc_max": float(ip.acc_max), "acc_min": float(ip.acc_min), "min_dis": float(ip.min_dis),
                          "acc_com": float(ip.acc_com), "thw": float(ip.thw), "vd": float(ip.vd)},
                "mobil_param": {"politeness_factor": float(mp.politeness_factor), "delta_a": float(mp.delta_a),
                                "bias_a": float(mp.bias_a)},
                "yielding_param": [float(p) for p in self.yielding_model.param],
                "behavior_model": self.behavior_model,
                "change_intention_threshold": float(self.change_intention_threshold),
                "random_yielding_seed": int(self.random_yielding_seed)}

    # ---- evaluation statistics (agent.py:140-183) ----
    def had_harsh_brake(self):
        h = self.acceleration_history
        return any(h[i] < -5 and h[i + 1] < -5 and i >= 3 for i in range(len(h) - 1))

    def was_merging_vehicle(self):
        return any("merging" in b for b in self.behavior_model_history)

    def num_lane_change(self):
        h = self.corridor_history
        return sum(1 for i in range(len(h) - 1) if h[i + 1] != h[i]) if len(h) > 1 else 0

    def average_utility(self):
        v_history = [s[2] for s in self.state_history]
        if len(v_history) == 0:
            return 0.
        vd = self.idm_param.vd
        u_sum = 0
        for v in v_history:
            u_sum += v / vd if v <= vd else max(1. - (v / vd - 1.), 0.9)
        return u_sum / len(v_history)

    def minimum_comfort(self):
        if len(self.acceleration_history) == 0:
            return 0.
        abs_acc = sorted((abs(a) for a in self.acceleration_history), reverse=True)
        n = max(int(0.2 * len(abs_acc)), 1)
        c_sum = 0
        for i in range(n):
            c_sum += 1 + abs_acc[i] / (-8.)
        return c_sum / n

    # ---- behaviour dispatch and state update (agent.py:175-208) ----
    def plan(self, observed_env):
        b = self.behavior_model
        if b == "idm":
            return behaviors.idm_behavior(observed_env, self)
        if b == "idm_mobil_lane_change":
            return behaviors.lane_change_behavior_mobil(observed_env, self, False)
        if b == "idm_mobil_lane_change_safe":
            return behaviors.lane_change_behavior_mobil(observed_env, self, True)
        if b == "merging_learned":
            return behaviors.cloned_merging_behavior(observed_env, self)
        if b == "merging_closest_gap_policy":
            return behaviors.merging_behavior_closest_gap(observed_env, self)
        if b == "lane_change_learned":
            return behaviors.lane_change_behavior_learned(observed_env, self)
        if b == "exit":
            return behaviors.exit_behavior_closest_gap(observed_env, self)
        if b == "exit_learned":
            return behaviors.cloned_exit_behavior(observed_env, self)

    def step(self, action, dt):
        if isinstance(action, DecoupledAction):
            self.x, self.y, self.v, self.yaw = step_along_path(action.ref_path, self, action.ax, action.vy, dt)
            self.a = 0. if self.v == 0 else action.ax
            self.vy = action.vy
            self.acceleration_history.append(self.a)
            self.hull = vehicle_pose_to_rect(self.x, self.y, self.yaw, self.l, self.w)
            self.state_history.append([self.x, self.y, self.v, self.a])


class _Statistics:
    """the members MergingAndExitStatistics and LaneChangeStatistics share (agent.py:282-496)"""
    _extra = ()

    def __init__(self):
        self.num_fallback = 0
        self.epoch_and_ego_id_history = []
        self.utility_ego, self.comfort_ego, self.utility_other, self.comfort_other = [], [], [], []

    @staticmethod
    def _mean(xs):
        return -1 if len(xs) == 0 else np.mean([x for x in xs if x != -10])

    def average_utility_ego(self):
        return self._mean(self.utility_ego)

    def average_comfort_ego(self):
        return self._mean(self.comfort_ego)

    def average_utility_other(self):
        return self._mean(self.utility_other)

    def average_comfort_other(self):
        return self._mean(self.comfort_other)

    def _files(self):
        return (("utility_ego", "_utility_ego.npy"), ("comfort_ego", "_comfort_ego.npy"), ("utility_other", "_utility_others.npy"),
                ("comfort_other", "_comfort_others.npy")) + self._extra + (("epoch_and_ego_id_history", "_epoch_and_ego_id_history.npy"),)

    def remove_invalid_data(self, index_list):               # agent.py:325-331 / 432-438
        """Drop the samples at the given positions from every per-sample list.  As in the reference the ego-utility list
        is cut first and the other lists are then read only up to ITS new length (its loops all range over
        len(self.utility_ego)), so removing k samples also drops the last k entries of the other lists."""
        self.utility_ego = [v for i, v in enumerate(self.utility_ego) if i not in index_list]
        keep = [i for i in range(len(self.utility_ego)) if i not in index_list]
        for attr, _ in self._files()[1:]:
            values = getattr(self, attr)
            setattr(self, attr, [values[i] for i in keep])

    def save(self, path, name):
        for attr, suffix in self._files():
            np.save(os.path.join(path, name + suffix), np.array(getattr(self, attr)))

    def load(self, path, name):
        for attr, suffix in self._files():
            setattr(self, attr, list(np.load(os.path.join(path, name + suffix))))
        for attr in ("utility_ego", "comfort_ego", "utility_other", "comfort_other") + tuple(a for a, _ in self._extra if a == "total_t"):
            setattr(self, attr, [x for x in getattr(self, attr) if x != -10])
        self.epoch_and_ego_id_history = [x for x in self.epoch_and_ego_id_history if x[0] != -1]
        assert len(self.utility_ego) == len(self.epoch_and_ego_id_history)

    def on_lane_number(self, num):
        id_range = {1: range(10), 2: range(10, 17), 3: range(19, 100)}.get(num, [])
        keep = [i for i in range(len(self.epoch_and_ego_id_history)) if self.epoch_and_ego_id_history[i][1] in id_range]
        for attr in ("utility_ego", "comfort_ego", "utility_other", "comfort_other"):
            setattr(self, attr, [getattr(self, attr)[i] for i in keep])

    @staticmethod
    def _last(xs):
        return round(xs[-1], 3) if len(xs) else 0.


class MergingAndExitStatistics(_Statistics):                 # agent.py:282-394
    _extra = (("total_t", "_finish_exit_t.npy"),)

    def __init__(self):
        super().__init__()
        self.num_total_agents = 0
        self.num_finish = 0
        self.total_t = []

    def average_t_finish_exit(self):
        return self._mean(self.total_t)

    def __repr__(self, average=False):
        if not average:
            return "{} agents, num finish {}, num fallback {}, average t {}, utility ego {}, comfort ego {} utility obj {}, comfort obj {}".format(
                self.num_total_agents, self.num_finish, self.num_fallback, round(self.total_t[-1], 3) if self.total_t else 0.,
                self._last(self.utility_ego), self._last(self.comfort_ego), self._last(self.utility_other), self._last(self.comfort_other))
        return ("{} agents, num finish {}, num fallback {}, average t {}, average utility ego {}, average comfort ego {}, "
                "average utility obj {}, average comfort obj {}").format(
            self.num_total_agents, self.num_finish, self.num_fallback, round(self.average_t_finish_exit(), 3),
            round(self.average_utility_ego(), 3), round(self.average_comfort_ego(), 3), round(self.average_utility_other(), 3),
            round(self.average_comfort_other(), 3))


class LaneChangeStatistics(_Statistics):                     # agent.py:397-496
    _extra = (("lane_id_history", "_lane_id_history.npy"),)

    def __init__(self):
        super().__init__()
        self.num_lane_change = 0
        self.lane_id_history = []

    def __repr__(self, average=False):
        head = str(len(self.utility_ego))
        if not average:
            return head + " sims, num lane change {}, num fallback {}, utility ego {}, comfort ego {} utility obj {}, comfort obj {}".format(
                self.num_lane_change, self.num_fallback, self._last(self.utility_ego), self._last(self.comfort_ego),
                self._last(self.utility_other), self._last(self.comfort_other))
        return head + (" sims, num lane change {}, num fallback {}, average utility ego {}, average comfort ego {}, "
                       "average utility obj {}, average comfort obj {}").format(
            self.num_lane_change, self.num_fallback, round(self.average_utility_ego(), 3), round(self.average_comfort_ego(), 3),
            round(self.average_utility_other(), 3), round(self.average_comfort_other(), 3))
