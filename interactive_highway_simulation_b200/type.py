"""Host mirror of the reference's parameter / yielding classes (type.py:9-137): DecoupledAction, IdmParameter,
RSSParameter, MOBILParameter, YieldingIntention, YieldingModel.  Same constructor orders, defaults, attribute names and
use of the process-global `random` stream as the reference (YieldingIntention re-seeds it, type.py:75-77).  The
geometric yielding features (distance of the merging vehicle's hull to its lane border, distance to the lane end) come
from the device (`Environment.dis_to_lane_border`, `Corridor.distance_to_corridor_end` -> include/mcsim_env.h).
"""
import math
import random

from .utility import thw_and_ttc


class DecoupledAction:                                       # type.py:9-15
    def __init__(self, ax, vy, ref_path):
        self.ax, self.vy, self.ref_path = ax, vy, ref_path


class Trajectory:                                            # type.py:18-21 (accepted by Agent.step, not implemented there either)
    def __init__(self, trajectory, time_step):
        self.trajectory, self.time_step = trajectory, time_step


class IdmParameter:                                          # type.py:24-33
    def __init__(self, acc_max=1.4, acc_min=-4., min_dis=2, acc_comfort=-2, thw=1.5, vd=None):
        self.acc_max, self.acc_min, self.min_dis, self.acc_com, self.thw, self.vd = acc_max, acc_min, min_dis, acc_comfort, thw, vd


class RSSParameter:                                          # type.py:36-43
    def __init__(self, r_time_ego=0.4, r_time_other=0.7, a_min=-8, a_max=-10, soft_acc=1.8, soft_dcc=1.2):
        self.r_t_ego, self.r_t_other, self.a_min, self.a_max = r_time_ego, r_time_other, a_min, a_max
        self.soft_acc, self.soft_dcc = soft_acc, soft_dcc


class MOBILParameter:                                        # type.py:46-50
    def __init__(self, politeness_factor=0.9, delta_a=0.5, bias_a=0.6):
        self.politeness_factor, self.delta_a, self.bias_a = politeness_factor, delta_a, bias_a


class YieldingIntention:                                     # type.py:70-90
    def __init__(self, idx, initial_probability, change_intention_threshold=0.625, random_intention_seed=None):
        self.id = idx
        self.initial_probability = initial_probability
        self.change_intention_threshold = change_intention_threshold
        if random_intention_seed:
            random.seed(random_intention_seed)
        self.yielding = random.random() <= initial_probability

    def update_intention(self, new_probability):
        if not self.yielding:
            if (1 - new_probability) <= self.change_intention_threshold * (1 - self.initial_probability):
                self.yielding = True
        else:
            if new_probability <= self.change_intention_threshold * self.initial_probability:
                self.yielding = False

    def __repr__(self):
        return "Yielding {}, initial_p: {}".format(self.yielding, self.initial_probability)


DEFAULT_YIELDING_PARAM = (1.70968573, 0.00582201, 0.40742229, 1.02390871, -0.32091885, 1.48127062, -0.69499426, -0.34547462,
                          -0.56578542, -1.911662, -0.03998742, 0.02613949, -0.02392805)


class YieldingModel:                                         # type.py:93-137
    def __init__(self, param=DEFAULT_YIELDING_PARAM):
        self.param = param

    def logistic_function(self, f):
        """f: the 12 features in the order of YieldingFeature (type.py:53-67); left-to-right sum like the reference"""
        p = self.param
        r = p[0]
        for k in range(12):
            r = r + p[k + 1] * f[k]
        r = max(min(100, r), -100)
        return 1 / (1 + math.exp(-r))

    def yielding_probability(self, idm_match, ego, ego_f, env, direction):
        if idm_match.dis + 0.5 * (ego.l + idm_match.l) < 0:
            return 0.
        v_kf = ego_f.v if ego_f else 20
        v_k = ego.v
        s_diff_kf_f = ego_f.dis if ego_f else 200
        s_diff_0_k = idm_match.dis - 0.5 * (ego.l + idm_match.l)
        v_0 = ego.v                                          # sic (type.py:122)
        d_0_lat = env.dis_to_lane_border(idm_match.id, direction)
        v_0_lat = idm_match.vy
        corridor_merging_vehicle = env.corridor_for_agent(idm_match.id)
        d_to_merging_lane_end = corridor_merging_vehicle.distance_to_corridor_end([idm_match.x, idm_match.y])
        t_to_merging_lane_end = d_to_merging_lane_end / max(v_0, 0.0001)
        predicted_remaining_length = d_to_merging_lane_end * (1 - v_k / max(v_0, 0.0001))
        thw, d_thw = thw_and_ttc(idm_match.dis - 0.5 * (ego.l + idm_match.l), ego.v, idm_match.v)
        return self.logistic_function((s_diff_kf_f, s_diff_0_k, v_0, d_0_lat, v_0_lat, v_k, v_kf, thw, d_thw,
                                       d_to_merging_lane_end, t_to_merging_lane_end, predicted_remaining_length))
