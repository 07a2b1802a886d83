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
        """Hysteresis around the first estimate p0: a yielding vehicle stops yielding once the probability has fallen to
        threshold * p0, a non-yielding one starts once the complement has fallen to threshold * (1 - p0)."""
        p0, k = self.initial_probability, self.change_intention_threshold
        if self.yielding:
            self.yielding = not (new_probability <= k * p0)
        else:
            self.yielding = (1 - new_probability) <= k * (1 - p0)

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
        """Probability that `ego` (on the main lane, own leader `ego_f` or None) yields to the vehicle `idm_match` that
        wants to move in from the side lane in `direction`; 0 once that vehicle is completely behind (type.py:111-137)."""
        half_lengths = 0.5 * (ego.l + idm_match.l)
        if idm_match.dis + half_lengths < 0:
            return 0.
        gap = idm_match.dis - half_lengths
        own_v = ego.v                      # the reference feeds the yielding vehicle's speed as BOTH v_0 and v_k (type.py:122)
        v_floor = max(own_v, 0.0001)
        border = env.dis_to_lane_border(idm_match.id, direction)
        to_end = env.corridor_for_agent(idm_match.id).distance_to_corridor_end([idm_match.x, idm_match.y])
        thw, d_thw = thw_and_ttc(gap, own_v, idm_match.v)
        return self.logistic_function((ego_f.dis if ego_f else 200,          # s_diff_kf_f   gap to the own leader
                                       gap,                                  # s_diff_0_k
                                       own_v,                                # v_0
                                       border,                               # d_0_lat       its hull to the lane border
                                       idm_match.vy,                         # v_0_lat
                                       own_v,                                # v_k
                                       ego_f.v if ego_f else 20,             # v_kf
                                       thw, d_thw,
                                       to_end,                               # d_to_merging_lane_end
                                       to_end / v_floor,                     # t_to_merging_lane_end
                                       to_end * (1 - own_v / v_floor)))      # predicted_remaining_length
