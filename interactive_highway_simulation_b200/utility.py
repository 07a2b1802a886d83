"""Host mirror of the scalar helpers of the reference's utility.py that the OUTER loop evaluates once per vehicle and
step (utility.py:12-49,83-109,129-146).  The geometry of utility.py (step_along_path's projection) runs on the device
(`mce_step_along_path`, include/mcsim_env.h); the rollouts evaluate their own IDM / RSS as device functions
(csrc/rollout_kernel.cuh) — this file is only the reference's Python arithmetic around a single outer-loop vehicle,
kept operation for operation so the outer simulation reproduces the reference's trace.
"""
import math
import random
from copy import deepcopy


def distance(p1, p2):                                                 # utility.py:8-9
    return math.sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)


def vehicle_pose_to_rect(pos_x, pos_y, yaw, length, width):           # utility.py:12-20
    head = [pos_x + length / 2 * math.cos(yaw), pos_y + length / 2 * math.sin(yaw)]
    tair = [pos_x - length / 2 * math.cos(yaw), pos_y - length / 2 * math.sin(yaw)]
    head_left = [head[0] - width / 2 * math.sin(yaw), head[1] + width / 2 * math.cos(yaw)]
    head_right = [head[0] + width / 2 * math.sin(yaw), head[1] - width / 2 * math.cos(yaw)]
    tair_left = [tair[0] - width / 2 * math.sin(yaw), tair[1] + width / 2 * math.cos(yaw)]
    tair_right = [tair[0] + width / 2 * math.sin(yaw), tair[1] - width / 2 * math.cos(yaw)]
    return [head_left, head_right, tair_right, tair_left]


def idm_acc_f_b(fs, b, ego_v, v_limit, idm_param, obey_v_limit=True):  # utility.py:23-49
    acc_max, acc_min, min_dis = idm_param.acc_max, idm_param.acc_min, idm_param.min_dis
    acc_com, thw = idm_param.acc_com, idm_param.thw
    vd = min(idm_param.vd, 1.1 * v_limit) if obey_v_limit else idm_param.vd
    free_road_term = acc_max * (1 - (ego_v / vd) ** 4)
    d_term = 0
    d_terms_f = []
    for f in fs:
        d_desire = (0.7 * (min_dis + ego_v * thw) + ego_v * (ego_v - f.v) / (2 * math.sqrt(-acc_max * acc_com)))
        d_terms_f.append(-acc_max * (d_desire / max(1, f.dis)) ** 2)
    d_term += min(d_terms_f) if d_terms_f else 0
    if b:
        d_desire = (0.7 * (min_dis + b.v * thw) + b.v * (b.v - ego_v) / (2 * math.sqrt(-acc_max * acc_com)))
        d_term += acc_max * (d_desire / max(1, -b.dis)) ** 2
    acc = free_road_term + d_term
    return max(min(acc_max, acc), acc_min)


def idm_acc_fs_bs(fs, bs, ego_v, v_limit, idm_param, obey_v_limit=True):   # utility.py:52-80: several followers, the strongest push counts
    acc_max, acc_com = idm_param.acc_max, idm_param.acc_com
    d_terms_b = []
    for b in bs:
        d_desire = (0.7 * (idm_param.min_dis + b.v * idm_param.thw) + b.v * (b.v - ego_v) / (2 * math.sqrt(-acc_max * acc_com)))
        d_terms_b.append(acc_max * (d_desire / max(1, -b.dis)) ** 2)
    vd = min(idm_param.vd, 1.1 * v_limit) if obey_v_limit else idm_param.vd
    free_road_term = acc_max * (1 - (ego_v / vd) ** 4)
    d_term = 0
    d_terms_f = []
    for f in fs:
        d_desire = (0.7 * (idm_param.min_dis + ego_v * idm_param.thw) + ego_v * (ego_v - f.v) / (2 * math.sqrt(-acc_max * acc_com)))
        d_terms_f.append(-acc_max * (d_desire / max(1, f.dis)) ** 2)
    d_term += min(d_terms_f) if d_terms_f else 0
    d_term += max(d_terms_b) if d_terms_b else 0
    return max(min(acc_max, free_road_term + d_term), idm_param.acc_min)


def offset_point(p, yaw, offset):                                     # utility.py:83-85
    return [p[0] - offset * math.sin(yaw), p[1] + offset * math.cos(yaw)]


def extend_line(line, d):                                             # utility.py:111-115
    k = d / distance(line[-1], line[-2])
    return [list(p) for p in line] + [[line[-1][0] + k * (line[-1][0] - line[-2][0]), line[-1][1] + k * (line[-1][1] - line[-2][1])]]


def extend_line_both_sides(line, d):                                  # utility.py:118-126
    kf, kb = d / distance(line[0], line[1]), d / distance(line[-1], line[-2])
    return [[line[0][0] + kf * (line[0][0] - line[1][0]), line[0][1] + kf * (line[0][1] - line[1][1])]] + [list(p) for p in line] + \
        [[line[-1][0] + kb * (line[-1][0] - line[-2][0]), line[-1][1] + kb * (line[-1][1] - line[-2][1])]]


def step_along_path(ref_path, ego, ax, vy, dt):                       # utility.py:88-98 — one device launch
    """ref_path: the corridor whose centre line is the reference's `DecoupledAction.ref_path` (behaviors.py)."""
    x, y, v, yaw = ref_path.device_map.step_along_path([ref_path.lane_index], [ego.x], [ego.y], [ego.v], [ax], [vy], dt)
    return float(x[0]), float(y[0]), float(v[0]), float(yaw[0])


def thw_and_ttc(d, ego_v, obj_v):                                     # utility.py:101-108
    thw = d / max(ego_v, 1e-10)
    if ego_v > 0:
        ttc = (ego_v - obj_v) / ego_v
    else:
        ttc = (ego_v - obj_v) / 1e-3
    return thw, ttc


def rss_dis(vf, vb, rt, a_min, a_max):                                # utility.py:129-130
    return max(vb * rt + vb * vb / (-2 * a_min) - vf * vf / (-2 * a_max), 0.)


def idm_p_for_merging_exit_agents(random_seed, idm_p_origin, is_truck):   # utility.py:133-141 (re-seeds `random`)
    random.seed(random_seed)
    random_number = random.random()
    idm_p_yielding = deepcopy(idm_p_origin)
    if not is_truck:
        idm_p_yielding.acc_max -= (0.6 + 0.4 * random_number)
        idm_p_yielding.acc_min += (0.8 + 0.4 * random_number)
    return idm_p_yielding
