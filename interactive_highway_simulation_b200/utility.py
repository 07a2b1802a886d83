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
    """Corners [front-left, front-right, rear-right, rear-left] of the vehicle's rectangle."""
    c, s = math.cos(yaw), math.sin(yaw)
    along = (length / 2 * c, length / 2 * s)          # centre -> front bumper
    across = (width / 2 * s, width / 2 * c)           # (minus x, plus y) components of centre line -> left side
    front = (pos_x + along[0], pos_y + along[1])
    rear = (pos_x - along[0], pos_y - along[1])

    def corner(p, left):
        return [p[0] - across[0], p[1] + across[1]] if left else [p[0] + across[0], p[1] - across[1]]
    return [corner(front, True), corner(front, False), corner(rear, False), corner(rear, True)]


def _idm_interaction(p, v_behind, v_ahead, gap):
    """(desired gap of the vehicle behind / actual gap, at least 1 m)^2 — the IDM interaction term without its sign and
    scale; the desired gap is 0.7 (min_dis + v thw) + v dv / (2 sqrt(-acc_max acc_com)) (utility.py:35,42)."""
    wish = (0.7 * (p.min_dis + v_behind * p.thw) + v_behind * (v_behind - v_ahead) / (2 * math.sqrt(-p.acc_max * p.acc_com)))
    return (wish / max(1, gap)) ** 2


def idm_acc_fs_bs(fs, bs, ego_v, v_limit, idm_param, obey_v_limit=True):   # utility.py:52-80
    """IDM acceleration with leaders `fs` (the most restrictive one counts) and followers `bs` (the strongest push
    counts); `dis` of a match is the bumper-to-bumper gap, positive ahead, negative behind."""
    p = idm_param
    vd = min(p.vd, 1.1 * v_limit) if obey_v_limit else p.vd
    acc = p.acc_max * (1 - (ego_v / vd) ** 4)
    pull = [-p.acc_max * _idm_interaction(p, ego_v, f.v, f.dis) for f in fs]
    push = [p.acc_max * _idm_interaction(p, b.v, ego_v, -b.dis) for b in bs]
    interaction = 0
    if pull:
        interaction += min(pull)
    if push:
        interaction += max(push)
    return max(min(p.acc_max, acc + interaction), p.acc_min)


def idm_acc_f_b(fs, b, ego_v, v_limit, idm_param, obey_v_limit=True):  # utility.py:23-49: at most one follower
    return idm_acc_fs_bs(fs, [b] if b else [], ego_v, v_limit, idm_param, obey_v_limit)


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
    return ref_path.device_map.step_one(ref_path.lane_index, ego.x, ego.y, ego.v, ax, vy, dt)


def thw_and_ttc(d, ego_v, obj_v):                                     # utility.py:101-108
    """Time headway over the gap `d` and the closing speed relative to the own speed (speeds below 1 mm/s count as 1 mm/s)."""
    return d / max(ego_v, 1e-10), (ego_v - obj_v) / (ego_v if ego_v > 0 else 1e-3)


def rss_dis(vf, vb, rt, a_min, a_max):                                # utility.py:129-130
    """RSS gap: what the rear vehicle covers in its reaction time and braking at a_min, less the leader's braking path at a_max."""
    rear_path = vb * rt + vb * vb / (-2 * a_min)
    return max(rear_path - vf * vf / (-2 * a_max), 0.)


def idm_p_for_merging_exit_agents(random_seed, idm_p_origin, is_truck):   # utility.py:133-141
    """The softened IDM parameters a vehicle yields with.  Re-seeds the global `random` stream like the reference — callers
    downstream depend on where that leaves it."""
    random.seed(random_seed)
    u = random.random()
    softened = deepcopy(idm_p_origin)
    if is_truck:
        return softened
    softened.acc_max -= (0.6 + 0.4 * u)
    softened.acc_min += (0.8 + 0.4 * u)
    return softened
