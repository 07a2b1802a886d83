"""The reference's marshalling layer (mcs_wrapper.py:8-283) on the device geometry of include/mcsim_env.h: same three
builders, same names, same arguments and return values

    merging_mcs_sim_from_env(env, agent, ego_corridor)      -> corridors, agents, possible_actions      (:24-75)
    exit_mcs_sim_from_env(env, agent, ego_corridor)         -> corridors, agents, possible_actions      (:78-190)
    lane_change_mcs_sim_from_env(env, agent, ego_corridor)  -> corridors, agents                        (:193-283)

`env` is an `outer_env.Environment`: the neighbour selection (FF / F / B / BB per side, front, follower, +-80 m on the
second lanes) reads the tables of the last match launch, and every XY -> Frenet conversion of one builder call — the
vehicles and the six corner points of each corridor — goes through ONE `mce_frenet` launch on the builder's reference
line (extend_line_both_sides(center, 1000), simplified).  The `vd` noise is drawn from numpy's global stream in the
reference's order (one `np.random.normal(0., 1)` per non-ego vehicle, in append order), so `np.random.seed(s)` gives
the reference's vehicles.  The parameter tables (idm_param / rss_param) are the reference's constants.
"""
import numpy as np

from .mc_sim_highway import Agent as AgentMS
from .mc_sim_highway import Corridor as CorridorMS
from .mc_sim_highway import IdmParam, Line as LineMS, Point as PointMS, RSSParam, YieldingParam


def idm_param(ego_l):                                        # mcs_wrapper.py:8-13
    if ego_l > 7:
        return IdmParam(0.8, -1.5, 2., -2, -8., 1.2)
    return IdmParam(1.6, -3.0, 2., -2, -8., 1.2)


def rss_param(is_merging):                                   # mcs_wrapper.py:16-21
    if is_merging:
        return RSSParam(0.7, 0.4, -8., -10., 1.8, 1.2)
    return RSSParam(0.7, 0.4, -6., -6., 1.8, 1.2)


def _ego_params(agent):
    ip, mp = agent.idm_param, agent.mobil_param
    idm_p_ego = IdmParam(ip.acc_max, ip.acc_min, ip.min_dis, ip.acc_com, -8., ip.thw)
    yp = agent.yielding_param
    yp_ego = YieldingParam(yp[0], yp[1], yp[2], yp[3], mp.politeness_factor, mp.delta_a, mp.bias_a)
    return idm_p_ego, yp_ego


class _Plan:
    """Collects what one builder call needs converted, in the reference's append order, and converts it in one launch."""

    def __init__(self, env, ref_corridor_id):
        self.env = env
        self.ref_lane = env.map.index_of[ref_corridor_id]
        self.px, self.py = [], []
        self.vehicles = []        # (point index, obj, vd, behavior, rss is_merging)
        self.corridors = []       # (corridor, first point index, limit or None)

    def _pt(self, x, y):
        self.px.append(float(x)); self.py.append(float(y))
        return len(self.px) - 1

    def vehicle(self, obj, vd, behavior):
        self.vehicles.append((self._pt(obj.x, obj.y), obj, vd, behavior))

    def corridor(self, c, limit=None):
        first = len(self.px)
        for p in (c.left[0], c.left[-1], c.right[0], c.right[-1], c.center[0], c.center[-1]):
            self._pt(p[0], p[1])
        self.corridors.append((c, first, limit))

    def convert(self):
        self.s, self.d = self.env.map.device_map.frenet(self.ref_lane, self.px, self.py)

    def corridor_ms(self, k, x_limit=None):
        """Corridor.to_frenet_corridor / to_frenet_corridor_limit_length (type.py:207-237): the far points take the
        lateral offset of the near ones."""
        c, f, _ = self.corridors[k]
        s, d = self.s, self.d
        far = [float(s[f + 1]), float(s[f + 3]), float(s[f + 5])]
        if x_limit is not None:
            far = [min(v, x_limit) for v in far]
        return CorridorMS(c.id, c.type,
                          LineMS(PointMS(float(s[f]), float(d[f])), PointMS(far[0], float(d[f]))),
                          LineMS(PointMS(float(s[f + 2]), float(d[f + 2])), PointMS(far[1], float(d[f + 2]))),
                          LineMS(PointMS(float(s[f + 4]), float(d[f + 4])), PointMS(far[2], float(d[f + 4]))), c.v_limit)


def _others(plan, yp):
    """AgentMS of every collected non-ego vehicle, drawing the vd noise in append order."""
    out = []
    for k, obj, vd, behavior in plan.vehicles:
        noisy = vd + np.random.normal(0., 1)
        out.append(AgentMS(obj.id, float(plan.s[k]), float(plan.d[k]), obj.v, obj.vy, noisy, obj.a, obj.l, obj.w,
                           idm_param(obj.l), rss_param(behavior == "merging"), yp, behavior))
    return out


def merging_mcs_sim_from_env(env, agent, ego_corridor):      # mcs_wrapper.py:24-75
    idm_p_ego, yp_ego = _ego_params(agent)
    yp = YieldingParam()
    possible_actions = [-1]
    left_corridor_id = env.map.corridors[ego_corridor.id].left_id
    left_corridor = env.map.corridors[left_corridor_id]
    plan = _Plan(env, left_corridor_id)                      # reference line: the main lane next to the ramp
    ego_pt = plan._pt(agent.x, agent.y)
    ego_behavior = "merging" if "merging" in agent.behavior_model else "idm"
    neighbor = env.neighbor_around_agent(agent.id)
    for obj in neighbor.lefts_and_front:
        if obj:
            plan.vehicle(obj, obj.vd, "merging" if "merging" in agent.behavior_model else "idm_lc")
    for obj in reversed(neighbor.lefts):
        if obj and len(possible_actions) < 4:
            possible_actions.append(obj.id)
    plan.corridor(ego_corridor)
    plan.corridor(left_corridor)
    ll_corridor_id = left_corridor.left_id
    if ll_corridor_id is not None:
        for obj in env.sorted_agents_on_corridor_within_range_of_vehicle(ll_corridor_id, agent.id, 80):
            plan.vehicle(obj, obj.idm_param.vd, "idm_lc")
        plan.corridor(env.map.corridors[ll_corridor_id])
    plan.convert()
    ego_agent = AgentMS(agent.id, float(plan.s[ego_pt]), float(plan.d[ego_pt]), agent.v, agent.vy, agent.idm_param.vd, agent.a,
                        agent.l, agent.w, idm_p_ego, rss_param(True), yp_ego, ego_behavior)
    agents = [ego_agent] + _others(plan, yp)
    corridors = [plan.corridor_ms(k) for k in range(len(plan.corridors))]
    return corridors, agents, possible_actions


def _center_and_sides(env, agent, ego_corridor, plan, with_gaps):
    """The part exit_ and lane_change_mcs_sim_from_env share (mcs_wrapper.py:120-189 == :219-282): front, front of the
    front, follower, then left / second-left / right / second-right lanes.  Returns possible_actions."""
    possible_actions = [-1]
    ego_front = env.front_agent(agent.id)
    ego_follow = env.following_agent(agent.id)
    if ego_front:
        plan.vehicle(ego_front, ego_front.vd, "idm_lc")
        ego_ff = env.front_agent(ego_front.id)
        if ego_ff:
            plan.vehicle(ego_ff, ego_ff.vd, "idm_lc")
    if ego_follow:
        plan.vehicle(ego_follow, ego_follow.vd, "idm_lc")
    neighbors = env.neighbor_around_agent(agent.id)
    left_corridor_id = env.map.corridors[ego_corridor.id].left_id
    if left_corridor_id is not None:
        left_corridor = env.map.corridors[left_corridor_id]
        for obj in neighbors.lefts:
            if obj:
                plan.vehicle(obj, obj.vd, "idm_lc")
        plan.corridor(left_corridor)
        ll_corridor_id = left_corridor.left_id
        if ll_corridor_id is not None:
            for obj in env.sorted_agents_on_corridor_within_range_of_vehicle(ll_corridor_id, agent.id, 80):
                plan.vehicle(obj, obj.idm_param.vd, "idm_lc")
            plan.corridor(env.map.corridors[ll_corridor_id])
    if with_gaps:
        for obj in reversed(neighbors.rights):
            if obj and len(possible_actions) < 4:
                possible_actions.append(obj.id)
    right_corridor_id = env.map.corridors[ego_corridor.id].right_id
    if right_corridor_id is not None:
        right_corridor = env.map.corridors[right_corridor_id]
        for obj in neighbors.rights:
            if obj:
                plan.vehicle(obj, obj.vd, "merging" if right_corridor.type == "merging" else "idm_lc")
        plan.corridor(right_corridor)
        rr_corridor_id = right_corridor.right_id
        if rr_corridor_id is not None:
            rr_corridor = env.map.corridors[rr_corridor_id]
            for obj in env.sorted_agents_on_corridor_within_range_of_vehicle(rr_corridor_id, agent.id, 80):
                plan.vehicle(obj, obj.idm_param.vd, "merging" if rr_corridor.type == "merging" else "idm_lc")
            plan.corridor(rr_corridor)
    return possible_actions


def exit_mcs_sim_from_env(env, agent, ego_corridor):         # mcs_wrapper.py:78-190
    idm_p_ego, yp_ego = _ego_params(agent)
    yp = YieldingParam()
    right_corridor_id = env.map.corridors[ego_corridor.id].right_id
    exit_lane_id = -1
    while right_corridor_id is not None:                     # :96-103 find the exit lane to the right
        right_corridor = env.map.corridors[right_corridor_id]
        if right_corridor.type == "exit":
            exit_lane_id = right_corridor_id
            break
        right_corridor_id = right_corridor.right_id
    if exit_lane_id == -1:
        print("Exit mcs wrapper: no exit lane found!")
    exit_corridor = env.map.corridors[exit_lane_id]
    plan = _Plan(env, ego_corridor.id)
    ego_pt = plan._pt(agent.x, agent.y)
    plan.corridor(exit_corridor)                             # index 0: only its far centre point is used (:113-115)
    plan.corridor(ego_corridor)                              # index 1: the length-limited centre corridor
    possible_actions = _center_and_sides(env, agent, ego_corridor, plan, with_gaps=True)
    plan.convert()
    ego_s = float(plan.s[ego_pt])
    exit_ms = plan.corridor_ms(0)
    max_length = min(ego_corridor.v_limit ** 2 / 2.5 ** 2,
                     exit_ms.m.p2.x - (abs(exit_ms.id - ego_corridor.id) - 1) * 50 - ego_s)
    ego_agent = AgentMS(agent.id, ego_s, float(plan.d[ego_pt]), agent.v, agent.vy, agent.idm_param.vd, agent.a, agent.l, agent.w,
                        idm_p_ego, rss_param(True), yp_ego, "exit")
    agents = [ego_agent] + _others(plan, yp)
    corridors = [plan.corridor_ms(1, x_limit=ego_s + max_length)] + [plan.corridor_ms(k) for k in range(2, len(plan.corridors))]
    return corridors, agents, possible_actions


def lane_change_mcs_sim_from_env(env, agent, ego_corridor):  # mcs_wrapper.py:193-283
    idm_p_ego, yp_ego = _ego_params(agent)
    rp = agent.rss_param
    rss_p_ego = RSSParam(rp.r_t_other, rp.r_t_ego, rp.a_min, rp.a_max, rp.soft_acc, rp.soft_dcc)
    yp = YieldingParam()
    plan = _Plan(env, ego_corridor.id)
    ego_pt = plan._pt(agent.x, agent.y)
    plan.corridor(ego_corridor)
    _center_and_sides(env, agent, ego_corridor, plan, with_gaps=False)
    plan.convert()
    ego_behavior = "merging" if "merging" in agent.behavior_model else "idm"
    ego_agent = AgentMS(agent.id, float(plan.s[ego_pt]), float(plan.d[ego_pt]), agent.v, agent.vy, agent.idm_param.vd, agent.a,
                        agent.l, agent.w, idm_p_ego, rss_p_ego, yp_ego, ego_behavior)
    agents = [ego_agent] + _others(plan, yp)
    corridors = [plan.corridor_ms(k) for k in range(len(plan.corridors))]
    return corridors, agents
