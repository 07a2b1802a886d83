"""The reference's marshalling layer (mcs_wrapper.py:8-283) on the device geometry of include/mcsim_env.h: same three
builders, same names, same arguments and return values

    merging_mcs_sim_from_env(env, agent, ego_corridor)      -> corridors, agents, possible_actions      (:24-75)
    exit_mcs_sim_from_env(env, agent, ego_corridor)         -> corridors, agents, possible_actions      (:78-190)
    lane_change_mcs_sim_from_env(env, agent, ego_corridor)  -> corridors, agents                        (:193-283)

`env` is an `outer_env.Environment`.  A builder call is ONE launch (`mce_build_rollout_scene`, include/mcsim_env.h): the
neighbour selection (FF / F / B / BB per side, front, front of the front, follower, +-80 m on the second lanes) over the
resident tables of the last match at the vehicles' current positions, every XY -> Frenet conversion — the vehicles and the
six corner points of each corridor — on the builder's reference line (extend_line_both_sides(center, 1000), simplified),
the generic parameters and the candidate gaps.  The `vd` noise is numpy's global stream in the reference's order (one
`np.random.normal(0., 1)` per non-ego vehicle, in append order: `Environment.marshal` draws ahead and rewinds), so
`np.random.seed(s)` gives the reference's vehicles.  These functions turn the records back into the reference's value
objects; the behaviours (behaviors.py) skip that and hand the marshalled scene straight to the decision kernels.
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


def _records(env, agent, kind):
    """One builder call on the device (outer_env.Environment.marshal -> mce_build_rollout_scene) turned back into the
    value objects the reference's builders return."""
    from . import _abi as abi
    cs, nc, ag, na, actions, _ = env.marshal(agent, kind, records=True, scene=False)
    lane_names = {v: k for k, v in abi.LANE_TYPES.items()}
    beh_names = {v: k for k, v in abi.BEHAVIORS.items()}
    corridors = []
    for c in cs[:nc]:
        corridors.append(CorridorMS(c.id, lane_names.get(c.type, "other"),
                                    LineMS(PointMS(c.left[0], c.left[1]), PointMS(c.left[2], c.left[3])),
                                    LineMS(PointMS(c.right[0], c.right[1]), PointMS(c.right[2], c.right[3])),
                                    LineMS(PointMS(c.center[0], c.center[1]), PointMS(c.center[2], c.center[3])), c.v_limit))
    agents = []
    for a in ag[:na]:
        i, r, y = a.idm, a.rss, a.yp
        agents.append(AgentMS(a.id, a.x, a.y, a.v, a.vy, a.vd, a.a, a.l, a.w, IdmParam(i[0], i[2], i[1], i[3], i[4], i[5]),
                              RSSParam(r[0], r[1], r[2], r[3], r[4], r[5]), YieldingParam(y[0], y[1], y[2], y[3], y[4], y[5], y[6]),
                              beh_names[a.behavior]))
    return corridors, agents, actions


def merging_mcs_sim_from_env(env, agent, ego_corridor):      # mcs_wrapper.py:24-75
    from .outer_env import BUILD_MERGING
    return _records(env, agent, BUILD_MERGING)


def exit_mcs_sim_from_env(env, agent, ego_corridor):         # mcs_wrapper.py:78-190
    from .outer_env import BUILD_EXIT
    if not any(c.type == "exit" for c in env.map.corridors.values()):
        print("Exit mcs wrapper: no exit lane found!")
    return _records(env, agent, BUILD_EXIT)


def lane_change_mcs_sim_from_env(env, agent, ego_corridor):  # mcs_wrapper.py:193-283
    from .outer_env import BUILD_LANE_CHANGE
    corridors, agents, _ = _records(env, agent, BUILD_LANE_CHANGE)
    return corridors, agents
