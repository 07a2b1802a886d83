"""The learned and closest-gap behaviours of the reference (behaviors.py:89-176) from the OUTER environment down:
`outer_env.Environment` (device lane match / neighbour tables) -> `mcs_wrapper` builders (device Frenet conversion) ->
all candidate commands of the decision in ONE rollout launch (`decisions`) -> learned model -> the single-step
behaviour kernel.  Same function names, arguments `(observed_model, agent)` and side effects on the agent
(`current_decision`) as the reference; the returned DecoupledAction carries the ego corridor (whose centre line is the
reference's `ref_path`) for `DeviceMap.step_along_path`.

The plain-IDM / MOBIL behaviours of the outer loop (behaviors.py:10-86) are here too: `idm_behavior` is the reference's
per-vehicle host arithmetic (IDM, yielding intentions, RSS brake) over the device's neighbour tables and border / centre
distances; `lane_change_behavior_mobil` asks the single-step MOBIL kernel (`laneChangeMobilBehavior`).
"""
from . import _abi as abi
from . import backend
from . import decisions
from . import mc_sim_highway as ms
from .mcs_wrapper import exit_mcs_sim_from_env, lane_change_mcs_sim_from_env, merging_mcs_sim_from_env  # noqa: F401  (reference's import list)
from .outer_env import BUILD_EXIT, BUILD_LANE_CHANGE, BUILD_MERGING
from .type import DecoupledAction, YieldingIntention
from .utility import idm_acc_f_b, idm_p_for_merging_exit_agents, rss_dis


def idm_behavior(observed_model, agent):                     # behaviors.py:10-67
    corridor = observed_model.corridor_for_agent(agent.id)
    corridors = observed_model.map.corridors
    front = observed_model.front_agent(agent.id)
    ax = idm_acc_f_b([front] if front else [], None, agent.v, corridor.v_limit, agent.idm_param)
    # lateral: attach to the centre line
    signed_distance = corridor.centerSimplified.signed_distance([agent.x, agent.y])
    vy_abs = min(max(abs(signed_distance) * 1.0, 0.1), 1.3)
    vy = -signed_distance / (abs(signed_distance) + 1e-10) * vy_abs
    # candidates to yield to: vehicles on a merging lane to the right, exiting vehicles on a main lane to the left
    matches = []
    if corridor.right_id is not None and corridors[corridor.right_id].type == "merging":
        nb = observed_model.neighbor_around_agent(agent.id)
        matches.extend(v for v in (nb.right_bb, nb.right_b, nb.right_f, nb.right_ff) if v is not None)
    if corridor.left_id is not None and corridors[corridor.left_id].type == "main":
        for _, other in observed_model.matched_agents_sorted_by_arc_length[corridor.left_id]:
            if "exit" in other.behavior_model:
                matches.append(observed_model.agent_to_idm_match(agent.id, other.id))
    intentions = agent.yielding_intention_to_others
    for v in matches:
        direction = "right" if "exit" in v.agent.behavior_model else "left"
        p = agent.yielding_model.yielding_probability(v, agent, front, observed_model, direction)
        if v.id not in intentions:
            intentions[v.id] = YieldingIntention(v.id, p, change_intention_threshold=agent.change_intention_threshold,
                                                 random_intention_seed=agent.random_yielding_seed)
        else:
            intentions[v.id].update_intention(p)
    wanted = [v.id for v in matches]
    for idx in list(intentions):
        if idx not in wanted:
            intentions.pop(idx)
    yielding_to = [v for v in matches if intentions[v.id].yielding]
    if yielding_to:
        ax = min(ax, idm_acc_f_b(yielding_to, None, agent.v, corridor.v_limit,
                                 idm_p_for_merging_exit_agents(agent.random_yielding_seed, agent.idm_param, agent.l > 7)))
    rp = agent.rss_param
    if front and rss_dis(front.v, agent.v, rp.r_t_ego, rp.a_min, rp.a_max) > front.dis - 0.5 * (front.l + agent.l):
        ax = rp.a_min
    return DecoupledAction(ax, vy, corridor)


class _Marshalled:
    """The inner simulator's scene of one vehicle, built on the device by the environment (mcs_wrapper's builder as ONE
    launch, outer_env.Environment.marshal) and handed to the decision / rollout kernels without passing through Python
    value objects; `actions` = the builder's candidate gap ids."""

    def __init__(self, observed_model, agent, kind, poke_commit=False):
        _, nc, _, na, self.actions, handle = observed_model.marshal(agent, kind, poke_commit=poke_commit)
        self.scene = backend.DeviceScene.adopt(handle, na, nc)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.scene.close()


def lane_change_behavior_mobil(observed_model, agent, require_rss_safety):    # behaviors.py:70-86
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    action_idm = idm_behavior(observed_model, agent)
    with _Marshalled(observed_model, agent, BUILD_LANE_CHANGE, poke_commit=True) as m:
        out = m.scene.decide_mobil(agent.id, action_idm.ax, action_idm.vy, bool(require_rss_safety), ms._next_seed())
    agent.commit_lane_change, agent.commit_keep_lane_step, agent.target_lane_id = \
        abi.COMMIT_NAMES[out.commit], out.commit_keep_lane_step, out.target_lane_id
    return DecoupledAction(out.ax, out.vy, ego_corridor)


def lane_change_behavior_learned(observed_model, agent):     # behaviors.py:89-115
    possible_actions = ["keep_lane", "dcc", "acc"]
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    corridors = observed_model.map.corridors
    if ego_corridor.left_id and corridors[ego_corridor.left_id].type == "main":       # sic: lane id 0 is falsy
        possible_actions.append("left")
    if ego_corridor.right_id and corridors[ego_corridor.right_id].type == "main":
        possible_actions.append("right")
    with _Marshalled(observed_model, agent, BUILD_LANE_CHANGE) as m:
        action, decision, _, _ = decisions.lane_change_decision_on(m.scene, agent.id, possible_actions, agent.current_decision,
                                                                   agent.learned_lane_change_model, 20, ms.SimParameter(0.3, 40, 30))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def cloned_merging_behavior(observed_model, agent):          # behaviors.py:118-137
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    with _Marshalled(observed_model, agent, BUILD_MERGING) as m:
        action, decision, _, _ = decisions.gap_decision_on(m.scene, agent.id, m.actions, "left", agent.learned_merging_model, 30,
                                                           ms.SimParameter(0.3, 40, 10))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def merging_behavior_closest_gap(observed_model, agent):     # behaviors.py:140-145
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    with _Marshalled(observed_model, agent, BUILD_MERGING) as m:
        action = m.scene.decide_closest_gap(agent.id, abi.ACTIONS["left"])
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def exit_behavior_closest_gap(observed_model, agent):        # behaviors.py:148-153
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    with _Marshalled(observed_model, agent, BUILD_EXIT) as m:
        action = m.scene.decide_closest_gap(agent.id, abi.ACTIONS["right"])
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def cloned_exit_behavior(observed_model, agent):             # behaviors.py:156-176
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    with _Marshalled(observed_model, agent, BUILD_EXIT) as m:
        action, decision, _, _ = decisions.gap_decision_on(m.scene, agent.id, m.actions, "right", agent.learned_merging_model, 20,
                                                           ms.SimParameter(0.3, 40, 10))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)
