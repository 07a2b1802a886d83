"""The learned and closest-gap behaviours of the reference (behaviors.py:89-176) from the OUTER environment down:
`outer_env.Environment` (device lane match / neighbour tables) -> `mcs_wrapper` builders (device Frenet conversion) ->
all candidate commands of the decision in ONE rollout launch (`decisions`) -> learned model -> the single-step
behaviour kernel.  Same function names, arguments `(observed_model, agent)` and side effects on the agent
(`current_decision`) as the reference; the returned DecoupledAction carries the ego corridor (whose centre line is the
reference's `ref_path`) for `DeviceMap.step_along_path`.

The MOBIL / plain-IDM behaviours of the outer loop (behaviors.py:10-86) stay with the reference's Python: they are
per-vehicle host arithmetic around the yielding model, not part of the rollout path.
"""
from . import decisions
from . import mc_sim_highway as ms
from .mcs_wrapper import exit_mcs_sim_from_env, lane_change_mcs_sim_from_env, merging_mcs_sim_from_env


class DecoupledAction:                                       # type.py:9-15
    def __init__(self, ax, vy, ref_path):
        self.ax, self.vy, self.ref_path = ax, vy, ref_path


def lane_change_behavior_learned(observed_model, agent):     # behaviors.py:89-115
    possible_actions = ["keep_lane", "dcc", "acc"]
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    corridors = observed_model.map.corridors
    if ego_corridor.left_id and corridors[ego_corridor.left_id].type == "main":       # sic: lane id 0 is falsy
        possible_actions.append("left")
    if ego_corridor.right_id and corridors[ego_corridor.right_id].type == "main":
        possible_actions.append("right")
    mcs_corridors, mcs_agents = lane_change_mcs_sim_from_env(observed_model, agent, ego_corridor)
    action, decision, _, _ = decisions.lane_change_decision(mcs_corridors, mcs_agents, agent.id, possible_actions,
                                                            agent.current_decision, agent.learned_lane_change_model, 20,
                                                            ms.SimParameter(0.3, 40, 30))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def cloned_merging_behavior(observed_model, agent):          # behaviors.py:118-137
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    mcs_corridors, mcs_agents, possible_actions = merging_mcs_sim_from_env(observed_model, agent, ego_corridor)
    action, decision, _, _ = decisions.gap_decision(mcs_corridors, mcs_agents, agent.id, possible_actions, "left",
                                                    agent.learned_merging_model, 30, ms.SimParameter(0.3, 40, 10))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def merging_behavior_closest_gap(observed_model, agent):     # behaviors.py:140-145
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    mcs_corridors, mcs_agents, _ = merging_mcs_sim_from_env(observed_model, agent, ego_corridor)
    action = ms.closestGapMergingBehavior(ms.Environment(ms.MapMultiLane(mcs_corridors), mcs_agents), agent.id, "left")
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def exit_behavior_closest_gap(observed_model, agent):        # behaviors.py:148-153
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    mcs_corridors, mcs_agents, _ = exit_mcs_sim_from_env(observed_model, agent, ego_corridor)
    action = ms.closestGapMergingBehavior(ms.Environment(ms.MapMultiLane(mcs_corridors), mcs_agents), agent.id, "right")
    return DecoupledAction(action.ax, action.vy, ego_corridor)


def cloned_exit_behavior(observed_model, agent):             # behaviors.py:156-176
    ego_corridor = observed_model.corridor_for_agent(agent.id)
    mcs_corridors, mcs_agents, possible_actions = exit_mcs_sim_from_env(observed_model, agent, ego_corridor)
    action, decision, _, _ = decisions.gap_decision(mcs_corridors, mcs_agents, agent.id, possible_actions, "right",
                                                    agent.learned_merging_model, 20, ms.SimParameter(0.3, 40, 10))
    agent.current_decision = decision
    return DecoupledAction(action.ax, action.vy, ego_corridor)
