"""One learned decision = one launch.

The reference evaluates a decision as 3-5 sequential simulationMultiThread calls (one per candidate command), a host-side
learned model over the 7 features of each, and a final single-step behaviour call (behaviors.py:89-115 lane change,
:118-137 merging, :156-176 exit).  These functions keep that contract at the inner-simulator boundary — inputs are what
the reference's mcs_wrapper builders return (corridors, AgentMS list, candidate commands) — but send all candidates
through ONE kernel launch over one device-resident scene.
"""
from . import _abi as abi
from . import learned_model as lm
from . import mc_sim_highway as ms


def _features_on(scene, ego_id, commands, n_per_group, sim_param):
    """All candidate commands [(action, gap id), ...] in one launch over a device-resident scene; consumes one seed of the
    drop-in's sequence per command, like simulationMultiThreadBatch."""
    cands = (abi.Candidate * len(commands))(*[abi.Candidate(int(ego_id), abi.ACTIONS.get(a, abi.ACT_INVALID), int(g), 0)
                                              for a, g in commands])
    base = [ms._next_seed() for _ in commands][0]
    feats, _ = scene.rollouts(cands, sim_param._c(), int(n_per_group), base)
    return [lm.feature_vector(ms.BehaviorFeature(f)) for f in feats]


def lane_change_decision_on(scene, ego_id, possible_actions, last_decision, model, n_per_group=20, sim_param=None):
    """lane_change_decision over a marshalled scene that is already on the device (outer_env.Environment.marshal)."""
    features = _features_on(scene, ego_id, [(a, -2) for a in possible_actions], n_per_group, sim_param or ms.SimParameter(0.3, 40, 30))
    q_values, decision = model.inference(features, possible_actions, last_decision)
    out = scene.decide_fixed_direction(int(ego_id), abi.ACTIONS.get(decision, abi.ACT_INVALID))
    return ms.Action(out.ax, out.vy), decision, q_values, features


def gap_decision_on(scene, ego_id, possible_gaps, direction, model, n_per_group=30, sim_param=None):
    """gap_decision over a marshalled scene that is already on the device."""
    features = _features_on(scene, ego_id, [(direction, g) for g in possible_gaps], n_per_group, sim_param or ms.SimParameter(0.3, 40, 10))
    index, q_values = model.inference(features)
    decision = possible_gaps[index]
    out = scene.decide_fixed_gap(int(ego_id), abi.ACTIONS[direction], int(decision))
    return ms.Action(out.ax, out.vy), decision, q_values, features


def lane_change_decision(corridors, mcs_agents, ego_id, possible_actions, last_decision="initial", model=None,
                         n_per_group=20, sim_param=None):
    """behaviors.py:89-115.  -> (Action, decision, q_values, features)"""
    model = model or lm.LearnedLaneChangeModel()
    sim_param = sim_param or ms.SimParameter(0.3, 40, 30)
    mcs_map = corridors if isinstance(corridors, ms.MapMultiLane) else ms.MapMultiLane(corridors)
    feats = ms.simulationMultiThreadBatch(n_per_group, sim_param, mcs_map, mcs_agents, ego_id,
                                          [(a, -2) for a in possible_actions])
    features = [lm.feature_vector(f) for f in feats]
    q_values, decision = model.inference(features, possible_actions, last_decision)
    action = ms.fixedDirectionLaneChangeBehavior(ms.Environment(mcs_map, mcs_agents), ego_id, decision)
    return action, decision, q_values, features


def gap_decision(corridors, mcs_agents, ego_id, possible_gaps, direction, model=None, n_per_group=30, sim_param=None):
    """behaviors.py:118-137 (direction "left", n 30) and :156-176 (direction "right", n 20).
    -> (Action, chosen gap id, q_values, features)"""
    model = model or lm.LearnedMergingModel()
    sim_param = sim_param or ms.SimParameter(0.3, 40, 10)
    mcs_map = corridors if isinstance(corridors, ms.MapMultiLane) else ms.MapMultiLane(corridors)
    feats = ms.simulationMultiThreadBatch(n_per_group, sim_param, mcs_map, mcs_agents, ego_id,
                                          [(direction, g) for g in possible_gaps])
    features = [lm.feature_vector(f) for f in feats]
    index, q_values = model.inference(features)
    decision = possible_gaps[index]          # index -1 selects the last entry, as in the reference (behaviors.py:132)
    action = ms.fixedGapMergingBehavior(ms.Environment(mcs_map, mcs_agents), ego_id, direction, decision)
    return action, decision, q_values, features
