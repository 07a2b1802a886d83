"""The drop-in `mc_sim_highway` module: the names and shapes the reference's Python relies on (CPU part), and its
results against the oracle through the reference's own call pattern (GPU part)."""
import ctypes as C
import importlib
import sys

import pytest

import interactive_highway_simulation_b200 as ihs
from interactive_highway_simulation_b200 import _abi as abi
from scenes import ACTIONS, probe_scene

# every name the reference imports from mc_sim_highway (behaviors.py:5-7, mcs_wrapper.py:3-5, type.py:4-6)
REFERENCE_IMPORTS = ["Point", "Line", "Corridor", "Agent", "MapMerging", "IdmParam", "RSSParam", "Vehicle", "YieldingParam",
                     "SimParameter", "MapMultiLane", "Action", "Simulation", "simulationMultiThread",
                     "fixedDirectionLaneChangeBehavior", "fixedGapMergingBehavior", "closestGapMergingBehavior",
                     "laneChangeMobilBehavior", "Environment"]
# every name the binding registers (strings in the reference binary's .rodata)
BINDING_NAMES = REFERENCE_IMPORTS + ["ActionWithType", "BehaviorFeature", "State", "IdmMatch", "Indicator", "PositionHistory",
                                     "simulationMerging", "simulationMergingOnce", "simulationMergingMultiThread",
                                     "getMergingAction", "getMergingActionClosestGap", "getMergingActionClosestGapId", "rssSafe",
                                     "rssSafeRelax"]


def module():
    return ihs.install_as_mc_sim_highway()


def test_install_registers_the_reference_module_name():
    m = module()
    assert sys.modules["mc_sim_highway"] is m
    assert importlib.import_module("mc_sim_highway") is m
    for n in BINDING_NAMES:
        assert hasattr(m, n), n


def test_value_types_follow_the_binding():
    m = module()
    idm = m.IdmParam(1.6, -3.0, 2.0, -2, -8.0, 1.2)                     # mcs_wrapper.py:9  (ctor order: accMax accMin minDis ...)
    assert idm._memory_order() == (1.6, 2.0, -3.0, -2.0, -8.0, 1.2)     # memory order: accMax minDis accMin accCom aE thw
    rss = m.RSSParam(0.7, 0.4, -8.0, -10.0, 1.8, 1.2)
    assert rss._memory_order() == (0.7, 0.4, -8.0, -10.0, 1.8, 1.2)
    assert m.YieldingParam()._memory_order() == (-0.5, 1.81, -4.8, -1.1, 0.9, 0.5, 0.51)
    a = m.Agent(7, 1.0, 2.0, 3.0, 0.1, 30.0, -0.5, 4.8, 2.0, idm, rss, m.YieldingParam(), "idm_lc")
    assert (a.id, a.targetLaneId, a.commitToDecision, a.commitKeepLaneStep) == (7, -1, "not_decided", 0)
    assert len(a.statesHistory) == 1 and a.statesHistory[0].v == 3.0
    a.setDesiredV(25.0)
    a.targetLaneId, a.commitToDecision, a.commitKeepLaneStep = 2, "left", 4      # behaviors.py:77-79
    o = abi.Agent()
    a._fill(o)
    assert (o.vd, o.target_lane_id, o.commit, o.commit_keep_lane_step) == (25.0, 2, abi.COMMITS["left"], 4)
    assert list(o.idm) == [1.6, 2.0, -3.0, -2.0, -8.0, 1.2]
    c = m.Corridor(3, "main", m.Line(m.Point(0, 3.5), m.Point(100, 3.5)), m.Line(m.Point(0, 0), m.Point(100, 0)),
                   m.Line(m.Point(0, 1.75), m.Point(100, 1.75)), 27.0)
    c.setLeftAndRightCorridorId(4, None)
    mp = m.MapMultiLane([c])
    assert mp.corridors[0].vLimit == 27.0 and mp.corridors[0].center.p2.x == 100.0
    sp = m.SimParameter(0.3, 40, 30)
    assert (sp.dt, sp.steps, sp.minSteps) == (0.3, 40, 30)
    with pytest.raises(NotImplementedError):
        m.MapMerging()


def build(m, s):
    corridors = [m.Corridor(c["id"], c["type"], m.Line(m.Point(c["x0"], c["left_y"]), m.Point(c["x1"], c["left_y"])),
                            m.Line(m.Point(c["x0"], c["right_y"]), m.Point(c["x1"], c["right_y"])),
                            m.Line(m.Point(c["x0"], c["center_y"]), m.Point(c["x1"], c["center_y"])), c["v_limit"])
                 for c in s.corridors]
    agents = []
    for a in s.agents:
        i, r, y = a["idm"], a["rss"], a["yp"]
        agents.append(m.Agent(a["id"], a["x"], a["y"], a["v"], a["vy"], a["vd"], a["a"], a["l"], a["w"],
                              m.IdmParam(i[0], i[2], i[1], i[3], i[4], i[5]), m.RSSParam(*r), m.YieldingParam(*y), a["behavior"]))
    return m.MapMultiLane(corridors), agents


@pytest.mark.gpu
def test_lane_change_decision_like_behaviors_py(oracle, backend_lib):
    """behaviors.py:89-115 call pattern: simulationMultiThread per action, then fixedDirectionLaneChangeBehavior."""
    m = module()
    s = probe_scene()
    mcs_map, mcs_agents = build(m, s)
    sim_param = m.SimParameter(0.3, 40, 30)
    m.seed(123)
    feats = [m.simulationMultiThread(20, sim_param, mcs_map, mcs_agents, 0, action, -2) for action in ACTIONS]
    m.seed(123)
    batch = m.simulationMultiThreadBatch(20, sim_param, mcs_map, mcs_agents, 0, [(a, -2) for a in ACTIONS])
    for c, action in enumerate(ACTIONS):
        _, ost = oracle.rollouts(s, 0, action, -2, 0.3, 40, 30, 123 + c, 0, 160)
        want = oracle.reduce(ost, 8, 20, 40)
        for f in m.BehaviorFeature._FIELDS:
            assert getattr(feats[c], f) == getattr(want, f), (action, f)
            assert getattr(batch[c], f) == getattr(want, f), (action, f)
        assert feats[c].fromNSims == 160
    act = m.fixedDirectionLaneChangeBehavior(m.Environment(mcs_map, mcs_agents), 0, "keep_lane")
    want = abi.Action()
    cs, as_ = s.c_corridors(), s.c_agents()
    assert oracle.lib.orc_decide_fixed_direction(cs, len(cs), as_, len(as_), 0, abi.ACTIONS["keep_lane"], 0, C.byref(want)) == 0
    assert (act.ax, act.vy) == (want.ax, want.vy)


@pytest.mark.gpu
def test_simulation_class_run_and_history(oracle, backend_lib):
    m = module()
    s = probe_scene()
    mcs_map, mcs_agents = build(m, s)
    m.seed(9)
    sim = m.Simulation(m.SimParameter(0.3, 40, 30), mcs_map, mcs_agents)
    assert sim.setEgoAgentIdAndDrivingBehavior(0, "left", -2)
    st = sim.run()
    o, _ = oracle.rollouts(s, 0, "left", -2, 0.3, 40, 30, 9, 0, 1, hist=True)
    assert (st.finish, st.fallBack, st.finishStep, st.utility, st.comfort) == \
        (bool(o[0]["finish"]), bool(o[0]["fallBack"]), o[0]["finishStep"], o[0]["utility"], o[0]["comfort"])
    hist = sim.agentsStatesHistory()
    for aid, h in o[0]["hist"].items():
        assert [(p.x, p.y, p.v, p.a) for p in hist[aid]] == h


@pytest.mark.gpu
def test_mobil_wrapper_like_behaviors_py(oracle, backend_lib):
    """behaviors.py:70-86: poke commit state into the ego's AgentMS, call laneChangeMobilBehavior, read it back."""
    m = module()
    s = probe_scene()
    mcs_map, mcs_agents = build(m, s)
    for a in mcs_agents:
        if a.id == 1:
            a.targetLaneId, a.commitToDecision, a.commitKeepLaneStep = -1, "keep_lane", 10
    m.seed(77)
    out = m.laneChangeMobilBehavior(m.Environment(mcs_map, mcs_agents), 1, m.Action(0.3, 0.0), True)
    s.agents[1]["commit"], s.agents[1]["keep_step"] = "keep_lane", 10
    cs, as_ = s.c_corridors(), s.c_agents()
    want = abi.ActionWithType()
    act = abi.Action(0.3, 0.0)
    assert oracle.lib.orc_decide_mobil(cs, len(cs), as_, len(as_), 1, C.byref(act), 1, 77, C.byref(want)) == 0
    assert (out.ax, out.vy, out.commitToDecision, out.commitKeepLaneStep, out.targetLaneId) == \
        (want.ax, want.vy, abi.COMMIT_NAMES[want.commit], want.commit_keep_lane_step, want.target_lane_id)


@pytest.mark.gpu
def test_decision_helpers_follow_behaviors_py(oracle, backend_lib):
    """decisions.lane_change_decision / gap_decision = behaviors.py:89-115 / :118-137 with one launch per decision: features,
    q-values, the discrete decision and the final action equal what the per-candidate reference flow gives (oracle)."""
    import random
    from interactive_highway_simulation_b200 import decisions, learned_model
    from scenes import random_scene
    m = module()
    s = probe_scene()
    mcs_map, mcs_agents = build(m, s)
    actions = ["keep_lane", "dcc", "acc", "left", "right"]
    m.seed(500)
    act, decision, q_values, features = decisions.lane_change_decision(mcs_map, mcs_agents, 0, actions, "keep_lane")
    want_features = []
    for c, a in enumerate(actions):
        _, ost = oracle.rollouts(s, 0, a, -2, 0.3, 40, 30, 500 + c, 0, 160)
        want_features.append(learned_model.feature_vector(oracle.reduce(ost, 8, 20, 40)))
    assert features == want_features
    wq, wd = learned_model.LearnedLaneChangeModel().inference(want_features, actions, "keep_lane")
    assert decision == wd and [[c, float(v)] for c, v in q_values] == [[c, float(v)] for c, v in wq]
    want = abi.Action()
    cs, as_ = s.c_corridors(), s.c_agents()
    assert oracle.lib.orc_decide_fixed_direction(cs, len(cs), as_, len(as_), 0, abi.ACTIONS[decision], 0, C.byref(want)) == 0
    assert (act.ax, act.vy) == (want.ax, want.vy)
    # merging: ego on the merging lane chooses among gaps
    rng = random.Random(3)
    s2, ego = random_scene(rng, n_agents=14, lane_types=["merging", "main", "main"], ego_behavior="merging", ego_lane=0)
    mcs_map, mcs_agents = build(m, s2)
    gaps = [-1] + [a["id"] for a in s2.agents if 3.5 <= a["y"] < 7.0][:3]
    m.seed(900)
    act, chosen, q_values, features = decisions.gap_decision(mcs_map, mcs_agents, ego, gaps, "left", n_per_group=30)
    want_features = []
    for c, g in enumerate(gaps):
        _, ost = oracle.rollouts(s2, ego, "left", g, 0.3, 40, 10, 900 + c, 0, 240)
        want_features.append(learned_model.feature_vector(oracle.reduce(ost, 8, 30, 40)))
    assert features == want_features
    idx, _ = learned_model.LearnedMergingModel().inference(want_features)
    assert chosen == gaps[idx]
    cs, as_ = s2.c_corridors(), s2.c_agents()
    assert oracle.lib.orc_decide_fixed_gap(cs, len(cs), as_, len(as_), ego, abi.ACTIONS["left"], chosen, 0, C.byref(want)) == 0
    assert (act.ax, act.vy) == (want.ax, want.vy)


@pytest.mark.gpu
def test_more_than_256_vehicles_is_a_clear_error(backend_lib):
    """The inner simulator holds up to MCS_MAX_AGENTS = 256 vehicles per rollout (one thread per vehicle of a 256-thread
    rollout); the reference has no such limit (its callers pass 10-25).  The drop-in says so instead of truncating."""
    from interactive_highway_simulation_b200 import mc_sim_highway as ms
    lanes = [ms.Corridor(k, "main", ms.Line(ms.Point(-1000, 3.5 * (k + 1)), ms.Point(9000, 3.5 * (k + 1))),
                         ms.Line(ms.Point(-1000, 3.5 * k), ms.Point(9000, 3.5 * k)),
                         ms.Line(ms.Point(-1000, 3.5 * k + 1.75), ms.Point(9000, 3.5 * k + 1.75)), 30.0) for k in range(2)]
    agents = [ms.Agent(i, 30.0 * (i // 2), 1.75 + 3.5 * (i % 2), 20.0, 0.0, 25.0, 0.0, 4.8, 2.0, ms.IdmParam(), ms.RSSParam(),
                       ms.YieldingParam(), "idm" if i == 0 else "idm_lc") for i in range(257)]
    with pytest.raises(backend_lib.McsimError) as e:
        ms.simulationMultiThread(1, ms.SimParameter(0.3, 5, 5), ms.MapMultiLane(lanes), agents, 0, "keep_lane", -2)
    assert e.value.code == -3 and "256" in str(e.value)
    f = ms.simulationMultiThread(1, ms.SimParameter(0.3, 5, 5), ms.MapMultiLane(lanes), agents[:256], 0, "keep_lane", -2)
    assert f.fromNSims == 8
