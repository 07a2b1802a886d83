"""configs[0..3] end to end at the module boundary.  tests/golden/scenario_*.json.gz hold every call the reference's own
Python (simulation_multilane / environment / agent / behaviors / mcs_wrapper, unchanged) made into `mc_sim_highway` while
stepping its shipped test_scene and one epoch of each evaluation script's random scene, together with the answer of the
reference's own machine code (generator: tests/golden/make_scenario_golden.py).  Replaying the calls and getting the
same answers bit for bit means the outer simulation evolves identically: decisions, actions and trajectories.

CPU: through the oracle restatement.  GPU: through the product's drop-in module (the C ABI, the CUDA kernels)."""
import ctypes as C
import glob
import gzip
import json
import os

import pytest

import interactive_highway_simulation_b200 as ihs
from interactive_highway_simulation_b200 import _abi as abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENARIOS = sorted(glob.glob(os.path.join(GOLDEN, "scenario_*.json.gz")))
FEATURES = ("fallBackRate", "successRate", "utility", "comfort", "expectedStepRatio", "utilityObjects", "comfortObjects", "fromNSims")


def load(path):
    with gzip.open(path, "rt") as f:
        return json.load(f)


def c_scene(rec):
    cs = (abi.Corridor * len(rec["corridors"]))()
    for o, c in zip(cs, rec["corridors"]):
        o.id, o.type = c[0], abi.LANE_TYPES.get(c[1], abi.LANE_OTHER)
        o.left[:] = c[2:6]; o.right[:] = c[6:10]; o.center[:] = c[10:14]; o.v_limit = c[14]
    as_ = (abi.Agent * len(rec["agents"]))()
    for o, a in zip(as_, rec["agents"]):
        o.id = a[0]
        o.x, o.y, o.v, o.vy, o.vd, o.a, o.l, o.w = a[1:9]
        o.idm[:] = a[9:15]; o.rss[:] = a[15:21]; o.yp[:] = a[21:28]
        o.behavior = abi.BEHAVIORS.get(a[28], abi.BEH_OTHER)
        o.commit, o.commit_keep_lane_step, o.target_lane_id = abi.COMMITS[a[29]], a[30], a[31]
    return cs, as_


def module_scene(m, rec):
    corridors = [m.Corridor(c[0], c[1], m.Line(m.Point(c[2], c[3]), m.Point(c[4], c[5])), m.Line(m.Point(c[6], c[7]), m.Point(c[8], c[9])),
                            m.Line(m.Point(c[10], c[11]), m.Point(c[12], c[13])), c[14]) for c in rec["corridors"]]
    agents = []
    for a in rec["agents"]:
        i = a[9:15]
        ag = m.Agent(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], m.IdmParam(i[0], i[2], i[1], i[3], i[4], i[5]),
                     m.RSSParam(*a[15:21]), m.YieldingParam(*a[21:28]), a[28])
        ag.commitToDecision, ag.commitKeepLaneStep, ag.targetLaneId = a[29], a[30], a[31]
        agents.append(ag)
    return m.MapMultiLane(corridors), agents


def test_fixtures_cover_the_configurations():
    tags = [os.path.basename(p) for p in SCENARIOS]
    for t in ("test_scene", "lane_change_learned", "merging_learned", "eval_merging_learned", "eval_lane_change_learned",
              "eval_exit_learned", "eval_exit_closest_gap"):
        assert "scenario_%s.json.gz" % t in tags
    n = {}
    for p in SCENARIOS:
        for c in load(p)["calls"]:
            n[c["fn"]] = n.get(c["fn"], 0) + 1
    assert n["simulationMultiThread"] >= 100 and n["laneChangeMobilBehavior"] >= 500
    assert n["fixedGapMergingBehavior"] >= 10 and n["closestGapMergingBehavior"] >= 40 and n["fixedDirectionLaneChangeBehavior"] >= 10


@pytest.mark.parametrize("path", SCENARIOS, ids=lambda p: os.path.basename(p)[9:-8])
def test_oracle_replays_the_reference_scenarios(oracle, path):
    for call in load(path)["calls"]:
        cs, as_ = c_scene(call["scene"])
        out, fn = call["out"], call["fn"]
        if fn == "simulationMultiThread":
            n = call["n"]
            cand = abi.Candidate(call["ego"], abi.ACTIONS.get(call["action"], abi.ACT_INVALID), call["gap"], 0)
            sp = abi.SimParam(*call["sim"])
            st = (abi.Status * (8 * n))()
            assert oracle.lib.orc_rollouts(cs, len(cs), as_, len(as_), C.byref(cand), C.byref(sp), call["seed"], 0, 8 * n, st, None, None) == 0
            f = oracle.reduce(st, 8, n, sp.steps)
            for k in FEATURES:
                assert getattr(f, k) == out[k], (fn, call["ego"], call["action"], call["gap"], k)
        elif fn == "laneChangeMobilBehavior":
            r = abi.ActionWithType()
            act = abi.Action(call["ax"], call["vy"])
            assert oracle.lib.orc_decide_mobil(cs, len(cs), as_, len(as_), call["id"], C.byref(act), int(call["require_rss"]),
                                               call["seed"], C.byref(r)) == 0
            assert (r.ax, r.vy, abi.COMMIT_NAMES[r.commit], r.commit_keep_lane_step, r.target_lane_id) == \
                (out["ax"], out["vy"], out["commit"], out["keep"], out["target"]), call["id"]
        else:
            r = abi.Action()
            if fn == "fixedDirectionLaneChangeBehavior":
                rc = oracle.lib.orc_decide_fixed_direction(cs, len(cs), as_, len(as_), call["id"],
                                                           abi.ACTIONS.get(call["decision"], abi.ACT_INVALID), 0, C.byref(r))
            elif fn == "fixedGapMergingBehavior":
                rc = oracle.lib.orc_decide_fixed_gap(cs, len(cs), as_, len(as_), call["id"], abi.ACTIONS[call["direction"]],
                                                     call["gap"], 0, C.byref(r))
            else:
                rc = oracle.lib.orc_decide_closest_gap(cs, len(cs), as_, len(as_), call["id"], abi.ACTIONS[call["direction"]], 0, C.byref(r))
            assert rc == 0 and (r.ax, r.vy) == (out["ax"], out["vy"]), (fn, call["id"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", SCENARIOS, ids=lambda p: os.path.basename(p)[9:-8])
def test_dropin_module_replays_the_reference_scenarios(backend_lib, path):
    m = ihs.install_as_mc_sim_highway()
    for call in load(path)["calls"]:
        mcs_map, agents = module_scene(m, call["scene"])
        out, fn = call["out"], call["fn"]
        m.seed(call["seed"])
        if fn == "simulationMultiThread":
            f = m.simulationMultiThread(call["n"], m.SimParameter(*call["sim"]), mcs_map, agents, call["ego"], call["action"], call["gap"])
            for k in FEATURES:
                assert getattr(f, k) == out[k], (fn, call["ego"], call["action"], call["gap"], k)
        elif fn == "laneChangeMobilBehavior":
            r = m.laneChangeMobilBehavior(m.Environment(mcs_map, agents), call["id"], m.Action(call["ax"], call["vy"]), call["require_rss"])
            assert (r.ax, r.vy, r.commitToDecision, r.commitKeepLaneStep, r.targetLaneId) == \
                (out["ax"], out["vy"], out["commit"], out["keep"], out["target"]), call["id"]
        elif fn == "fixedDirectionLaneChangeBehavior":
            r = m.fixedDirectionLaneChangeBehavior(m.Environment(mcs_map, agents), call["id"], call["decision"])
            assert (r.ax, r.vy) == (out["ax"], out["vy"])
        elif fn == "fixedGapMergingBehavior":
            r = m.fixedGapMergingBehavior(m.Environment(mcs_map, agents), call["id"], call["direction"], call["gap"])
            assert (r.ax, r.vy) == (out["ax"], out["vy"])
        else:
            r = m.closestGapMergingBehavior(m.Environment(mcs_map, agents), call["id"], call["direction"])
            assert (r.ax, r.vy) == (out["ax"], out["vy"])
