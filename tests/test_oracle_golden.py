"""The oracle (oracle/mcsim_oracle.cpp, the CPU restatement) against the golden vectors produced by the reference's
own machine code (tests/golden/, generator: tests/golden/make_golden.py).  Bit-exact: every double is compared with ==."""
import ctypes as C

import pytest

import golden_io
from interactive_highway_simulation_b200 import _abi as abi
from scenes import STATUS_KEYS, same

ROLLOUT_CASES = golden_io.load_all("rollouts")
ORDER_CASES = golden_io.load_all("order")
DECISION_CASES = golden_io.load_all("decisions")


def test_fixtures_present():
    assert len(ROLLOUT_CASES) >= 30 and len(ORDER_CASES) >= 5 and len(DECISION_CASES) >= 1


@pytest.mark.parametrize("case", ROLLOUT_CASES, ids=golden_io.ids(ROLLOUT_CASES))
def test_rollouts_match_reference_binary(oracle, case):
    s = golden_io.scene_of(case)
    got, _ = oracle.rollouts(s, case["ego"], case["action"], case["gap"], case["dt"], case["steps"], case["min_steps"],
                             case["seed"], case["first"], case["count"], hist=True)
    for i, (g, r) in enumerate(zip(got, case["rollouts"])):
        if g["ok"] == 0 and r["ok"] == 0:
            continue   # setEgoAgentIdAndDrivingBehavior failed: runMTimes never runs such a rollout (so:0xc2935)
        for k in STATUS_KEYS:
            assert same(g[k], r[k]), (case["tag"], "rollout", i, k, g[k], r[k])
        if "hist" in r:
            for aid, hr in r["hist"].items():
                ho = g["hist"][int(aid)]
                assert len(ho) == len(hr), (case["tag"], i, aid, "history length")
                for t, (a, b) in enumerate(zip(ho, hr)):
                    assert tuple(a) == tuple(b), (case["tag"], "rollout", i, "agent", aid, "t", t, a, b)


@pytest.mark.parametrize("case", ORDER_CASES, ids=golden_io.ids(ORDER_CASES))
def test_iteration_orders_match_reference_binary(oracle, case):
    """std::unordered_map iteration orders the reference walks (agent plan/draw order, corridor first-match order)."""
    agents, corridors, _, _ = oracle.orders(golden_io.scene_of(case))
    assert agents == case["agents"]
    assert corridors == case["corridors"]


@pytest.mark.parametrize("case", DECISION_CASES, ids=golden_io.ids(DECISION_CASES))
def test_single_decisions_match_reference_binary(oracle, case):
    s = golden_io.scene_of(case)
    cs, as_ = s.c_corridors(), s.c_agents()
    for call in case["calls"]:
        t = call["cmd"].split()
        ans = call["answer"]
        if t[0] == "mobil":
            out = abi.ActionWithType()
            act = abi.Action(float(t[2]), float(t[3]))
            rc = oracle.lib.orc_decide_mobil(cs, len(cs), as_, len(as_), int(t[1]), C.byref(act), int(t[4]), case["key"],
                                             C.byref(out))
            assert rc == 0
            assert (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), call
            assert abi.COMMIT_NAMES[out.commit] == ans["commit"] and out.commit_keep_lane_step == int(ans["keep"])
            assert out.target_lane_id == int(ans["target"])
        elif t[0] == "fixed_gap":
            out = abi.Action()
            rc = oracle.lib.orc_decide_fixed_gap(cs, len(cs), as_, len(as_), int(t[1]), abi.ACTIONS[t[2]], int(t[3]),
                                                 case["key"], C.byref(out))
            assert rc == 0
            assert (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), call
        elif t[0] == "closest_gap":
            out = abi.Action()
            rc = oracle.lib.orc_decide_closest_gap(cs, len(cs), as_, len(as_), int(t[1]), abi.ACTIONS[t[2]], case["key"],
                                                   C.byref(out))
            assert rc == 0
            assert (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), call
        else:
            out = abi.Action()
            rc = oracle.lib.orc_decide_fixed_direction(cs, len(cs), as_, len(as_), int(t[1]), abi.ACTIONS[t[2]], case["key"],
                                                       C.byref(out))
            assert rc == 0
            assert (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), call


def test_time_to_reach_gap_matches_reference_binary(oracle):
    import math
    oracle.lib.orc_time_to_reach_gap.restype = C.c_double
    oracle.lib.orc_time_to_reach_gap.argtypes = [C.POINTER(C.c_double)]
    (case,) = golden_io.load_all("ttrg")
    for row in case["rows"]:
        g = oracle.lib.orc_time_to_reach_gap((C.c_double * 7)(*row["args"]))
        r = float(row["result"])
        assert g == r or (math.isnan(g) and math.isnan(r)), row
