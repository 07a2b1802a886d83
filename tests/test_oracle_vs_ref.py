"""Live cross-check: the oracle restatement against the reference's own machine code (oracle/_ref/ref_harness) on
freshly generated random scenes — wider than the committed fixtures.  Skipped where oracle/_ref is not built."""
import ctypes as C
import math
import random

import pytest

from oracle.pyoracle import RefHarness
from scenes import ACTIONS, STATUS_KEYS, has_side_lane, merge_exit_cases, probe_scene, random_scene, same


def compare(oracle, scene, ego, action, gap, dt, steps, min_steps, seed, first, count, math_mode="mcs"):
    ref = RefHarness(math=math_mode)
    ref.load(scene, dt, steps, min_steps)
    r = ref.rollouts(ego, action, gap, first, count, seed, hist=True)
    ref.close()
    oracle.set_math_mode(0 if math_mode == "mcs" else 1)
    try:
        o, _ = oracle.rollouts(scene, ego, action, gap, dt, steps, min_steps, seed, first, count, hist=True)
    finally:
        oracle.set_math_mode(0)
    for i in range(count):
        if r[i]["ok"] == 0 and o[i]["ok"] == 0:
            continue
        for k in STATUS_KEYS:
            assert same(r[i][k], o[i][k]), (action, "rollout", first + i, k, r[i][k], o[i][k])
        for aid, hr in r[i]["hist"].items():
            ho = o[i]["hist"][aid]
            assert len(hr) == len(ho)
            for t, (a, b) in enumerate(zip(hr, ho)):
                assert a == b, (action, "rollout", first + i, "agent", aid, "t", t, a, b)


@pytest.mark.parametrize("action", ACTIONS)
def test_probe_scene(oracle, ref_available, action):
    compare(oracle, probe_scene(), 0, action, -2, 0.3, 40, 30, 21, 100, 12)


@pytest.mark.parametrize("sc", range(8))
def test_random_lane_change_scenes(oracle, ref_available, sc):
    rng = random.Random(7000 + sc)
    n_lanes = rng.choice([2, 3, 4, 5])
    n_agents = rng.choice([5, 9, 14, 20, 31, 40])
    s, ego = random_scene(rng, n_lanes, n_agents)
    dt, steps, min_steps = rng.choice([0.2, 0.3]), rng.choice([25, 40]), rng.choice([10, 30])
    for action in ACTIONS:
        compare(oracle, s, ego, action, -2, dt, steps, min_steps, 300 + sc, 0, 4)


def test_platform_libm_mode(oracle, ref_available):
    """With the platform libm on both sides (the reference's stock configuration) the restatement is still
    bit-identical: the pinned exp/pow of include/mcsim_math.h are not what makes the two agree."""
    compare(oracle, probe_scene(), 0, "left", -2, 0.3, 40, 30, 5, 0, 8, math_mode="libm")


def test_reference_known_answer_vector(ref_available):
    """glibc srand(1), runMTimes M=20 on the probe scene — the binary reproduces the captured vector."""
    ref = RefHarness(math="libm")
    ref.load(probe_scene(), 0.3, 40, 30)
    ref._send("rng libc_after_build 1")
    line, _ = ref.raw("runm 20 0 keep_lane -2", tag="runm")
    ref.close()
    assert "utility=0.85061929957464044" in line and "comfort=0.83534452498773037" in line
    assert "utilityObjects=0.80775693906906376" in line and "comfortObjects=0.9081216235292624" in line


@pytest.mark.parametrize("case", merge_exit_cases(24), ids=lambda c: c[0])
def test_merging_and_exit_rollouts(oracle, ref_available, case):
    """customizedMergingBehavior / mergingBehaviorClosestGap / rssSafeRelax / timeToReachGap / yielding intentions."""
    tag, s, ego, direction, gaps, seed = case
    for gap in gaps:
        compare(oracle, s, ego, direction, gap, 0.3, 40, 10, seed, 0, 3)


@pytest.mark.parametrize("case", merge_exit_cases(16, seed0=500), ids=lambda c: c[0])
def test_single_merging_decisions(oracle, ref_available, case):
    """fixedGapMergingBehavior / closestGapMergingBehavior (python_converter wrappers so:0xb9cc0) on a fresh Environment."""
    from interactive_highway_simulation_b200 import _abi as abi
    tag, s, ego, direction, gaps, seed = case
    rng = random.Random(seed)
    cs, as_ = s.c_corridors(), s.c_agents()
    ids = [a["id"] for a in s.agents]
    who_all = [a["id"] for a in s.agents if has_side_lane(s, a, direction)]   # the binary crashes without a side lane
    ref = RefHarness()
    ref.load(s, 0.3, 40, 10)
    try:
        for who in rng.sample(who_all, min(4, len(who_all))):
            for gap in [-1, -2] + rng.sample(ids, min(4, len(ids))):
                ans = ref.decide(5, "fixed_gap %d %s %d" % (who, direction, gap))
                out = abi.Action()
                rc = oracle.lib.orc_decide_fixed_gap(cs, len(cs), as_, len(as_), who, abi.ACTIONS[direction], gap, 5, C.byref(out))
                assert rc == 0 and (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), (tag, who, gap, ans)
            ans = ref.decide(5, "closest_gap %d %s" % (who, direction))
            out = abi.Action()
            rc = oracle.lib.orc_decide_closest_gap(cs, len(cs), as_, len(as_), who, abi.ACTIONS[direction], 5, C.byref(out))
            assert rc == 0 and (out.ax, out.vy) == (float(ans["ax"]), float(ans["vy"])), (tag, who, ans)
    finally:
        ref.close()


def test_time_to_reach_gap(oracle, ref_available):
    """timeToReachGap so:0xe1630 on 4000 random argument tuples covering its three regimes and their boundaries."""
    oracle.lib.orc_time_to_reach_gap.restype = C.c_double
    oracle.lib.orc_time_to_reach_gap.argtypes = [C.POINTER(C.c_double)]
    ref = RefHarness()
    rng = random.Random(1)
    try:
        for _ in range(4000):
            m = rng.random()
            xE, vE = rng.uniform(0, 300), rng.uniform(0, 35)
            if m < 0.3:
                xB = xE + rng.uniform(-5, 100); xF = xB + rng.uniform(0, 60)
            elif m < 0.6:
                xF = xE - rng.uniform(-5, 100); xB = xF - rng.uniform(0, 60)
            else:
                xB = xE - rng.uniform(0, 30); xF = xE + rng.uniform(0, 30)
            if rng.random() < 0.05: xB = xE
            if rng.random() < 0.05: xF = xE
            vB, vF, vT = rng.uniform(0, 35), rng.uniform(0, 35), rng.uniform(5, 40)
            if rng.random() < 0.1: xB, vB = -1e10, 0.0
            if rng.random() < 0.1: xF, vF = 1e10, 100.0
            p = [xB, vB, xF, vF, xE, vE, vT]
            line, _ = ref.raw("ttrg " + " ".join(repr(x) for x in p))
            r = float(line.split()[1])
            g = oracle.lib.orc_time_to_reach_gap((C.c_double * 7)(*p))
            assert r == g or (math.isnan(r) and math.isnan(g)), (p, r, g)
    finally:
        ref.close()
