#!/usr/bin/env python
"""Regenerates tests/golden/*.json from the REFERENCE'S OWN MACHINE CODE (oracle/_ref/ref_harness, which dlopens
the reference's mc_sim_highway.so).  Run in the build container (needs oracle/_ref, i.e. /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the scene, the call and what the reference binary returned: per-rollout StatusOneSim, the
number of rand() draws, and the full agentsStatesHistory() of the first rollouts.  Random draws are injected through
the interposed rand() as the counter stream of include/mcsim_rng.h; exp/pow are the pinned definitions of
include/mcsim_math.h ("math": "mcs") unless stated otherwise.  Floats are written with repr() and round-trip exactly.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle.pyoracle import RefHarness  # noqa: E402
import scenes  # noqa: E402


def scene_dict(s):
    return {"corridors": s.corridors, "agents": s.agents}


def rollout_case(tag, s, ego, action, gap, dt, steps, min_steps, seed, first, count, n_hist=1, math="mcs"):
    ref = RefHarness(math=math)
    ref.load(s, dt, steps, min_steps)
    rs = ref.rollouts(ego, action, gap, first, count, seed, hist=True)
    ref.close()
    for i, r in enumerate(rs):
        if i >= n_hist:
            r.pop("hist", None)
        else:
            r["hist"] = {str(k): [list(t) for t in v] for k, v in r["hist"].items()}
    return {"tag": tag, "kind": "rollouts", "scene": scene_dict(s), "ego": ego, "action": action, "gap": gap, "dt": dt,
            "steps": steps, "min_steps": min_steps, "seed": seed, "first": first, "count": count, "math": math,
            "rollouts": rs}


def order_case(tag, s):
    ref = RefHarness()
    ref.load(s, 0.3, 40, 30)
    agents, corridors = ref.order()
    ref.close()
    return {"tag": tag, "kind": "order", "scene": scene_dict(s), "agents": agents, "corridors": corridors}


def decision_cases(tag, s, ids, key=99):
    """The four Python-facing single-decision wrappers on a fresh Environment."""
    ref = RefHarness()
    ref.load(s, 0.3, 40, 30)
    out = []
    for i in ids:
        for req in (0, 1):
            out.append({"cmd": "mobil %d 0.3 0.0 %d" % (i, req), "answer": ref.decide(key, "mobil %d 0.3 0.0 %d" % (i, req))})
        y = [a["y"] for a in s.agents if a["id"] == i][0]
        top = max(c["left_y"] for c in s.corridors)
        bottom = min(c["right_y"] for c in s.corridors)
        for d in scenes.ACTIONS:
            # without a neighbour lane on that side the binary returns an uninitialised stack slot as ax (so:0xf6263)
            if (d == "left" and y > top - 3.5) or (d == "right" and y < bottom + 3.5):
                continue
            out.append({"cmd": "fixed_dir %d %s" % (i, d), "answer": ref.decide(key, "fixed_dir %d %s" % (i, d))})
    ref.close()
    return {"tag": tag, "kind": "decisions", "scene": scene_dict(s), "key": key, "calls": out}


def merging_decision_cases(tag, s, d, gaps, key=5):
    """fixedGapMergingBehavior / closestGapMergingBehavior for every vehicle that has a lane on that side."""
    ref = RefHarness()
    ref.load(s, 0.3, 40, 10)
    out = []
    for a in s.agents:
        if not scenes.has_side_lane(s, a, d):
            continue
        for gap in [-2] + list(gaps):
            cmd = "fixed_gap %d %s %d" % (a["id"], d, gap)
            out.append({"cmd": cmd, "answer": ref.decide(key, cmd)})
        cmd = "closest_gap %d %s" % (a["id"], d)
        out.append({"cmd": cmd, "answer": ref.decide(key, cmd)})
    ref.close()
    return {"tag": tag, "kind": "decisions", "scene": scene_dict(s), "key": key, "calls": out}


def ttrg_case():
    """timeToReachGap on a fixed list of argument tuples (regimes: gap ahead / behind / alongside, and boundaries)."""
    rng = random.Random(77)
    ref = RefHarness()
    rows = []
    for i in range(300):
        xE, vE = rng.uniform(0, 300), rng.uniform(0, 35)
        m = i % 3
        if m == 0:
            xB = xE + rng.uniform(-5, 100); xF = xB + rng.uniform(0, 60)
        elif m == 1:
            xF = xE - rng.uniform(-5, 100); xB = xF - rng.uniform(0, 60)
        else:
            xB = xE - rng.uniform(0, 30); xF = xE + rng.uniform(0, 30)
        if i % 17 == 0: xB = xE
        if i % 19 == 0: xF = xE
        vB, vF, vT = rng.uniform(0, 35), rng.uniform(0, 35), rng.uniform(5, 40)
        if i % 11 == 0: xB, vB = -1e10, 0.0
        if i % 13 == 0: xF, vF = 1e10, 100.0
        p = [xB, vB, xF, vF, xE, vE, vT]
        line, _ = ref.raw("ttrg " + " ".join(repr(x) for x in p))
        rows.append({"args": p, "result": line.split()[1]})
    ref.close()
    return {"tag": "time_to_reach_gap", "kind": "ttrg", "rows": rows}


def known_answer_libc():
    """probe scene, glibc srand(1) before agent construction and before each action, runMTimes M=20,
    platform libm — the reference's stock configuration."""
    out = []
    for action in ("keep_lane", "left", "right"):
        ref = RefHarness(math="libm")
        ref.load(scenes.probe_scene(), 0.3, 40, 30)
        ref._send("rng libc_after_build 1")
        line, _ = ref.raw("runm 20 0 %s -2" % action, tag="runm")
        ref.close()
        out.append({"action": action, "line": line})
    return {"tag": "known_answer_libc_srand1", "kind": "runm_libc", "calls": out}


def main():
    cases = []
    p = scenes.probe_scene()
    cases.append(order_case("probe_order", p))
    for act in scenes.ACTIONS:
        cases.append(rollout_case("probe_%s" % act, p, 0, act, -2, 0.3, 40, 30, 7, 0, 16, n_hist=2))
    cases.append(decision_cases("probe_decisions", p, [0, 1, 2, 5]))
    for tag, s, ego, dt, steps, min_steps, seed in scenes.lane_change_cases(6, sizes=(5, 9, 14, 20, 31, 33)):
        cases.append(order_case(tag + "_order", s))
        for act in scenes.ACTIONS:
            cases.append(rollout_case("%s_%s" % (tag, act), s, ego, act, -2, dt, steps, min_steps, seed, 0, 4,
                                      n_hist=1 if act in ("keep_lane", "left") else 0))
    for tag, s, ego, d, gaps, seed in scenes.merge_exit_cases(8, sizes=(4, 7, 12, 20)):
        for gi, gap in enumerate(gaps):
            cases.append(rollout_case("%s_gap%d" % (tag, gi), s, ego, d, gap, 0.3, 40, 10, seed, 0, 4, n_hist=1 if gi == 0 else 0))
        cases.append(merging_decision_cases(tag + "_decisions", s, d, gaps))
    cases.append(ttrg_case())
    cases.append(known_answer_libc())
    extra = os.path.join(HERE, "extra_cases.py")
    if os.path.exists(extra):
        import importlib.util
        spec = importlib.util.spec_from_file_location("extra_cases", extra)
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        cases.extend(m.cases(sys.modules[__name__]))
    for c in cases:
        with open(os.path.join(HERE, c["tag"] + ".json"), "w") as f:
            json.dump(c, f, separators=(",", ":"))
    print("wrote %d fixtures" % len(cases))


if __name__ == "__main__":
    main()
