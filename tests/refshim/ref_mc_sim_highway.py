"""TEST INFRASTRUCTURE ONLY — a `mc_sim_highway` module backed by THE REFERENCE'S OWN MACHINE CODE.

Value classes are the drop-in's plain-Python ones; every compute function is answered by oracle/_ref/ref_harness (which
dlopens the reference's mc_sim_highway.so) with the random-stream convention of the product's drop-in (one seed per call,
seed() resets the sequence), and every call is appended to `RECORD` with its inputs and the reference's answer.
"""
import itertools

from interactive_highway_simulation_b200 import _abi as abi
from interactive_highway_simulation_b200.mc_sim_highway import (  # noqa: F401  (re-exported value types)
    Point, Line, Corridor, MapMultiLane, IdmParam, RSSParam, YieldingParam, Agent, Environment, Action, ActionWithType,
    SimParameter, BehaviorFeature, StatusOneSim, State, IdmMatch, MapMerging, Vehicle, Indicator, PositionHistory)
from oracle.pyoracle import Oracle, RefHarness

RECORD = []
_seed_base = 1
_calls = itertools.count()
_ref = None
_orc = None


def seed(s):
    global _seed_base, _calls
    _seed_base = int(s) & 0xFFFFFFFFFFFFFFFF
    _calls = itertools.count()


def _next_seed():
    return (_seed_base + next(_calls)) & 0xFFFFFFFFFFFFFFFF


def _harness():
    global _ref, _orc
    if _ref is None:
        _ref = RefHarness(math="mcs")
        _orc = Oracle(0)
    return _ref


def scene_record(map_, agents):
    cs = [[c.id, c.type, c.left.p1.x, c.left.p1.y, c.left.p2.x, c.left.p2.y, c.right.p1.x, c.right.p1.y, c.right.p2.x,
           c.right.p2.y, c.center.p1.x, c.center.p1.y, c.center.p2.x, c.center.p2.y, c.vLimit] for c in map_.corridors]
    as_ = [[a.id, a.x, a.y, a.vx, a.vy, a.vd, a.a, a.l, a.w] + list(a.idm._memory_order()) + list(a.rss._memory_order()) +
           list(a.yp._memory_order()) + [a.behavior, a.commitToDecision, int(a.commitKeepLaneStep), int(a.targetLaneId)]
           for a in agents]
    return {"corridors": cs, "agents": as_}


def scene_text(rec):
    f = lambda v: repr(float(v))
    out = ["clear"]
    for c in rec["corridors"]:
        out.append("corridor %d %s %s" % (c[0], c[1], " ".join(f(v) for v in c[2:])))
    for a in rec["agents"]:
        line = "agent %d %s %s" % (a[0], " ".join(f(v) for v in a[1:28]), a[28])
        if a[29] != "not_decided" or a[30] != 0 or a[31] != -1:
            line += " %s %d %d" % (a[29], a[30], a[31])
        out.append(line)
    return "\n".join(out) + "\n"


def _load(rec, sp):
    h = _harness()
    h._send(scene_text(rec))
    h._send("sim %r %d %d" % (float(sp.dt), sp.steps, sp.minSteps))
    return h


def simulationMultiThread(n, simParam, map_, agents, egoId, action, gapId):
    rec = scene_record(map_, agents)
    s = _next_seed()
    h = _load(rec, simParam)
    count = 8 * int(n)
    rs = h.rollouts(int(egoId), str(action), int(gapId), 0, count, s)
    st = (abi.Status * count)()
    for o, r in zip(st, rs):
        o.ok, o.finish, o.fallBack, o.finishStep = r["ok"], r["finish"], r["fallBack"], r["finishStep"]
        o.utility, o.comfort, o.utilityObjects, o.comfortObjects = r["utility"], r["comfort"], r["utilityObjects"], r["comfortObjects"]
    f = _orc.reduce(st, 8, int(n), simParam.steps)
    out = BehaviorFeature(f)
    RECORD.append({"fn": "simulationMultiThread", "scene": rec, "n": int(n), "sim": [simParam.dt, simParam.steps, simParam.minSteps],
                   "ego": int(egoId), "action": str(action), "gap": int(gapId), "seed": s,
                   "out": {k: getattr(out, k) for k in BehaviorFeature._FIELDS}})
    return out


def simulationMultiThreadBatch(n, simParam, map_, agents, egoId, commands):
    return [simulationMultiThread(n, simParam, map_, agents, egoId, a, g) for a, g in commands]


def _decide(env, cmd, fn, extra, seed_value=0):
    rec = scene_record(env.map, env.agents)
    h = _load(rec, SimParameter(0.3, 40, 30))
    ans = h.decide(seed_value, cmd)
    out = {"ax": float(ans["ax"]), "vy": float(ans["vy"])}
    if "commit" in ans:
        out.update(commit=ans["commit"], keep=int(ans["keep"]), target=int(ans["target"]))
    r = {"fn": fn, "scene": rec, "seed": seed_value, "out": out}
    r.update(extra)
    RECORD.append(r)
    return out


def laneChangeMobilBehavior(env, agentId, idmAction, requireRss):
    s = _next_seed()
    o = _decide(env, "mobil %d %r %r %d" % (int(agentId), float(idmAction.ax), float(idmAction.vy), 1 if requireRss else 0),
                "laneChangeMobilBehavior", {"id": int(agentId), "ax": float(idmAction.ax), "vy": float(idmAction.vy),
                                            "require_rss": bool(requireRss)}, s)
    return ActionWithType(o["ax"], o["vy"], o["commit"], o["keep"], o["target"])


def fixedDirectionLaneChangeBehavior(env, agentId, decision):
    o = _decide(env, "fixed_dir %d %s" % (int(agentId), decision), "fixedDirectionLaneChangeBehavior",
                {"id": int(agentId), "decision": str(decision)})
    return Action(o["ax"], o["vy"])


def fixedGapMergingBehavior(env, agentId, direction, gapId):
    o = _decide(env, "fixed_gap %d %s %d" % (int(agentId), direction, int(gapId)), "fixedGapMergingBehavior",
                {"id": int(agentId), "direction": str(direction), "gap": int(gapId)})
    return Action(o["ax"], o["vy"])


def closestGapMergingBehavior(env, agentId, direction):
    o = _decide(env, "closest_gap %d %s" % (int(agentId), direction), "closestGapMergingBehavior",
                {"id": int(agentId), "direction": str(direction)})
    return Action(o["ax"], o["vy"])


class Simulation:   # the reference's Python never instantiates the inner Simulation (simulation_multilane shadows the name)
    def __init__(self, *a, **k):
        raise NotImplementedError
