"""Drop-in for the reference's precompiled boost-python module ``mc_sim_highway``.

Same class names, constructor argument orders, attributes and free functions as the binding table in the reference's
mc_sim_highway.so (``init_module_mc_sim_highway`` 0xc2bf0; strings in .rodata) — the names the reference's Python imports
at behaviors.py:5-7, mcs_wrapper.py:3-5 and type.py:4-6.  Every compute call goes through the C ABI of libmcsim_b200.so
(include/mcsim.h) to the sm_100a kernels; there is no CPU implementation behind this module.

Make it importable under the reference's name with::

    import interactive_highway_simulation_b200 as ihs
    ihs.install_as_mc_sim_highway()          # sys.modules["mc_sim_highway"] = this module
    from behaviors import *                  # the reference's own Python now runs on the GPU path

Differences a caller can observe (all documented in DESIGN.md):
  * random draws come from a counter-based stream per rollout (include/mcsim_rng.h) instead of the process-global libc
    ``rand()``; ``seed(s)`` fixes the stream, results are reproducible and independent of thread interleaving;
  * ``simulationMultiThreadBatch`` (extension) evaluates all candidate actions of one decision in a single launch.
"""
import ctypes as _C

from . import _abi as _abi
from . import backend as _backend

__all__ = [
    "Point", "Line", "Corridor", "MapMultiLane", "IdmParam", "RSSParam", "YieldingParam", "Agent", "Environment", "Action",
    "ActionWithType", "SimParameter", "BehaviorFeature", "StatusOneSim", "State", "IdmMatch", "Simulation",
    "simulationMultiThread", "simulationMultiThreadBatch", "laneChangeMobilBehavior", "fixedDirectionLaneChangeBehavior",
    "fixedGapMergingBehavior", "closestGapMergingBehavior", "seed", "set_device",
    # legacy single-lane API: names only (imported at mcs_wrapper.py:4, never called; arithmetic lives in an absent library)
    "MapMerging", "Vehicle", "Indicator", "PositionHistory", "simulationMerging", "simulationMergingOnce",
    "simulationMergingMultiThread", "getMergingAction", "getMergingActionClosestGap", "getMergingActionClosestGapId",
    "rssSafe", "rssSafeRelax",
]

# ------------------------------------------------------------------------------------------------ value types


class Point:
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = float(x), float(y)

    def __repr__(self):
        return "Point(%r, %r)" % (self.x, self.y)


class Line:
    def __init__(self, p1=None, p2=None):
        self.p1 = p1 if p1 is not None else Point()
        self.p2 = p2 if p2 is not None else Point()


class Corridor:
    """Corridor(id, type, left, right, center, vLimit) — mc_sim_highway.so make_holder<6>; fields at Corridor+0..0x98."""

    def __init__(self, id, type, left, right, center, vLimit):
        self.id, self.type = int(id), str(type)
        self.left, self.right, self.center = left, right, center
        self.vLimit = float(vLimit)
        self.leftId = self.rightId = None

    # the binding registers the three lines under the attribute names l / r / m (mcs_wrapper.py:114 reads `.m.p2.x`)
    l = property(lambda self: self.left, lambda self, v: setattr(self, "left", v))
    r = property(lambda self: self.right, lambda self, v: setattr(self, "right", v))
    m = property(lambda self: self.center, lambda self, v: setattr(self, "center", v))

    def setLeftAndRightCorridorId(self, left_id, right_id):
        # MapMultiLane recomputes both from the lanes' centre offsets (0x10a780-0x10a839); kept for API parity
        self.leftId, self.rightId = left_id, right_id


class MapMultiLane:
    def __init__(self, corridors):
        self.corridors = list(corridors)


class IdmParam:
    """IdmParam(accMax, accMin, minDis, accCom, aEmergency, thw): constructor order, not memory order."""

    def __init__(self, accMax=1.6, accMin=-3.0, minDis=2.0, accCom=-2.0, aEmergency=-8.0, thw=1.2):
        self.accMax, self.accMin, self.minDis = float(accMax), float(accMin), float(minDis)
        self.accCom, self.aEmergency, self.thw = float(accCom), float(aEmergency), float(thw)

    def _memory_order(self):
        return (self.accMax, self.minDis, self.accMin, self.accCom, self.aEmergency, self.thw)

    # attribute names of the binding (constructor keywords are the long names above)
    am = property(lambda self: self.accMax, lambda self, v: setattr(self, "accMax", float(v)))
    amin = property(lambda self: self.accMin, lambda self, v: setattr(self, "accMin", float(v)))
    minD = property(lambda self: self.minDis, lambda self, v: setattr(self, "minDis", float(v)))
    aCom = property(lambda self: self.accCom, lambda self, v: setattr(self, "accCom", float(v)))
    aE = property(lambda self: self.aEmergency, lambda self, v: setattr(self, "aEmergency", float(v)))
    t = property(lambda self: self.thw, lambda self, v: setattr(self, "thw", float(v)))


class RSSParam:
    def __init__(self, rTimeOther=0.7, rTimeEgo=0.4, aMinBrake=-6.0, aMaxBrake=-6.0, aSoft=1.8, dSoft=1.2):
        self.rTimeOther, self.rTimeEgo = float(rTimeOther), float(rTimeEgo)
        self.aMinBrake, self.aMaxBrake = float(aMinBrake), float(aMaxBrake)
        self.aSoft, self.dSoft = float(aSoft), float(dSoft)

    def _memory_order(self):
        return (self.rTimeOther, self.rTimeEgo, self.aMinBrake, self.aMaxBrake, self.aSoft, self.dSoft)

    rTOther = property(lambda self: self.rTimeOther, lambda self, v: setattr(self, "rTimeOther", float(v)))
    rTEgo = property(lambda self: self.rTimeEgo, lambda self, v: setattr(self, "rTimeEgo", float(v)))
    aMin = property(lambda self: self.aMinBrake, lambda self, v: setattr(self, "aMinBrake", float(v)))
    aMax = property(lambda self: self.aMaxBrake, lambda self, v: setattr(self, "aMaxBrake", float(v)))


class YieldingParam:
    """YieldingParam() (defaults at .rodata 0x12ebe0) or YieldingParam(bias, a, b, c, politenessFactor, deltaA, biasA)."""

    def __init__(self, bias=-0.5, a=1.81, b=-4.8, c=-1.1, politenessFactor=0.9, deltaA=0.5, biasA=0.51):
        self.bias, self.a, self.b, self.c = float(bias), float(a), float(b), float(c)
        self.politenessFactor, self.deltaA, self.biasA = float(politenessFactor), float(deltaA), float(biasA)

    def _memory_order(self):
        return (self.bias, self.a, self.b, self.c, self.politenessFactor, self.deltaA, self.biasA)

    pf = property(lambda self: self.politenessFactor, lambda self, v: setattr(self, "politenessFactor", float(v)))
    da = property(lambda self: self.deltaA, lambda self, v: setattr(self, "deltaA", float(v)))
    ba = property(lambda self: self.biasA, lambda self, v: setattr(self, "biasA", float(v)))


class State:
    def __init__(self, x=0.0, y=0.0, v=0.0, a=0.0):
        self.x, self.y, self.v, self.a = x, y, v, a

    def __repr__(self):
        return "State(x=%r, y=%r, v=%r, a=%r)" % (self.x, self.y, self.v, self.a)


class IdmMatch:
    def __init__(self, id=-1, dis=0.0, v=0.0):
        self.id, self.dis, self.v = int(id), float(dis), float(v)


class Agent:
    """Agent(id, x, y, vx, vy, vd, a, l, w, IdmParam, RSSParam, YieldingParam, behavior) — make_holder<13>."""

    def __init__(self, id, x, y, vx, vy, vd, a, l, w, idm, rss, yp, behavior):
        self.id = int(id)
        self.x, self.y, self.vx, self.vy = float(x), float(y), float(vx), float(vy)
        self.vd, self.a, self.l, self.w = float(vd), float(a), float(l), float(w)
        self.idm, self.rss, self.yp = idm, rss, yp
        self.behavior = str(behavior)
        # fields the reference's Python pokes after construction (behaviors.py:75-80)
        self.targetLaneId = -1
        self.commitToDecision = "not_decided"
        self.commitKeepLaneStep = 0
        self.statesHistory = [State(self.x, self.y, self.vx, self.a)]

    def setDesiredV(self, v):
        self.vd = float(v)

    def _fill(self, o):
        o.id = self.id
        o.behavior = _abi.BEHAVIORS.get(self.behavior, _abi.BEH_OTHER)
        o.x, o.y, o.v, o.vy, o.vd, o.a, o.l, o.w = self.x, self.y, self.vx, self.vy, self.vd, self.a, self.l, self.w
        o.idm[:] = self.idm._memory_order()
        o.rss[:] = self.rss._memory_order()
        o.yp[:] = self.yp._memory_order()
        o.commit = _abi.COMMITS.get(self.commitToDecision, _abi.COMMITS["not_decided"])
        o.commit_keep_lane_step = int(self.commitKeepLaneStep)
        o.target_lane_id = int(self.targetLaneId)


class Action:
    def __init__(self, ax=0.0, vy=0.0):
        self.ax, self.vy = float(ax), float(vy)

    def __repr__(self):
        return "Action(ax=%r, vy=%r)" % (self.ax, self.vy)


class ActionWithType:
    def __init__(self, ax=0.0, vy=0.0, commitToDecision="not_decided", commitKeepLaneStep=0, targetLaneId=-1):
        self.ax, self.vy = ax, vy
        self.commitToDecision, self.commitKeepLaneStep, self.targetLaneId = commitToDecision, commitKeepLaneStep, targetLaneId


class SimParameter:
    def __init__(self, dt=0.3, steps=40, minSteps=30):
        self.dt, self.steps, self.minSteps = float(dt), int(steps), int(minSteps)

    def _c(self):
        return _abi.SimParam(self.dt, self.steps, self.minSteps)


class BehaviorFeature:
    _FIELDS = ("fallBackRate", "successRate", "utility", "comfort", "expectedStepRatio", "utilityObjects", "comfortObjects",
               "fromNSims")

    def __init__(self, c=None):
        for f in self._FIELDS:
            setattr(self, f, getattr(c, f) if c is not None else 0)

    def __repr__(self):
        return "BehaviorFeature(%s)" % ", ".join("%s=%r" % (f, getattr(self, f)) for f in self._FIELDS)


class StatusOneSim:
    _FIELDS = ("finish", "fallBack", "utility", "comfort", "utilityObjects", "comfortObjects", "finishStep")

    def __init__(self, c=None):
        for f in self._FIELDS:
            setattr(self, f, getattr(c, f) if c is not None else 0)
        if c is not None:
            self.finish, self.fallBack = bool(c.finish), bool(c.fallBack)


# ------------------------------------------------------------------------------------------------ device plumbing
_device = 0
_seed_base = 1            # libc's default rand() state is srand(1); a fresh process is reproducible like the reference's
_calls = 0                # calls made since seed(): call k draws from the stream of seed _seed_base + k
_cache = {"key": None, "scene": None}


def seed(s):
    """Fix the random stream of every following call (the reference never seeds libc rand())."""
    global _seed_base, _calls
    _seed_base = int(s) & 0xFFFFFFFFFFFFFFFF
    _calls = 0


def set_device(ordinal):
    global _device
    _device = int(ordinal)
    _drop_cache()


def _drop_cache():
    if _cache["scene"] is not None:
        _cache["scene"].close()
    _cache["key"], _cache["scene"] = None, None


def _next_seed():
    # one seed per call; the candidates of a batch take consecutive seeds, exactly what mcs_rollouts gives them (seed + c)
    global _calls
    _calls += 1
    return (_seed_base + _calls - 1) & 0xFFFFFFFFFFFFFFFF


def _peek_seed():
    """the seed the next call will take (the device-side walk of the outer step makes its laneChangeMobilBehavior calls with
    consecutive seeds from here and reports how many it made: _advance_seeds)"""
    return (_seed_base + _calls) & 0xFFFFFFFFFFFFFFFF


def _advance_seeds(k):
    global _calls
    _calls += int(k)


def _pack(map_, agents):
    cs = (_abi.Corridor * len(map_.corridors))()
    for o, c in zip(cs, map_.corridors):
        o.id = c.id
        o.type = _abi.LANE_TYPES.get(c.type, _abi.LANE_OTHER)
        o.left[:] = [c.left.p1.x, c.left.p1.y, c.left.p2.x, c.left.p2.y]
        o.right[:] = [c.right.p1.x, c.right.p1.y, c.right.p2.x, c.right.p2.y]
        o.center[:] = [c.center.p1.x, c.center.p1.y, c.center.p2.x, c.center.p2.y]
        o.v_limit = c.vLimit
    as_ = (_abi.Agent * len(agents))()
    for o, a in zip(as_, agents):
        a._fill(o)
    return cs, as_


def _scene(map_, agents):
    """Device-resident scene for (map, agents); the 3-5 calls of one decision share one upload (content-keyed)."""
    cs, as_ = _pack(map_, agents)
    key = (bytes(cs), bytes(as_), _device)
    if _cache["key"] != key:
        _drop_cache()
        _cache["scene"] = _backend.DeviceScene(cs, as_, device=_device)
        _cache["key"] = key
    return _cache["scene"]


def _action_enum(name):
    return _abi.ACTIONS.get(name, _abi.ACT_INVALID)


# ------------------------------------------------------------------------------------------------ free functions
def simulationMultiThreadBatch(n, simParam, map_, agents, egoId, commands):
    """Extension: all candidate commands [(action, gapId), ...] of one decision in ONE launch.  Returns one
    BehaviorFeature per command, each identical to what simulationMultiThread returns for that command alone when the
    calls are made in the same order after the same seed()."""
    scene = _scene(map_, agents)
    cands = (_abi.Candidate * len(commands))(*[_abi.Candidate(int(egoId), _action_enum(a), int(g), 0) for a, g in commands])
    base = [_next_seed() for _ in commands][0]
    feats, _ = scene.rollouts(cands, simParam._c(), int(n), base)
    return [BehaviorFeature(f) for f in feats]


def simulationMultiThread(n, simParam, map_, agents, egoId, action, gapId):
    """simulationMultiThread(n, SimParameter, MapMultiLane, list[Agent], egoId, action, gapId) -> BehaviorFeature:
    8 groups x n rollouts of the ego command, mean of the per-group means (runMultiThreads 0xbb910)."""
    return simulationMultiThreadBatch(n, simParam, map_, agents, egoId, [(action, gapId)])[0]


class Environment:
    """Environment(MapMultiLane, list[Agent]) — agents, map, reset(), setEgoAgentIdAndDrivingBehavior()."""

    def __init__(self, map_, agents):
        self.map = map_
        self.agents = list(agents)
        self._ego = None

    def reset(self):
        pass

    def setEgoAgentIdAndDrivingBehavior(self, egoId, action, gapId=-2):
        self._ego = (int(egoId), str(action), int(gapId))
        return any(a.id == int(egoId) for a in self.agents)


def laneChangeMobilBehavior(env, agentId, idmAction, requireRss):
    """laneChangeBehaviorMOBILWrapper 0xb9960 -> ActionWithType (behaviors.py:81-85)."""
    out = _scene(env.map, env.agents).decide_mobil(int(agentId), idmAction.ax, idmAction.vy, bool(requireRss), _next_seed())
    return ActionWithType(out.ax, out.vy, _abi.COMMIT_NAMES[out.commit], out.commit_keep_lane_step, out.target_lane_id)


def fixedDirectionLaneChangeBehavior(env, agentId, decision):
    out = _scene(env.map, env.agents).decide_fixed_direction(int(agentId), _action_enum(decision))
    return Action(out.ax, out.vy)


def fixedGapMergingBehavior(env, agentId, direction, gapId):
    out = _scene(env.map, env.agents).decide_fixed_gap(int(agentId), _action_enum(direction), int(gapId))
    return Action(out.ax, out.vy)


def closestGapMergingBehavior(env, agentId, direction):
    out = _scene(env.map, env.agents).decide_closest_gap(int(agentId), _action_enum(direction))
    return Action(out.ax, out.vy)


class Simulation:
    """Simulation(SimParameter, MapMultiLane, list[Agent]) — run(), runMultiTimes(n), reset(), agentsStatesHistory(),
    setEgoAgentIdAndDrivingBehavior(egoId, action, gapId) (0x10d9d0 / 0x10e790 / 0x10f8f0)."""

    def __init__(self, simParam, map_, agents):
        self.param, self.map, self.agents = simParam, map_, list(agents)
        self._ego = (-1, "keep_lane", -2)
        self._hist = None
        self._seed = _next_seed()
        self._runs = 0

    def setEgoAgentIdAndDrivingBehavior(self, egoId, action, gapId=-2):
        self._ego = (int(egoId), str(action), int(gapId))
        return any(a.id == int(egoId) for a in self.agents)

    def reset(self):
        self._hist = None

    def run(self):
        scene = _scene(self.map, self.agents)
        cand = _abi.Candidate(self._ego[0], _action_enum(self._ego[1]), self._ego[2], 0)
        sp = self.param._c()
        buf, ns, st = scene.trajectories(cand, sp, self._seed, self._runs)
        self._runs += 1
        stride = (sp.steps + 1) * 4
        self._hist = {a.id: [State(*buf[i * stride + 4 * t: i * stride + 4 * t + 4]) for t in range(ns)]
                      for i, a in enumerate(self.agents)}
        return StatusOneSim(st)

    def runMultiTimes(self, n):
        """One group of n rollouts (runMTimes / Simulation::runMultipleTimes semantics: plain means over the group)."""
        scene = _scene(self.map, self.agents)
        cand = (_abi.Candidate * 1)(_abi.Candidate(self._ego[0], _action_enum(self._ego[1]), self._ego[2], 0))
        sp = self.param._c()
        st = scene.rollouts_range(cand, sp, self._seed, self._runs, int(n))
        self._runs += int(n)
        return BehaviorFeature(_backend.reduce_features(st, 1, 1, int(n), sp.steps)[0])

    def agentsStatesHistory(self):
        if self._hist is None:
            return {a.id: list(a.statesHistory) for a in self.agents}
        return self._hist


# ------------------------------------------------------------------------------------------------ legacy names
class _Legacy:
    """The single-lane merging API of the reference binary keeps its arithmetic in libmc_sim_highway.so.0, which the
    reference repository does not ship; its Python never calls it.  The names exist so that the reference's imports work."""
    _name = "legacy"

    def __init__(self, *a, **k):
        raise NotImplementedError("%s belongs to the reference's legacy single-lane API, which is not part of the rollout "
                                  "path (its implementation is absent from the reference repository)" % self._name)


def _legacy_class(name):
    return type(name, (_Legacy,), {"_name": name})


def _legacy_fn(name):
    def fn(*a, **k):
        raise NotImplementedError("%s belongs to the reference's legacy single-lane API, which is not part of the rollout "
                                  "path (its implementation is absent from the reference repository)" % name)
    fn.__name__ = name
    return fn


MapMerging, Vehicle, Indicator, PositionHistory = (_legacy_class(n) for n in ("MapMerging", "Vehicle", "Indicator", "PositionHistory"))
simulationMerging, simulationMergingOnce, simulationMergingMultiThread = (_legacy_fn(n) for n in (
    "simulationMerging", "simulationMergingOnce", "simulationMergingMultiThread"))
getMergingAction, getMergingActionClosestGap, getMergingActionClosestGapId = (_legacy_fn(n) for n in (
    "getMergingAction", "getMergingActionClosestGap", "getMergingActionClosestGapId"))
rssSafe, rssSafeRelax = _legacy_fn("rssSafe"), _legacy_fn("rssSafeRelax")
