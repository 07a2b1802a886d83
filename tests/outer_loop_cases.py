"""Shared by the outer-loop tests: the scenario records of tests/golden/scenario_*.json.gz (written by the reference's own
Python, tests/golden/make_scenario_golden.py) as free-running cases for the product's outer loop
(interactive_highway_simulation_b200.simulation_multilane / outer_env / agent / behaviors).

CPU suite: the host logic alone — the device geometry is stood in for by the oracle (`OracleDeviceMap`, test
infrastructure) and the inner simulator by the answers recorded from the reference (`RecordedInnerSimulator`), which also
checks that every boundary call the product's outer loop makes carries the inputs the reference's call carried.
GPU suite: nothing is stood in for."""
import ctypes as C
import gzip
import json
import os
import random

import numpy as np

from oracle import env_oracle as eo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
SCENE_DIR = os.path.join(GOLDEN, "test_scene_epoch0")
TRACE_TOL = 1e-7     # metres / m/s over <= 100 steps: geometry differs from the reference's hypot / arctan2 by ~1e-13 per step


def load_scenario(tag):
    with gzip.open(os.path.join(GOLDEN, "scenario_%s.json.gz" % tag), "rt") as f:
        return json.load(f)


class OracleDeviceMap:
    """outer_env.DeviceMap with every answer computed by oracle/env_oracle.py (CPU tests only)."""

    def __init__(self, corridors, device=0):
        cs = list(corridors)
        pos = {c.id: k for k, c in enumerate(cs)}
        self.pm = eo.PreparedMap([{"id": c.id, "left": c.left, "right": c.right, "center": c.center,
                                   "left_index": pos[c.left_id] if c.left_id is not None else -1,
                                   "right_index": pos[c.right_id] if c.right_id is not None else -1} for c in cs])
        self.n_lanes = len(cs)
        self.center_cache = {}
        self._m = None

    def close(self):
        pass

    def match(self, x, y, n_scenes=1, arc_on_lanes=False):
        from interactive_highway_simulation_b200.outer_env import MatchTables
        assert n_scenes == 1
        xs, ys = [float(v) for v in np.ravel(x)], [float(v) for v in np.ravel(y)]
        lane, arc, order, start = eo.match(self.pm, xs, ys)
        self._m = (lane, arc, order, start)
        n = self._n = len(xs)
        t = MatchTables()
        t.n_agents, t.n_scenes = n, 1
        t.lane, t.arc = np.array([lane], np.int32), np.array([arc])
        t.order, t.start = np.array([order], np.int32), np.array([start], np.int32)
        t.arc_on_lane = np.array([eo.arc_on_lanes(self.pm, xs, ys)]) if arc_on_lanes else None
        t.nb_index, t.nb_dis = np.full((1, 10, n), -1, np.int32), np.zeros((1, 10, n))
        for i in range(n):
            a, b = eo.neighbors(self.pm, lane, arc, order, start, i, xs[i], ys[i])
            t.nb_index[0, :, i], t.nb_dis[0, :, i] = a, b
        return t

    def neighbors(self, scene, slot, qx, qy):
        q = len(slot)
        idx, dis = np.empty((10, q), np.int32), np.empty((10, q))
        for k in range(q):
            a, b = eo.neighbors(self.pm, *self._m, int(slot[k]), float(qx[k]), float(qy[k]))
            idx[:, k], dis[:, k] = a, b
        return idx, dis

    def frenet(self, ref_lane, x, y):
        out = [self.pm.ref[int(ref_lane)].to_frenet_frame(float(a), float(b)) for a, b in zip(x, y)]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])

    def project_center(self, lane, x, y):
        out = [eo.project_and_distance(self.pm.center[int(lane)], float(a), float(b)) for a, b in zip(x, y)]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])

    def step_along_path(self, ref_lane, x, y, v, ax, vy, dt):
        out = [eo.step_along_path(self.pm, int(l), float(a), float(b), float(c), float(d), float(e), dt)
               for l, a, b, c, d, e in zip(ref_lane, x, y, v, ax, vy)]
        return [np.array([o[k] for o in out]) for k in range(4)]

    def step_one(self, ref_lane, x, y, v, ax, vy, dt):
        return eo.step_along_path(self.pm, int(ref_lane), float(x), float(y), float(v), float(ax), float(vy), dt)

    def center_distance(self, lane, x, y):
        out = [eo.center_distance(self.pm, int(l), float(a), float(b)) for l, a, b in zip(lane, x, y)]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])

    def border_distance(self, lane, side, hull):
        return np.array([eo.border_distance(self.pm, int(l), int(s), h) for l, s, h in zip(lane, side, hull)])

    def set_lane_info(self, ids, types, v_limits):
        self.info = [{"id": int(i), "type": int(t), "v_limit": float(v)} for i, t, v in zip(ids, types, v_limits)]

    def build_scene(self, rq, ids, attrs, noise, device=0, records=True, scene=True):
        """mce_build_rollout_scene answered by the oracle's restatement of the builders (oracle/env_oracle.py build_scene);
        the "scene handle" is the pair of records, which RecordedInnerSimulator's stand-in scene checks and answers from."""
        from interactive_highway_simulation_b200 import _abi as abi
        if isinstance(ids, int):            # raw addresses of the environment's mirror arrays
            n = self._n
            ids = np.ctypeslib.as_array((C.c_int32 * n).from_address(ids))
            attrs = np.ctypeslib.as_array((C.c_double * (8 * n)).from_address(attrs)).reshape(8, n)
        n = len(ids)
        xs, ys = [float(v) for v in attrs[0]], [float(v) for v in attrs[1]]
        per = [dict(v=float(attrs[2, i]), vy=float(attrs[3, i]), a=float(attrs[4, i]), l=float(attrs[5, i]), w=float(attrs[6, i]),
                    vd=float(attrs[7, i])) for i in range(n)]
        ego = dict(behavior=rq.ego_behavior, vd=rq.ego_vd, idm=tuple(rq.ego_idm), rss=tuple(rq.ego_rss), yp=tuple(rq.ego_yp),
                   commit=rq.ego_commit, keep=rq.ego_keep_step, target=rq.ego_target_lane)
        cs, ag, actions = eo.build_scene(self.pm, self.info, self._m, xs, ys, per, [int(i) for i in ids], rq.kind, rq.ego_slot, ego,
                                         bool(rq.others_merging), [float(v) for v in noise] if noise is not None else None)
        c_arr, a_arr = (abi.Corridor * 8)(), (abi.Agent * 48)()
        for o, c in zip(c_arr, cs):
            o.id, o.type, o.v_limit = c[0], c[1], c[5]
            o.left[:], o.right[:], o.center[:] = c[2], c[3], c[4]
        for o, a in zip(a_arr, ag):
            o.id, o.behavior = a["id"], a["behavior"]
            o.x, o.y, o.v, o.vy, o.vd, o.a, o.l, o.w = a["x"], a["y"], a["v"], a["vy"], a["vd"], a["a"], a["l"], a["w"]
            o.idm[:], o.rss[:], o.yp[:] = a["idm"], a["rss"], a["yp"]
            o.commit, o.commit_keep_lane_step, o.target_lane_id = a["commit"], a["keep"], a["target"]
        return c_arr, len(cs), a_arr, len(ag), actions, ((c_arr, len(cs), a_arr, len(ag)) if scene else None)

    @staticmethod
    def create_scene(corridors, n_corridors, agents, n_agents, device=0):
        return (corridors, n_corridors, agents, n_agents)      # what RecordedInnerSimulator's stand-in scene checks


class _Rec:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class RecordedInnerSimulator:
    """Stands in for the free functions of the drop-in module: answers with what the reference's machine code answered
    at the same position of the call sequence, after checking that the call is the recorded call (same function, same
    vehicle, same command, every vehicle of the marshalled scene where the reference had it)."""

    def __init__(self, ms, calls, tol=1e-6):
        self.ms, self.calls, self.k, self.tol = ms, calls, 0, tol

    def _next(self, fn, env_map, agents, **want):
        assert self.k < len(self.calls), "more boundary calls than the reference made"
        c = self.calls[self.k]
        self.k += 1
        assert c["fn"] == fn, (self.k, c["fn"], fn)
        for key, v in want.items():
            assert c[key] == v, (self.k, fn, key, c[key], v)
        rec = c["scene"]
        assert [a[0] for a in rec["agents"]] == [a.id for a in agents], (self.k, fn)
        assert [k[0] for k in rec["corridors"]] == [k.id for k in env_map.corridors]
        for r, a in zip(rec["agents"], agents):
            got = [a.x, a.y, a.vx, a.vy, a.vd, a.a, a.l, a.w]
            assert all(abs(g - w) <= self.tol for g, w in zip(got, r[1:9])), (self.k, fn, a.id, got, r[1:9])
            assert a.behavior == r[28] and (a.commitToDecision, a.commitKeepLaneStep, a.targetLaneId) == tuple(r[29:32])
        return c

    def _next_scene(self, fn, handle, **want):
        """the same check for a scene the product's fast path marshalled itself (records from build_scene)"""
        from interactive_highway_simulation_b200 import _abi as abi
        cs, nc, ag, na = handle
        beh = {v: k for k, v in abi.BEHAVIORS.items()}
        agents = [_Rec(id=a.id, x=a.x, y=a.y, vx=a.v, vy=a.vy, vd=a.vd, a=a.a, l=a.l, w=a.w, behavior=beh[a.behavior],
                       commitToDecision=abi.COMMIT_NAMES[a.commit], commitKeepLaneStep=a.commit_keep_lane_step,
                       targetLaneId=a.target_lane_id) for a in ag[:na]]
        return self._next(fn, _Rec(corridors=[_Rec(id=c.id) for c in cs[:nc]]), agents, **want)

    def install(self, monkeypatch):
        ms = self.ms
        sim = self
        from interactive_highway_simulation_b200 import _abi as abi, backend
        act_names = {v: k for k, v in abi.ACTIONS.items()}

        class RecordedScene:
            """stands in for backend.DeviceScene over a marshalled scene: every decision / rollout call is the recorded one"""

            def __init__(self, handle):
                self.h = handle

            def close(self):
                pass

            def decide_mobil(self, agent_id, ax, vy, require_rss, seed):
                c = sim._next_scene("laneChangeMobilBehavior", self.h, id=int(agent_id), require_rss=bool(require_rss))
                assert abs(ax - c["ax"]) <= sim.tol and abs(vy - c["vy"]) <= sim.tol, (sim.k, ax, c["ax"])
                o = c["out"]
                return _Rec(ax=o["ax"], vy=o["vy"], commit=abi.COMMITS[o["commit"]], commit_keep_lane_step=o["keep"], target_lane_id=o["target"])

            def decide_fixed_direction(self, agent_id, action):
                c = sim._next_scene("fixedDirectionLaneChangeBehavior", self.h, id=int(agent_id), decision=act_names[action])
                return _Rec(ax=c["out"]["ax"], vy=c["out"]["vy"])

            def decide_fixed_gap(self, agent_id, direction, gap):
                c = sim._next_scene("fixedGapMergingBehavior", self.h, id=int(agent_id), direction=act_names[direction], gap=int(gap))
                return _Rec(ax=c["out"]["ax"], vy=c["out"]["vy"])

            def decide_closest_gap(self, agent_id, direction):
                c = sim._next_scene("closestGapMergingBehavior", self.h, id=int(agent_id), direction=act_names[direction])
                return _Rec(ax=c["out"]["ax"], vy=c["out"]["vy"])

            def rollouts(self, cands, sim_param, n, seed, want_status=False):
                out = []
                for cd in cands:
                    c = sim._next_scene("simulationMultiThread", self.h, n=int(n), ego=int(cd.ego_id), action=act_names[cd.action],
                                        gap=int(cd.gap_id))
                    out.append(_Rec(**c["out"]))
                return out, None

        monkeypatch.setattr(backend.DeviceScene, "adopt", classmethod(lambda cls, handle, na, nc: RecordedScene(handle)))

        def batch(n, sim_param, map_, agents, ego, commands):
            out = []
            for action, gap in commands:
                c = self._next("simulationMultiThread", map_, agents, n=int(n), ego=int(ego), action=action, gap=int(gap))
                f = ms.BehaviorFeature.__new__(ms.BehaviorFeature)
                f.__dict__.update(c["out"])
                out.append(f)
            return out

        def mobil(env, agent_id, idm_action, require_rss):
            c = self._next("laneChangeMobilBehavior", env.map, env.agents, id=int(agent_id), require_rss=bool(require_rss))
            assert abs(idm_action.ax - c["ax"]) <= self.tol and abs(idm_action.vy - c["vy"]) <= self.tol, (self.k, idm_action.ax, c["ax"])
            o = c["out"]
            return ms.ActionWithType(o["ax"], o["vy"], o["commit"], o["keep"], o["target"])

        def fixed_direction(env, agent_id, decision):
            c = self._next("fixedDirectionLaneChangeBehavior", env.map, env.agents, id=int(agent_id), decision=decision)
            return ms.Action(c["out"]["ax"], c["out"]["vy"])

        def fixed_gap(env, agent_id, direction, gap):
            c = self._next("fixedGapMergingBehavior", env.map, env.agents, id=int(agent_id), direction=direction, gap=int(gap))
            return ms.Action(c["out"]["ax"], c["out"]["vy"])

        def closest_gap(env, agent_id, direction):
            c = self._next("closestGapMergingBehavior", env.map, env.agents, id=int(agent_id), direction=direction)
            return ms.Action(c["out"]["ax"], c["out"]["vy"])

        monkeypatch.setattr(ms, "simulationMultiThreadBatch", batch)
        monkeypatch.setattr(ms, "laneChangeMobilBehavior", mobil)
        monkeypatch.setattr(ms, "fixedDirectionLaneChangeBehavior", fixed_direction)
        monkeypatch.setattr(ms, "fixedGapMergingBehavior", fixed_gap)
        monkeypatch.setattr(ms, "closestGapMergingBehavior", closest_gap)


def seed_streams(seed):
    random.seed(seed)
    np.random.seed(seed)


def compare_trace(env, trace_step, step, tol=TRACE_TOL):
    """every vehicle where the reference had it after this step (x, y, v, a)"""
    assert sorted(str(i) for i in env.agents) == sorted(trace_step), ("vehicles in the scene", step)
    worst = 0.0
    for i, a in env.agents.items():
        w = trace_step[str(i)]
        for g, r in zip((a.x, a.y, a.v, a.a), w):
            worst = max(worst, abs(g - r))
            assert abs(g - r) <= tol, ("step", step, "vehicle", i, (a.x, a.y, a.v, a.a), w)
    return worst
