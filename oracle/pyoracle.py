"""TEST INFRASTRUCTURE ONLY — Python drivers for the two checkers under oracle/:

* ``Oracle``      ctypes binding of oracle/libmcsim_oracle.so (the C++ restatement)
* ``RefHarness``  subprocess driver of oracle/_ref/ref_harness (the reference's own machine code)

plus ``SceneSpec``, a plain description of a scene that can be handed to either of them and to
the product.  Nothing under interactive_highway_simulation_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

from interactive_highway_simulation_b200 import _abi as abi

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmcsim_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_BIN = os.path.join(REF_DIR, "ref_harness")
REF_SO = os.path.join(REF_DIR, "mc_sim_highway.so")

LANE_NAMES = {v: k for k, v in abi.LANE_TYPES.items()}
BEH_NAMES = {v: k for k, v in abi.BEHAVIORS.items()}
ACT_NAMES = {v: k for k, v in abi.ACTIONS.items()}

DEFAULT_IDM = (1.6, 2.0, -3.0, -2.0, -8.0, 1.2)          # memory order: accMax minDis accMin accCom aE thw
DEFAULT_RSS = (0.7, 0.4, -6.0, -6.0, 1.8, 1.2)
DEFAULT_YP = (-0.5, 1.81, -4.8, -1.1, 0.9, 0.5, 0.51)


class SceneSpec:
    """corridors: dicts {id,type,left_y,right_y,center_y,x0,x1,v_limit}; agents: dicts (see add_agent)."""

    def __init__(self):
        self.corridors = []
        self.agents = []

    def add_corridor(self, id, type, left_y, right_y, center_y, v_limit, x0=-1000.0, x1=2000.0):
        self.corridors.append(dict(id=id, type=type, left_y=left_y, right_y=right_y, center_y=center_y,
                                   v_limit=v_limit, x0=x0, x1=x1))
        return self

    def add_agent(self, id, x, y, v, vd, behavior, vy=0.0, a=0.0, l=4.8, w=2.0, idm=DEFAULT_IDM, rss=DEFAULT_RSS,
                  yp=DEFAULT_YP, commit="not_decided", keep_step=0, target_lane=-1):
        self.agents.append(dict(id=id, x=x, y=y, v=v, vy=vy, vd=vd, a=a, l=l, w=w, idm=tuple(idm), rss=tuple(rss),
                                yp=tuple(yp), behavior=behavior, commit=commit, keep_step=keep_step,
                                target_lane=target_lane))
        return self

    # ---- C arrays for the oracle / product
    def c_corridors(self):
        arr = (abi.Corridor * len(self.corridors))()
        for o, c in zip(arr, self.corridors):
            o.id = c["id"]; o.type = abi.LANE_TYPES.get(c["type"], abi.LANE_OTHER)
            o.left[:] = [c["x0"], c["left_y"], c["x1"], c["left_y"]]
            o.right[:] = [c["x0"], c["right_y"], c["x1"], c["right_y"]]
            o.center[:] = [c["x0"], c["center_y"], c["x1"], c["center_y"]]
            o.v_limit = c["v_limit"]
        return arr

    def c_agents(self):
        arr = (abi.Agent * len(self.agents))()
        for o, a in zip(arr, self.agents):
            o.id = a["id"]; o.behavior = abi.BEHAVIORS.get(a["behavior"], abi.BEH_OTHER)
            o.x, o.y, o.v, o.vy, o.vd, o.a, o.l, o.w = a["x"], a["y"], a["v"], a["vy"], a["vd"], a["a"], a["l"], a["w"]
            o.idm[:] = a["idm"]; o.rss[:] = a["rss"]; o.yp[:] = a["yp"]
            o.commit = abi.COMMITS[a["commit"]]; o.commit_keep_lane_step = a["keep_step"]
            o.target_lane_id = a["target_lane"]
        return arr

    # ---- text for the reference harness
    def harness_text(self):
        f = lambda v: repr(float(v))
        out = ["clear"]
        for c in self.corridors:
            vals = [c["x0"], c["left_y"], c["x1"], c["left_y"], c["x0"], c["right_y"], c["x1"], c["right_y"],
                    c["x0"], c["center_y"], c["x1"], c["center_y"], c["v_limit"]]
            out.append("corridor %d %s %s" % (c["id"], c["type"], " ".join(f(v) for v in vals)))
        for a in self.agents:
            vals = [a["x"], a["y"], a["v"], a["vy"], a["vd"], a["a"], a["l"], a["w"]] + list(a["idm"]) + \
                list(a["rss"]) + list(a["yp"])
            line = "agent %d %s %s" % (a["id"], " ".join(f(v) for v in vals), a["behavior"])
            if a["commit"] != "not_decided" or a["keep_step"] != 0 or a["target_lane"] != -1:
                line += " %s %d %d" % (a["commit"], a["keep_step"], a["target_lane"])
            out.append(line)
        return "\n".join(out) + "\n"


def _kv(line):
    d = {}
    for tok in line.split():
        if "=" in tok:
            k, v = tok.split("=", 1)
            d[k] = v
    return d


class RefHarness:
    """One ref_harness process; commands are written to its stdin, one answer line (plus optional
    hist/agent lines) is read back per command."""

    def __init__(self, math="mcs", trace=None):
        if not (os.path.exists(REF_BIN) and os.path.exists(REF_SO)):
            raise FileNotFoundError("oracle/_ref not built (run oracle/ref_harness/build_ref.sh where /root/reference exists)")
        env = dict(os.environ)
        if trace:
            env["REF_TRACE_DRAWS"] = trace
        self.p = subprocess.Popen([REF_BIN, REF_SO], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True,
                                  bufsize=1, env=env)
        self._send("math %s" % math)

    def _send(self, text):
        self.p.stdin.write(text if text.endswith("\n") else text + "\n")
        self.p.stdin.flush()

    def _answer(self, tag):
        extra = []
        while True:
            line = self.p.stdout.readline()
            if not line:
                raise RuntimeError("ref_harness died")
            if line.startswith(tag + " "):
                return line.strip(), extra
            extra.append(line.strip())

    def load(self, scene, dt, steps, min_steps):
        self._send(scene.harness_text())
        self._send("sim %r %d %d" % (float(dt), steps, min_steps))

    @staticmethod
    def _status(line):
        d = _kv(line)
        return dict(ok=int(d.get("ok", 1)), finish=int(d["finish"]), fallBack=int(d["fallBack"]), finishStep=int(d["finishStep"]),
                    utility=float(d["utility"]), comfort=float(d["comfort"]), utilityObjects=float(d["utilityObjects"]),
                    comfortObjects=float(d["comfortObjects"]), draws=int(d["draws"]))

    @staticmethod
    def _hist(extra):
        h = {}
        for l in extra:
            if l.startswith("hist "):
                t = l.split()
                aid = int(t[1].split("=")[1]); n = int(t[2].split("=")[1])
                vals = [float(v) for v in t[3:3 + 4 * n]]
                h[aid] = [tuple(vals[4 * i:4 * i + 4]) for i in range(n)]
        return h

    def rollouts(self, ego, action, gap, first, count, seed, hist=False):
        self._send("rollouts %d %s %d %d %d %d%s" % (ego, action, gap, first, count, seed, " hist" if hist else ""))
        out = []
        for _ in range(count):
            line, extra = self._answer("rollout")
            st = self._status(line)
            if hist:
                st["hist"] = self._hist(extra)
            out.append(st)
        return out

    def run_counter(self, key, ego, action, gap, hist=False, dump=False):
        self._send("rng counter %d" % key)
        self._send("run %d %s %d%s%s" % (ego, action, gap, " hist" if hist else "", " dump" if dump else ""))
        line, extra = self._answer("run")
        st = self._status(line)
        st["hist"] = self._hist(extra)
        st["dump"] = [_kv(l) for l in extra if l.startswith("agent ")]
        return st

    def order(self):
        self._send("order")
        line, _ = self._answer("order")
        t = line.split()
        i = t.index("corridors")
        return [int(v) for v in t[2:i]], [int(v) for v in t[i + 1:]]

    def decide(self, key, cmd):
        self._send("rng counter %d" % key)
        self._send(cmd)
        line, _ = self._answer(cmd.split()[0])
        return _kv(line)

    def raw(self, cmd, tag=None):
        self._send(cmd)
        line, extra = self._answer(tag or cmd.split()[0])
        return line, extra

    def close(self):
        try:
            self.p.stdin.close()
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()


class Oracle:
    def __init__(self, math_mode=0):
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.orc_rollouts.restype = C.c_int
        L.orc_rollouts.argtypes = [C.POINTER(abi.Corridor), C.c_int, C.POINTER(abi.Agent), C.c_int, C.POINTER(abi.Candidate),
                                   C.POINTER(abi.SimParam), C.c_uint64, C.c_int64, C.c_int64, C.POINTER(abi.Status),
                                   C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.orc_rollouts_mt.restype = C.c_int
        L.orc_rollouts_mt.argtypes = [C.POINTER(abi.Corridor), C.c_int, C.POINTER(abi.Agent), C.c_int, C.POINTER(abi.Candidate),
                                      C.POINTER(abi.SimParam), C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.POINTER(abi.Status)]
        L.orc_reduce_features.restype = C.c_int
        L.orc_reduce_features.argtypes = [C.POINTER(abi.Status), C.c_int, C.c_int, C.c_int, C.POINTER(abi.Feature)]
        L.orc_scene_orders.restype = C.c_int
        L.orc_scene_orders.argtypes = [C.POINTER(abi.Corridor), C.c_int, C.POINTER(abi.Agent), C.c_int] + [C.POINTER(C.c_int32)] * 4
        L.orc_set_math_mode.argtypes = [C.c_int]
        for name, extra in (("orc_decide_mobil", [C.c_int32, C.POINTER(abi.Action), C.c_int, C.c_uint64, C.POINTER(abi.ActionWithType)]),
                            ("orc_decide_fixed_direction", [C.c_int32, C.c_int32, C.c_uint64, C.POINTER(abi.Action)]),
                            ("orc_decide_fixed_gap", [C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(abi.Action)]),
                            ("orc_decide_closest_gap", [C.c_int32, C.c_int32, C.c_uint64, C.POINTER(abi.Action)])):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(abi.Corridor), C.c_int, C.POINTER(abi.Agent), C.c_int] + extra
        L.orc_set_math_mode(math_mode)

    def set_math_mode(self, m):
        self.lib.orc_set_math_mode(m)

    def rollouts(self, scene, ego, action, gap, dt, steps, min_steps, seed, first, count, hist=False, threads=0):
        cs, as_ = scene.c_corridors(), scene.c_agents()
        cand = abi.Candidate(ego, abi.ACTIONS.get(action, abi.ACT_INVALID), gap, 0)
        sp = abi.SimParam(dt, steps, min_steps)
        st = (abi.Status * count)()
        n = len(as_)
        states = nst = None
        if hist:
            states = (C.c_double * (count * n * (steps + 1) * 4))()
            nst = (C.c_int32 * count)()
        if threads > 0:
            rc = self.lib.orc_rollouts_mt(cs, len(cs), as_, n, C.byref(cand), C.byref(sp), seed, first, count, threads, st)
        else:
            rc = self.lib.orc_rollouts(cs, len(cs), as_, n, C.byref(cand), C.byref(sp), seed, first, count, st, states, nst)
        if rc != 0:
            raise RuntimeError("oracle error %d" % rc)
        out = []
        for i in range(count):
            s = st[i]
            d = dict(ok=s.ok, finish=s.finish, fallBack=s.fallBack, finishStep=s.finishStep, utility=s.utility,
                     comfort=s.comfort, utilityObjects=s.utilityObjects, comfortObjects=s.comfortObjects,
                     draws=s.draws, executedSteps=s.executedSteps)
            if hist:
                h = {}
                for ai, a in enumerate(scene.agents):
                    base = (i * n + ai) * (steps + 1) * 4
                    h[a["id"]] = [tuple(states[base + 4 * t: base + 4 * t + 4]) for t in range(nst[i])]
                d["hist"] = h
            out.append(d)
        return out, st

    def reduce(self, st, groups, m, steps):
        f = abi.Feature()
        self.lib.orc_reduce_features(st, groups, m, steps, C.byref(f))
        return f

    def orders(self, scene):
        cs, as_ = scene.c_corridors(), scene.c_agents()
        ao = (C.c_int32 * len(as_))(); co = (C.c_int32 * len(cs))()
        li = (C.c_int32 * len(cs))(); ri = (C.c_int32 * len(cs))()
        rc = self.lib.orc_scene_orders(cs, len(cs), as_, len(as_), ao, co, li, ri)
        if rc != 0:
            raise RuntimeError("oracle error %d" % rc)
        return ([scene.agents[i]["id"] for i in ao], [scene.corridors[i]["id"] for i in co], list(li), list(ri))
