"""Shared by the environment-geometry tests: the golden states of tests/golden/env_states.json.gz as oracle inputs."""
import gzip
import json
import os

from oracle import env_oracle as eo

TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))
with gzip.open(os.path.join(HERE, "golden", "env_states.json.gz"), "rt") as f:
    STATES = json.load(f)
with gzip.open(os.path.join(HERE, "golden", "env_border.json.gz"), "rt") as f:
    BORDER = json.load(f)          # tests/golden/make_border_golden.py


def corridor_dicts(rec):
    """corridors of a record with neighbour ids turned into array positions (-1 = None)"""
    pos = {c["id"]: k for k, c in enumerate(rec["corridors"])}
    return [{"id": c["id"], "type": c["type"], "left": c["left"], "right": c["right"], "center": c["center"], "v_limit": c["v_limit"],
             "left_index": pos[c["left_id"]] if c["left_id"] is not None else -1,
             "right_index": pos[c["right_id"]] if c["right_id"] is not None else -1} for c in rec["corridors"]]


def prepared(rec):
    pm = eo.PreparedMap(corridor_dicts(rec))
    ids = [a["id"] for a in rec["agents"]]
    return pm, ids, [a["x"] for a in rec["agents"]], [a["y"] for a in rec["agents"]]


def slots(ids):
    return {a: i for i, a in enumerate(ids)}


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class DuckAgent:
    """A vehicle with the attributes of the reference's Agent that the environment and the builders read."""

    def __init__(self, r):
        self.id, self.x, self.y, self.v, self.vy, self.a, self.l, self.w = (r[k] for k in ("id", "x", "y", "v", "vy", "a", "l", "w"))
        self.behavior_model = r["behavior_model"]
        i, s, m = r["idm"], r["rss"], r["mobil"]
        self.idm_param = _NS(vd=i[0], acc_max=i[1], acc_min=i[2], min_dis=i[3], acc_com=i[4], thw=i[5])
        self.rss_param = _NS(r_t_other=s[0], r_t_ego=s[1], a_min=s[2], a_max=s[3], soft_acc=s[4], soft_dcc=s[5])
        self.mobil_param = _NS(politeness_factor=m[0], delta_a=m[1], bias_a=m[2])
        self.yielding_param = list(r["yielding"])
        self.corridor_history = []

    def set_behavior_model(self, b):
        self.behavior_model = b
