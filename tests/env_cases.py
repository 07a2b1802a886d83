"""Shared by the environment-geometry tests: the golden states of tests/golden/env_states.json.gz as oracle inputs."""
import gzip
import json
import os

from oracle import env_oracle as eo

TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))
with gzip.open(os.path.join(HERE, "golden", "env_states.json.gz"), "rt") as f:
    STATES = json.load(f)


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
