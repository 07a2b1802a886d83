"""Loading of the tests/golden fixtures (written by tests/golden/make_golden.py from the reference binary)."""
import glob
import gzip
import json
import os

from oracle.pyoracle import SceneSpec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_all(kind=None):
    """*.json (make_golden.py) and large_*.json.gz (make_large_golden.py: BASELINE-size scenes)."""
    out = []
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json"))) + sorted(glob.glob(os.path.join(GOLDEN_DIR, "large_*.json.gz")))
    for p in paths:
        with (gzip.open(p, "rt") if p.endswith(".gz") else open(p)) as f:
            c = json.load(f)
        if kind is None or c["kind"] == kind:
            out.append(c)
    return out


def scene_of(case):
    s = SceneSpec()
    s.corridors = [dict(c) for c in case["scene"]["corridors"]]
    s.agents = []
    for a in case["scene"]["agents"]:
        a = dict(a)
        for k in ("idm", "rss", "yp"):
            a[k] = tuple(a[k])
        s.agents.append(a)
    return s


def ids(cases):
    return [c["tag"] for c in cases]
