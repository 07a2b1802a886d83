"""Kernel time of the small-scene decisions on bundle_kernel at every width (rollouts per warp) and on rollout_kernel
(diagnostic).  usage: bundle_sweep.py [lib]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 1: backend.LIB_PATH = os.path.abspath(sys.argv[1])
acts = ["keep_lane", "dcc", "acc", "left", "right"]


def time_it(sc, cands, sp, R, seed):
    best = 1e9
    for i in range(4):
        ms, vs = sc.rollouts_device(cands, sp, seed, 0, R); best = min(best, ms)
    return best, vs


def sweep(name, sc, cands, sp, R, seed):
    out = []
    os.environ["MCS_KERNEL"] = "rollout"
    ms, vs = time_it(sc, cands, sp, R, seed)
    out.append("rollout_kernel %.3f" % ms)
    os.environ["MCS_KERNEL"] = "bundle"
    for w in (None, 0, 1, 2, 3, 4, 5):
        if w is None: os.environ.pop("MCS_BUNDLE_RW_LOG2", None)
        else: os.environ["MCS_BUNDLE_RW_LOG2"] = str(w)
        ms, vs2 = time_it(sc, cands, sp, R, seed)
        assert vs2 == vs
        out.append("%s %.3f" % ("auto" if w is None else "rw%d" % (1 << w), ms))
    os.environ.pop("MCS_BUNDLE_RW_LOG2", None)
    print("%-70s %s" % (name, "  ".join(out)))


c, a, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(c, a)
cg = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
for R in (64, 128, 512, 2048):
    sweep("configs[1] merging 40 veh, 4 gaps x %d, 40 steps" % R, sc, cg, abi.SimParam(0.3, 40, 10), R, 1)
sc.close()
c, a, ego, gaps = synthetic.merging_scene(20, 1, n_main=2)
sc = backend.DeviceScene(c, a)
cg = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
sweep("reference size merging 20 veh, %d gaps x 240, 40 steps" % len(gaps), sc, cg, abi.SimParam(0.3, 40, 10), 240, 1)
sc.close()
c, a, ego = synthetic.lane_change_scene(24, 3, seed=3)
sc = backend.DeviceScene(c, a)
cl = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[x], -2, 0) for x in acts])
for R in (160, 2048):
    sweep("reference size lane change 24 veh, 5 actions x %d, 40 steps" % R, sc, cl, abi.SimParam(0.3, 40, 30), R, 2)
sc.close()
c, a, ego = synthetic.lane_change_scene(64, 4, seed=3)
sc = backend.DeviceScene(c, a)
cl = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[x], -2, 0) for x in acts])
sweep("lane change 64 veh, 5 actions x 512, 40 steps", sc, cl, abi.SimParam(0.3, 40, 30), 512, 2)
sc.close()
