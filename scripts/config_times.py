"""Kernel time of one decision at each BASELINE.json configuration's size (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 1: backend.LIB_PATH = os.path.abspath(sys.argv[1])
def run(name, sc, cands, sp, R, seed):
    best = 1e9
    for i in range(3):
        ms, vs = sc.rollouts_device(cands, sp, seed, 0, R); best = min(best, ms)
    print("%-78s %8.3f ms  %.3e vehicle-steps/s" % (name, best, vs / best * 1e3))
acts = ["keep_lane", "dcc", "acc", "left", "right"]
c, a, ego, gaps = synthetic.merging_scene(40, 0)
sc = backend.DeviceScene(c, a)
run("configs[1] merging: 40 vehicles, 4 gaps x 512 rollouts, 40 steps", sc, (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps]), abi.SimParam(0.3, 40, 10), 512, 1)
sc.close()
c, a, ego = synthetic.lane_change_scene(100, 4, seed=1)
sc = backend.DeviceScene(c, a)
cands = (abi.Candidate * 5)(*[abi.Candidate(ego, abi.ACTIONS[x], -2, 0) for x in acts])
run("configs[2] lane change: 100 vehicles, 5 actions x 160 rollouts, 40 steps", sc, cands, abi.SimParam(0.3, 40, 30), 160, 2)
run("configs[2] lane change: 100 vehicles, 5 actions x 512 rollouts, 40 steps", sc, cands, abi.SimParam(0.3, 40, 30), 512, 2)
sc.close()
c, a, ego, gaps = synthetic.exit_scene(200, 0)
sc = backend.DeviceScene(c, a)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps])
run("configs[3] exit: 200 vehicles, %d gaps x 4096 rollouts, 50 steps" % len(gaps), sc, cands, abi.SimParam(0.3, 50, 10), 4096, 3)
run("configs[3] exit: 200 vehicles, %d gaps x 512 rollouts, 50 steps" % len(gaps), sc, cands, abi.SimParam(0.3, 50, 10), 512, 3)
sc.close()
c, a, ego = synthetic.lane_change_scene(256, 4, seed=0)
sc = backend.DeviceScene(c, a)
run("configs[4] sweep: 256 vehicles x 100 steps, 16384 rollouts", sc, (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0)), abi.SimParam(0.3, 100, 100), 16384, 5)
sc.close()
