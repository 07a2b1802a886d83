"""One launch of the configs[3]-shaped decision (exit scene, 200 vehicles) for ncu (diagnostic).  usage: profile_exit.py [rollouts per gap]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
c, a, ego, gaps = synthetic.exit_scene(200, 0)
sc = backend.DeviceScene(c, a)
cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["right"], g, 0) for g in gaps])
for i in range(3):
    ms, vs = sc.rollouts_device(cands, abi.SimParam(0.3, 50, 10), 3, 0, R)
print("exit decision %d gaps x %d rollouts: %.3f ms, %d vehicle-steps" % (len(gaps), R, ms, vs))
sc.close()
