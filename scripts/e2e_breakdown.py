"""Where does the end-to-end call spend its time? (diagnostic, not a bench)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
corridors, agents, ego = synthetic.lane_change_scene(256, 4, seed=0)
cand = (abi.Candidate * 1)(abi.Candidate(ego, abi.ACTIONS["keep_lane"], -2, 0))
sp = abi.SimParam(0.3, 100, 100)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for it in range(4):
    t0 = time.perf_counter(); sc = backend.DeviceScene(corridors, agents); t1 = time.perf_counter()
    feats, _ = sc.rollouts(cand, sp, R // 8, 5 + it); t2 = time.perf_counter()
    feats, _ = sc.rollouts(cand, sp, R // 8, 9 + it); t3 = time.perf_counter()
    ms, vs = sc.rollouts_device(cand, sp, 5 + it, 0, R); t4 = time.perf_counter()
    sc.close(); t5 = time.perf_counter()
    print("create %.2f ms  rollouts#1 %.2f ms  rollouts#2 %.2f ms  device %.2f ms (kernel %.2f)  close %.2f ms" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), ms, 1e3 * (t5 - t4)))
