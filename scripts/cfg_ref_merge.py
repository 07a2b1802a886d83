"""Reference-size merging decision (20 vehicles, 4 gaps x 160 rollouts) at the forced thread counts (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
for nveh, R in ((20, 160), (20, 64), (40, 160), (60, 160)):
    corridors, agents, ego, gaps = synthetic.merging_scene(nveh, 0)
    sc = backend.DeviceScene(corridors, agents)
    cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
    sp = abi.SimParam(0.3, 40, 30)
    best = 1e9
    for i in range(5):
        ms, vs = sc.rollouts_device(cands, sp, 55, 0, R); best = min(best, ms)
    print(os.environ.get("MCS_ROLLOUT_THREADS"), "%d vehicles %d x %d: %.3f ms" % (nveh, len(gaps), R, best))
    sc.close()
