"""Kernel time of gap decisions on exit / merging scenes of several sizes (which block shape serves 33-64 / 65-128 / 129-256
vehicle merging rollouts best).  usage: merge_width_times.py [alternative libmcsim_b200.so]   (diagnostic; needs a GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 1: backend.LIB_PATH = os.path.abspath(sys.argv[1])
for kind, n, R in (("exit", 60, 4096), ("exit", 100, 4096), ("exit", 100, 512), ("exit", 120, 2048), ("merging", 100, 2048), ("exit", 160, 2048)):
    if kind == "exit":
        c, a, ego, gaps = synthetic.exit_scene(n, 0); act = "right"
    else:
        c, a, ego, gaps = synthetic.merging_scene(n, 0); act = "left"
    sc = backend.DeviceScene(c, a)
    cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS[act], g, 0) for g in gaps])
    best = 1e9
    for i in range(3):
        ms, vs = sc.rollouts_device(cands, abi.SimParam(0.3, 50, 10), 3, 0, R); best = min(best, ms)
    print("%-8s %3d vehicles, %d gaps x %4d rollouts x 50 steps  %8.3f ms  %.3e vehicle-steps/s" % (kind, n, len(gaps), R, best, vs / best * 1e3))
    sc.close()
