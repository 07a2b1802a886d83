"""Kernel time of merging decisions with narrow (64-thread) rollouts at several launch sizes (block-shape tuning).  usage: merge_narrow_times.py [alternative .so]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interactive_highway_simulation_b200 import backend, synthetic, _abi as abi
if len(sys.argv) > 1: backend.LIB_PATH = os.path.abspath(sys.argv[1])
for n, R in ((40, 128), (40, 256), (40, 512), (40, 1024), (40, 4096), (60, 512), (60, 4096)):
    c, a, ego, gaps = synthetic.merging_scene(n, 0)
    sc = backend.DeviceScene(c, a)
    cands = (abi.Candidate * len(gaps))(*[abi.Candidate(ego, abi.ACTIONS["left"], g, 0) for g in gaps])
    best = 1e9
    for i in range(4):
        ms, vs = sc.rollouts_device(cands, abi.SimParam(0.3, 40, 10), 1, 0, R); best = min(best, ms)
    print("merging %3d vehicles, %d gaps x %4d rollouts x 40 steps  %8.3f ms" % (n, len(gaps), R, best))
    sc.close()
