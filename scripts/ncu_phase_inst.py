"""Warp / thread instructions of an ncu report by source-line range (phase) of rollout_kernel.cuh, per warp and step.
usage: ncu_phase_inst.py src.csv (from `ncu -i rep --page source --print-source cuda,sass --csv`) warps_steps"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
norm = float(sys.argv[2]) if len(sys.argv) > 2 else 8192 * 100 * 8
cur = None; hdr = None
per = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit() and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        g = lambda k: int(d.get(k, "0")) if d.get(k, "0").isdigit() else 0
        p = per[(cur, int(r[0]))]
        p[0] += g("Instructions Executed"); p[1] += g("Thread Instructions Executed"); p[2] += g("# Samples")
phases = [  # (name, file, lo, hi)
    ("mcs_div", "rollout_kernel.cuh", 76, 87), ("idm_acc", "rollout_kernel.cuh", 156, 210), ("search", "rollout_kernel.cuh", 212, 241),
    ("rss_safe", "rollout_kernel.cuh", 243, 262), ("lc_prob", "rollout_kernel.cuh", 264, 274), ("centering", "rollout_kernel.cuh", 276, 283),
    ("rss_emerg", "rollout_kernel.cuh", 285, 294), ("side_lf", "rollout_kernel.cuh", 300, 309), ("mobil_geo", "rollout_kernel.cuh", 321, 414),
    ("mobil_decide", "rollout_kernel.cuh", 438, 501), ("sync/scan", "rollout_kernel.cuh", 537, 586), ("init", "rollout_kernel.cuh", 615, 712),
    ("loophead", "rollout_kernel.cuh", 713, 728), ("partA", "rollout_kernel.cuh", 729, 776), ("scan1+front", "rollout_kernel.cuh", 777, 797),
    ("tasks", "rollout_kernel.cuh", 798, 814), ("drawcount", "rollout_kernel.cuh", 815, 829), ("partB", "rollout_kernel.cuh", 830, 921),
    ("integrate", "rollout_kernel.cuh", 922, 940), ("match", "rollout_kernel.cuh", 941, 964), ("rebuild", "rollout_kernel.cuh", 965, 1047),
    ("termination", "rollout_kernel.cuh", 1048, 1076), ("stats", "rollout_kernel.cuh", 1077, 1172)]
agg = collections.OrderedDict((p[0], [0, 0, 0]) for p in phases)
other = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), v in per.items():
    hit = False
    for name, pf, lo, hi in phases:
        if f == pf and lo <= ln <= hi:
            a = agg[name]; hit = True; break
    if not hit: a = other[f]
    for i in range(3): a[i] += v[i]
ti = sum(v[0] for v in per.values()); ts = sum(v[2] for v in per.values())
print("total warp inst per warp-step %.0f" % (ti / norm))
print("%-14s %9s %6s %6s %6s" % ("phase", "inst/w-st", "inst%", "thr", "smp%"))
for name, v in list(agg.items()) + list(other.items()):
    if v[0] == 0: continue
    print("%-14s %9.1f %6.1f %6.1f %6.1f" % (name, v[0] / norm, 100 * v[0] / ti, v[1] / v[0], 100 * v[2] / ts))
