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
import os, re
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "interactive_highway_simulation_b200", "csrc", "rollout_kernel.cuh")
marks = [  # (phase, regex of the line that starts it) in file order
    ("mcs_div", r"double mcs_div\("), ("smem", r"void carve_smem\("), ("idm_acc", r"double idm_acc\("), ("search", r"int lane_upper\("),
    ("rss_safe", r"bool rss_safe\("), ("lc_prob", r"void lane_change_probability\("), ("centering", r"double centering_vy\("),
    ("rss_emerg", r"bool rss_emergency\("), ("side_lf", r"void side_leader_follower\("), ("mobil_geo", r"void mobil_sub_task\("),
    ("mobil_decide", r"void mobil_decide\("), ("cust_lc", r"void customized_lane_change\("), ("sync/scan", r"void sub_sync\("),
    ("init", r"void __launch_bounds__"), ("loophead", r"^  while \(true\) \{"), ("helper", r"if \(kHelper && helper\) \{"),
    ("partA", r"=+ plan, part A"), ("scan1+front", r"// compaction: queue MOBIL tasks"), ("tasks", r"// The seven sub-tasks of a queued vehicle"),
    ("drawcount", r"// draw counts:"), ("partB", r"=+ plan, part B"), ("integrate", r"=+ integrate"), ("match", r"=+ matchAgentsOnCorridor"),
    ("rebuild", r"^    if \(anyChanged\) \{"), ("termination", r"=+ termination"), ("stats", r"=+ outcome statistics")]
lines = open(SRC).read().split("\n")
phases, at = [], 0
for name, rx in marks:
    for i in range(at, len(lines)):
        if re.search(rx, lines[i]):
            phases.append([name, "rollout_kernel.cuh", i + 1, len(lines)]); at = i + 1; break
for a, b in zip(phases, phases[1:]): a[3] = b[2] - 1
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
