"""Bucket an ncu report's samples / instructions by kernel phase (line ranges of rollout_kernel.cuh found by markers).
usage: ncu_buckets.py report.ncu-rep [rollout_kernel.cuh as it was when the report was captured]"""
import csv, subprocess, sys, io, re, os
rep = sys.argv[1]
src = open(sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "interactive_highway_simulation_b200/csrc/rollout_kernel.cuh")).read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
marks = [("helpers:idm_acc", find("double idm_acc(")), ("helpers:search", find("int lane_upper(")), ("helpers:rss/prob/center", find("bool rss_safe(")),
         ("mobil_geometry", find("void mobil_geometry(")), ("mobil_decide", find("void mobil_decide(")), ("customized_lane_change", find("void customized_lane_change(")),
         ("scan", find("int block_exclusive_scan")), ("kernel:init", find("rollout_kernel(const SceneDev")), ("partA", find("plan, part A")),
         ("tasks", find("compaction: queue MOBIL tasks")), ("draws", find("// draw counts")), ("partB", find("plan, part B")),
         ("integrate", find("================= integrate")), ("match", find("================= matchAgentsOnCorridor")),
         ("termination", find("================= termination")), ("stats", find("================= outcome statistics")), ("end", len(src) + 1)]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit() and len(r) > 8 and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        g = lambda k: int(d.get(k, "0")) if d.get(k, "0").isdigit() else 0
        line = int(r[0])
        if cur == "rollout_kernel.cuh":
            b = [m[0] for m in marks if m[1] <= line][-1] if line >= marks[0][1] else "smem/carve"
        else:
            b = cur
        a = agg.setdefault(b, [0, 0, 0, 0])
        a[0] += g("# Samples"); a[1] += g("Instructions Executed"); a[2] += g("stall_barrier"); a[3] += g("stall_no_inst")
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("%-28s %8s %8s %8s %8s" % ("bucket", "samples%", "inst%", "barrier%", "noinst%"))
for b, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %8.1f %8.1f %8.1f %8.1f" % (b, 100 * a[0] / ts, 100 * a[1] / ti, 100 * a[2] / ts, 100 * a[3] / ts))
