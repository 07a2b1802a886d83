"""Per-source-line warp instructions per warp-step. usage: ncu_line_inst.py src.csv file lo hi [norm]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
fn, lo, hi = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
norm = float(sys.argv[5]) if len(sys.argv) > 5 else 8192 * 100 * 8
cur = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit() and hdr and cur == fn and lo <= int(r[0]) <= hi:
        d = dict(zip(hdr[4:], r[4:]))
        g = lambda k: int(d.get(k, "0")) if d.get(k, "0").isdigit() else 0
        i = g("Instructions Executed")
        if i: print("%5d %7.1f thr=%5.1f smp=%6d bar=%6d | %s" % (int(r[0]), i / norm, g("Thread Instructions Executed") / i, g("# Samples"), g("stall_barrier"), r[1].strip()[:110]))
