"""Aggregate an ncu report's per-instruction samples by CUDA source line.  usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0].isdigit() and len(r) > 8 and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        def g(k):
            v = d.get(k, "0"); return int(v) if v.isdigit() else 0
        agg.append(dict(file=cur, line=int(r[0]), src=r[1].strip()[:100], smp=g("# Samples"), inst=g("Instructions Executed"),
                        thr=d.get("Avg. Threads Executed"), bar=g("stall_barrier"), wait=g("stall_wait"), ssb=g("stall_short_sb"),
                        lsb=g("stall_long_sb"), noi=g("stall_no_inst"), br=g("stall_branch_resolving"), math=g("stall_math"), sel=g("stall_selected"), nsel=g("stall_not_selected")))
tot = sum(a["smp"] for a in agg); toti = sum(a["inst"] for a in agg)
print("total samples", tot, "total warp inst", toti)
for k in ("bar", "wait", "ssb", "lsb", "noi", "br", "math", "sel", "nsel"):
    print("  %-5s %5.1f%%" % (k, 100.0 * sum(a[k] for a in agg) / tot))
agg.sort(key=lambda a: -a["smp"])
for a in agg[:top]:
    print("%5.1f%% smp %5.1f%% inst thr=%-4s bar=%-6d wait=%-6d ssb=%-6d lsb=%-6d %s:%d %s" % (
        100 * a["smp"] / tot, 100 * a["inst"] / toti, a["thr"], a["bar"], a["wait"], a["ssb"], a["lsb"], a["file"], a["line"], a["src"]))
