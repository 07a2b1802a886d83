"""Out-of-line calls (division / sqrt slow paths) a kernel actually executed, by source line (diagnostic).
usage: ncu_calls.py report.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
def page(src):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", src, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
addr2line = {}; cur_file = cur_line = None
for r in page("cuda,sass"):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] and r[0].isdigit(): cur_line = (cur_file, int(r[0]), r[1].strip()[:90]); continue
    if r[0] == "" and len(r) > 3 and r[2].startswith("0x"): addr2line[r[2]] = cur_line
rows = page("sass"); hdr = rows[1]; data = rows[2:]
ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isrc = hdr.index("Source")
total = sum(int(r[ie]) for r in data)
print("total warp instructions", total)
for i, r in enumerate(data):
    if "CALL" in r[isrc] and int(r[ie]) > 0:
        where = next((addr2line[data[j][0]] for j in range(i, max(0, i - 40), -1) if data[j][0] in addr2line), None)
        print("%12d calls (%.1f threads) -> %s  at %s" % (int(r[ie]), int(r[it]) / int(r[ie]), r[isrc].split()[-1], where))
# instructions that belong to no source line (the callees)
known = set(addr2line)
orphan = sum(int(r[ie]) for r in data if r[0] not in known)
print("instructions outside any source line (callee bodies): %d = %.1f%%" % (orphan, 100.0 * orphan / total))
