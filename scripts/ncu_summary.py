"""Selected raw metrics of the (single) kernel in an ncu report -> JSON on stdout (diagnostic; fills profiles/*_ncu_full.json).
usage: ncu_summary.py report.ncu-rep [launch index, default 0]"""
import csv, io, json, re, subprocess, sys
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
                  r"launch__(block_size|grid_size|occupancy_limit_(registers|shared_mem|warps)|registers_per_thread|shared_mem_per_block)|"
                  r"sm__inst_executed\.avg\.per_cycle_elapsed|sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warp_latency_per_inst_issued\.ratio|smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
                  r"smsp__cycles_active\.avg|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"smsp__sass_thread_inst_executed_op_(dadd|dmul|dfma)_pred_on\.sum|sm__sass_inst_executed_op_(shared|local|global)_(ld|st)\.sum)$")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
names, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
res = {"kernel": vals[names.index("Kernel Name")]}
for n, u, v in zip(names, units, vals):
    if KEEP.match(n):
        res[n] = {"unit": u, "value": v}
json.dump(res, sys.stdout, indent=1)
print()
