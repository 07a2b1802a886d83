"""throughput-config kernel time for several library builds (diagnostic). usage: quick_tp.py lib1.so lib2.so ..."""
import sys, os, subprocess
here = os.path.dirname(os.path.abspath(__file__))
for lib in sys.argv[1:]:
    out = subprocess.run([sys.executable, os.path.join(here, "kernel_time.py"), lib], capture_output=True, text=True).stdout
    print(lib, "|", " | ".join(l.strip() for l in out.splitlines() if "config" in l or "decision" in l))
