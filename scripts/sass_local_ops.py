"""Attribute local-memory instructions (LDL / STL: spills, stack arrays) of one kernel to source lines (diagnostic).
usage: sass_local_ops.py <nvdisasm -g -c output> <substring of the mangled kernel name> [top]"""
import re, sys, collections
path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cur = None; fn = None; cnt = collections.Counter(); n_inst = 0
for line in open(path):
    m = re.match(r'\.text\.(\S+):', line)
    if m: fn = m.group(1); continue
    if fn is None or want not in fn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.search(r'^\s+/\*[0-9a-f]+\*/', line):
        n_inst += 1
        m = re.search(r'\b(LDL|STL)\b', line)
        if m and cur: cnt[cur + (m.group(1),)] += 1
print(sum(cnt.values()), "local ops in", n_inst, "instructions")
for k, v in cnt.most_common(top): print(v, k)
