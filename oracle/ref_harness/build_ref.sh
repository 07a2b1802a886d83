#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/: stub libraries + the harness that
# executes the reference's precompiled inner simulator (the reference ships no C++
# source for it; the binary is the only authoritative statement of the algorithm).
# Needs /root/reference (present in the build container only); outputs go to
# oracle/_ref/ which is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../_ref"
REF_SO="${REF_SO:-/root/reference/mc_sim_highway/mc_sim_highway.so}"
if [ ! -f "$REF_SO" ]; then
  echo "build_ref: $REF_SO not present; keeping prebuilt oracle/_ref if any" >&2
  exit 0
fi
mkdir -p "$OUT"
cp -f "$REF_SO" "$OUT/mc_sim_highway.so"
chmod u+w "$OUT/mc_sim_highway.so"
python3 "$HERE/gen_stubs.py" "$REF_SO" "$OUT/stubs.s"
# all stubs live in the python stub library; the other two only need to exist under the right soname
gcc -shared -fPIC -o "$OUT/libpython3.6m.so.1.0" "$OUT/stubs.s" -Wl,-soname,libpython3.6m.so.1.0
echo "" | gcc -shared -fPIC -x c - -o "$OUT/libmc_sim_highway.so.0" -Wl,-soname,libmc_sim_highway.so.0
g++ -O1 -shared -fPIC -o "$OUT/libboost_python3-py36.so.1.65.1" "$HERE/boost_holder.cpp" \
    -Wl,-soname,libboost_python3-py36.so.1.65.1
g++ -std=c++17 -O2 -g -rdynamic -o "$OUT/ref_harness" "$HERE/harness.cpp" \
    -ldl -lpthread
rm -f "$OUT/stubs.s"
echo "build_ref: built $OUT/ref_harness"
