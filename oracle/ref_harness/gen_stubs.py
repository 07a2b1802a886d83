#!/usr/bin/env python3
"""Generate assembler stubs for every symbol the reference binary imports from
libraries that do not exist in this image (CPython 3.6, Boost.Python 1.65.1 and
libmc_sim_highway.so.0).  TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Every undefined FUNC becomes a function returning a pointer to 4 KB of zeroed
memory, every undefined OBJECT becomes 1 KB of zeroed, 64-byte aligned data.
The five boost::python::instance_holder members are NOT stubbed: they get a real
implementation in boost_holder.cpp (the value_holder<T> the binding code
constructs derives from it).

usage: gen_stubs.py <reference .so> <out.s>
"""
import subprocess, sys

REAL = ("instance_holder",)          # implemented in boost_holder.cpp

def main(so, out):
    txt = subprocess.run(["readelf", "-W", "--dyn-syms", so], check=True,
                         capture_output=True, text=True).stdout
    funcs, objs = [], []
    for line in txt.splitlines():
        f = line.split()
        if len(f) < 8 or f[6] != "UND":
            continue
        name, kind, bind = f[7], f[3], f[4]
        if "@" in name or bind == "WEAK":     # versioned: real glibc/libstdc++
            continue
        if any(r in name for r in REAL):
            continue
        (funcs if kind == "FUNC" else objs).append(name)
    with open(out, "w") as fh:
        fh.write("\t.bss\n\t.balign 64\nstub_dummy:\n\t.zero 4096\n")
        for o in objs:
            fh.write(f"\t.globl {o}\n\t.type {o},@object\n\t.balign 64\n{o}:\n\t.zero 1024\n\t.size {o},1024\n")
        fh.write("\t.text\n")
        for fn in funcs:
            fh.write(f"\t.globl {fn}\n\t.type {fn},@function\n{fn}:\n"
                     f"\tlea stub_dummy(%rip),%rax\n\tret\n\t.size {fn},.-{fn}\n")
        fh.write('\t.section .note.GNU-stack,"",@progbits\n')
    print(f"stubs: {len(funcs)} functions, {len(objs)} objects -> {out}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
