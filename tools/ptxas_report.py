#!/usr/bin/env python
"""Registers / spills of every kernel of one translation unit (nvcc -Xptxas -v), demangled and shortened.

    python tools/ptxas_report.py ek_ops_fused_tqp.cu [--exact] [-DEK_...=..] [--grep PATTERN]
"""
import os
import re
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "earthkit-meteo_b200", "csrc")


def main():
    args = sys.argv[1:]
    src = args.pop(0)
    lean = "0" if "--exact" in args else "1"
    pat = None
    if "--grep" in args:
        pat = re.compile(args[args.index("--grep") + 1])
    extra = [a for a in args if a.startswith("-D")]
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC",
           f"-DEK_LEAN_MATH={lean}", "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", "/dev/null"] + extra
    out = subprocess.run(cmd, capture_output=True, text=True).stderr
    out = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Compiling entry function '(.*)' for", ln)
        if m:
            name = m.group(1)
            name = re.sub(r"^void ", "", name)
            name = re.sub(r"\(ek::InArgs.*$|\((?:const )?(?:ek::|\(anonymous).*$", "", name)
            name = name.replace("ek::fastm::", "").replace("(unsigned int)", "").replace("(int)", "").replace("(bool)", "")
            name = re.sub(r", ek::exactm::[^>]*>(?:, \d+>)?", "", name)  # drop the exact twin
            cur = name
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m and cur:
            stack, ss, sl = m.groups()
            continue_line = (stack, ss, sl)
            spill = continue_line
            continue
        m = re.search(r"Used (\d+) registers", ln)
        if m and cur:
            line = f"{int(m.group(1)):4d} regs  stack {spill[0]:>4}  spill st/ld {spill[1]:>4}/{spill[2]:<4}  {cur}"
            if pat is None or pat.search(cur):
                print(line)
            cur = None


if __name__ == "__main__":
    main()
