#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs (registers / smem / spills per kernel) from seganygaussians_b200/lib/obj."""
import glob, os, re, subprocess, sys

def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n

def main(d):
    for f in sorted(glob.glob(os.path.join(d, "*.ptxas.log"))):
        txt = open(f).read()
        blocks = re.split(r"ptxas info\s+: Compiling entry function '", txt)[1:]
        for b in blocks:
            name = b.split("'")[0]
            regs = re.search(r"Used (\d+) registers", b)
            smem = re.search(r"(\d+) bytes smem", b)
            spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
            dn = demangle(name)
            dn = re.sub(r"\(.*", "", dn).replace("void sagars::", "")
            print(f"{os.path.basename(f)[:-10]:16s} {dn:60s} regs={regs.group(1) if regs else '?':>3s} "
                  f"smem={smem.group(1) if smem else '0':>6s} spill={spill.group(1) if spill else '?'}/{spill.group(2) if spill else '?'}")

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "seganygaussians_b200", "lib", "obj"))
