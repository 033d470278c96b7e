#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel, stall reasons, and a SASS region breakdown (contiguous
instructions with the same executed count) -- the view used for the notes under profiles/."""
import csv
import subprocess
import sys
from collections import Counter

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("=" * 100)
        print(d.get("Kernel Name", "?")[:110])
        for w in WANT:
            if w in d:
                print(f"  {w:72s} {d[w]}")
        st = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))


def regions(rep, kernel_regex, min_pct=0.8):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    ia, ie, isamp, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    data = [r for r in data if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    tots = sum(int(r[isamp]) for r in data) or 1
    print(f"-- {kernel_regex}: {tot} warp instructions, {tots} samples")
    groups, cur = [], None
    for i, r in enumerate(data):
        e, s_ = int(r[ie]), int(r[isamp])
        toks = r[ia].split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        if cur and abs(e - cur["e"]) <= 0.03 * max(e, cur["e"], 1):
            cur["n"] += 1; cur["sum"] += e; cur["s"] += s_; cur["ops"].append(op); cur["end"] = i
            cur["thr"] += float(r[ith] or 0) * e
        else:
            cur = dict(start=i, end=i, e=e, n=1, sum=e, s=s_, ops=[op], thr=float(r[ith] or 0) * e)
            groups.append(cur)
    for g in groups:
        if 100 * g["sum"] / tot >= min_pct or 100 * g["s"] / tots >= min_pct:
            c = Counter(o.split(".")[0] for o in g["ops"]).most_common(7)
            thr = g["thr"] / max(g["sum"], 1)
            print(f"  [{g['start']:4d}-{g['end']:4d}] n={g['n']:3d} exec={g['e']:>10d} instr%={100 * g['sum'] / tot:5.1f} "
                  f"samp%={100 * g['s'] / tots:5.1f} thr={thr:4.1f} {c}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw(rep)
    for k in sys.argv[2:]:
        regions(rep, k)
