#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: headline metrics, stall mix, opcode mix,
top stalled SASS lines.  Usage: tools/ncu_summary.py report.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys


SKIP = 0


def page(rep, name):
    """one kernel of the report (the SKIP-th captured launch)"""
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", "--launch-skip", str(SKIP), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    global SKIP
    rep = sys.argv[1]
    SKIP = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
            "sm__cycles_elapsed.avg.per_second", "smsp__warps_eligible.avg.per_cycle_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
    print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, vals):
        if h in want:
            print(f"{h:75s} {v} {u}")
    src = page(rep, "source")
    sh = src[1]
    ix = {h: i for i, h in enumerate(sh)}
    stalls = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    opc = collections.Counter()
    ops = collections.Counter()
    ts = ti = 0
    recs = []
    for n, r in enumerate(src[2:]):
        try:
            s = int(r[ix["# Samples"]] or 0)
            ie = int(r[ix["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        ts += s
        ti += ie
        for st in stalls:
            try:
                tot[st] += int(r[ix[st]] or 0)
            except ValueError:
                pass
        toks = [x for x in r[ix["Source"]].split() if not x.startswith("@")]
        name = toks[0].split(".")[0] if toks else "?"
        opc[name] += ie
        ops[name] += s
        recs.append((s, n, ie, r))
    print(f"\n== stall mix ({ts} samples, {ti} warp-instructions)")
    for st, v in tot.most_common(9):
        print(f"{st:28s} {100.0 * v / max(ts, 1):5.1f}%")
    print("\n== opcode mix (share of executed warp-instructions | share of samples)")
    for k, v in opc.most_common(18):
        print(f"{k:10s} {100.0 * v / max(ti, 1):5.1f}% | {100.0 * ops[k] / max(ts, 1):5.1f}%")
    print("\n== top stalled SASS lines")
    for s, n, ie, r in sorted(sorted(recs, key=lambda x: -x[0])[:25], key=lambda x: x[1]):
        st = {h: int(r[ix[h]] or 0) for h in stalls}
        big = max(st.items(), key=lambda x: x[1])
        print(f"{n:5d} samples={s:6d} exec={ie:9d} {r[ix['Source']][:70]:70s} {big[0]}={big[1]}")


if __name__ == "__main__":
    main()
