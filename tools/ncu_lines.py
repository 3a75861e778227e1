#!/usr/bin/env python
"""Per-CUDA-source-line cost of a kernel from an .ncu-rep (needs -lineinfo): executed warp
instructions and stall samples per line.  Usage: tools/ncu_lines.py report.ncu-rep [top-N]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, recs = "?", None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r and r[0].isdigit():
            try:
                recs.append((fname, int(r[0]), r[1].strip(), int(r[hdr["Instructions Executed"]]), int(r[hdr["# Samples"]])))
            except ValueError:
                pass
    ti = sum(x[3] for x in recs) or 1
    ts = sum(x[4] for x in recs) or 1
    print(f"total warp-instructions {ti}, samples {ts}")
    for f, ln, src, ie, s in sorted(recs, key=lambda x: -x[3])[:top]:
        print(f"{100.0 * ie / ti:5.1f}% inst {100.0 * s / ts:5.1f}% smpl  {f}:{ln:<4d} {src[:110]}")


if __name__ == "__main__":
    main()
