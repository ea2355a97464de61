#!/usr/bin/env python
"""Per-source-line view of one `ncu --set full --import-source on` report: share of warp-stall
samples and of executed instructions for the hottest CUDA source lines, plus the dominant stall
reasons per line.  Usage: ncu_lines.py report.ncu-rep [min_share_percent]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    floor = float(sys.argv[2]) / 100.0 if len(sys.argv) > 2 else 0.01
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    width = len(hdr)
    n_col, i_col = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines = []
    for r in rows[hi + 1:]:
        if r and r[0].isdigit():  # a source line (fields counted from the end: the text may hold commas)
            lines.append((int(r[0]), r[1], int(r[n_col - width] or 0), int(r[i_col - width] or 0),
                          {h: int(r[i - width] or 0) for i, h in stalls}))
    ts = sum(l[2] for l in lines) or 1
    ti = sum(l[3] for l in lines) or 1
    print("%5s %8s %8s  %-46s %s" % ("line", "samples", "instr", "top stall reasons", "source"))
    for ln, text, smp, ins, st in sorted(lines, key=lambda l: -l[2]):
        if smp < floor * ts:
            continue
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        why = " ".join("%s=%d%%" % (k[6:], 100 * v / max(smp, 1)) for k, v in top)
        print("%5d %7.1f%% %7.1f%%  %-46s %s" % (ln, 100 * smp / ts, 100 * ins / ti, why, text.strip()[:90]))


if __name__ == "__main__":
    main()
