"""Rank CUDA source lines of an `ncu --page source --print-source cuda,sass --csv` dump.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_lines.py src.csv [top]
"""
import csv
import os
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr, fname, out = None, "", []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or r[2] != "-":
            continue
        i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = []
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and r[j] not in ("", "0"):
                stall.append((float(r[j]), h[6:]))
        out.append((float(r[i_s] or 0), float(r[i_i] or 0), fname, r[0], r[1].strip(), sorted(stall, reverse=True)[:3]))
    ts = sum(o[0] for o in out) or 1
    ti = sum(o[1] for o in out) or 1
    print("total samples %.0f, warp instructions %.0f" % (ts, ti))
    for s, i, f, ln, src, st in sorted(out, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%-4s %-78s %s" % (100 * s / ts, 100 * i / ti, f[:14], ln, src[:78],
                                                        " ".join("%s=%.0f" % (k, v) for v, k in st)))


if __name__ == "__main__":
    main()
