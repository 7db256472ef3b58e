#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line.

usage: ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python profiles/ncu_by_line.py src.csv [top_n]
Prints, per (file, line): executed warp instructions, stall samples and the share of each.
"""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file, inst, samp, text = None, collections.Counter(), collections.Counter(), {}
    hdr = None
    line_no = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0].strip():  # a CUDA source line row
            line_no = int(r[0])
            text[(cur_file, line_no)] = r[1].strip()
        if r[2].strip():  # a SASS row (has an address)
            try:
                inst[(cur_file, line_no)] += int(r[i_inst] or 0)
                samp[(cur_file, line_no)] += int(r[i_samp] or 0)
            except ValueError:
                pass
    ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s}  source")
    for key, n in inst.most_common(top):
        print(f"{key[0] + ':' + str(key[1]):28s} {100 * n / ti:6.2f} {100 * samp[key] / ts:6.2f}  {text.get(key, '')[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
