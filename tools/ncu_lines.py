#!/usr/bin/env python
"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py <rep> <kernel-regex> [top_n]
Prints, per file:line, executed warp instructions, stall samples and the top stall reasons."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--print-source=sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, agg = None, None, []
    seen_kernel = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            if seen_kernel is None:
                seen_kernel = r[1]
            elif r[1] != seen_kernel and False:
                break
            continue
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            hdr_list = r
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":  # a source line summary row
            def num(k):
                try:
                    return float(r[hdr[k]].replace(",", ""))
                except Exception:
                    return 0.0
            stalls = sorted(((num(h), h) for h in hdr_list if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:3]
            agg.append((num("Instructions Executed"), num("Warp Stall Sampling (All Samples)"), cur_file, r[0], r[1].strip()[:90],
                        ", ".join("%s=%d" % (h[6:], v) for v, h in stalls if v > 0)))
    tot_i = sum(a[0] for a in agg) or 1
    tot_s = sum(a[1] for a in agg) or 1
    print("total warp instructions %.0f, stall samples %.0f" % (tot_i, tot_s))
    print("---- by stall samples")
    for a in sorted(agg, key=lambda a: -a[1])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%s  %s   [%s]" % (100 * a[1] / tot_s, 100 * a[0] / tot_i, a[2], a[3], a[4], a[5]))
    print("---- by executed instructions")
    for a in sorted(agg, key=lambda a: -a[0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%s  %s   [%s]" % (100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2], a[3], a[4], a[5]))


if __name__ == "__main__":
    main()
