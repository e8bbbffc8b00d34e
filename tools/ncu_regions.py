#!/usr/bin/env python
"""Instruction / stall-sample share per source region of tile_render_fwd_kernel (developer tool).
Regions are delimited by marker comments `// @region name` in the .cu / .cuh sources."""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_regions(fn):
    path = os.path.join(ROOT, "prometheus_b200", "csrc", fn)
    if not os.path.exists(path):
        return []
    out, cur = [], None
    for i, line in enumerate(open(path), 1):
        m = re.search(r"// @region (\S+)", line)
        if m:
            cur = m.group(1)
        out.append(cur)
    return out


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                          "--print-source=sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, agg, regions = None, None, {}, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            if cur not in regions:
                regions[cur] = load_regions(cur)
            continue
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            hl = r
            continue
        if hdr is None or len(r) < 10 or r[0] == "":
            continue
        # the source-text column may itself contain commas / quotes (inline asm): index the numeric columns from the end
        off = len(r) - len(hl)
        num = lambda k: float((r[hdr[k] + off] or "0").replace(",", "")) if r[hdr[k] + off] not in ("", "-") else 0.0
        ln = int(r[0])
        reg = regions[cur][ln - 1] if regions[cur] and ln - 1 < len(regions[cur]) and regions[cur][ln - 1] else cur
        a = agg.setdefault(reg, [0.0, 0.0, {}])
        a[0] += num("Instructions Executed")
        a[1] += num("Warp Stall Sampling (All Samples)")
        for h in hl:
            if h.startswith("stall_") and "Not Issued" not in h:
                a[2][h[6:]] = a[2].get(h[6:], 0.0) + num(h)
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        top = sorted(a[2].items(), key=lambda x: -x[1])[:3]
        print("%-26s inst %5.1f%%  samples %5.1f%%   %s" % (k, 100 * a[0] / ti, 100 * a[1] / ts,
                                                            ", ".join("%s=%.0f%%" % (n, 100 * v / max(a[1], 1)) for n, v in top)))


if __name__ == "__main__":
    main()
