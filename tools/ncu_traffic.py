#!/usr/bin/env python
"""Regenerate profiles/traffic.json from `ncu --set full` captures, stamped with the kernel sources they describe.

    python tools/ncu_traffic.py gpurun_out/a.ncu-rep[::suffix] [gpurun_out/b.ncu-rep ...]

`::suffix` is appended to the kernel names of that report (e.g. the training-mode capture of a kernel that also has an
inference-mode entry).

For every kernel name in the reports (last launch wins): dram__bytes_read.sum + dram__bytes_write.sum and
smsp__inst_executed.sum per launch. `src_sha` = prometheus_b200.kernel_source_sha() of the tree the capture was made
from (the .ncu-rep comes back from the same snapshot gpurun ran, so the working tree must be unchanged since).
bench.py copies the dominant kernel's entry into roofline.traffic only when src_sha matches the sources it runs."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def num(s):
    return float(s.replace(",", "") or 0)


def short(name):
    """pm::tile_render_fwd_kernel(const unsigned long *, ...) -> tile_render_fwd_kernel; keeps template arguments"""
    name = name.split("(")[0]
    name = re.sub(r"^void\s+", "", name)
    return name.replace("pm::", "").strip()


def main():
    from prometheus_b200._lib import kernel_source_sha

    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch from `ncu --set full` "
                       "(tools/ncu_traffic.py); bench.py uses an entry only while src_sha matches the kernel sources it runs",
           "src_sha": kernel_source_sha(), "reports": [os.path.basename(p.split("::")[0]) for p in sys.argv[1:]],
           "_warp_instructions": {}}
    for arg in sys.argv[1:]:
        rep, _, suffix = arg.partition("::")
        # base units: the raw page otherwise scales every column to its own unit (Mbyte here, Gbyte there)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, data = rows[0], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in data:
            k = short(r[idx["Kernel Name"]]) + ("[%s]" % suffix if suffix else "")
            out[k] = int(num(r[idx["dram__bytes_read.sum"]]) + num(r[idx["dram__bytes_write.sum"]]))
            if "smsp__inst_executed.sum" in idx:
                out["_warp_instructions"][k] = int(num(r[idx["smsp__inst_executed.sum"]]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
