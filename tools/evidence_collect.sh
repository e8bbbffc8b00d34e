#!/bin/bash
# Turn the gpurun_out/TAG_* files of tools/evidence_run.sh into the committed summaries under profiles/ (run HERE,
# on the unchanged working tree the capture was made from — traffic.json is stamped with the kernel sources' hash):
#   tools/evidence_collect.sh TAG
set -e
TAG=${1:-r2}
cd "$(dirname "$0")/.."
O=gpurun_out
P=profiles
for f in bench_C2_N1 bench_C1_N1 bench_C3_N1 bench_reference; do
  [ -s $O/${TAG}_$f.json ] && cp $O/${TAG}_$f.json $P/${TAG}_$f.json
done
cp $O/${TAG}_launches_C2.csv $P/${TAG}_launches_C2.csv
python - "$P/${TAG}_launches_C2.csv" > $P/${TAG}_launches_C2_summary.txt <<'EOF'
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]; ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", r[ik])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("ncu launch list (gpu__time_duration.sum, --clock-control none): launches, total us, share of all launches, avg us")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-60s %5d %12.1f %6.1f%% %10.2f" % (k[:60], a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / a[0] / 1e3))
EOF
python tools/ncu_summary.py $O/${TAG}_full.ncu-rep > $P/${TAG}_kernels_ncu_summary.txt
python tools/ncu_summary.py $O/${TAG}_fwd.ncu-rep > $P/${TAG}_tile_render_inference_ncu_summary.txt
python tools/ncu_lines.py $O/${TAG}_fwd.ncu-rep tile_render 45 > $P/${TAG}_tile_render_lines.txt
python tools/ncu_regions.py $O/${TAG}_fwd.ncu-rep tile_render > $P/${TAG}_tile_render_regions.txt
python tools/ncu_lines.py $O/${TAG}_full.ncu-rep composite_bwd 45 > $P/${TAG}_composite_bwd_lines.txt
python tools/ncu_regions.py $O/${TAG}_full.ncu-rep composite_bwd > $P/${TAG}_composite_bwd_regions.txt
python tools/ncu_traffic.py $O/${TAG}_fwd.ncu-rep $O/${TAG}_full.ncu-rep::bwd > /dev/null
# SASS evidence of the Blackwell-specific instructions (TMA bulk copies, mbarrier, two-wide fp32)
{
  echo "cuobjdump -sass prometheus_b200/libprometheus_gs.so: instruction counts per kernel"
  cuobjdump -sass prometheus_b200/libprometheus_gs.so | awk '
    /Function :/ { fn = $3 }
    /UBLKCP/ { c[fn, "UBLKCP"]++ } /SYNCS/ { c[fn, "SYNCS"]++ } /FFMA2/ { c[fn, "FFMA2"]++ } /FMUL2/ { c[fn, "FMUL2"]++ }
    /FADD2/ { c[fn, "FADD2"]++ } /UTMA/ { c[fn, "UTMA"]++ } /MUFU.EX2/ { c[fn, "MUFU.EX2"]++ }
    END { for (k in c) { split(k, a, SUBSEP); printf "%-90s %-10s %d\n", substr(a[1], 1, 90), a[2], c[k] } }' | sort
} > $P/${TAG}_sass_counts.txt
nvcc --version | tail -1 >> $P/${TAG}_sass_counts.txt
echo "collected into $P/${TAG}_*"
