#!/bin/bash
# One gpurun call that produces the round's evidence (run ON the GPU box, from the repo root):
#   tools/evidence_run.sh TAG      -> gpurun_out/TAG_*
# bench lines (C2 default with C5 / C4 sub-runs; C1; C3), the ncu launch list of the bench command, and one
# `ncu --set full` capture of every kernel of a forward + backward step (read here with tools/ncu_summary.py,
# tools/ncu_traffic.py, tools/ncu_lines.py). Numbers printed under ncu are never bench values.
TAG=${1:-r2}
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_C2_N1.json 2> $O/${TAG}_bench.err || exit 1
python bench.py --config C1 --steps 20 --warmup 5 --no-extra > $O/${TAG}_bench_C1_N1.json 2>> $O/${TAG}_bench.err
python bench.py --config C3 --steps 10 --warmup 3 --no-extra > $O/${TAG}_bench_C3_N1.json 2>> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_C2.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-graph --no-half-maps > $O/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"unproject_|camera_prep|project_|tile_scan|bin_kernel|tile_render|composite_" --launch-skip 55 --launch-count 11 -o $O/${TAG}_full -f \
    python tools/stage_bench.py --bwd --steps 2 > $O/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tile_render" --launch-skip 5 --launch-count 1 -o $O/${TAG}_fwd -f \
    python tools/stage_bench.py --steps 2 > $O/${TAG}_ncu_fwd.log 2>&1
tail -n 2 $O/${TAG}_ncu_full.log $O/${TAG}_ncu_fwd.log
