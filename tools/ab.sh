#!/bin/bash
# A/B timing of alternative builds on the GPU box: tools/ab.sh [--bwd] variant1 variant2 ...   ("default" = in-tree .so)
# Each variant is timed twice, interleaved, with tools/stage_bench.py; prints total ms/step and the two dominant kernels.
ARGS=""
if [ "$1" == "--bwd" ]; then ARGS="--bwd"; shift; fi
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" == "default" ]; then LIB=""; else LIB="gpurun_variants/$v.so"; fi
    PM_GS_LIB=$LIB python tools/stage_bench.py --steps 20 $ARGS | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages_ms']
print('%-14s total %.4f  tile_render %.4f  project %.4f  scatter %.4f  %s' % ('$v', d['ms_per_step'], s.get('tile_render_fwd',0), s.get('project_fwd',0) + s.get('project_scatter_fwd',0), s.get('bin_scatter',0), ('bwd %.4f pbwd %.4f' % (s.get('composite_bwd',0), s.get('project_bwd',0))) if d['bwd'] else ''))"
  done
done
