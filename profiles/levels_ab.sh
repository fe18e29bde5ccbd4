#!/bin/bash
# A/B of the level kernels on config 2 (1000 x 5 Mb): tiled + arena-direct (default), tiled + gather, first-generation kernels
run() { python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('$1', 'step_ms %.3f' % b['ms_per_step'], {k: round(v,3) for k,v in b['stages_ms'].items()})"; }
run default
PGR_B200_ALWAYS_GATHER=1 run tiled_with_gather
PGR_B200_ALWAYS_GATHER=1 PGR_B200_LEVELS_UNTILED=1 run first_generation
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"level_|block_scan|block_chunk|gather_l0|l0_kernel" -c 40 --csv \
    --log-file gpurun_out/launches_levels.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
grep -i "kernel" gpurun_out/launches_levels.csv | awk -F'","' '{print $5, $NF}' | tail -24
