#!/bin/bash
# Per-kernel HBM evidence for every kernel family of the library: duration + DRAM bytes of each launch (ncu, no clock
# control), on reduced-size runs of config 2 (100 contigs) and configs 3/4/5 (scale 0.2).  Summarised by
# profiles/kernel_table.py into profiles/r1_kernel_table.txt.  Run under gpurun.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
K='regex:^(l0_|level_|block_|gather_|replay_|patch_|splice_|sketch_|pair_|tuple_|csr_|rs_|lookup_|qpair_|hit_|chain_|seg_|assemble_|adj_|frag_|weight_|sid_count|scan_|iota_|set_u64|add_frg|dest_keys|max_span)'
ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/kt_config2.csv \
    python bench.py --contigs 100 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
PGR_B200_BENCH_FRAGS=1 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/kt_configs345.csv \
    python bench_configs.py --configs 3,4,5 --scale 0.2 > gpurun_out/kt_configs345.log 2>&1
ls -la gpurun_out/kt_*.csv
