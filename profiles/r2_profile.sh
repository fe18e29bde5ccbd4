#!/bin/bash
# Round-2 profiling pass (run under gpurun): one full ncu capture of the dominant kernel, the launch list of a small bench step
# (e2e and assembly-like legs included: unpack_kernel, cluster kernels), and per-launch duration + DRAM bytes for configs 3/4/5.
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --set full --clock-control none --import-source on -k regex:l0_kernel -c 1 -f -o gpurun_out/r2b_l0_kernel \
    python bench.py --contigs 100 --steps 1 --warmup 1 --no-e2e --no-cpu --no-index --no-assembly > /dev/null 2> gpurun_out/r2b_ncu_l0.err
ncu --metrics $M --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_bench.csv \
    python bench.py --contigs 100 --steps 2 --warmup 1 --no-cpu --no-index > /dev/null 2> gpurun_out/r2b_ncu_launches.err
K='regex:^(l0_|level_|block_|gather_|replay_|patch_|splice_|sketch_|pair_|tuple_|csr_|rs_|os_|lookup_|qpair_|hit_|chain_|seg_|assemble_|adj_|smp_|frag_|weight_|sid_count|scan_|iota_|set_u64|add_frg|dest_keys|max_span|cluster_|unpack_|part_|sample_|splitters_|qsort_)'
PGR_B200_BENCH_FRAGS=1 ncu --metrics $M --clock-control none -k "$K" --csv --log-file gpurun_out/r2b_kt_configs345.csv \
    python bench_configs.py --configs 3,4,5 --scale 0.2 > gpurun_out/r2b_kt_configs345.log 2>&1
ls -la gpurun_out/r2b_*
