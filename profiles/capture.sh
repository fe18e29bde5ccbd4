#!/bin/bash
# One GPU session: GPU parity tests, the bench line, the ncu launch list and one full ncu capture of l0_kernel.
# Usage (under gpurun): bash profiles/capture.sh <tag>
tag=${1:-run}
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; cat gpurun_out/bench_n1_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --contigs 100 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:l0_kernel -c 1 -f -o gpurun_out/l0_$tag \
    python bench.py --contigs 100 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/
